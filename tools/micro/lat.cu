// Development aid: dependent-issue latencies on sm_100a for the instructions the in-place a-trous chain is made of.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dadd(double *o, double a, long long *t) { double x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = __dadd_rn(x, a); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_dmul(double *o, double a, long long *t) { double x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = __dmul_rn(x, a); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_dadd4(double *o, double a, long long *t) { double x = a, y = a + 1, z = a + 2, w = a + 3; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { x = __dadd_rn(x, a); y = __dadd_rn(y, a); z = __dadd_rn(z, a); w = __dadd_rn(w, a); } long long t1 = clock64(); o[threadIdx.x] = x + y + z + w; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_fadd(float *o, float a, long long *t) { float x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = __fadd_rn(x, a); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_ffma(float *o, float a, long long *t) { float x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = __fmaf_rn(x, a, a); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_fadd2(float *o, float a, long long *t) { unsigned long long x, b; asm("mov.b64 %0, {%1,%1};" : "=l"(x) : "f"(a)); b = x; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(b)); long long t1 = clock64(); float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x)); o[threadIdx.x] = lo + hi; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_cvt(float *o, float a, long long *t) { float x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) { double d = (double)x; d = __dadd_rn(d, 1.0); x = (float)d; } long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_lds(float *o, int a, long long *t) { __shared__ int s[1024]; for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i + a) & 1023; __syncwarp(); int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = s[x]; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_ldg(float *o, const int *g, long long *t) { int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __ldg(g + x); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_ldcg(float *o, const int *g, long long *t) { int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(g + x) : "memory"); x = v; } long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_rcp(float *o, float a, long long *t) { float x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = 1.0f / x; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
__global__ void k_imad(int *o, int a, long long *t) { int x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; i++) x = x * a + i; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) *t = t1 - t0; }
// store -> visible to another SM: ping-pong between two CTAs through L2
__global__ void k_pingpong(volatile int *flag, long long *t, int iters) {
  if (threadIdx.x) return; long long t0 = clock64();
  for (int i = 0; i < iters; i++) { if (blockIdx.x == 0) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(2 * i + 1) : "memory"); int v; do { asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag + 32) : "memory"); } while (v != 2 * i + 2); }
    else { int v; do { asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v != 2 * i + 1); asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flag + 32), "r"(2 * i + 2) : "memory"); } }
  long long t1 = clock64(); if (blockIdx.x == 0) *t = t1 - t0; }
int main() { void *o; long long *t; int *g; cudaMalloc(&o, 4096); cudaMallocManaged(&t, 8); cudaMalloc(&g, 4096 * 4); int h[4096]; for (int i = 0; i < 4096; i++) h[i] = (i + 32) & 4095; cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
#define RUN(name, call) call; cudaDeviceSynchronize(); call; cudaDeviceSynchronize(); printf("%-10s %.2f cycles/op\n", name, (double)*t / N);
  RUN("dadd", (k_dadd<<<1, 32>>>((double *)o, 1.5, t))) RUN("dmul", (k_dmul<<<1, 32>>>((double *)o, 1.0000001, t))) RUN("dadd x4", (k_dadd4<<<1, 32>>>((double *)o, 1.5, t)))
  RUN("fadd", (k_fadd<<<1, 32>>>((float *)o, 1.5f, t))) RUN("ffma", (k_ffma<<<1, 32>>>((float *)o, 0.5f, t))) RUN("fadd2", (k_fadd2<<<1, 32>>>((float *)o, 1.5f, t)))
  RUN("cvt f-d-f", (k_cvt<<<1, 32>>>((float *)o, 1.5f, t))) RUN("lds", (k_lds<<<1, 32>>>((float *)o, 1, t))) RUN("ldg L1", (k_ldg<<<1, 32>>>((float *)o, g, t)))
  RUN("ld L2", (k_ldcg<<<1, 32>>>((float *)o, g, t))) RUN("1/x", (k_rcp<<<1, 32>>>((float *)o, 1.5f, t))) RUN("imad", (k_imad<<<1, 32>>>((int *)o, 3, t)))
  int *flag; cudaMalloc(&flag, 1024); cudaMemset(flag, 0, 1024); k_pingpong<<<2, 32>>>(flag, t, 1000); cudaDeviceSynchronize(); printf("pingpong   %.1f cycles/round trip (2 hand-offs)\n", (double)*t / 1000);
  return 0; }
