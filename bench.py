#!/usr/bin/env python
"""bench.py — headline benchmark of the per-frame ray tracing path (BASELINE.json: Mrays/s & frames/s, Dragon
1920x1080 1 spp, 1/2/4/8 B200 next to the host-CPU reference path).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (the oracle restatement; the C#
                                                           reference cannot run here) on the box's host cores

A "step" is one frame of the hot path (TryFlipAndBlit: raygen -> trace -> TAA -> a-trous -> exposure -> cells).
`value` times K frames enqueued back to back with the scene resident in HBM (CUDA events on the launching stream), frames
in flight (one GPU: ycge_pipeline_config; N GPUs: frames in parallel over the ranks).  `e2e` times K frames submitted one
by one from the host -- SetCamera + submit per step, every frame's cells copied to pinned host memory inside the region,
the host a few frames ahead of the arrivals; `e2e_synchronous` is the strict drop-in call (SetCamera + TryFlipAndBlit,
one frame's latency per step).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FB_W, FB_H, SS = 480, 135, 4  # 1920x1080 internal (hiW = fbW*ss, hiH = fbH*2*ss, RaytraceRenderer.cs:86-87)
SCENE = "dragon"              # real xyzrgb_dragon.obj when present in assets/, else the procedural stand-in


def uses_bench_pose(scene_name):
    """The single-mesh scenes put the mesh at (0, 0.5, 1) behind the default camera: they are timed from the bench pose.
    Every other scene (the museum, the all-meshes scene, the voxel worlds, the primitive scenes) from its own default camera."""
    return scene_name in ("cow", "bunny", "teapot", "dragon") or scene_name.startswith("knot")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=SCENE)
    ap.add_argument("--fb", default=f"{FB_W}x{FB_H}")
    ap.add_argument("--ss", type=int, default=SS)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-seconds", type=float, default=240.0, help="cap of the timed part of --impl reference (complete frames)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled during the timed region (B200_PROFILING.md): NVML every 5 ms when pynvml is
    importable (a bench run lasts a fraction of a second), else nvidia-smi every 100 ms."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            R = n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown, n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap
            while not self._halt.is_set():
                try:
                    sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if rs & r else "Not Active" for r in R])
                except Exception:
                    pass
                self._halt.wait(0.005)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def algorithmic_bytes(st, W, H, fb_w, fb_h, ss):
    """bytes the reference's own data structures would move for the counted traversal events."""
    b_trav = 40 * st["top_nodes_popped"] + 40 * st["mesh_nodes_popped"] + 4 * st["leaf_refs"] + 48 * st["tris_tested"] + 32 * st["prims_tested"] + 8 * st["dda_cells"]
    step = max(2, 2 * ss)
    b_trace = b_trav + W * H * 41
    b_img = W * H * (41 + 87 + 3 * 53 + 12) + (W // step) * (H // step) * 13 + fb_w * fb_h * 32
    return b_trav, b_trace, b_trav + b_img


def cpu_leg(args, fb_w, fb_h, ss, seconds, steps=None, warmup=1):
    """The reference's CPU path on the workload itself: the same scene, pose and FULL internal resolution; the sample is bounded in
    FRAMES (each one a complete TryFlipAndBlit), never in pixels.  `steps`: render exactly that many frames (the reference arm,
    capped by `seconds`); else as many as fit `seconds` (>= 2).
    kind "reference": oracle/_ref/libycge_ref.so -- the reference's OWN C# source text, rewritten into C++ syntactically at build time
    (oracle/ref_transpile.py) and compiled; its thread pools run on all host cores as the reference's do (trace striped over the cores;
    TAA, a-trous, exposure serial, RaytraceRenderer.cs:218-227).  It has no ray counter (neither has the reference): rays per frame are
    counted by the oracle restatement on the same pose (identical frames, tests/test_reference_transpiled.py).
    kind "port": the oracle restatement itself, when the transpiled library was not built (no reference checkout at build time)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import yetanotherconsolegameengine_b200 as pkg
    from oracle_binding import Oracle
    import ref_binding
    cores = os.cpu_count() or 1
    scene = pkg.HostScene(args.scene)
    pose = pkg.BENCH_POSE if uses_bench_pose(args.scene) else None
    use_ref = ref_binding.available() and scene.n_textures == 0 and not os.environ.get("YCGE_CPU_PORT")
    if use_ref:
        o = Oracle(scene, fb_w, fb_h, ss)
        if pose:
            o.set_camera(*pose)
        o.render_frame(threads=cores)
        o.render_frame(threads=cores)
        rays_per_frame = o.stats()["rays"]  # frame 2 (frames differ by a fraction of a percent: the RNG stream is per frame)
        o.close()
        r = ref_binding.RefRenderer(scene, fb_w, fb_h, ss, threads=cores)
        if pose:
            r.set_camera(*pose)
        for _ in range(warmup):
            r.render_frame()
        frames, t0 = 0, time.perf_counter()
        while True:
            r.render_frame()
            frames += 1
            el = time.perf_counter() - t0
            if (steps is not None and frames >= steps) or (frames >= 2 and el >= seconds) or frames >= 64:
                break
        el = time.perf_counter() - t0
        r.close()
        return {"value": rays_per_frame * frames / el / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                "sample": f"{frames} complete frames (after {warmup} untimed) of the workload itself at the full {fb_w}x{fb_h} cells ss={ss} ({fb_w*ss}x{fb_h*2*ss} px) through oracle/_ref: the "
                          f"reference's own C# sources transpiled to C++ at build time; thread pools on all {cores} cores (trace), TAA / a-trous / exposure serial as in the reference",
                "frames_per_s": frames / el, "ms_per_stage": None, "seconds": el, "frames": frames}
    o = Oracle(scene, fb_w, fb_h, ss)
    if pose:
        o.set_camera(*pose)
    for _ in range(warmup):
        o.render_frame(threads=cores)
    rays, frames, t0 = 0, 0, time.perf_counter()
    stage = {"ms_trace": 0.0, "ms_taa": 0.0, "ms_atrous": 0.0, "ms_exposure": 0.0, "ms_cells": 0.0}
    while True:
        o.render_frame(threads=cores)
        st = o.stats()
        rays += st["rays"]
        frames += 1
        for k in stage:
            stage[k] += st[k]
        el = time.perf_counter() - t0
        if (steps is not None and frames >= steps) or (frames >= 2 and el >= seconds) or frames >= 64:
            break
    el = time.perf_counter() - t0
    return {"value": rays / el / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{frames} complete frames (after {warmup} untimed) of the workload itself: same scene/pose at the full {fb_w}x{fb_h} cells ss={ss} "
                      f"({fb_w*ss}x{fb_h*2*ss} px), reference threading (trace on all {cores} cores; TAA, a-trous, exposure serial, RaytraceRenderer.cs:218-227)",
            "frames_per_s": frames / el, "ms_per_stage": {k: v / frames for k, v in stage.items()}, "seconds": el, "frames": frames}


def main():
    args = parse()
    fb_w, fb_h = (int(x) for x in args.fb.lower().split("x"))
    ss = max(1, args.ss)
    W, H = fb_w * ss, fb_h * 2 * ss
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"{args.scene} scene, {W}x{H} internal ({fb_w}x{fb_h} cells, ss={ss}), 1 spp, reference bounce constants, TAA + a-trous + auto-exposure, {'bench pose' if uses_bench_pose(args.scene) else 'default pose'}"

    if args.impl == "reference":
        if rank != 0:
            return 0
        # every step is one complete frame of the headline configuration on all host cores; the run is capped at
        # --ref-seconds of CPU time (a 1080p frame costs ~3 s on 16 cores, so the driver's 20 + 5 steps fit)
        leg = cpu_leg(args, fb_w, fb_h, ss, seconds=args.ref_seconds, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
        line = {"impl": "reference", "metric": "Mrays/s", "value": leg["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "steps_run": leg["frames"], "ms_per_step": 1e3 * leg["seconds"] / leg["frames"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "frames_per_s": leg["frames_per_s"], "same_config": True,
                "config": {"workload": workload, "note": "the reference's CPU path on all host cores (cpu_baseline.kind says which build of it: its own sources transpiled, or the restatement); every step is a complete frame at the workload's full resolution"},
                "stage_ms": leg["ms_per_stage"],
                "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": leg["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import yetanotherconsolegameengine_b200 as pkg
    from yetanotherconsolegameengine_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the ray tracing path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries the ONE JSON line: NCCL prints its version banner there when the first communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n = max(1, world)
    rc = 0

    # what a scene switch costs (RaytraceEntity.cs:234-246): the host builds the scene (parse, normalise, both SAH trees),
    # then the renderer is created and everything is uploaded and flattened; timed once, here, on the box's host
    t_sw = time.perf_counter()
    scene = pkg.HostScene(args.scene)
    scene_switch = {"host_build_ms": 1e3 * (time.perf_counter() - t_sw)}
    pose = pkg.BENCH_POSE if uses_bench_pose(args.scene) else scene.default_camera()[:3]
    stream = torch.cuda.Stream()
    pinned = torch.empty((fb_h * fb_w * api.CELL_DTYPE.itemsize,), dtype=torch.uint8, pin_memory=True)
    cells = pinned.numpy().view(api.CELL_DTYPE).reshape(fb_h, fb_w)
    h2d = C.sizeof(C.c_float) * 32 + 64  # FrameConsts + launch parameters; the scene stays resident
    d2h = fb_w * fb_h * api.CELL_DTYPE.itemsize

    if n == 1:
        torch.cuda.synchronize()
        t_sw = time.perf_counter()
        r = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local_rank)
        torch.cuda.synchronize()
        scene_switch["create_and_upload_ms"] = 1e3 * (time.perf_counter() - t_sw)
        if scene.n_meshes > 0:
            # the mesh trees once more, by the library's host builder and ON THE DEVICE (same trees, node for node);
            # the timed frames below are rendered from the device-built ones
            r.rebuild_meshes_on_device()  # untimed: the first call sizes the scratch arena
            tm = r.rebuild_meshes_on_device(also_time_host_builder=True)
            scene_switch["mesh_trees_host_builder_ms"] = tm["host"]
            scene_switch["mesh_trees_device_builder_ms"] = tm["device"]
        r.SetCamera(*pose)
        r.set_stream(stream.cuda_stream)
        # untimed: one frame with the reference-defined event counters (feeds the algorithmic-bytes roofline)
        r.render_frame_stats()
        st_events = r.stats()
        r.reset_history()
        with torch.cuda.stream(stream):
            r.render_frames_async(max(3, args.warmup))
        r.wait()
        st0 = r.stats()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(stream)
        r.render_frames_async(args.steps)
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        st1 = r.stats()
        serial = {"ms_per_step": ms / args.steps, "frames_per_s": args.steps / (ms / 1e3), "mrays_per_s": (st1["rays_total"] - st0["rays_total"]) / (ms / 1e3) / 1e6}
        # per-stage device times of a steady-state frame on the strictly serial schedule (events recorded by the library on
        # the same stream; with frames in flight the kernels of different frames overlap and a per-kernel time means little)
        stage_ms = {k: st1[k] for k in ("ms_trace", "ms_taa", "ms_atrous", "ms_atrous_chain", "ms_exposure", "ms_cells", "ms_total")}
        launches_per_frame = st1["kernel_launches"]
        # `value`: the same K frames with up to `slots` frames in flight on this GPU (ycge_pipeline_config; bit-identical)
        slots = int(os.environ.get("YCGE_SLOTS", "3"))
        if slots > 1:
            r.pipeline_config(slots)
            r.render_frames_async(max(3, args.warmup))
            r.wait()
            st0 = r.stats()
            torch.cuda.synchronize()
            ev0.record(stream)
            r.render_frames_async(args.steps)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            st1 = r.stats()
        clocks = sampler.stop()
        rays_timed = st1["rays_total"] - st0["rays_total"]
        # streaming end to end: camera in, frame enqueued, cells copied to pinned host memory as part of the frame; the host
        # waits for frame N - slots + 1 before it submits frame N + 1 (every frame's cells land on the host inside the region)
        ring = [torch.empty((fb_h * fb_w * api.CELL_DTYPE.itemsize,), dtype=torch.uint8, pin_memory=True) for _ in range(max(1, slots))]
        ring_np = [t.numpy().view(api.CELL_DTYPE).reshape(fb_h, fb_w) for t in ring]
        def stream_frames(k):
            ids = []
            for i in range(k):
                if len(ids) == max(1, slots):
                    r.frame_wait(ids.pop(0))
                r.SetCamera(*pose)
                ids.append(r.submit_frame(ring_np[i % len(ring_np)]))
            for fid in ids:
                r.frame_wait(fid)
        stream_frames(max(3, args.warmup))
        sts0 = r.stats()
        t0 = time.perf_counter()
        stream_frames(args.steps)
        stream_s = time.perf_counter() - t0
        sts1 = r.stats()
        streaming = {"value": (sts1["rays_total"] - sts0["rays_total"]) / stream_s / 1e6, "unit": "Mrays/s", "frames_per_s": args.steps / stream_s,
                     "frames_in_flight": slots, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "api": "SetCamera + ycge_submit_frame per step, ycge_frame_wait %d frames later" % (slots - 1)}
        # e2e: the public IConsoleRenderer call per step, camera in, cells out to pinned host memory
        for _ in range(3):
            r.SetCamera(*pose)
            r.TryFlipAndBlit(cells)
        st2 = r.stats()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r.SetCamera(*pose)
            r.TryFlipAndBlit(cells)
        e2e_s = time.perf_counter() - t0
        st3 = r.stats()
        e2e_rays = st3["rays_total"] - st2["rays_total"]
        r.close()
    else:
        # one process per GPU: contiguous row tiles of cell rows, scene replicated, boundary rows of the in-place pass
        # handed rank -> rank+1, exposure samples all-reduced, cell tiles gathered on rank 0 (sharding.py)
        from yetanotherconsolegameengine_b200 import sharding
        tiles = [sharding.tile_rows(r, n, fb_h) for r in range(n)]
        with torch.cuda.stream(stream):
            # untimed set-up: four rounds of load balancing.  The trace cost of a row is very uneven (sky vs mesh), so equal
            # tiles leave most ranks waiting for the one that holds the mesh; tiles are re-cut from the measured trace times.
            for balance_round in range(5):
                row0, rows = tiles[rank]
                b = sharding.CudaTileBackend(scene, fb_w, fb_h, ss, row0, rows, local_rank)
                sr = sharding.ShardedRenderer(b, rank, n, fb_w, fb_h, peers=not os.environ.get("YCGE_NO_PEERS"), tiles=tiles)
                sr.SetCamera(*pose)
                if balance_round == 4 or os.environ.get("YCGE_NO_BALANCE"):
                    break
                for _ in range(3):
                    sr.render_device()
                torch.cuda.synchronize()
                tr = [None] * n
                dist.all_gather_object(tr, float(b.r.stats()["ms_trace"]))
                tiles = sharding.balanced_tiles(tiles, tr, fb_h, min_rows=max(1, -(-4 // (2 * ss))))
                sr.close()
                del sr, b
            peer_handoff = sr.peer_handoff
            # the event counters of the whole frame come from an unsharded frame on rank 0's GPU (untimed)
            st_events = None
            if rank == 0:
                r1 = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local_rank)
                r1.SetCamera(*pose)
                r1.render_frame_stats()
                st_events = r1.stats()
                r1.close()
            # `value`: the asynchronous path.  Default: frames in parallel (FRONT on row tiles, BACK + FINISH of whole frames
            # round-robin over the ranks, sharding.FrameParallelRenderer); YCGE_MULTI=rowpipe: row tiles with the wavefront
            # handed from rank to rank and frames pipelined over the ranks (sharding.ShardedRenderer.render_pipelined)
            mode = os.environ.get("YCGE_MULTI", "frames")
            pipelined = sr.peer_handoff and not os.environ.get("YCGE_NO_PIPELINE")
            fp, front_tiles = None, None
            if mode == "frames":
                tr = [None] * n
                for _ in range(3):
                    sr.render_device()
                torch.cuda.synchronize()
                dist.all_gather_object(tr, float(b.r.stats()["ms_trace"]))
                front_tiles = sharding.balanced_tiles(tiles, tr, fb_h, per_row_ms=0.001)
                fp = sharding.FrameParallelRenderer(scene, rank, n, fb_w, fb_h, ss, local_rank, tiles=front_tiles, back_slots=int(os.environ.get("YCGE_BACK_SLOTS", "2")))
                fp.SetCamera(*pose)
                run = lambda k: fp.render(k)
            elif pipelined:
                run = lambda k: sr.render_pipelined(k)
            else:
                run = lambda k: [sr.render_device() for _ in range(k)]
            # warm-up: every back slot of every rank, every staging-ring entry and every NCCL channel is used at least twice
            # before the timed region (the frame-parallel pipeline is N x back slots frames deep)
            back_slots = fp.S if fp is not None else 2
            n_warm = max(3, args.warmup, 2 * n * back_slots + 2)
            run(n_warm)
            torch.cuda.synchronize()
            st0 = b.r.stats()
            sampler = ClockSampler(local_rank)
            sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            ev0.record(stream)
            run(args.steps)
            ev1.record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            ms_local = ev0.elapsed_time(ev1)
            clocks = sampler.stop()
            fp_stage = None
            if fp is not None:
                fs = fp.b.front_r.stats()
                fp_stage = {"front_ms_trace": round(fs["ms_trace"], 3), "front_ms_taa": round(fs["ms_taa"], 3)}
            # per-stage times of the lock-step row-tile path (the e2e path below)
            for _ in range(2):
                sr.render_device()
            torch.cuda.synchronize()
            st1 = b.r.stats()
            stage_ms = {k: st1[k] for k in ("ms_trace", "ms_taa", "ms_atrous", "ms_atrous_chain", "ms_exposure", "ms_cells", "ms_total")}
            launches_per_frame = st1["kernel_launches"]
            # e2e: camera in, assembled cells out to pinned host memory on rank 0, every step
            def e2e_step():
                sr.SetCamera(*pose)
                g = sr.render_device()
                if rank == 0:
                    for rr, (t0_, tn) in enumerate(sr.tiles):
                        nb = tn * fb_w * api.CELL_DTYPE.itemsize
                        pinned[t0_ * fb_w * 32:t0_ * fb_w * 32 + nb].copy_(g[rr][:nb], non_blocking=True)
                torch.cuda.current_stream().synchronize()
            for _ in range(3):
                e2e_step()
            st2 = b.r.stats()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            dist.barrier()
            e2e_local = time.perf_counter() - t0
            st3 = b.r.stats()
            # streaming end to end on the frame-parallel path: camera in on every rank per frame, every frame's cells copied
            # to pinned host memory on rank 0, the host at most N x back slots + 2 frames ahead of the arrivals
            stream_local = None
            if fp is not None:
                n_ring = n * fp.S + 2  # as many host buffers as frames can be in flight (N ranks x back slots), or the host's pacing caps them
                ring = [torch.empty((fb_h * fb_w * api.CELL_DTYPE.itemsize,), dtype=torch.uint8, pin_memory=True) for _ in range(n_ring)] if rank == 0 else None
                cam = lambda f: fp.SetCamera(*pose)
                fp.render(n_warm, set_camera=cam, host_ring=ring)
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                fp.render(args.steps, set_camera=cam, host_ring=ring)
                torch.cuda.synchronize()
                dist.barrier()
                stream_local = time.perf_counter() - t0
            # ---- parity of the objects just timed (untimed): the next frames of the sharded paths against an unsharded context on
            # rank 0's GPU that has rendered the same number of frames from the same start (static camera); cells bit for bit
            parity = {}
            n_chk = 2
            if fp is not None:
                done = fp.frame
                got = fp.render(n_chk, collect=True)
                torch.cuda.synchronize()
                if rank == 0:
                    full = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local_rank)
                    full.SetCamera(*pose)
                    if done:
                        full.render_frames_async(done)
                        full.wait()
                    parity["frames_in_parallel"] = all(fp.cells_host(g).tobytes() == full.TryFlipAndBlit().tobytes() for g in got)
                    parity["frames_in_parallel_checked"] = "frames %d..%d" % (done + 1, done + n_chk)
                    full.close()
            done = b.r.stats()["frames"]
            got = [sr.TryFlipAndBlit() for _ in range(n_chk)]
            if rank == 0:
                full = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local_rank)
                full.SetCamera(*pose)
                if done:
                    full.render_frames_async(int(done))
                    full.wait()
                parity["row_tiles_lock_step"] = all(g.tobytes() == full.TryFlipAndBlit().tobytes() for g in got)
                parity["row_tiles_lock_step_checked"] = "frames %d..%d" % (done + 1, done + n_chk)
                full.close()
            dist.barrier()
            # ---- the same N GPUs behind ONE context of the library (ycge_config.n_devices, include/ycge.h): what the C# host gets
            # from a single ycge_create.  Rank 0 drives all GPUs from one thread while the other ranks wait at the barrier.
            in_library = None
            torch.cuda.synchronize()
            store = dist.distributed_c10d._get_default_store() # a HOST-side wait: an NCCL barrier would keep a kernel spinning on every other GPU
            if rank != 0:
                store.wait(["ycge_in_library_done"])
            if rank == 0 and os.environ.get("YCGE_NO_INLIB"):
                store.set("ycge_in_library_done", "1")
            if rank == 0 and not os.environ.get("YCGE_NO_INLIB"):
                try:
                    ml = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, devices=list(range(n)))
                    ml.SetCamera(*pose)
                    depth = 2 * n
                    ring_ml = [torch.empty((fb_h * fb_w * api.CELL_DTYPE.itemsize,), dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
                    ring_ml_np = [t_.numpy().view(api.CELL_DTYPE).reshape(fb_h, fb_w) for t_ in ring_ml]
                    def ml_frames(k):
                        ids = []
                        for i_ in range(k):
                            if len(ids) == depth:
                                ml.frame_wait(ids.pop(0))
                            ml.SetCamera(*pose)
                            ids.append(ml.submit_frame(ring_ml_np[i_ % depth]))
                        for fid in ids:
                            ml.frame_wait(fid)
                    ml_frames(n_warm)
                    t0 = time.perf_counter()
                    ml_frames(args.steps)
                    ml_s = time.perf_counter() - t0
                    done = n_warm + args.steps
                    full = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local_rank)
                    full.SetCamera(*pose)
                    full.render_frames_async(done)
                    full.wait()
                    same = all(ml.TryFlipAndBlit().tobytes() == full.TryFlipAndBlit().tobytes() for _ in range(2))
                    full.close()
                    ml.close()
                    in_library = {"frames_per_s": args.steps / ml_s, "value": st_events["rays"] * args.steps / ml_s / 1e6, "unit": "Mrays/s", "parity_vs_one_device": same,
                                  "d2h_bytes_per_step": d2h, "api": "ONE ycge_create with n_devices = %d, SetCamera + ycge_submit_frame per step from one host thread, ycge_frame_wait %d frames later; "
                                                                    "every frame's cells copied to pinned host memory" % (n, depth)}
                    parity["in_library"] = same
                except Exception as e:  # noqa: BLE001
                    in_library = {"error": str(e)[:300]}
                store.set("ycge_in_library_done", "1")
            dist.barrier()
        # max over ranks of the device time; rays summed over ranks (halo rows are traced redundantly and counted as
        # work done — rays/frame of the UNSHARDED frame is what the metric divides by, so use the unsharded count)
        all_stage = [None] * n
        dist.all_gather_object(all_stage, dict({k: round(v, 3) for k, v in stage_ms.items()}, **(fp_stage or {})))
        t = torch.tensor([ms_local, e2e_local, stream_local or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        streaming = None
        if stream_local:
            stream_s = float(t[2])
        rays_frame = torch.tensor([st_events["rays"] if rank == 0 else 0], dtype=torch.int64, device="cuda")
        dist.broadcast(rays_frame, 0)
        rays_timed = int(rays_frame[0]) * args.steps
        e2e_rays = int(rays_frame[0]) * args.steps
        if stream_local:
            streaming = {"value": e2e_rays / stream_s / 1e6, "unit": "Mrays/s", "frames_per_s": args.steps / stream_s, "frames_in_flight": "host at most %d frames ahead; %d back slots per rank" % (n * fp.S + 2, fp.S),
                         "h2d_bytes_per_step": h2d * n, "d2h_bytes_per_step": d2h,
                         "api": "SetCamera on every rank + FrameParallelRenderer.render per step, every frame's cells copied to pinned host memory on rank 0"}
        if fp is not None:
            fp.close()
        sr.close()
    fps = args.steps / (ms / 1e3)
    mrays = rays_timed / (ms / 1e3) / 1e6
    e2e_mrays = e2e_rays / e2e_s / 1e6

    # roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    b_trav, b_trace, b_frame = algorithmic_bytes(st_events, W, H, fb_w, fb_h, ss)
    # the dominant single kernel: the trace megakernel or the wavefront kernel of the in-place à-trous pass
    rows_here = H if n == 1 else min(H, tiles[0][1] * 2 * ss + 16)
    if stage_ms["ms_trace"] >= stage_ms["ms_atrous_chain"]:
        kname, kbytes, kms = "trace_kernel", b_trace * rows_here // H, stage_ms["ms_trace"]
    else:
        kname, kbytes, kms = "atrous_wave_kernel (wavefront of the in-place a-trous iteration)", W * rows_here * 53, stage_ms["ms_atrous_chain"]
    achieved = kbytes / (kms / 1e3) / 1e9 if kms > 0 else 0.0
    # DRAM traffic of that kernel per launch: only from an `ncu --set full` capture of THIS build of the library and this
    # command (tools/summarize_ncu.py records the sha256 of the library's SOURCES -- nvcc's output is not byte-reproducible -- beside
    # dram__bytes_read.sum + dram__bytes_write.sum);
    # a capture of any other build is not this run's traffic -> null
    traffic, traffic_src = None, None
    try:
        import hashlib
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        sha = api.library_source_digest()
        if n == 1 and cap.get("source_sha256") == sha and cap.get("workload") == f"{args.scene} {fb_w}x{fb_h} ss={ss}":
            key = next(k for k in cap["kernels"] if k.startswith("trace_" if kname == "trace_kernel" else kname.split(" ")[0]))
            traffic = float(cap["kernels"][key]["dram_bytes"])
            traffic_src = "profiles/ncu_traffic.json (%s; dram__bytes_read.sum + dram__bytes_write.sum of %s, same library sources, sha256 %s)" % (cap.get("captured", "?"), key, sha[:12])
    except Exception:
        pass
    # both candidates side by side (the dominant one is the `roofline` object itself)
    both = []
    for nm, by, tm in (("trace_kernel", b_trace * rows_here // H, stage_ms["ms_trace"]), ("atrous_wave_kernel", W * rows_here * 53, stage_ms["ms_atrous_chain"])):
        gbs = by / (tm / 1e3) / 1e9 if tm > 0 else 0.0
        both.append({"kernel": nm, "kernel_ms": tm, "algorithmic_bytes_per_launch": by, "achieved": gbs, "frac": gbs / peak})
    roofline = {"bound": "hbm", "kernel": kname, "candidates": both, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": kbytes, "kernel_ms": kms,
                "frame": {"algorithmic_bytes": b_frame, "achieved_GBps": b_frame * fps / 1e9, "frac": b_frame * fps / 1e9 / peak},
                "note": "algorithmic bytes = SURVEY 8(d) per-pixel figures of the reference's own layout; both candidate kernels are latency-bound (divergent traversal; serial wavefront), "
                        "not HBM-bound: see DESIGN.md and profiles/"}

    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        leg = cpu_leg(args, fb_w, fb_h, ss, seconds=args.cpu_seconds)
        cpu = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample", "frames_per_s", "ms_per_stage")}

    sync_e2e = {"value": e2e_mrays, "unit": "Mrays/s", "frames_per_s": args.steps / e2e_s, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "SetCamera + TryFlipAndBlit per step, synchronous (the drop-in call; latency of one frame)"}
    if rank == 0:
        line = {"metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": n, "steps": args.steps, "warmup": max(3, args.warmup) if n == 1 else n_warm,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "frames_per_s": fps, "rays_per_frame": rays_timed / args.steps, "mpaths_per_s": W * H * fps / 1e6,
                "config": {"workload": workload, "scene": scene.name, "triangles": scene.counts()["triangles"], "parallelism": f"row-tiles x{n}" + ((", %d frames in flight on the GPU (value, e2e); strictly serial (e2e_synchronous, serial_schedule, stage_ms, roofline kernel time)" % slots) if n == 1 else (", frames in parallel: FRONT on row tiles, BACK + FINISH of whole frames round-robin over the ranks (value, e2e); row tiles in lock step" if mode == "frames" else
                            (", frames pipelined over ranks (value); lock-step" if pipelined else "")) + (", peer hand-off (e2e_synchronous)" if peer_handoff else ", NCCL send/recv hand-off (e2e_synchronous)")),
                           "l2": "per-frame working set (8 float4 image planes = %d MB) exceeds the 126 MB L2; no explicit flush" % (W * H * 128 // (1 << 20))},
                "stage_ms": stage_ms, **({"stage_ms_ranks": all_stage, "peer_handoff": peer_handoff, "frame_pipelining": pipelined, "multi_gpu_mode": mode, "tiles_cell_rows": [t[1] for t in tiles],
                                              "front_tiles_cell_rows": [t[1] for t in front_tiles] if front_tiles else None} if n > 1 else {}),
                # e2e: frames in flight (the host submits frame N while earlier frames finish; every frame's cells are copied to pinned
                # host memory inside the timed region).  e2e_synchronous: the strict drop-in call, one frame's latency per step.
                # on N GPUs the call a user makes is ONE ycge_create over all of them (in_library); the one-process-per-GPU NCCL path is e2e_per_process
                "e2e": ({"value": in_library["value"], "unit": "Mrays/s", "frames_per_s": in_library["frames_per_s"], "h2d_bytes_per_step": h2d * n, "d2h_bytes_per_step": d2h, "api": in_library["api"]}
                        if (n > 1 and in_library and "value" in in_library) else (streaming if streaming else sync_e2e)),
                **({"e2e_per_process": streaming} if (n > 1 and streaming) else {}), "e2e_synchronous": sync_e2e,
                **({"serial_schedule": serial, "frames_in_flight": slots} if n == 1 else {}),
                "gpu_launches": launches_per_frame * args.steps * n,
                **({"parity_vs_unsharded": bool(parity) and all(v for k, v in parity.items() if not k.endswith("_checked")), "parity_detail": parity} if n > 1 else {}),
                **({"in_library": in_library} if n > 1 else {}),
                "scene_switch_ms": scene_switch,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
        if n > 1 and not line["parity_vs_unsharded"]:
            sys.stderr.write("bench.py: the sharded frame differs from the unsharded frame: %r\n" % (parity,))
            rc = 3
    if dist is not None:
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
