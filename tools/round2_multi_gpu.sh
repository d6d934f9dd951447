#!/bin/bash
# Multi-GPU bench lines of round 2: bash tools/round2_multi_gpu.sh N   (under gpurun --gpus N)
set -u
N=$1
mkdir -p gpurun_out
run() { # tag, bench arguments
    tag=$1; shift
    timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" --steps 20 --warmup 5 > gpurun_out/r02d_bench_${tag}_n$N.json 2> gpurun_out/r02d_bench_${tag}_n$N.err
    python - "$tag" "$N" <<'PY'
import json, sys
tag, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/r02d_bench_{tag}_n{n}.json").read().strip().splitlines()[-1])
    print(tag, "N", n, "frames/s", round(d["frames_per_s"], 1), "Mrays/s", round(d["value"], 1), "e2e", round(d["e2e"]["frames_per_s"], 1), "sync", round(d["e2e_synchronous"]["frames_per_s"], 1),
          "parity", d.get("parity_vs_unsharded"), "warmup", d.get("warmup"))
except Exception as e:
    print(tag, "FAILED", e)
PY
}
if [ "${ONLY:-}" != "second" ]; then run dragon; fi
if [ "${ONLY:-}" = "first" ]; then exit 0; fi
if [ "$N" = "8" ]; then run c5_dragon_4k --scene dragon --fb 480x135 --ss 8; else run c4_voxel_world --scene voxel_world --fb 320x90 --ss 8; fi
