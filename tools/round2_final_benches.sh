#!/bin/bash
# Final bench lines of round 2 on ONE B200 (about 6 minutes): the BASELINE configurations other than the default workload, and the
# reference arm on the default workload.  /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_final_benches.sh'
set -u
mkdir -p gpurun_out
run() { # tag, bench arguments
    tag=$1; shift
    timeout 300 python bench.py "$@" --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02c_bench_$tag.json 2> gpurun_out/r02c_bench_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02c_bench_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "frames/s", round(d["frames_per_s"], 1), "Mrays/s", round(d["value"], 1), "e2e", round(d["e2e"]["frames_per_s"], 1), "sync", round(d["e2e_synchronous"]["frames_per_s"], 1),
          "stage", {k: round(v, 3) for k, v in d["stage_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e)
PY
}
run c1_cornell --scene cornell --fb 240x135 --ss 1
run c2_mirror_spheres --scene mirror_spheres
for sc in cylinders_disks_triangles boxes bunny teapot all_meshes; do run c3_$sc --scene $sc; done
run c4_voxel_world --scene voxel_world --fb 320x90 --ss 8
run c4_voxel_island --scene voxel_island --fb 320x90 --ss 8
run c5_dragon_4k --scene dragon --fb 480x135 --ss 8
run museum --scene museum
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err
tail -c 700 gpurun_out/r02c_bench_reference.json
