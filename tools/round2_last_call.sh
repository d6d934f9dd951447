#!/bin/bash
# Last GPU call of round 2 on ONE B200: the GPU suite, the profile round of the shipped library (bench line, ncu launch list, ncu --set
# full), an ncu capture of the trace kernel on the voxel world, and the bench lines of the configurations the occupancy bytes touch.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r02g_gpu_suite.log
bash tools/profile_round.sh r02g > gpurun_out/profile_round_r02g.log 2>&1; tail -c 600 gpurun_out/bench_r02g.json; echo
timeout 200 ncu --set full --clock-control none --import-source on -k regex:trace_stream -s 5 -c 1 -f -o gpurun_out/prof_r02g_voxel python tools/stage_times.py voxel_world:320:90:8 > gpurun_out/r02g_ncu_voxel.log 2>&1; tail -2 gpurun_out/r02g_ncu_voxel.log
for spec in "c4_voxel_world --scene voxel_world --fb 320x90 --ss 8" "c4_voxel_island --scene voxel_island --fb 320x90 --ss 8" "museum --scene museum"; do
    set -- $spec; tag=$1; shift
    timeout 200 python bench.py "$@" --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02g_bench_$tag.json 2> gpurun_out/r02g_bench_$tag.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/r02g_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['frames_per_s'],1), 'fps', round(d['value'],1), 'Mrays/s e2e', round(d['e2e']['frames_per_s'],1), 'sync', round(d['e2e_synchronous']['frames_per_s'],1), {k: round(v,3) for k,v in d['stage_ms'].items()})
" || echo "$tag FAILED"
done
