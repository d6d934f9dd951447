#!/bin/bash
# First GPU call of the next round (one B200, ~12 min): everything written after round 1's GPU minutes ran out.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/round2_first_gpu_call.sh'
# 1. the parity tests on new inputs (island world, all-meshes scene, museum, day/night, animated entities);
# 2. the whole GPU suite again (the host library changed since it last ran there, the device library did not);
# 3. bench lines for the scenes that did not exist yet: the island world at config C4's resolution, the museum at 1080p;
# 4. config C3's scenes, which have parity tests but no bench line yet.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_new_inputs_gpu.py -q -m gpu > gpurun_out/r02_new_inputs.log 2>&1; echo "new inputs: rc $?"
tail -5 gpurun_out/r02_new_inputs.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_suite.log 2>&1; echo "gpu suite: rc $?"
tail -3 gpurun_out/r02_gpu_suite.log
timeout 300 python bench.py --scene voxel_island --fb 320x90 --ss 8 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_c4_island.json 2> gpurun_out/r02_bench_c4_island.err
timeout 300 python bench.py --scene museum --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_museum.json 2> gpurun_out/r02_bench_museum.err
# 4. SURVEY 8(d) config C3 (showcase scenes, bunny, teapot at 1080p), one line each
for sc in cylinders_disks_triangles boxes bunny teapot all_meshes; do
    timeout 200 python bench.py --scene $sc --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_c3_$sc.json 2> gpurun_out/r02_bench_c3_$sc.err
    tail -c 300 gpurun_out/r02_bench_c3_$sc.json; echo
done
tail -c 600 gpurun_out/r02_bench_c4_island.json; echo; tail -c 600 gpurun_out/r02_bench_museum.json
