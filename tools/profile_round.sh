#!/bin/bash
# Runs on the GPU box (under gpurun): parity probe, bench line, ncu launch list, ncu --set full of the top kernels.
# Outputs land in gpurun_out/ (copied into profiles/ by hand once read).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python bench.py --steps 32 --warmup 4 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'trace|atrous|taa_kernel|cells_kernel|exposure' -s 40 -c 12 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
tail -c 1500 gpurun_out/bench_$TAG.json
