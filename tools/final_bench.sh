#!/bin/bash
# Final round numbers without the ncu passes (those are in tools/profile_orb.sh): our arm, front end only, reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_r1.csv &
SMI=$!
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
kill $SMI
python bench.py --steps 20 --warmup 3 --no-lba --no-cpu-baseline > gpurun_out/bench_r1_final_frontend_only.json 2>> gpurun_out/bench_r1_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 39 -c 13 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 3 --batch 64 --min-area 7100 --no-cpu-baseline --no-lba > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lba_launches_r1_final.csv python tools/lba_time.py stereo 1 > gpurun_out/lba_ncu.log 2>&1
