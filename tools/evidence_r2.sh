#!/bin/bash
# Round-2 evidence under gpurun: full GPU suite, compute-sanitizer (memcheck / racecheck / synccheck) over tests/sanitize_cases.py,
# ncu launch lists (front-end step, one 16-window local-BA batch) and one `--set full` capture of the hot kernels.
# Outputs land in gpurun_out/; summaries are copied to profiles/ (see profiles/README.md).
set -x
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r2_gpu_suite.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2_$tool.log python tests/sanitize_cases.py > gpurun_out/r2_${tool}_stdout.log 2>&1
  echo "$tool exit $?"; tail -2 gpurun_out/r2_${tool}_stdout.log; tail -3 gpurun_out/r2_$tool.log
done
export B200_BENCH_REPEATS=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_frontend.csv \
    python bench.py --steps 2 --warmup 3 --batch 64 --min-area 7100 --no-cpu-baseline --no-lba > gpurun_out/r2_ncu_frontend.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_lba.csv \
    python tools/lba_time.py stereo 1 16 > gpurun_out/r2_ncu_lba.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'fast_cells|describe_kernel|resize_kernel|topk_tc_kernel|resolve_kernel|select_kernel' -s 36 -c 24 -o gpurun_out/r2_prof_frontend -f \
    python bench.py --steps 2 --warmup 3 --batch 64 --min-area 7100 --no-cpu-baseline --no-lba > gpurun_out/r2_ncu_full_frontend.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'landmark_kernel|pose_rows|schur_mma|chol_solve|backsub' -s 20 -c 8 -o gpurun_out/r2_prof_lba -f \
    python tools/lba_time.py stereo 1 16 > gpurun_out/r2_ncu_full_lba.log 2>&1
# the round's bench lines: our arm (default flags), front end only, the reference arm
unset B200_BENCH_REPEATS
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
kill $SMI
timeout 600 python bench.py --no-lba --no-cpu-baseline --no-tracking > gpurun_out/r2_bench_frontend_only.json 2>> gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
python tools/bench_brief.py gpurun_out/r2_bench.json; python tools/bench_brief.py gpurun_out/r2_bench_frontend_only.json; tail -c 400 gpurun_out/r2_bench_reference.json
ls -la gpurun_out | tail -20
