#!/bin/bash
# ncu recipe (B200_PROFILING.md): launch list of one bench run + one full capture of each extractor kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --batch 64 --min-area 7100 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fast_cells|blur_kernel|describe_kernel|resize_kernel|topk_kernel|resolve_kernel|select_kernel' \
    -s 39 -c 13 -o gpurun_out/prof_r1 -f python bench.py --steps 2 --warmup 3 --batch 64 --min-area 7100 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
