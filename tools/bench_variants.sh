#!/bin/bash
# bench.py under a few scheduling variants of the local-BA pipeline (stream priority, windows per batch, batches in flight)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" B200_BENCH_REPEATS=3 timeout 300 python bench.py --no-cpu-baseline --no-tracking > gpurun_out/r2_var_$name.json 2> gpurun_out/r2_var_$name.err
  echo "== $name: $*"; python tools/bench_brief.py gpurun_out/r2_var_$name.json | head -2
}
run base X=1
run prio_high B200_LBA_PRIORITY=high
run prio_normal B200_LBA_PRIORITY=normal
run b8_f2 B200_BENCH_LBA_BATCH_STEPS=2 B200_BENCH_LBA_INFLIGHT=2
run b16_f2 B200_BENCH_LBA_BATCH_STEPS=4 B200_BENCH_LBA_INFLIGHT=2
run b4_f2 B200_BENCH_LBA_INFLIGHT=2
run b4_f8 B200_BENCH_LBA_INFLIGHT=8
