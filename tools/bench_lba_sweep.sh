# usage: bash tools/bench_lba_sweep.sh "<batch_steps> <inflight> [env...]" ...   (measurement recipe; not part of the product)
for cfg in "$@"; do
  set -- $cfg
  bs=$1; inf=$2; shift 2
  echo "== BATCH_STEPS=$bs INFLIGHT=$inf $@"
  env B200_BENCH_REPEATS=3 B200_BENCH_LBA_BATCH_STEPS=$bs B200_BENCH_LBA_INFLIGHT=$inf "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f ms/step %.2f e2e ms/step %.2f launches %d reps %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], [round(x,2) for x in d['repeat_stats']['ms_per_step']]))
    else: print(l[:200])
"
done
