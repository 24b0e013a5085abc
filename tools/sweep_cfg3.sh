#!/bin/bash
# cfg3 (compact step kernel): launch bounds x refill CTAs
run() { python bench.py --workload cfg3 --only --no-cpu-baseline --steps 256 --warmup 5 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('  us/launch %.2f  %.2f G steps/s  rollout %.1f G' % (b['ms_per_step']*1e3, b['value']/1e9, b['rollout']['value']/1e9))"; }
for lib in default ab/lib_mb6.so ab/lib_mb8.so; do
  for r in 0 64 128 256; do
    echo "lib=$lib refill_ctas=$r"
    if [ $lib = default ]; then CW_REFILL_CTAS=$r run; else CW_LIB_PATH=$lib CW_REFILL_CTAS=$r run; fi
  done
done
