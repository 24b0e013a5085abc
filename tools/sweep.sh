#!/bin/bash
# usage: scripts_sweep.sh <workload> ; sweeps launch tunables of cw_env_kernel, prints env-steps/s
w=$1
for F in 2 3 4; do for G in 0 4 7 8 14 16; do for C in 0 3 2; do
  r=$(CW_FRAME_BUFFERS=$F CW_GROUP=$G CW_CTAS_PER_SM=$C timeout 200 python bench.py --workload $w --steps 2560 --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))")
  echo "$w F=$F G=$G C=$C : $r"
done; done; done
