#!/bin/bash
# usage: tools/sweep2.sh <workload> <steps> [extra bench args] : (F, G) map of cw_env_kernel
w=$1; k=$2; shift 2
for F in 2 3 4 5 6; do for G in 7 10 12 14 16; do
  r=$(CW_FRAME_BUFFERS=$F CW_GROUP=$G timeout 200 python bench.py --workload $w --steps $k --warmup 16 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))")
  echo "$w $* F=$F G=$G : $r"
done; done
