#!/bin/bash
# usage: tools/sweep.sh <workload> [steps] ; sweeps launch tunables of cw_env_kernel, prints env-steps/s
w=$1; k=${2:-2560}
for F in 2 3 4; do for G in 0 4 8 12; do for FS in 1 4; do
  r=$(CW_FRAME_BUFFERS=$F CW_GROUP=$G CW_FIRST_SPLIT=$FS timeout 200 python bench.py --workload $w --steps $k --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))")
  echo "$w F=$F G=$G FS=$FS : $r"
done; done; done
