#!/bin/bash
# chained launches: ring depth F x group size G (cfg2 and cfg4), one box
run() { # label, args
  r=$(timeout 200 python bench.py ${@:2} --warmup 16 --no-cpu-baseline --no-e2e --no-unchained --no-incremental 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))")
  echo "$1 : $r"
}
for F in 2 3; do for G in 4 7 10 14; do
  export CW_FRAME_BUFFERS=$F CW_GROUP=$G
  run "cfg2 F=$F G=$G" --workload cfg2 --steps 12800
done; done
for F in 3 4; do for G in 8 16; do
  export CW_FRAME_BUFFERS=$F CW_GROUP=$G
  run "cfg4 F=$F G=$G" --workload cfg4 --steps 512
done; done
unset CW_FRAME_BUFFERS CW_GROUP
run "cfg2 auto" --workload cfg2 --steps 12800
run "cfg4 auto" --workload cfg4 --steps 512
