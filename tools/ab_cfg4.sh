#!/bin/bash
# A/B on one box (independent launches): builds given as tools/ab/libcw_<tag>.so vs current
run() { # label, lib, args
  r=$(CW_LIB_PATH=$2 timeout 200 python bench.py ${@:3} --warmup 16 --no-cpu-baseline --no-e2e --no-chain 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))")
  echo "$1 : $r"
}
for tag in "$@"; do
  run "cfg4 $tag" tools/ab/libcw_$tag.so --workload cfg4 --steps 512
  run "cfg2 $tag" tools/ab/libcw_$tag.so --workload cfg2 --steps 12800
done
run "cfg4 current" gym_craftingworld_b200/libcw_b200.so --workload cfg4 --steps 512
run "cfg2 current" gym_craftingworld_b200/libcw_b200.so --workload cfg2 --steps 12800
