#!/bin/bash
# chained vs independent launches across batch sizes / workloads (one box)
run() { # label, args
  r=$(timeout 200 python bench.py ${@:2} --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); u=d.get('unchained'); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']), ('| unchained %.1f M/s %.3f' % (u['value']/1e6, u['roofline_frac'])) if u else '')")
  echo "$1 : $r"
}
run "cfg2        " --workload cfg2 --steps 12800
run "cfg4 ring 2 " --workload cfg4 --steps 640 --ring 2
run "cfg4 ring 2 no timeout" --workload cfg4 --steps 640 --ring 2 --max-steps 100000
run "N=65536" --workload cfg2 --envs 65536 --steps 1600 --ring 4
