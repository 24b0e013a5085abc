#!/bin/bash
# 20-step windows (the driver's setting) vs steady state over group size G / CTAs per SM / first-frame split, cfg2, one box
run() { # label
  r=$(timeout 200 python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline --no-e2e --no-unchained --no-incremental --no-closed-loop 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('20-step %.2f us frac %.3f | steady %.2f us frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['steady_state']['ms_per_step']*1e3, d['steady_state']['roofline_frac']))")
  echo "$1 : $r"
}
run "auto"
for G in 4 5 6 7 8 10; do export CW_GROUP=$G; run "G=$G"; done; unset CW_GROUP
for C in 3 5; do export CW_CTAS_PER_SM=$C; run "ctas/SM=$C"; done; unset CW_CTAS_PER_SM
for S in 1 2 8; do export CW_FIRST_SPLIT=$S; run "first_split=$S"; done; unset CW_FIRST_SPLIT
run "auto again"
