#!/bin/bash
# ring depth F vs batch size (auto G), 21x21 pixel worlds
for N in 2048 4096 8192 16384 32768 65536 131072; do for F in 2 3 4; do
  k=$((40000000 / N)); [ $k -gt 12800 ] && k=12800
  r=$(CW_FRAME_BUFFERS=$F timeout 200 python bench.py --workload cfg2 --envs $N --steps $k --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M/s %.2f us frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))")
  echo "N=$N F=$F : $r"
done; done
