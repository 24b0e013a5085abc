#!/bin/bash
# host transport anatomy: delta-step time vs worker threads, with and without the frame patching (CW_HOST_NOPATCH)
for n in 4096 16384; do
for t in 1 2 4 8 16; do
  for np in 0 1; do
    echo "== N=$n threads=$t nopatch=$np"
    CW_HOST_TRACE=1 CW_HOST_THREADS=$t CW_HOST_NOPATCH=$np python tools/e2e_probe.py $n 1500 2>&1 | grep -E "delta|trace"
  done
done
done
