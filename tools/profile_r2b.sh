#!/bin/bash
# second batch of round-2 ncu captures: the kernels added after profile_r2.sh (one gpurun call, 1 GPU).  Plain run first, ncu only if it exited 0.
T="python tools/ncu_targets.py"
O=gpurun_out
FULL="ncu --set full --clock-control none --import-source on"
cap() { name=$1; shift; target=$1; shift; timeout 120 $T $target > $O/plain_$name.log 2>&1 && timeout 400 "$@" $T $target > $O/ncu_$name.log 2>&1; echo "$name rc=$?"; }
cap full_pipe_step pipe $FULL -k regex:cw_step_snap_kernel -s 30 -c 2 -o $O/r2_step_snap_kernel
cap full_pipe_render pipe $FULL -k regex:cw_env_kernel -s 30 -c 2 -o $O/r2_env_kernel_pipe
cap full_compact_chained compact_chained $FULL -k regex:cw_step_chained -s 30 -c 2 -o $O/r2_step_chained_kernel
cap full_compact compact $FULL -k regex:cw_step_kernel -s 30 -c 2 -o $O/r2b_step_kernel_compact
B="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
timeout 300 $B > $O/plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2b_launches_bench.csv $B > $O/ncu_bench.log 2>&1; echo "launches rc=$?"
ls -la $O/*.ncu-rep $O/r2b_*.csv
