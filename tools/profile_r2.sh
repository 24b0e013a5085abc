#!/bin/bash
# round-2 ncu captures (one gpurun call, 1 GPU).  Every target first runs plain; ncu only if that exited 0.
T="python tools/ncu_targets.py"
O=gpurun_out
FULL="ncu --set full --clock-control none --import-source on"
DRAM="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none --csv"
cap() { name=$1; shift; target=$1; shift; $T $target > $O/plain_$name.log 2>&1 && "$@" $T $target > $O/ncu_$name.log 2>&1; echo "$name rc=$?"; }
# (ii) DRAM traffic of consecutive ring launches with the caches left alone: >= 8 launches of the chained kernel at cfg2
cap traffic_cfg2 chained_cfg2 $DRAM -k regex:cw_env_kernel -s 30 -c 16 --log-file $O/r2_traffic_chained_cfg2_nocachectl.csv
cap traffic_cfg5 chained_cfg5 $DRAM -k regex:cw_env_kernel -s 8 -c 8 --log-file $O/r2_traffic_chained_cfg5_nocachectl.csv
# (iii) full sets
cap full_cfg2 chained_cfg2 $FULL -k regex:cw_env_kernel -s 40 -c 2 -o $O/r2_env_kernel_chained_cfg2
cap full_cfg5 chained_cfg5 $FULL -k regex:cw_env_kernel -s 10 -c 2 -o $O/r2_env_kernel_chained_cfg5
cap full_delta delta $FULL -k regex:cw_delta_kernel -s 30 -c 2 -o $O/r2_delta_kernel
cap full_incremental incremental $FULL -k "regex:cw_step_kernel|cw_env_kernel" -s 40 -c 4 -o $O/r2_incremental_step_edit_and_list
cap full_compact compact $FULL -k regex:cw_step_kernel -s 30 -c 2 -o $O/r2_step_kernel_compact
cap full_closed closed $FULL -k regex:cw_frame_policy -s 20 -c 2 -o $O/r2_frame_policy
# (i) launch list of the bench command at the driver's settings (device legs only)
B="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
$B > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_bench.csv $B > $O/ncu_bench.log 2>&1; echo "launches rc=$?"
ls -la $O/*.ncu-rep $O/r2_*.csv
