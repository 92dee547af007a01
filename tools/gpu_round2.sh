#!/bin/bash
# Round-2 profiling pass (under gpurun, ONE GPU): launch lists and full captures of the kernels added or
# re-measured this round.  Numbers printed by a run under ncu are never bench values.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/${TAG}_train_launches.csv python tools/train_launches.py > $OUT/${TAG}_ncu_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bnn_mh_kernel" -s 60 -c 1 \
    -o $OUT/${TAG}_bnn_mh_prof -f python bench.py --config cfg3bnn --steps 1 --warmup 1 > $OUT/${TAG}_ncu_bnn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hmc_kernel" -s 3 -c 1 \
    -o $OUT/${TAG}_hmc_prof -f python tools/hmc_bench.py --steps 3 --reps 0 > $OUT/${TAG}_ncu_hmc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"causal_mh_tc16" -s 3 -c 1 \
    -o $OUT/${TAG}_mh_tc_prof -f python bench.py --steps 1 --warmup 1 > $OUT/${TAG}_ncu_mh.log 2>&1
ls -la $OUT/${TAG}_*
