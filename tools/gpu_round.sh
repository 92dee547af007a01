#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), launch list, one full ncu capture of the sampler.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err
timeout 600 python bench.py --engine simt > $OUT/${TAG}_bench_1gpu_simt.json 2> $OUT/${TAG}_bench_1gpu_simt.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"causal_mh" -s 2 -c 1 \
    -o $OUT/${TAG}_mh_prof -f python bench.py --steps 1 --warmup 1 > $OUT/${TAG}_ncu_full.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"causal_effect" -c 1 \
    -o $OUT/${TAG}_effect_prof -f python bench.py --steps 1 --warmup 1 > $OUT/${TAG}_ncu_full_effect.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_smoke.log | tail -2; cat $OUT/${TAG}_bench_1gpu.json
