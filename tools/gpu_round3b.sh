#!/bin/bash
# Reduced end-of-round pass (the full GPU test suite of the same code is in the pass before): smoke, both bench arms,
# launch list and a full capture of the sampler kernel.  Numbers printed by a run under ncu are never bench values.
TAG=${1:-r03f}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1
python bench.py > $OUT/${TAG}_bench_cfg3.json 2> $OUT/${TAG}_bench_cfg3.err
python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"causal_mh_tc16" -s 3 -c 1 \
    -o $OUT/${TAG}_mh_tc_prof -f python bench.py --steps 1 --warmup 1 > $OUT/${TAG}_ncu_mh.log 2>&1
tail -2 $OUT/${TAG}_smoke.log
python - <<P
import json
for f in ("bench_cfg3", "bench_reference_arm"):
    try:
        d = json.load(open("$OUT/${TAG}_%s.json" % f)); print(f, d["value"], d.get("ms_per_step"), d["e2e"]["value"], d["e2e"].get("per_step_ms"))
    except Exception as e:
        print(f, "ERR", e)
P
