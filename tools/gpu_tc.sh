#!/bin/bash
# tensor-core path: parity tests + bench lines
TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest tensorcore"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -q -s -x > $OUT/pytest_tc.log 2>&1; echo "exit $?" >> $OUT/pytest_tc.log; grep -E "max\|err|passed|failed|Error|error|exit" $OUT/pytest_tc.log | head -60
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
for prec in bf16x3 bf16; do
for cfg in cfg2 cfg1_true_default cfg3_dense cfg5_rgcn default_batch_100k_nodes; do
  timeout 300 python bench.py --config $cfg --precision $prec --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_${cfg}_$prec.json 2> $OUT/bench_${cfg}_$prec.err; echo "$cfg $prec exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${cfg}_$prec.json").read().strip().splitlines()[-1])
    print("   ms/step %.4f hot %.4f value %.3e e2e_ms %.3f frac %.4f plan %s" % (d["ms_per_step"], d["ms_per_step_hot_l2"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"]["plan"]))
except Exception as ex:
    print("   parse failed", ex); print(open("$OUT/bench_${cfg}_$prec.err").read()[-800:])
PY
done; done
