#!/bin/bash
# tensor-core kernel experiment: parity first (stop on failure), then timing
OUT=gpurun_out/${1:-tcx}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py tests/test_gpu_dropout.py -m gpu -x -q > $OUT/pytest.log 2>&1; rc=$?
tail -4 $OUT/pytest.log
[ $rc -ne 0 ] && { grep -E "Error|error|assert|timeout" $OUT/pytest.log | head -20; echo "tests failed, stopping"; exit 0; }
for cfg in cfg2 cfg1_true_default cfg3_dense; do
  timeout 200 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err
  python -c "import json; d=json.loads(open('$OUT/bench_$cfg.json').read().strip().splitlines()[-1]); print('$cfg ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['config']['plan'])" || tail -3 $OUT/bench_$cfg.err
done
GGNN_TC_NO_COMPACT=1 timeout 200 python bench.py --config cfg2 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 no-compact ms', round(d['ms_per_step'],4))"
timeout 100 python tools/tc_phase_timing.py cfg2 bf16x3 | tail -4
