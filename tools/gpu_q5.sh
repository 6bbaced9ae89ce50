#!/bin/bash
OUT=gpurun_out/${OUT_TAG:-q5}; mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_prepared.py tests/test_gpu_refgraph.py tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_readout.py tests/test_gpu_checkpoint.py tests/test_gpu_odd_hidden.py tests/test_gpu_dropout.py -m gpu -q > $OUT/pytest_dense_and_prepared.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_dense_and_prepared.log; tail -6 $OUT/pytest_dense_and_prepared.log
timeout 200 python bench.py --config cfg3_dense --steps 20 --warmup 3 --no-other-configs --no-train-step --no-cpu-baseline > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "bench exit $?"
python - <<PY
import json
c=json.loads([l for l in open("$OUT/bench_cfg3.json") if l.startswith("{")][-1]); print("cfg3 ms", round(c["ms_per_step"],4), "e2e", round(c["e2e"]["ms_per_step"],4), c["engine"]["plan"][:80])
PY
