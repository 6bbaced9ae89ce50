#!/bin/bash
OUT=gpurun_out/${OUT_TAG:-q2}; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_prepared.py -m gpu -q -s 2>&1 | tail -40 > $OUT/pytest_new.log; tail -25 $OUT/pytest_new.log
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-train-step > $OUT/bench_cfg2_short.json 2> $OUT/bench_cfg2_short.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_cfg2_short.json") if l.startswith("{")][-1])
for k in ("ms_per_step","e2e","e2e_readout","e2e_pipelined","e2e_producer_thread"):
    v=d.get(k); print(k, v if not isinstance(v,dict) else round(v["ms_per_step"],4))
PY
tail -5 $OUT/bench_cfg2_short.err
