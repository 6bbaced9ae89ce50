#!/bin/bash
TAG=${1:-b1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
    print("cfg2 ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), "train_prop", round(d["train_propagation"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4))
    print("train_step_dp", {k: v for k, v in d["train_step_dp"].items() if k != "what"})
    for k, v in d["configs"].items():
        print(k, "ms", round(v["ms_per_step"], 4), "hot", round(v["ms_per_step_hot_l2"], 4), "e2e", round(v["e2e"]["ms_per_step"], 3), "frac", round(v["roofline"]["frac"], 4), v["plan"][:60])
    print("cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as ex:
    print("parse failed", ex); print(open("$OUT/bench_default.err").read()[-3000:])
PY
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "full suite exit $?"; tail -5 $OUT/pytest_gpu.log
