#!/bin/bash
# Round-2 (second session) evidence pass: full GPU suite, smoke, host-build timing on the box's CPU, the default bench line, the reference arm.
OUT=gpurun_out/${OUT_TAG:-fin3}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -5 $OUT/smoke.log
( GGNN_HOST_TIMING=1 python tools/host_build_time.py cfg4 2>&1 | grep -E "OpenMP|median"; echo "--- auto"; python tools/host_build_time.py; echo "--- GGNN_HOST_THREADS=1"; GGNN_HOST_THREADS=1 python tools/host_build_time.py ) > $OUT/host_build_time.txt 2>&1; cat $OUT/host_build_time.txt
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference exit $?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_default.json") if l.startswith("{")][-1])
print("cfg2 ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "readout", round(d["e2e_readout"]["ms_per_step"],4),
      "pipelined", round(d["e2e_pipelined"]["ms_per_step"],4), "producer", round(d["e2e_producer_thread"]["ms_per_step"],4), "train", round(d["train_propagation"]["ms_per_step"],4),
      "dp", round(d["train_step_dp"]["ms_per_step"],4), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["configs"].items():
    print(k, "ms", round(v["ms_per_step"],4), "frac", round(v["roofline"]["frac"],4), "e2e", round(v["e2e"]["ms_per_step"],4))
r=json.loads([l for l in open("$OUT/bench_reference.json") if l.startswith("{")][-1]); print("reference", r["value"], r.get("cpu_baseline",{}).get("cores"))
PY
tail -3 $OUT/bench_default.err
