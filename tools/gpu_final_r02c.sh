#!/bin/bash
# Last evidence pass of round 2: full GPU suite, smoke, default bench line, reference arm (NUMA placements), a torchrun-like environment on one GPU.
OUT=gpurun_out/${OUT_TAG:-fin4}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference exit $?"
OMP_NUM_THREADS=1 LOCAL_WORLD_SIZE=8 GGNN_HOST_TIMING=1 timeout 300 python bench.py --config cfg3_dense --steps 20 --warmup 3 --no-other-configs --no-train-step --no-cpu-baseline > $OUT/bench_cfg3_torchrun_env.json 2> $OUT/bench_cfg3_torchrun_env.err; echo "cfg3 exit $?"; grep -m1 OpenMP $OUT/bench_cfg3_torchrun_env.err
python - <<PY
import json
def last(f): return json.loads([l for l in open(f) if l.startswith("{")][-1])
d=last("$OUT/bench_default.json")
print("cfg2 ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "readout", round(d["e2e_readout"]["ms_per_step"],4),
      "pipelined", round(d["e2e_pipelined"]["ms_per_step"],4), "producer", round(d["e2e_producer_thread"]["ms_per_step"],4), "train", round(d["train_propagation"]["ms_per_step"],4),
      "dp", round(d["train_step_dp"]["ms_per_step"],4), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["configs"].items():
    print(k, "ms", round(v["ms_per_step"],4), "frac", round(v["roofline"]["frac"],4), "e2e", round(v["e2e"]["ms_per_step"],4))
r=last("$OUT/bench_reference.json"); print("reference", r["value"], r["cpu_baseline"]["sample"][-260:])
c=last("$OUT/bench_cfg3_torchrun_env.json"); print("cfg3 under OMP_NUM_THREADS=1 LOCAL_WORLD_SIZE=8: ms", round(c["ms_per_step"],4), "e2e", round(c["e2e"]["ms_per_step"],4))
PY
tail -2 $OUT/bench_default.err
