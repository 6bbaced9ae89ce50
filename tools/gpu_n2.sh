#!/bin/bash
OUT=gpurun_out/${OUT_TAG:-n2}; mkdir -p $OUT
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "cfg2 value %.3e" % d["value"], "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), "producer", round(d["e2e_producer_thread"]["ms_per_step"], 3), "pipelined", round(d["e2e_pipelined"]["ms_per_step"], 3))
    print("train_step_dp", {k: v for k, v in d["train_step_dp"].items() if k != "what"})
    for k, v in d["configs"].items():
        print(k, "value %.3e" % v["value"], "ms", round(v["ms_per_step"], 4), "e2e", round(v["e2e"]["ms_per_step"], 3), "frac", round(v["roofline"]["frac"], 4), v["scaling"][:12])
except Exception as ex:
    print("parse failed", ex); print(open("$OUT/bench_n2.err").read()[-3000:])
PY
tail -3 $OUT/bench_n2.err
