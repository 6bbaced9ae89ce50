#!/bin/bash
# Round-2 final evidence pass: full GPU suite, smoke, the default bench line, per-config lines, ncu summaries.
OUT=gpurun_out/${1:-fin2}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
for cfg in default_batch_100k_nodes; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-other-configs --no-train-step > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err
done
timeout 600 python bench.py --config cfg4 --steps 20 --warmup 5 --no-other-configs --no-train-step > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err
timeout 600 python bench.py --config cfg2 --precision fp32 --steps 20 --warmup 5 --no-other-configs --no-train-step --no-cpu-baseline > $OUT/bench_cfg2_fp32.json 2> $OUT/bench_cfg2_fp32.err
GGNN_BENCH_SHARD=0,8 timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/bench_cfg4_shard_1of8.json 2> $OUT/bench_cfg4_shard.err
timeout 300 python tools/stream_trace.py cfg4 > $OUT/stream_trace_cfg4.txt 2>&1
bash tools/gpu_profile_r02.sh $(basename $OUT)/ncu > /dev/null 2>&1
rm -f $OUT/ncu/*.ncu-rep
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms", round(d.get("ms_per_step",0),4), "value %.3e" % d["value"], "e2e", round(d["e2e"].get("ms_per_step",0),4) if "ms_per_step" in d["e2e"] else "", "frac", round(d.get("roofline",{}).get("frac",0),4), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as ex: print(f, "unparsed", ex)
PY
ls $OUT $OUT/ncu; du -sh $OUT
