#!/bin/bash
TAG=${1:-st}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q > $OUT/pytest_new.log 2>&1; echo "stream tests exit $?"; tail -3 $OUT/pytest_new.log
timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4.txt 2>&1; grep -E "^edge|^gate|^cand" $OUT/trace_cfg4.txt | head -6 | cut -c1-200
GGNN_BENCH_SHARD=0,8 timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4_shard8.txt 2>&1
grep -E "^edge|^gate|^cand" $OUT/trace_cfg4_shard8.txt | head -3 | cut -c1-200
grep -A14 "edge launch" $OUT/trace_cfg4_shard8.txt | head -16
grep -A10 "cand launch" $OUT/trace_cfg4.txt | head -12
for cfg in cfg4 cfg5_rgcn; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "$cfg exit $?"
  python -c "import json; d=json.loads(open('$OUT/bench_$cfg.json').read().strip().splitlines()[-1]); print('  ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'train_ms', round(d['train_propagation']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['engine']['plan'])" || tail -15 $OUT/bench_$cfg.err
done
GGNN_BENCH_SHARD=0,8 timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/bench_cfg4_shard8.json 2> $OUT/bench_cfg4_shard8.err
python -c "import json; d=json.loads(open('$OUT/bench_cfg4_shard8.json').read().strip().splitlines()[-1]); print('cfg4 1/8 shard on one GPU: ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4), d['engine']['plan'])" || tail -5 $OUT/bench_cfg4_shard8.err
