#!/bin/bash
TAG=${1:-st2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q > $OUT/pytest_new.log 2>&1; echo "stream tests exit $?"; tail -3 $OUT/pytest_new.log
timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4.txt 2>&1; cat $OUT/trace_cfg4.txt | head -40
for cfg in ${BENCH_CFGS:-cfg4}; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "$cfg exit $?"
  python -c "import json; d=json.loads(open('$OUT/bench_$cfg.json').read().strip().splitlines()[-1]); print('  ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'train_ms', round(d['train_propagation']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['config']['plan'])" || tail -15 $OUT/bench_$cfg.err
done
