#!/bin/bash
OUT=gpurun_out/${1:-t}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_refgraph.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest.log
timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4.txt 2>&1; grep -E "^edge" $OUT/trace_cfg4.txt | head -2 | cut -c1-330
for cfg in cfg4 cfg5_rgcn; do
timeout 300 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err
python -c "import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('$cfg ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -3 $OUT/b.err
done
GGNN_BENCH_SHARD=0,8 timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err
python -c "import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('cfg4 1/8 shard ms', round(d['ms_per_step'],4))"
