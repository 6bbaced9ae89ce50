#!/bin/bash
OUT=gpurun_out/${1:-tc4}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_refgraph.py tests/test_gpu_dropout.py tests/test_gpu_backward.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "tests exit $?"; tail -3 $OUT/pytest.log
for gb in 2 4; do
for cfg in cfg2 cfg1_true_default cfg3_dense default_batch_100k_nodes; do
  GGNN_TC_GBUFS=$gb timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err
  python -c "import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('gbufs=$gb $cfg: ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -3 $OUT/b.err
done; done
