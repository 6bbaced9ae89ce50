#!/bin/bash
OUT=gpurun_out/exp; mkdir -p $OUT
run() { # name env...
  python -c "import json,sys; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('$1: ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4))" 2>/dev/null || tail -3 $OUT/b.err
}
for ks in 1 2 4; do
  GGNN_TS_KSTEPS=$ks GGNN_BENCH_SHARD=0,8 timeout 200 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err; run "cfg4 shard8 KS=$ks"
  GGNN_TS_KSTEPS=$ks timeout 200 python bench.py --config cfg5_rgcn --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err; run "cfg5 KS=$ks"
  GGNN_TS_KSTEPS=$ks timeout 200 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err; run "cfg4 KS=$ks"
done
GGNN_TC_STREAM=1 timeout 200 python bench.py --config default_batch_100k_nodes --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err; run "100k stream default KS"
GGNN_TC_STREAM=1 GGNN_TS_KSTEPS=2 timeout 200 python bench.py --config default_batch_100k_nodes --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err; run "100k stream KS=2"
