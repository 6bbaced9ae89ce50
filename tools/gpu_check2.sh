#!/bin/bash
OUT=gpurun_out/${1:-c2}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "full suite exit $?"; tail -4 $OUT/pytest_gpu.log
bash tools/gpu_hosttime.sh > $OUT/ht.log 2>&1; grep -B6 "cfg2 run_sparse_host call 5" $OUT/ht.log; grep -B6 "cfg4 run_sparse_host call 5" $OUT/ht.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json; d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print('cfg2 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'e2e_readout', round(d['e2e_readout']['ms_per_step'],4), 'pipe', round(d['e2e_pipelined']['ms_per_step'],4), 'dp', round(d['train_step_dp']['ms_per_step'],3))" || tail -5 $OUT/bench.err
