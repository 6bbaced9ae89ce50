#!/bin/bash
OUT=gpurun_out/${1:-dn}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "dense" > $OUT/pytest.log 2>&1; echo "dense tests exit $?"; tail -2 $OUT/pytest.log
for nt in 1 2 4 8; do
GGNN_HOST_THREADS=$nt timeout 300 python bench.py --config cfg3_dense --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err
python -c "import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('threads=$nt cfg3 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'pipe', round(d['e2e_pipelined']['ms_per_step'],4))" || tail -3 $OUT/b.err
done
timeout 300 python bench.py --config cfg3_dense --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-train-step > $OUT/b.json 2> $OUT/b.err
python -c "import json; d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]); print('default cfg3 ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
