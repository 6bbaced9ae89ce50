#!/bin/bash
# Quick guarded check: new tests first, then the full GPU suite, then two short bench lines.
TAG=${1:-chk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest ${NEW_TESTS:-tests/test_gpu_readout.py tests/test_gpu_dropout.py} -m gpu -x -q > $OUT/pytest_new.log 2>&1; rc=$?
tail -25 $OUT/pytest_new.log
[ $rc -ne 0 ] && { echo "new tests failed ($rc), stopping"; exit 0; }
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "full suite exit $?"; tail -5 $OUT/pytest_gpu.log
for cfg in cfg2 cfg3_dense cfg5_rgcn; do
  timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "$cfg exit $?"
  python -c "import json; d=json.loads(open('$OUT/bench_$cfg.json').read().strip().splitlines()[-1]); print('  ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'pipe_ms', round(d['e2e_pipelined']['ms_per_step'],3), 'train_ms', round(d['train_propagation']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'readout', d.get('readout'))" || tail -5 $OUT/bench_$cfg.err
done
