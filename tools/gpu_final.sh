#!/bin/bash
# Round-end style run: full GPU test suite, smoke, bench lines for every BASELINE config, ncu evidence.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
echo "== bench (auto precision)"
for cfg in cfg2 cfg1_true_default cfg3_dense cfg4 cfg5_rgcn default_batch_100k_nodes; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "$cfg exit $?"
  python -c "import sys,json; d=json.loads(open('$OUT/bench_$cfg.json').read().strip().splitlines()[-1]); print('  ', d['dtype'][:8], 'ms', round(d['ms_per_step'],4), 'value %.3e' % d['value'], 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'pipe_ms', round(d['e2e_pipelined']['ms_per_step'],3), 'train_ms', round(d['train_propagation']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'cpu %.3e' % d['cpu_baseline']['value'], d['config']['plan'])" || tail -3 $OUT/bench_$cfg.err
done
echo "== bench cfg2 fp32 (FFMA path) and default invocation"
timeout 300 python bench.py --config cfg2 --precision fp32 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_cfg2_fp32.json 2> $OUT/bench_cfg2_fp32.err
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 600 $OUT/bench_default.json; echo
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 400 $OUT/bench_reference.json; echo
echo "== ncu launch list (default bench)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full (tc kernel, cfg2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ggnn_fwd_tc -s 3 -c 1 -o $OUT/prof_tc_cfg2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
timeout 100 python tools/tc_trace.py cfg2 > $OUT/tc_trace_cfg2.txt 2>&1
timeout 100 python tools/tc_phase_timing.py cfg2 bf16x3 > $OUT/tc_phase_timing_cfg2.txt 2>&1
ls -la $OUT | head -40
