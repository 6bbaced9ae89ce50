#!/bin/bash
# One gpurun call: parity tests, probes, bench lines for every BASELINE config, ncu evidence.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
if [ -x profiles/probes/tc_probe ]; then echo "== tc_probe"; timeout 120 profiles/probes/tc_probe > $OUT/tc_probe.log 2>&1; echo "probe exit $?" >> $OUT/tc_probe.log; cat $OUT/tc_probe.log; fi
echo "== bench"
for cfg in cfg2 cfg1_true_default cfg3_dense cfg4 cfg5_rgcn default_batch_100k_nodes; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "$cfg exit $?"; tail -c 1500 $OUT/bench_$cfg.json | head -c 1500; echo
done
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2>&1
echo "== ncu launch list (cfg2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full (fused kernel, cfg2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ggnn_fwd -s 3 -c 2 -o $OUT/prof_cfg2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
