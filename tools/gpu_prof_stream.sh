#!/bin/bash
# ncu launch list + one full capture per streaming kernel variant at cfg4; forced-stream bench lines of the other large configs
TAG=${1:-sp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg4.csv python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/launches_cfg4.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:60] + (" " + r[gi] if gi is not None else "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
for k, (n, t) in agg.items(): print("%-90s n=%3d avg=%9.1f ns" % (k, n, t / n))
PY
if [ -n "$FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ggnn_stream_kernel -s 30 -c 3 -o $OUT/prof_stream_cfg4 -f python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
fi
for cfg in ${BENCH_CFGS:-cfg5_rgcn default_batch_100k_nodes}; do
  for fs in 0 1; do
  GGNN_TC_STREAM=$fs timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_${cfg}_s$fs.json 2> $OUT/bench_${cfg}_s$fs.err; echo "$cfg stream=$fs exit $?"
  python -c "import json; d=json.loads(open('$OUT/bench_${cfg}_s$fs.json').read().strip().splitlines()[-1]); print('  ms', round(d['ms_per_step'],4), 'hot', round(d['ms_per_step_hot_l2'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['config']['plan'])" || tail -15 $OUT/bench_${cfg}_s$fs.err
  done
done
