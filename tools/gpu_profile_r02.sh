#!/bin/bash
# Round-2 evidence: launch lists + one `ncu --set full` capture per kernel family, summarised to text on the box (the reports are 16 MB each).
OUT=gpurun_out/${1:-prof2}; mkdir -p $OUT
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
nvidia-smi > $OUT/nvidia_smi.txt 2>&1
full() {  # name kernel-regex skip count  bench-args...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 900 $NCU --set full --import-source on -k regex:$rx -s $skip -c $cnt -o $OUT/$name -f $B "$@" > $OUT/$name.log 2>&1
  echo "ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt $B $*" > $OUT/$name.summary.txt
  python tools/ncu_summary.py $OUT/$name.ncu-rep >> $OUT/$name.summary.txt 2>&1
  python tools/ncu_hotspots.py $OUT/$name.ncu-rep 14 > $OUT/$name.hotspots.txt 2>&1
  [ "$KEEP" = "$name" ] || rm -f $OUT/$name.ncu-rep
}
echo "== launch lists"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/launches_default_bench.csv $B > $OUT/l1.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $OUT/launches_cfg4.csv $B --config cfg4 --no-train-step > $OUT/l2.log 2>&1
KEEP=prof_stream_cfg4
full prof_stream_cfg4 ggnn_stream_kernel 30 3 --config cfg4 --no-train-step
full prof_tc_cfg2 ggnn_fwd_tc 3 1 --no-train-step
full prof_ffma_cfg4 ggnn_fwd_ffma 2 1 --config cfg4 --precision fp32 --no-train-step
GGNN_TC_STREAM=0 full prof_tc_global_cfg5 ggnn_fwd_tc 10 1 --config cfg5_rgcn --no-train-step
full prof_bwd_gemms_cfg2 gemm_ 40 2 --no-train-step
cuobjdump -sass gated_graph_neural_network_samples_b200/libggnn_b200.so 2>/dev/null | python -c "
import sys,re,collections
cur=None; cnt=collections.defaultdict(collections.Counter)
for l in sys.stdin:
    m=re.search(r'Function : (\S+)', l)
    if m: cur=m.group(1); continue
    m=re.search(r'^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m and cur: cnt[cur][m.group(2).split('.')[0]]+=1
for f,c in cnt.items():
    if 'stream_kernel' in f or 'fwd_tc_kernel' in f:
        print(f[:110]); print('   ', {k:c[k] for k in ('UTCHMMA','UTCBAR','UBLKCP','LDTM','STTM','UTCATOMSWS','LDGSTS','SYNCS','FFMA','MUFU') if c[k]})
" > $OUT/sass_mnemonics.txt 2>&1
rm -f $OUT/*.log
ls -la $OUT; du -sh $OUT
