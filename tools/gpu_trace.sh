#!/bin/bash
OUT=gpurun_out/${1:-tr}; mkdir -p $OUT
timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4.txt 2>&1
GGNN_BENCH_SHARD=0,8 timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4_shard8.txt 2>&1
grep -A45 "edge launch" $OUT/trace_cfg4_shard8.txt | head -50
grep -A20 "cand launch" $OUT/trace_cfg4_shard8.txt | head -24
