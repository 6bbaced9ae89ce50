#!/bin/bash
OUT=gpurun_out/${1:-tr2}; mkdir -p $OUT
timeout 300 python tools/stream_trace.py cfg4 > $OUT/trace_cfg4.txt 2>&1; grep -E "^edge" $OUT/trace_cfg4.txt | head -3 | cut -c1-330
