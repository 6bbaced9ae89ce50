#!/bin/bash
OUT=gpurun_out/${1:-san}; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stream.py -m gpu -x -q -k "eight_edge_types and 20" > $OUT/sanitizer.log 2>&1
grep -E "Invalid|Error|error|at 0x|by thread|Address|========= " $OUT/sanitizer.log | head -40
tail -5 $OUT/sanitizer.log
