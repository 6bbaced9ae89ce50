#!/bin/bash
OUT=gpurun_out/${1:-t}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stream.py -m gpu -q -s > $OUT/pytest.log 2>&1; echo "exit $?"; grep -E "passed|failed|FAILED|Error|max\|err" $OUT/pytest.log | tail -30
