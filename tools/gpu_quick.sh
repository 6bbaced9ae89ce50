#!/bin/bash
# Quick GPU check: the test files given as arguments first (verbose), then the whole GPU suite.
OUT=gpurun_out/${OUT_TAG:-quick}; mkdir -p $OUT
if [ $# -gt 0 ]; then
  timeout 400 python -m pytest "$@" -m gpu -q -s 2>&1 | tail -60 > $OUT/pytest_new.log; tail -40 $OUT/pytest_new.log
fi
if [ -z "$SKIP_FULL" ]; then
  timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
fi
