#!/bin/bash
# GPU iteration for the pre-processing kernels and the blocked block inversion of k_onchip
set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_preprocess.py -m gpu -x -q > $O/pre_tests.log 2>&1; echo "pre tests rc=$? $(tail -1 $O/pre_tests.log)"
grep -E "^(FAILED|E   )" $O/pre_tests.log | head -12
bash tools/r2_iter.sh ${1:-it2}
