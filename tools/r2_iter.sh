#!/bin/bash
# One optimisation iteration on the GPU box: parity tests of the on-chip family, timing of the headline workload, phase clocks.
# usage (through gpurun): bash tools/r2_iter.sh [tag] [quick]
set -u
TAG=${1:-it}; O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
if [ "${2:-}" != "quick" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_hard.py -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/${TAG}_tests.log)"
  grep -E "^(FAILED|E   )" $O/${TAG}_tests.log | head -8
fi
timeout 300 python tools/r2_probe.py 2>&1 | grep -E "PROBE|rror" | tail -2
timeout 300 python tools/r2_probe.py dome_16x20 4096 2>&1 | grep -E "PROBE|rror" | tail -2
TNO_ONCHIP_LAYOUT=256,1,0 timeout 300 python tools/r2_probe.py dome_16x20 4096 2>&1 | grep -E "PROBE|rror" | tail -2
timeout 300 python tools/r2_probe.py crossvault_disc10 32768 2>&1 | grep -E "PROBE|rror" | tail -2
TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 2>&1 | grep "phase clocks" | tail -1
