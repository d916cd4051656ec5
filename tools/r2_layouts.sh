#!/bin/bash
# GPU parity tests of the on-chip family under forced layouts / variants (developer knobs), so that the 512-thread kernel,
# the two-group CTA, the 4 x 4 tiles, the element-wise dense variant and the warp-local variants all see the golden cases
set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
T="tests/test_gpu_parity.py tests/test_gpu_parity_hard.py"
run() { echo "== $*"; env "$@" timeout 900 python -m pytest $T -m gpu -x -q -k "not interleaved" 2>&1 | tail -2; }
run TNO_ONCHIP_LAYOUT=512,1,0
run TNO_ONCHIP_LAYOUT=256,2,0
run TNO_ONCHIP_LAYOUT=256,1,1
run TNO_ONCHIP_TILE4=1
run TNO_ONCHIP_DENSE_ELEMENTS=1
run TNO_ONCHIP_WARP_LOCAL=1 TNO_ONCHIP_WARP_FACTOR=1
run TNO_ONCHIP_WARP_LOCAL=1 TNO_ONCHIP_WARP_FACTOR=1 TNO_ONCHIP_LAYOUT=512,1,0
run TNO_ONCHIP_NO_DMMA=1
run TNO_DBG=512
