#!/bin/bash
# ncu --set full capture of k_onchip on the headline workload (2 048 candidates), after the same command ran plainly
set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python tools/r2_probe.py crossvault_disc20 2048 2>&1 | grep PROBE && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_onchip --launch-skip 3 -c 1 -f -o $O/r2b_onchip \
   python tools/r2_probe.py crossvault_disc20 2048 > $O/r2b_ncu_onchip.log 2>&1; echo "ncu rc=$?"
ls -la $O/r2b_onchip.ncu-rep
