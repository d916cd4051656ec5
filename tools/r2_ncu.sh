#!/bin/bash
# ncu --set full capture of k_onchip on the headline workload (2 048 candidates) + single-candidate phase clocks per layout
set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
LY=${1:-256,1,0}
export TNO_ONCHIP_LAYOUT=$LY
for L2 in "256,1,0" "256,3,1" "256,2,0"; do
  TNO_ONCHIP_LAYOUT=$L2 TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 1 2>&1 | grep "phase clocks" | tail -1
done
timeout 300 python tools/r2_probe.py crossvault_disc20 2048 2>&1 | grep PROBE && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_onchip --launch-skip 3 -c 1 -f -o $O/r2_onchip \
   python tools/r2_probe.py crossvault_disc20 2048 > $O/r2_ncu_onchip.log 2>&1; echo "ncu rc=$?"
ls -la $O/r2_onchip.ncu-rep
