#!/bin/bash
# phase clocks of k_onchip (CTA 0, 5th candidate) on the headline workload and the dome
set -u
export PYTHONUNBUFFERED=1
for a in "crossvault_disc20 4096" "dome_16x20 4096" "crossvault_disc10 8192"; do
  TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py $a 2>&1 | grep -E "phase clocks|tno smem|PROBE" | tail -3
done
