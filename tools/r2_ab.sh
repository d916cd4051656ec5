#!/bin/bash
# A/B timing of library variants (tools/build_variant.py) on the headline workload: usage r2_ab.sh variant...
export PYTHONUNBUFFERED=1
timeout 300 python tools/r2_probe.py 2>&1 | grep -E "PROBE|rror" | tail -2
TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 2>&1 | grep "phase clocks" | tail -1
for v in "$@"; do
  echo "--- variant $v"
  TNO_LIB_VARIANT=$v timeout 300 python tools/r2_probe.py 2>&1 | grep -E "PROBE|rror" | tail -2
  TNO_LIB_VARIANT=$v TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 2>&1 | grep "phase clocks" | tail -1
done
