#!/bin/bash
# developer experiments on k_onchip's dense tile steps (results are wrong by construction; timing only)
export PYTHONUNBUFFERED=1
for d in 0 1 2 3 4 7 8; do
  TNO_DBG=$d TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 2>&1 | grep -E "phase clocks" | tail -1 | sed -e 's/.*factor: dense=\([0-9]*\).*/dbg '$d' dense=\1/'
done
