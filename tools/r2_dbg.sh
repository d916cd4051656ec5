#!/bin/bash
# developer experiments on k_onchip (results are wrong by construction; timing only): TNO_DBG bits
#   1 / 2: skip the Z / S updates of the dense tile steps, 4: no reciprocal, 8: empty tile steps,
#   16: skip the item loops of the staged factor chunks, 32: do not stream the chunks,
#   64: skip the stores of the envelope rows of J, 128: skip the stores of its dz/dzb block
export PYTHONUNBUFFERED=1
for d in "$@"; do
  TNO_DBG=$d TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 2>&1 | grep -E "phase clocks|PROBE" | tail -2 | sed -e "s/^/dbg $d /" -e 's/ | factor.*//'
done
