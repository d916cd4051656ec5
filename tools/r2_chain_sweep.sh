#!/bin/bash
# sweep of the chain partition knobs (symbolic analysis) against the step time of k_onchip
export PYTHONUNBUFFERED=1
WL=${1:-crossvault_disc20}; N=${2:-16384}
for cw in 4 5 6 7 8 12 20; do for mg in 8 16; do
  TNO_CHAIN_WIDTH=$cw TNO_CHAIN_MERGE=$mg timeout 120 python tools/r2_probe.py $WL $N 2>&1 | grep -E "PROBE|rror" | tail -1 | sed -e 's/kernel=auto//' -e 's/fsum.*//'
done; done
