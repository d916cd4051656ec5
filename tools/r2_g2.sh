#!/bin/bash
# Re-entry check of the restored state on one B200: all GPU tests, smoke, headline probe.
set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2_g2_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r2_g2_tests.log)"
grep -E "^(FAILED|E   )" $O/r2_g2_tests.log | head -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_g2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python tools/r2_probe.py 2>&1 | grep -E "PROBE|rror" | tail -2
timeout 300 python tools/r2_probe.py crossvault_disc40 4096 2>&1 | grep -E "PROBE|rror" | tail -2
