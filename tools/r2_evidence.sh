#!/bin/bash
# Round-2 evidence on one B200 (through gpurun): GPU tests, smoke, bench lines, launch list and one `ncu --set full`
# capture of k_onchip.  Every ncu command runs only after the same command exited 0 plainly.
set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > $O/r2_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > $O/r2_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r2_tests.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --workload crossvault_disc40 --batch 4096 --steps 10 --no-cpu-baseline --no-sweeps > $O/r2_bench_disc40.json 2> $O/r2_bench_disc40.err; echo "disc40 rc=$?"
A="--steps 2 --warmup 3 --no-cpu-baseline --no-sweeps"
timeout 300 python bench.py $A > $O/r2_plain_onchip.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_onchip.csv \
    python bench.py $A > $O/r2_ncu_onchip.log 2>&1; echo "launch list onchip rc=$?"
timeout 300 python tools/r2_probe.py crossvault_disc20 2048 2>&1 | grep PROBE && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_onchip --launch-skip 3 -c 1 -f -o $O/r2_onchip \
   python tools/r2_probe.py crossvault_disc20 2048 > $O/r2_ncu_full_onchip.log 2>&1; echo "ncu full rc=$?"
TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 2>&1 | grep -E "phase clocks|factor ops|factor program" | tail -3 > $O/r2_phase_clocks.log
TNO_ONCHIP_NO_DMMA=1 timeout 300 python tools/r2_probe.py 2>&1 | grep PROBE > $O/r2_ab_dmma.log
timeout 300 python tools/r2_probe.py 2>&1 | grep PROBE >> $O/r2_ab_dmma.log
TNO_ONCHIP_NO_DMMA=1 timeout 300 python tools/r2_probe.py dome_16x20 4096 2>&1 | grep PROBE >> $O/r2_ab_dmma.log
timeout 300 python tools/r2_probe.py dome_16x20 4096 2>&1 | grep PROBE >> $O/r2_ab_dmma.log
TNO_ONCHIP_DENSE_ELEMENTS=1 timeout 300 python tools/r2_probe.py 2>&1 | grep PROBE >> $O/r2_ab_dmma.log
TNO_ONCHIP_WARP_LOCAL=1 TNO_ONCHIP_WARP_FACTOR=1 timeout 300 python tools/r2_probe.py 2>&1 | grep PROBE >> $O/r2_ab_dmma.log
TNO_ONCHIP_NO_ITEM_SORT=1 timeout 300 python tools/r2_probe.py 2>&1 | grep PROBE >> $O/r2_ab_dmma.log
./tools/latbench > $O/r2_latbench.log 2>&1
cat $O/r2_ab_dmma.log
ls -la $O | tail -24
