#!/bin/bash
# Round evidence on one B200 (run through gpurun from the repo root): tests, smoke, bench lines, ncu launch lists and
# one `ncu --set full` capture per dominant kernel.  Every ncu command runs only after the same command exited 0 plainly.
# Outputs land in gpurun_out/; the summaries that are kept are copied to profiles/ by hand afterwards.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1

timeout 700 python -m pytest tests -m gpu -x -q > $O/ev_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/ev_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/ev_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/ev_smoke.log

timeout 600 python bench.py > $O/ev_bench_1gpu.json 2> $O/ev_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/ev_bench_reference.json 2> $O/ev_bench_reference.err; echo "reference rc=$?"
timeout 400 python bench.py --workload crossvault_disc40 --batch 8192 --e2e-batch 256 --steps 5 --warmup 3 --no-cpu-baseline --no-sweeps \
  > $O/ev_bench_disc40.json 2> $O/ev_bench_disc40.err; echo "disc40 rc=$?"

# ---- launch lists ----
A="--steps 2 --warmup 1 --no-cpu-baseline --no-sweeps"
timeout 300 python bench.py $A > $O/ev_plain_onchip.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches_onchip.csv \
    python bench.py $A > $O/ev_ncu_onchip.log 2>&1; echo "launch list onchip rc=$?"
B="--workload crossvault_disc40 --batch 2048 --e2e-batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-sweeps"
timeout 300 python bench.py $B > $O/ev_plain_disc40.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/ev_launches_interleaved.csv \
    python bench.py $B > $O/ev_ncu_disc40.log 2>&1; echo "launch list disc40 rc=$?"

# ---- full captures ----
C="--batch 2048 --e2e-batch 256 --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps"
timeout 300 python bench.py $C > $O/ev_plain_c.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_onchip --launch-skip 2 -c 1 -f -o $O/ev_onchip \
    python bench.py $C > $O/ev_ncu_full_onchip.log 2>&1; echo "full onchip rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:k_forward_level -c 2 -f -o $O/ev_forward_level \
    python bench.py $B > $O/ev_ncu_full_fwd.log 2>&1; echo "full forward rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:k_factor_level --launch-skip 8 -c 3 -f -o $O/ev_factor_level \
    python bench.py $B > $O/ev_ncu_full_fac.log 2>&1; echo "full factor rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:k_solve_level_seg --launch-skip 150 -c 2 -f -o $O/ev_solve_seg \
    python bench.py $B > $O/ev_ncu_full_seg.log 2>&1; echo "full seg rc=$?"
ls -la $O | tail -30
