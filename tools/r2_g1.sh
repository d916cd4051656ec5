set -u
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2_g1_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/r2_g1_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2_g1_tests.log
python -c "
from compas_tno_b200 import _lib
print('DFMA peak TFLOP/s', _lib.device_peak('fp64_fma'))
print('DMMA peak TFLOP/s', _lib.device_peak('fp64_mma'))
" > $O/r2_g1_peaks.log 2>&1
timeout 300 python tools/r2_probe.py > $O/r2_g1_probe_base.log 2>&1
TNO_ONCHIP_512X2=1 timeout 300 python tools/r2_probe.py > $O/r2_g1_probe_512x2.log 2>&1
TNO_PHASE_CLOCKS=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 > $O/r2_g1_clocks_base.log 2>&1
TNO_PHASE_CLOCKS=1 TNO_ONCHIP_512X2=1 timeout 300 python tools/r2_probe.py crossvault_disc20 4096 > $O/r2_g1_clocks_512x2.log 2>&1
tail -3 $O/r2_g1_tests.log; cat $O/r2_g1_peaks.log; grep PROBE $O/r2_g1_probe_*.log
