#!/bin/bash
# bench.py under torchrun on N GPUs of one box (the way the driver launches it); usage: r2_scale.sh N
set -u
N=${1:-2}; O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 40 --warmup 3 > $O/scale_r2_$N.json 2> $O/scale_r2_$N.err; echo "rc=$?"
tail -c 600 $O/scale_r2_$N.json; tail -5 $O/scale_r2_$N.err
