#!/bin/bash
# GPU session r2r (8 GPUs): BASELINE configs[2] with the relative parity bar, configs[4] through the batched hybrid entry points
mkdir -p gpurun_out
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.out 2> gpurun_out/$name.err
  echo "$name rc=$?"; grep '^{' gpurun_out/$name.out > gpurun_out/$name.json; head -c 1200 gpurun_out/$name.json; echo; tail -n 2 gpurun_out/$name.err
}
run r2r_config5_n8 8 --config 5 --steps 10 --warmup 3
run r2r_config3_n8 8 --config 3 --steps 10 --warmup 3
run r2r_config5_n2 2 --config 5 --steps 10 --warmup 3
