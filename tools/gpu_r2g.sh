#!/bin/bash
# GPU session r2g (8 GPUs): NCCL shard-group test, BASELINE configs[1] / [2] / [4] sharded over 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_vecsim_sharded.py -x -q -m gpu -k "nccl or packed" > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -n 6 gpurun_out/r2g_tests.log
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; head -c 1800 gpurun_out/$name.json; echo; tail -n 3 gpurun_out/$name.err
}
run r2g_bench_n8 8 --steps 20 --warmup 3
run r2g_config3_n8 8 --config 3 --steps 10 --warmup 3
run r2g_config5_n8 8 --config 5 --steps 10 --warmup 3
run r2g_bench_n2 2 --steps 20 --warmup 3
run r2g_bench_n8_weak 8 --steps 20 --warmup 3 --scaling weak --no-parity
