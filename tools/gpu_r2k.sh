#!/bin/bash
# GPU session r2k (8 GPUs): NCCL shard-group tests, BASELINE configs[2] and [4] sharded over 8 GPUs, configs[1] strong + weak scaling
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 500 python -m pytest tests/test_vecsim_sharded.py -x -q -m gpu -k "nccl or packed" > gpurun_out/r2k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2k_tests.log
tail -n 5 gpurun_out/r2k_tests.log
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; head -c 1500 gpurun_out/$name.json; echo; tail -n 2 gpurun_out/$name.err
}
run r2k_config3_n8 8 --config 3 --steps 10 --warmup 3
run r2k_config5_n8 8 --config 5 --steps 10 --warmup 3
run r2k_bench_n8 8 --steps 20 --warmup 3 --no-cpu-baseline
run r2k_bench_n8_weak 8 --steps 20 --warmup 3 --scaling weak --no-parity --no-cpu-baseline --no-postings
run r2k_config3_n2 2 --config 3 --rows3 12500000 --steps 10 --warmup 3
