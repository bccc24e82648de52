#!/bin/bash
# GPU session r2e (2 GPUs): NCCL shard group test, full-size parity tests, bench at N=2 (strong / weak), configs 3 and 5 at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_vecsim_sharded.py tests/test_vecsim_fullsize.py -x -q -m gpu > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -n 15 gpurun_out/r2e_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err ) 2> gpurun_out/r2e_n2.time
echo "n2 strong rc=$?"; tail -n 3 gpurun_out/r2e_n2.time; head -c 2500 gpurun_out/r2e_bench_n2.json; echo; tail -n 5 gpurun_out/r2e_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --scaling weak --no-parity > gpurun_out/r2e_bench_n2_weak.json 2> gpurun_out/r2e_bench_n2_weak.err
echo "n2 weak rc=$?"; head -c 1200 gpurun_out/r2e_bench_n2_weak.json; echo; tail -n 3 gpurun_out/r2e_bench_n2_weak.err
timeout 900 $TR bench.py --gpus 2 --config 3 --steps 10 --warmup 3 > gpurun_out/r2e_config3_n2.json 2> gpurun_out/r2e_config3_n2.err
echo "config3 rc=$?"; head -c 2500 gpurun_out/r2e_config3_n2.json; echo; tail -n 5 gpurun_out/r2e_config3_n2.err
timeout 600 $TR bench.py --gpus 2 --config 5 --steps 10 --warmup 3 > gpurun_out/r2e_config5_n2.json 2> gpurun_out/r2e_config5_n2.err
echo "config5 n2 rc=$?"; head -c 2500 gpurun_out/r2e_config5_n2.json; echo; tail -n 5 gpurun_out/r2e_config5_n2.err
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2e_config5_n1.json 2> gpurun_out/r2e_config5_n1.err
echo "config5 n1 rc=$?"; head -c 2500 gpurun_out/r2e_config5_n1.json; echo; tail -n 5 gpurun_out/r2e_config5_n1.err
