#!/bin/bash
# GPU session r2c: posting-path tests (fused batch, batch decode, term cache), full default bench, launch list, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_postings_parity.py tests/test_hybrid_filtered.py -x -q -m gpu > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -n 12 gpurun_out/r2c_tests.log
( time timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2> gpurun_out/r2c_bench.time
echo "bench rc=$?"; tail -n 3 gpurun_out/r2c_bench.time; tail -n 5 gpurun_out/r2c_bench.err; head -c 6000 gpurun_out/r2c_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --no-postings --no-parity --no-cpu-baseline --steps 2 --warmup 3 --sustained-seconds 0 > gpurun_out/r2c_ncu_bench.log 2>&1
echo "ncu rc=$?"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err ) 2> gpurun_out/r2c_ref.time
echo "ref rc=$?"; tail -n 3 gpurun_out/r2c_ref.time; cat gpurun_out/r2c_ref.json | head -c 3000; tail -n 5 gpurun_out/r2c_ref.err
