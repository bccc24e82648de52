#!/bin/bash
# GPU session r2f: full suite (incl. full-size parity tests, timeout polling, hybrid state machine, 16-bit fixed bound),
# racecheck / memcheck, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -n 25 gpurun_out/r2f_tests.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_vecsim_coarse.py::test_coarse_path_is_exact" -x -q -m gpu -k "70000-128-40-10-1" > gpurun_out/r2f_racecheck_coarse.log 2>&1
echo "racecheck coarse rc=$?"; tail -n 6 gpurun_out/r2f_racecheck_coarse.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" -x -q -m gpu -k "0" > gpurun_out/r2f_racecheck_postings.log 2>&1
echo "racecheck postings rc=$?"; tail -n 6 gpurun_out/r2f_racecheck_postings.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" "tests/test_postings_parity.py::test_batch_decode_of_many_lists_matches_the_oracle_reader" -x -q -m gpu -k "0 or 1" > gpurun_out/r2f_memcheck_postings.log 2>&1
echo "memcheck postings rc=$?"; tail -n 6 gpurun_out/r2f_memcheck_postings.log
timeout 300 python bench.py --no-postings --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['sustained']['ms_per_step'], d['parity_at_config'])
PY
timeout 400 python bench.py --config 3 --rows3 6250000 --steps 10 --warmup 3 > gpurun_out/r2f_config3_shard.json 2> gpurun_out/r2f_config3_shard.err
echo "config3 shard rc=$?"; head -c 2000 gpurun_out/r2f_config3_shard.json; echo; tail -n 3 gpurun_out/r2f_config3_shard.err
timeout 300 env VECSIM_B200_FIXED=0 python bench.py --config 3 --rows3 6250000 --steps 10 --warmup 3 --no-parity > gpurun_out/r2f_config3_shard_adaptive.json 2> gpurun_out/r2f_config3_shard_adaptive.err
echo "config3 adaptive rc=$?"; head -c 900 gpurun_out/r2f_config3_shard_adaptive.json; echo
