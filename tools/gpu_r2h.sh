#!/bin/bash
# GPU session r2h (1 GPU): full -m gpu suite on HEAD (slop / in-order included), default bench line, launch list,
# full ncu capture of the main pass, sanitizer passes over the coarse and fused posting routes
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -n 30 gpurun_out/r2h_tests.log
timeout 600 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; head -c 6000 gpurun_out/r2h_bench.json; echo; tail -n 5 gpurun_out/r2h_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > gpurun_out/r2h_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:coarse_qtmem -s 6 -c 1 -f -o gpurun_out/r2h_main_pass python bench.py --no-cpu-baseline --no-postings --no-parity --steps 2 --warmup 3 > gpurun_out/r2h_ncu_main.log 2>&1
echo "ncu main rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_and -c 1 -f -o gpurun_out/r2h_fused_and python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > gpurun_out/r2h_ncu_fused.log 2>&1
echo "ncu fused rc=$?"; tail -n 3 gpurun_out/r2h_ncu_fused.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_vecsim_coarse.py::test_coarse_path_is_exact" -x -q -m gpu -k "70000-128-40-10-1" > gpurun_out/r2h_racecheck_coarse.log 2>&1
echo "racecheck coarse rc=$?"; tail -n 6 gpurun_out/r2h_racecheck_coarse.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" -x -q -m gpu -k "0" > gpurun_out/r2h_racecheck_postings.log 2>&1
echo "racecheck postings rc=$?"; tail -n 6 gpurun_out/r2h_racecheck_postings.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" "tests/test_postings_parity.py::test_batch_decode_of_many_lists_matches_the_oracle_reader" -x -q -m gpu -k "0 or 1" > gpurun_out/r2h_memcheck_postings.log 2>&1
echo "memcheck postings rc=$?"; tail -n 6 gpurun_out/r2h_memcheck_postings.log
