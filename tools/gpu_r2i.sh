#!/bin/bash
# GPU session r2i (1 GPU): full -m gpu suite (slop factor, union child order, HAMMING, fused search with the window pre-pass),
# racecheck of the fused posting route after the fold fix, bench, full ncu capture of the MAIN pass (demangled kernel name)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/r2i_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -n 40 gpurun_out/r2i_tests.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" -x -q -m gpu -k "0 or 1" > gpurun_out/r2i_racecheck_postings.log 2>&1
echo "racecheck postings rc=$?"; tail -n 6 gpurun_out/r2i_racecheck_postings.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest "tests/test_postings_parity.py::test_legacy_scorers_divide_by_the_slop_of_the_hit" "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" -x -q -m gpu -k "and3 or or3 or phrase or chains" > gpurun_out/r2i_memcheck_postings.log 2>&1
echo "memcheck postings rc=$?"; tail -n 6 gpurun_out/r2i_memcheck_postings.log
timeout 600 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'])
b=d['bm25_intersect']; print(b['value'], b['ms_per_query_set'], b['roofline'], b['e2e']['value'], b['parity_at_config'], b.get('sequential_route_agrees'))
PY
tail -n 5 gpurun_out/r2i_bench.err
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'coarse_qtmem_kernel<\(bool\)0, 3, 0, 1>|coarse_qtmem_kernel<0, 3, 0, 1>' -s 3 -c 1 -f -o gpurun_out/r2i_main_pass python bench.py --no-cpu-baseline --no-postings --no-parity --steps 2 --warmup 3 > gpurun_out/r2i_ncu_main.log 2>&1
echo "ncu main rc=$?"; tail -n 4 gpurun_out/r2i_ncu_main.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_and -c 1 -f -o gpurun_out/r2i_fused_and python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > gpurun_out/r2i_ncu_fused.log 2>&1
echo "ncu fused rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused --csv --log-file gpurun_out/r2i_fused_launches.csv python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /dev/null 2>&1
echo "fused launches rc=$?"
