#!/bin/bash
# GPU session r2q (1 GPU): cp.async window staging in the posting kernels: parity tests, racecheck / memcheck, bench (incl. the
# clustered-corpus leg with contiguous clusters), fused launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_postings_parity.py tests/test_hybrid_filtered.py tests/test_boundary_harness.py -q -m gpu > gpurun_out/r2q_tests.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/r2q_tests.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" "tests/test_postings_parity.py::test_intersection_random_skewed_lists" -x -q -m gpu -k "0 or 1 or skewed" > gpurun_out/r2q_racecheck_postings.log 2>&1
echo "racecheck rc=$?"; tail -n 4 gpurun_out/r2q_racecheck_postings.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" "tests/test_postings_parity.py::test_intersection_random_skewed_lists" -x -q -m gpu -k "0 or 1 or skewed" > gpurun_out/r2q_memcheck_postings.log 2>&1
echo "memcheck rc=$?"; tail -n 4 gpurun_out/r2q_memcheck_postings.log
timeout 700 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'])
print('clustered', d.get('clustered_corpus'))
b=d['bm25_intersect']; print(b['value'], b['ms_per_query_set'], b['roofline']['device_ms_per_query_set'], b['roofline']['frac'], b['e2e']['value'], b.get('parity_at_config'), b.get('sequential_route_agrees'))
PY
tail -n 5 gpurun_out/r2q_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused --csv --log-file gpurun_out/r2q_fused_launches.csv python bench.py --no-cpu-baseline --no-parity --no-clustered --steps 2 --warmup 3 > /dev/null 2>&1
echo "fused launches rc=$?"
