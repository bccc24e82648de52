#!/bin/bash
# GPU session r2l (1 GPU): full -m gpu suite (numeric leaves, wildcard children, expansion-sized unions, 13 codecs, DebugInfo),
# bench (fused search with lockstep searches + thresholded per-query top-N), fused launch list
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2l_tests.log
tail -n 25 gpurun_out/r2l_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'])
b=d['bm25_intersect']; print(b['value'], b['ms_per_query_set'], b['roofline']['device_ms_per_query_set'], b['roofline']['frac'], b['e2e']['value'], b['e2e'].get('decode_rate_postings_per_s'), b.get('parity_at_config'), b.get('sequential_route_agrees'))
PY
tail -n 5 gpurun_out/r2l_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused --csv --log-file gpurun_out/r2l_fused_launches.csv python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /dev/null 2>&1
echo "fused launches rc=$?"
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" -x -q -m gpu -k "0 or 1" > gpurun_out/r2l_racecheck_postings.log 2>&1
echo "racecheck postings rc=$?"; tail -n 4 gpurun_out/r2l_racecheck_postings.log
