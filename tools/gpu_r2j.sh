#!/bin/bash
# GPU session r2j (1 GPU): posting tests (13 codecs, wide field masks, large-window pivot search, compact candidate lists),
# racecheck of the fused route, bench, full ncu capture of the MAIN pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_postings_parity.py tests/test_boundary_harness.py tests/test_hybrid_filtered.py -q -m gpu --durations=8 > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -n 30 gpurun_out/r2j_tests.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_postings_parity.py::test_fused_batch_search_equals_the_per_query_chains" -x -q -m gpu -k "0 or 1" > gpurun_out/r2j_racecheck_postings.log 2>&1
echo "racecheck postings rc=$?"; tail -n 4 gpurun_out/r2j_racecheck_postings.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'])
b=d['bm25_intersect']; print(b['value'], b['ms_per_query_set'], b['roofline']['device_ms_per_query_set'], b['roofline']['frac'], b['e2e']['value'], b['e2e'].get('decode_rate_postings_per_s'), b['parity_at_config'], b.get('sequential_route_agrees'))
PY
tail -n 5 gpurun_out/r2j_bench.err
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'coarse_qtmem_kernel<\(bool\)0, \(int\)3, \(int\)0, \(int\)1>' -s 3 -c 1 -f -o gpurun_out/r2j_main_pass python bench.py --no-cpu-baseline --no-postings --no-parity --steps 2 --warmup 3 > gpurun_out/r2j_ncu_main.log 2>&1
echo "ncu main rc=$?"; tail -n 3 gpurun_out/r2j_ncu_main.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused --csv --log-file gpurun_out/r2j_fused_launches.csv python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /dev/null 2>&1
echo "fused launches rc=$?"
