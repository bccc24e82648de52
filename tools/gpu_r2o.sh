#!/bin/bash
# GPU session r2o (1 GPU): fused posting search after the child-count templating (5 CTAs / SM): parity tests, bench, a full ncu
# capture of fused_and_kernel with sources for the per-line view, clustered-corpus leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_postings_parity.py tests/test_hybrid_filtered.py tests/test_boundary_harness.py -q -m gpu > gpurun_out/r2o_tests.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/r2o_tests.log
timeout 700 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['parity_at_config'])
print('clustered', d.get('clustered_corpus'))
b=d['bm25_intersect']; print(b['value'], b['ms_per_query_set'], b['roofline']['device_ms_per_query_set'], b['roofline']['frac'], b['e2e']['value'], b.get('parity_at_config'), b.get('sequential_route_agrees'))
print('cpu', d.get('cpu_baseline'))
PY
tail -n 5 gpurun_out/r2o_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_and -c 1 -f -o gpurun_out/r2o_fused_and python bench.py --no-cpu-baseline --no-parity --no-clustered --steps 2 --warmup 3 > gpurun_out/r2o_ncu_fused.log 2>&1
echo "ncu fused rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused --csv --log-file gpurun_out/r2o_fused_launches.csv python bench.py --no-cpu-baseline --no-parity --no-clustered --steps 2 --warmup 3 > /dev/null 2>&1
echo "fused launches rc=$?"
