#!/bin/bash
# GPU session r2s (1 GPU): final state — smoke(), the whole -m gpu suite, the default bench line, the reference arm, the pair MMA probe
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 4 gpurun_out/r2s_smoke.log
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r2s_tests.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/r2s_tests.log
timeout 120 tools/mma_probe > gpurun_out/r2s_mma_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/r2s_mma_probe.log
timeout 700 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['parity_at_config'] and {k:d['parity_at_config'][k] for k in ('ids_equal','score_bits_equal','proven_by_tier')})
print('clustered', d.get('clustered_corpus'))
b=d['bm25_intersect']; print(b['value'], b['ms_per_query_set'], b['roofline']['device_ms_per_query_set'], b['roofline']['frac'], b['e2e']['value'], b.get('parity_at_config'))
print('cpu', d.get('cpu_baseline'), 'gpu_launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
PY
tail -n 3 gpurun_out/r2s_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err
echo "reference arm rc=$?"; head -c 900 gpurun_out/r2s_bench_reference.json; echo
