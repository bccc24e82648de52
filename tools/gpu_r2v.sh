#!/bin/bash
# GPU session r2v (1 GPU): final state of the round — smoke(), the whole -m gpu suite, the default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 4 gpurun_out/r2v_smoke.log
timeout 2400 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2v_tests.log 2>&1
echo "tests rc=$?"; tail -n 25 gpurun_out/r2v_tests.log
timeout 700 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['traffic'], d['parity_at_config'] and {k:d['parity_at_config'][k] for k in ('ids_equal','score_bits_equal','proven_by_tier')})
b=d['bm25_intersect']; print(b['value'], b['roofline']['device_ms_per_query_set'], b['roofline']['frac'], b['roofline']['traffic'], b['e2e']['value'], b.get('parity_at_config') and b['parity_at_config']['ids_equal'])
print('cpu', d.get('cpu_baseline',{}).get('value'), 'gpu_launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
PY
tail -n 3 gpurun_out/r2v_bench.err
