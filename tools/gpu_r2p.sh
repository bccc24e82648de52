#!/bin/bash
# GPU session r2p (1 GPU): CTA-pair main pass with the 14-deep ring of half-size stages vs the multicast build, same box
mkdir -p gpurun_out
export VECSIM_B200_PAIR=1
timeout 240 python -m pytest tests/test_vecsim_coarse.py tests/test_vecsim_fullsize.py -x -q -m gpu -k "300000-64-256 or fullsize or 2000000" > gpurun_out/r2p_pair_tests.log 2>&1
rc=$?; echo "pair tests rc=$rc"; tail -n 4 gpurun_out/r2p_pair_tests.log
if [ $rc -eq 0 ]; then
  timeout 400 python bench.py --no-cpu-baseline --no-postings --no-clustered --steps 30 --warmup 5 > gpurun_out/r2p_bench_pair.json 2> gpurun_out/r2p_bench_pair.err
  echo "bench pair rc=$?"
fi
export VECSIM_B200_PAIR=0
timeout 400 python bench.py --no-cpu-baseline --no-postings --no-clustered --no-parity --steps 30 --warmup 5 > gpurun_out/r2p_bench_nopair.json 2> gpurun_out/r2p_bench_nopair.err
echo "bench nopair rc=$?"
python - <<'PY'
import json
for f in ('pair','nopair'):
    try:
        d=json.load(open(f'gpurun_out/r2p_bench_{f}.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['sustained']['ms_per_step'], d['sustained']['clocks'], d['parity_at_config'] and {k:d['parity_at_config'][k] for k in ('ids_equal','score_bits_equal')})
    except Exception as e:
        print(f, 'no result', e)
PY
