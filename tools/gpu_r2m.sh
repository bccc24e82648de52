#!/bin/bash
# GPU session r2m (1 GPU): CTA-pair main pass (VECSIM_B200_PAIR=1: tcgen05.mma.cta_group::2) — parity first, then speed;
# batched hybrid path (test + configs[4] bench at N=1)
mkdir -p gpurun_out
export VECSIM_B200_PAIR=1
timeout 240 python -m pytest tests/test_vecsim_coarse.py -x -q -m gpu -k "300000-64-256" > gpurun_out/r2m_pair_first.log 2>&1
rc=$?; echo "pair first test rc=$rc"; tail -n 12 gpurun_out/r2m_pair_first.log
if [ $rc -eq 0 ]; then
  timeout 900 python -m pytest tests/test_vecsim_coarse.py tests/test_vecsim_fullsize.py tests/test_vecsim_parity.py -q -m gpu > gpurun_out/r2m_pair_tests.log 2>&1
  echo "pair tests rc=$?"; tail -n 8 gpurun_out/r2m_pair_tests.log
  timeout 400 python bench.py --no-cpu-baseline --no-postings --steps 30 --warmup 5 > gpurun_out/r2m_bench_pair.json 2> gpurun_out/r2m_bench_pair.err
  echo "bench pair rc=$?"
fi
export VECSIM_B200_PAIR=0
timeout 400 python bench.py --no-cpu-baseline --no-postings --steps 30 --warmup 5 > gpurun_out/r2m_bench_nopair.json 2> gpurun_out/r2m_bench_nopair.err
echo "bench nopair rc=$?"
python - <<'PY'
import json
for f in ('pair','nopair'):
    try:
        d=json.load(open(f'gpurun_out/r2m_bench_{f}.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['sustained']['ms_per_step'], d['sustained']['clocks'], d['parity_at_config'])
    except Exception as e:
        print(f, 'no result', e)
PY
timeout 300 python -m pytest tests/test_hybrid_filtered.py -q -m gpu > gpurun_out/r2m_hybrid_tests.log 2>&1
echo "hybrid tests rc=$?"; tail -n 5 gpurun_out/r2m_hybrid_tests.log
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2m_config5_n1.json 2> gpurun_out/r2m_config5_n1.err
echo "config5 rc=$?"; head -c 700 gpurun_out/r2m_config5_n1.json; echo; tail -n 3 gpurun_out/r2m_config5_n1.err
