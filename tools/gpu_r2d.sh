#!/bin/bash
# GPU session r2d: full GPU suite after the sample-pass rewrite + boundary layer; bench default vs single accumulator
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -n 25 gpurun_out/r2d_tests.log
for v in "default" "VECSIM_B200_ACC=1"; do
  if [ "$v" = "default" ]; then envs=""; else envs="$v"; fi
  echo "== $v" >> gpurun_out/r2d_bench.log
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-postings --no-parity --steps 30 --warmup 5 2>> gpurun_out/r2d_bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print(json.dumps({'ms_per_step':d['ms_per_step'],'value':d['value'],'e2e':d['e2e']['value'],'kernel_us':r['avg_launch_us'],'frac':r['frac'],'launches':d['gpu_launches'],'sustained':d['sustained'],'b1':d['single_query_as_served']['ms_per_query']}))
" >> gpurun_out/r2d_bench.log
done
cat gpurun_out/r2d_bench.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --no-postings --no-parity --no-cpu-baseline --steps 2 --warmup 3 --sustained-seconds 0 > gpurun_out/r2d_ncu_bench.log 2>&1
echo "ncu rc=$?"
