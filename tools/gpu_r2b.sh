#!/bin/bash
# GPU session r2b: coarse-route tests after the two-pass rewrite + kernel variants timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vecsim_coarse.py tests/test_vecsim_sharded.py -x -q -m gpu > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -n 15 gpurun_out/r2b_tests.log
for v in "default" "VECSIM_B200_ACC=1" "VECSIM_B200_FIXED=0" "VECSIM_B200_TIER2=0"; do
  if [ "$v" = "default" ]; then envs=""; else envs="$v"; fi
  echo "== $v" >> gpurun_out/r2b_bench.log
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-postings --steps 30 --warmup 5 2>> gpurun_out/r2b_bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print(json.dumps({'ms_per_step':d['ms_per_step'],'value':d['value'],'e2e':d['e2e']['value'],'kernel_us':r['avg_launch_us'],'frac':r['frac'],'launches':d['gpu_launches'],'agree':d['config'].get('host_device_results_agree'),'clocks':d['clocks'],'b1':d['single_query_as_served']['ms_per_query']}))
" >> gpurun_out/r2b_bench.log
done
cat gpurun_out/r2b_bench.log
