#!/bin/bash
# GPU session r2u (2 GPUs): configs[4] at N=2 with the N>1 parity check (per-shard reference answers merged), the default bench line at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 2 --config 5 --steps 10 --warmup 3 2> gpurun_out/r2u_config5_n2.err | grep '^{' > gpurun_out/r2u_config5_n2.json
echo "config5 n2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_config5_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('parity_at_config'))
PY
tail -n 3 gpurun_out/r2u_config5_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2u_bench_n2.err | grep '^{' > gpurun_out/r2u_bench_n2.json
echo "bench n2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['roofline']['frac'], d.get('parity_at_config') and {k:d['parity_at_config'][k] for k in ('ids_equal','score_bits_equal')}, d.get('clocks'))
PY
tail -n 3 gpurun_out/r2u_bench_n2.err
