#!/bin/bash
# GPU session r2n: why the CTA-pair build of the main pass does not launch (error text), then memcheck of it
mkdir -p gpurun_out
export VECSIM_B200_PAIR=1
timeout 200 python -m pytest tests/test_vecsim_coarse.py -x -q -m gpu -k "300000-64-256-10-1" > gpurun_out/r2n_pair.log 2>&1
echo "pair rc=$?"; grep -n "vecsim_b200:" gpurun_out/r2n_pair.log | head -5; tail -n 4 gpurun_out/r2n_pair.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_vecsim_coarse.py -x -q -m gpu -k "300000-64-256-10-1" > gpurun_out/r2n_pair_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -v "^=========     Host Frame\|^=========         in \|^=========$" gpurun_out/r2n_pair_memcheck.log | head -60
