#!/bin/bash
# GPU session r2w (1 GPU, short): the C-host harness with EXPLAINSCORE of a nested result, the nested-aggregate tests incl. II_SearchTopN over a nested child
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_boundary_harness.py tests/test_postings_nested.py -q -m gpu --tb=short > gpurun_out/r2w_tests.log 2>&1
echo "tests rc=$?"; tail -n 40 gpurun_out/r2w_tests.log
