#!/bin/bash
# GPU session r2x (1 GPU, short): the posting parity file + hybrid + C-API surface on the final tree (the constructors / union paths touched by the EXPLAINSCORE plumbing)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_postings_parity.py tests/test_hybrid_filtered.py tests/test_capi_surface.py tests/test_hybrid_state_machine.py -q -m gpu --tb=short > gpurun_out/r2x_tests.log 2>&1
echo "tests rc=$?"; tail -n 30 gpurun_out/r2x_tests.log
