#!/bin/bash
# GPU session r2t (1 GPU): nested aggregates — the new tests, the C-host harness (variant 5), the whole posting parity file, memcheck of one nested shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_postings_nested.py tests/test_boundary_harness.py -q -m gpu --tb=short > gpurun_out/r2t_nested.log 2>&1
echo "nested rc=$?"; tail -n 60 gpurun_out/r2t_nested.log
timeout 900 python -m pytest tests/test_postings_parity.py tests/test_hybrid_filtered.py -q -m gpu --tb=short > gpurun_out/r2t_parity.log 2>&1
echo "parity rc=$?"; tail -n 15 gpurun_out/r2t_parity.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_postings_nested.py -q -m gpu -k "test_phrase_over_nested_unions or (test_nested_aggregates and TFIDF and b_c) or test_union_of_terms" --tb=short > gpurun_out/r2t_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 12 gpurun_out/r2t_memcheck.log
