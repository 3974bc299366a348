#!/bin/bash
# compute-sanitizer memcheck over the newest kernels (binned pull, row-per-warp fill, CSR push, bitmap export, async hand-off)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bins or csr_frontier or traverse_to_host or export_bitmap or diagonal or single_entry" > gpurun_out/memcheck_pytest.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -20; tail -3 gpurun_out/memcheck.log
