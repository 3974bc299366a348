#!/bin/bash
# first GPU pass: sanitizer on a tiny case, parity tests, smoke, small + full bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== sanitizer (tiny)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "rmat_generator or build_bool or rowwise or heavy_rows" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/sanitizer.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench scale 20"; timeout 600 python bench.py --scale 20 --steps 5 --warmup 3 > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err; echo "rc=$?"; cat gpurun_out/bench_s20.json; tail -5 gpurun_out/bench_s20.err
echo "== bench scale 24"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s24.json 2> gpurun_out/bench_s24.err; echo "rc=$?"; cat gpurun_out/bench_s24.json; tail -5 gpurun_out/bench_s24.err
