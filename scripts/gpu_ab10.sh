#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"]/1e9,1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))'
for v in "" "--opt fill_kernel=0" "--opt pull_kernel=3" "--opt pull_kernel=3 --opt pull_grid=8" "--opt pull_kernel=3 --opt pull_grid=5" "--opt pull_kernel=3 --opt pull_grid=32" "--opt pull_grid=8" "--opt pull_grid=32" "--opt pull_kernel=3 --sources 64" "--opt pull_kernel=3 --sources 512"; do
  echo "-- $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
