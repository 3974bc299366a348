#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2), "csr", round(e["csr_handoff"]["value"]/1e9,1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["clocks"])'
for v in "" "--steps 30 --warmup 5" "--opt pull_kernel=0 --opt pull_grid=16" "--opt pull_grid=4" "--opt pull_grid=6" "--opt pull_grid=12" "--sources 64" "--sources 128" "--sources 512" "--sources 1024 --steps 5"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
