#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (all gpu)"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), round(e["ms_per_step"],2), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"])'
for v in "" "--opt fused_prep=0" "--sources 256"; do
  echo "-- $v"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
