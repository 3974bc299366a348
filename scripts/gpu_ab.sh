#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"]/1e9,1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d.get("cpu_baseline"))'
for v in "--pull-kernel 0" "--pull-kernel 1" "--pull-kernel 1 --sources 256" "--pull-kernel 0 --sources 256" "--pull-kernel 1 --sources 1024 --steps 3"; do
  echo "-- $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
echo "-- default full"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | python -c "$summ"
