#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"]/1e9,1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))'
for v in "--sources 64" "--sources 128" "--sources 256" "--sources 512" "--sources 1024 --steps 3" "--sources 256 --opt early_exit=2"; do
  echo "-- $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sources 256"
echo "== ncu pull W4"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_pull$' -s 1 -c 1 -f -o gpurun_out/prof_pull_w4 $B > gpurun_out/prof1.log 2>&1; echo rc=$?
