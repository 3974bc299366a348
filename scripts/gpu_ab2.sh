#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"]/1e9,1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))'
for v in "--opt unroll=1 --opt hints=0" "--opt unroll=2 --opt hints=0" "--opt unroll=4 --opt hints=0" "--opt unroll=4 --opt hints=1" "--pull-kernel 1" "--opt unroll=4 --sources 256" "--pull-kernel 1 --sources 256"; do
  echo "-- $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "== ncu unroll4"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bits_pull -s 1 -c 1 -f -o gpurun_out/prof_pull_u4 $B --opt unroll=4 --opt hints=0 > gpurun_out/prof1.log 2>&1; echo rc=$?
echo "== ncu unroll4 hints"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bits_pull -s 1 -c 1 -f -o gpurun_out/prof_pull_u4h $B --opt unroll=4 --opt hints=1 > gpurun_out/prof2.log 2>&1; echo rc=$?
echo "== ncu mp"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bits_pull_mp -s 1 -c 1 -f -o gpurun_out/prof_pull_mp $B --pull-kernel 1 > gpurun_out/prof3.log 2>&1; echo rc=$?
echo "== ncu fill"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bits_fill -s 1 -c 1 -f -o gpurun_out/prof_fill $B > gpurun_out/prof4.log 2>&1; echo rc=$?
