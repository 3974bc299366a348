#!/bin/bash
mkdir -p gpurun_out
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2))'
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for v in "--opt pull_kernel=3" "--opt pull_kernel=4" "--opt pull_kernel=4 --opt hints=0" "--opt pull_kernel=4 --opt pull_grid=8" "--opt pull_kernel=3 --sources 512" "--opt pull_kernel=4 --sources 512" "--opt pull_kernel=3 --sources 64" "--opt pull_kernel=4 --sources 64" "--opt pull_kernel=4 --sources 128" "--opt pull_kernel=4 --sources 1024 --steps 5"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr --opt pull_kernel=4"
echo "== full: pull mid"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_pull_mid$' -s 1 -c 1 -f -o gpurun_out/prof_r1c_mid $B > gpurun_out/p1.log 2>&1; echo rc=$?
echo "== full: pull small"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_pull_small$' -s 1 -c 1 -f -o gpurun_out/prof_r1c_small $B > gpurun_out/p2.log 2>&1; echo rc=$?
