#!/bin/bash
# mid-round profile of the new pull / fill kernels
mkdir -p gpurun_out
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2))'
for v in "" "--opt pull_grid=10" "--opt pull_grid=16"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr"
echo "== full: pull pipe"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_pull_pipe$' -s 1 -c 1 -f -o gpurun_out/prof_r1b_pull $B > gpurun_out/p1.log 2>&1; echo rc=$?
echo "== full: fill rows"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_fill_rows$' -s 1 -c 1 -f -o gpurun_out/prof_r1b_fill $B > gpurun_out/p2.log 2>&1; echo rc=$?
