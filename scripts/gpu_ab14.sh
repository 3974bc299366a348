#!/bin/bash
mkdir -p gpurun_out
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2))'
echo "== pytest (all gpu)"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for v in "" "--opt early_exit=0" "--opt early_exit=2" "--opt hints=0" "--opt unroll=2" "--sources 512" "--sources 512 --opt early_exit=2" "--sources 64" "--sources 64 --opt early_exit=2"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1c.csv $B > gpurun_out/launches_r1c.log 2>&1; echo rc=$?
