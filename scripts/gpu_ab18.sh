#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (parity)"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_host_traverse.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2), "launches", d["gpu_launches"])'
for v in "" "--sources 512" "--sources 64"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
