#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (all gpu)"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_gpu.log
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2), "csr", round(e["csr_handoff"]["value"]/1e9,1))'
for v in "" "--e2e-subbatches 1" "--e2e-subbatches 2" "--sources 512 --e2e-subbatches 4" "--sources 512 --e2e-subbatches 8"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
