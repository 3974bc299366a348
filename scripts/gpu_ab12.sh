#!/bin/bash
# A/B of landing-register budgets for the pipelined pull (variant libraries built with -DPIPE_MINB=n)
mkdir -p gpurun_out
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2))'
echo "== pytest variants on default lib"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants or chain or bitmap" 2>&1 | tail -2
for lib in "" build_variants/libb200grb_mb5.so build_variants/libb200grb_mb3.so build_variants/libb200grb_mb2.so; do
 for v in "" "--sources 512" "--sources 64"; do
  echo "-- lib=$lib $v"; B200GRB_LIB=${lib:+$PWD/$lib} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
 done
done
echo "-- default lib, second run (noise check)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err.log | tail -1 | python -c "$summ"
echo "-- default lib, steps 30"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/err.log | tail -1 | python -c "$summ"
