#!/bin/bash
mkdir -p gpurun_out
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2))'
echo "== pytest parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for lib in "" build_variants/libb200grb_mb3.so build_variants/libb200grb_mb5.so; do
 for v in "" "--opt unroll=8" "--opt unroll=2"; do
  echo "-- lib=$lib $v"; B200GRB_LIB=${lib:+$PWD/$lib} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
 done
done
for v in "--opt fill_kernel=1" "--sources 512" "--sources 512 --opt unroll=8" "--sources 64" "--sources 1024 --steps 5"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
