#!/bin/bash
# round 2, call D: parity of materialise v3 / ordered frontiers, then A/B bench lines
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_full_size.py > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-format csr"
for v in "fill_kernel=3" "fill_kernel=1" "perm_push=0" "small_split=1" "fill_kernel=1 --opt perm_push=0"; do
  tag=$(echo "$v" | tr -d ' ' | tr '=' '_' | tr -d '-')
  timeout 300 $B --opt $v > gpurun_out/r2d_bench_$tag.json 2> gpurun_out/r2d_bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2d_bench_{t}.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print(t, 'TTEPS %.3f ms %.2f e2e %.2f ms launches %d'%(d['value']/1e12,d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches']), {n:round(v['ms']/v['launches'],3) for n,v in k.items()})
except Exception as e: print(t,'ERR',e); print(open(f'gpurun_out/r2d_bench_{t}.err').read()[-600:])
PY
done
timeout 300 $B --sources 256 > gpurun_out/r2d_bench_s256.json 2> gpurun_out/r2d_bench_s256.err
timeout 300 $B --sources 1024 > gpurun_out/r2d_bench_s1024.json 2> gpurun_out/r2d_bench_s1024.err
python - <<'PY'
import json
for t in ('s256','s1024'):
    try:
        d=json.loads(open(f'gpurun_out/r2d_bench_{t}.json').read().strip().splitlines()[-1]); k=d['kernels']
        print(t, 'TTEPS %.3f ms %.2f'%(d['value']/1e12,d['ms_per_step']), {n:round(v['ms']/v['launches'],3) for n,v in k.items()})
    except Exception as e: print(t,'ERR',e)
PY
# the full default bench line (with the CPU baseline and the bitmap e2e arm) and the reference arm, short
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; tail -c 3000 gpurun_out/r2d_bench_default.json; tail -3 gpurun_out/r2d_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; tail -c 1500 gpurun_out/r2d_bench_ref.json; tail -3 gpurun_out/r2d_bench_ref.err
