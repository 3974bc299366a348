#!/bin/bash
# round 2, call A: parity of the new pull / count kernels, then A/B bench lines of the chain with kernel options
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_full_size.py > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-format csr"
for v in "pull_kernel=4" "pull_kernel=5" "pull_kernel=5 --opt unroll=8" "pull_kernel=5 --opt l2_window=67108864" "pull_kernel=5 --opt l2_window=33554432" "pull_kernel=5 --opt l2_window=33554432 --opt l2_reset=2" "pull_kernel=5 --opt l2_window=67108864 --opt l2_reset=2" "pull_kernel=5 --opt hints=0" "pull_kernel=5 --opt early_exit=2"; do
  tag=$(echo "$v" | tr -d ' ' | tr '=' '_' | tr -d '-')
  timeout 300 $B --opt $v > gpurun_out/r2a_bench_$tag.json 2> gpurun_out/r2a_bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2a_bench_{t}.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print(t, 'TTEPS %.3f ms %.2f'%(d['value']/1e12,d['ms_per_step']), {n:round(v['ms']/v['launches'],3) for n,v in k.items()})
except Exception as e: print(t,'ERR',e)
PY
done
# W=4 (256 sources) sanity
timeout 300 $B --sources 256 --opt pull_kernel=5 > gpurun_out/r2a_bench_s256.json 2> gpurun_out/r2a_bench_s256.err
timeout 300 $B --sources 1024 --opt pull_kernel=5 > gpurun_out/r2a_bench_s1024.json 2> gpurun_out/r2a_bench_s1024.err
python - <<'PY'
import json
for t in ('s256','s1024'):
    try:
        d=json.loads(open(f'gpurun_out/r2a_bench_{t}.json').read().strip().splitlines()[-1]); k=d['kernels']
        print(t, 'TTEPS %.3f ms %.2f'%(d['value']/1e12,d['ms_per_step']), {n:round(v['ms']/v['launches'],3) for n,v in k.items()})
    except Exception as e: print(t,'ERR',e)
PY
grep -h "l2_persist" gpurun_out/*.err | head -2
python -c "
import falkordb_b200 as fb
fb.init(); A=fb.rmat(10,8,1); A.prepare(True)
print('l2_persist_max', fb.get_stat('l2_persist_max'), 'l2_window_max', fb.get_stat('l2_window_max'))
"
