#!/bin/bash
# ncu evidence: launch list of one short bench run + full captures of the dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "== pytest (remaining)"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1; echo rc=$?
echo "== full: pull"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bits_pull -s 2 -c 2 -f -o gpurun_out/prof_pull $B > gpurun_out/prof_pull.log 2>&1; echo rc=$?
echo "== full: tiles"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bits_tiles -s 4 -c 2 -f -o gpurun_out/prof_tiles $B > gpurun_out/prof_tiles.log 2>&1; echo rc=$?
echo "== variants"
for v in "--pull-mode 0" "--sources 256" "--sources 1024 --steps 3" "--bits-mode 0 --sources 8 --steps 3"; do
  echo "-- $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','flops_per_step','nnz_out_per_step')}, {k: (round(v['ms']/d['steps'],3)) for k, v in d['kernels'].items()}, 'e2e', d['e2e']['value'])"
done
