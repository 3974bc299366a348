#!/bin/bash
# final round-1 evidence (after the early-hop paths): tests, smoke, both bench arms, launch list, top kernels
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo rc=$?; tail -2 gpurun_out/bench_r1f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'])"
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cut -c1-300 gpurun_out/bench_r1f_ref.json
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1f.csv $B > gpurun_out/launches_r1f.log 2>&1; echo rc=$?
for k in k_bits_pull_mid k_bits_pull_small k_bits_fill_rows; do
  echo "== full: $k"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^$k\$" -s 1 -c 1 -f -o gpurun_out/prof_r1f_$k $B > gpurun_out/p_$k.log 2>&1; echo rc=$?
done
rm -f gpurun_out/prof_r1c_* gpurun_out/prof_r1b_* 2>/dev/null
du -sh gpurun_out
