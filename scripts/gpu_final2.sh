#!/bin/bash
# round-1 evidence after the degree-binned pull / row-per-warp materialise / bitmap hand-off
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo rc=$?; cut -c1-1200 gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_r1c.err
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-700 gpurun_out/bench_r1c_ref.json
summ='import sys, json
d = json.loads(sys.stdin.read())
e = d["e2e"]
print({k: round(d[k],3) for k in ("value","ms_per_step")}, {k: (round(v["ms"]/d["steps"],3)) for k, v in d["kernels"].items()}, "e2e", round(e["value"]/1e9,1), e.get("result_format"), round(e["ms_per_step"],2), "csr", round(e["csr_handoff"]["value"]/1e9,1))'
for v in "--opt fill_kernel=2" "--sources 512" "--sources 64" "--sources 1024 --steps 5"; do
  echo "-- $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v 2> gpurun_out/err.log | tee gpurun_out/last.json | tail -1 | python -c "$summ" || tail -5 gpurun_out/err.log
done
echo "== bfs s24"; timeout 600 python bench.py --workload bfs --scale 24 --bfs-sources 16 --warmup 2 2>&1 | tail -1 | cut -c1-600
echo "== triangles s22"; timeout 600 python bench.py --workload triangles --scale 22 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1c.csv $B > gpurun_out/launches_r1c.log 2>&1; echo rc=$?
for k in k_bits_pull_mid k_bits_pull_small k_bits_fill_rows k_bits_count k_bits_pull_long k_bits_push; do
  echo "== full: $k"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^$k\$" -s 1 -c 1 -f -o gpurun_out/prof_r1c_$k $B > gpurun_out/p_$k.log 2>&1; echo rc=$?
done
