#!/bin/bash
# round-end evidence: tests, default bench (both arms), triangles / bfs workloads, ncu launch list + full captures
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?; cut -c1-900 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-700 gpurun_out/bench_ref.json
echo "== triangles s20"; timeout 300 python bench.py --workload triangles --scale 20 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-700
echo "== triangles s22"; timeout 600 python bench.py --workload triangles --scale 22 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-700
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/launches_final.log 2>&1; echo rc=$?
echo "== full: pull"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_pull$' -s 1 -c 1 -f -o gpurun_out/prof_final_pull $B > gpurun_out/p1.log 2>&1; echo rc=$?
echo "== full: fill"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_fill$' -s 1 -c 1 -f -o gpurun_out/prof_final_fill $B > gpurun_out/p2.log 2>&1; echo rc=$?
echo "== full: count"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_bits_count$' -s 1 -c 1 -f -o gpurun_out/prof_final_count $B > gpurun_out/p3.log 2>&1; echo rc=$?
