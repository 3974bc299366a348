#!/bin/bash
# round 2, call F: ordered-frontier parity, then ncu evidence: launch list of the chain bench + --set full per kernel
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hot_set_order or kernel_variants or pull_bins" > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
BQ="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches.csv $BQ > gpurun_out/r2f_launches_bench.log 2>&1
for k in k_pull_seg k_pull_small k_bits_fill_v3 k_bits_count_csa k_csr_push k_ordered_flops; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -o gpurun_out/r2f_prof_$k -f $BQ > gpurun_out/r2f_ncu_$k.log 2>&1
done
# BFS: per-kernel time of one sweep at scale 26
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_bfs_launches.csv python bench.py --workload bfs --scale 26 --bfs-sources 2 --warmup 1 --bfs-parity 0 > gpurun_out/r2f_bfs_launches.log 2>&1
# delta-sync kernels, pagerank, triangles (single GPU) -- bench lines with per-kernel rooflines
timeout 400 python bench.py --workload delta --steps 3 --warmup 1 > gpurun_out/r2f_delta.json 2> gpurun_out/r2f_delta.err; tail -c 2000 gpurun_out/r2f_delta.json; tail -2 gpurun_out/r2f_delta.err
timeout 400 python bench.py --workload pagerank --steps 3 --warmup 1 > gpurun_out/r2f_pagerank.json 2> gpurun_out/r2f_pagerank.err; tail -c 1500 gpurun_out/r2f_pagerank.json; tail -2 gpurun_out/r2f_pagerank.err
timeout 500 python bench.py --workload triangles --scale 22 --steps 3 --warmup 1 > gpurun_out/r2f_tri_s22.json 2> gpurun_out/r2f_tri_s22.err; tail -c 1500 gpurun_out/r2f_tri_s22.json; tail -2 gpurun_out/r2f_tri_s22.err
ls -la gpurun_out | grep r2f | head -40
