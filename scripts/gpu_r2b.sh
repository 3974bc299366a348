#!/bin/bash
# round 2, call B: the BFS engine, small to large, every step bounded in time and host memory
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
timeout 300 python -m pytest tests/test_dist_bfs.py tests/test_fp64.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
for sc in 20 22 24; do
  timeout 200 python bench.py --workload bfs --scale $sc --bfs-sources 8 --warmup 2 > gpurun_out/r2b_bfs_s$sc.json 2> gpurun_out/r2b_bfs_s$sc.err
  tail -c 1500 gpurun_out/r2b_bfs_s$sc.json; tail -2 gpurun_out/r2b_bfs_s$sc.err
done
nvidia-smi --query-gpu=memory.used --format=csv
timeout 400 python bench.py --workload bfs --scale 26 --bfs-sources 8 --warmup 2 --bfs-parity 1 > gpurun_out/r2b_bfs_s26.json 2> gpurun_out/r2b_bfs_s26.err
tail -c 1500 gpurun_out/r2b_bfs_s26.json; tail -3 gpurun_out/r2b_bfs_s26.err
