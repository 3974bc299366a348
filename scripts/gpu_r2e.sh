#!/bin/bash
# round 2, call E (--gpus 2): the partitioned BFS over NCCL against the oracle, then short 2-rank bench lines
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
DIST_CHECK_SCALE=16 timeout 300 $TR scripts/dist_check.py > gpurun_out/r2e_dist_check16.log 2>&1; tail -12 gpurun_out/r2e_dist_check16.log
DIST_CHECK_SCALE=20 timeout 300 $TR scripts/dist_check.py > gpurun_out/r2e_dist_check20.log 2>&1; tail -4 gpurun_out/r2e_dist_check20.log
timeout 400 $TR bench.py --gpus 2 --workload bfs --scale 24 --bfs-sources 8 --warmup 2 > gpurun_out/r2e_bfs_n2_s24.json 2> gpurun_out/r2e_bfs_n2_s24.err; tail -c 1800 gpurun_out/r2e_bfs_n2_s24.json; tail -3 gpurun_out/r2e_bfs_n2_s24.err
timeout 400 $TR bench.py --gpus 2 --workload bfs --scale 26 --bfs-sources 8 --warmup 2 --bfs-parity 1 > gpurun_out/r2e_bfs_n2_s26.json 2> gpurun_out/r2e_bfs_n2_s26.err; tail -c 1800 gpurun_out/r2e_bfs_n2_s26.json; tail -3 gpurun_out/r2e_bfs_n2_s26.err
timeout 400 $TR bench.py --gpus 2 --workload triangles --scale 22 --steps 3 --warmup 1 > gpurun_out/r2e_tri_n2_s22.json 2> gpurun_out/r2e_tri_n2_s22.err; tail -c 1500 gpurun_out/r2e_tri_n2_s22.json; tail -3 gpurun_out/r2e_tri_n2_s22.err
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_chain_n2.json 2> gpurun_out/r2e_chain_n2.err; tail -c 2500 gpurun_out/r2e_chain_n2.json; tail -3 gpurun_out/r2e_chain_n2.err
