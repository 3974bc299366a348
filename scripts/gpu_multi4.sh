#!/bin/bash
# multi-GPU evidence at N = 4 (one box): distributed BFS parity, chain weak scaling, BFS scale 26, triangles scale 24
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== dist bfs parity N=$N"; timeout 600 $TR scripts/dist_check.py 2>&1 | grep -E "src|DIST_BFS" | tail -5
echo "== chain N=$N"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/e.log | tail -1 > gpurun_out/scale_chain_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/scale_chain_n$N.json')); print({k: d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', round(d['e2e']['value']/1e9,1), d['e2e']['result_format'], d['clocks'])" || tail -5 gpurun_out/e.log
echo "== chain N=1 (same box)"; timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/e1.log | tail -1 > gpurun_out/scale_chain_n1.json
python -c "
import json; d=json.load(open('gpurun_out/scale_chain_n1.json')); print({k: d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', round(d['e2e']['value']/1e9,1))" || tail -5 gpurun_out/e1.log
echo "== bfs N=$N scale 26"; timeout 1200 $TR bench.py --gpus $N --workload bfs --scale 26 --bfs-sources 8 --warmup 1 2> gpurun_out/e2.log | tail -1 | tee gpurun_out/bfs_n${N}_s26.json | cut -c1-700; tail -2 gpurun_out/e2.log | cut -c1-300
echo "== triangles N=$N scale 24"; timeout 1200 $TR bench.py --gpus $N --workload triangles --scale 24 --steps 2 --warmup 1 2> gpurun_out/e4.log | tail -1 | tee gpurun_out/tri_n${N}_s24.json | cut -c1-700; tail -2 gpurun_out/e4.log | cut -c1-300
