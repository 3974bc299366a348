#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== dist bfs parity (N=$N)"; timeout 600 $TR scripts/dist_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|Setting OMP" | tail -8
echo "== bench chain N=$N"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/multi_err.log | tail -1 > gpurun_out/bench_chain_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/bench_chain_n$N.json')); print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/multi_err.log
echo "== bench bfs N=$N scale 24"; timeout 900 $TR bench.py --gpus $N --workload bfs --scale 24 --bfs-sources 8 --warmup 2 2> gpurun_out/multi_err2.log | tail -1 | tee gpurun_out/bench_bfs_n$N.json | cut -c1-600 || tail -5 gpurun_out/multi_err2.log
echo "== bench bfs N=1 scale 24 (same box)"; timeout 900 python bench.py --workload bfs --scale 24 --bfs-sources 8 --warmup 2 2>&1 | tail -1 | cut -c1-400
