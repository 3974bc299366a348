#!/bin/bash
mkdir -p gpurun_out
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
echo "== dist bfs parity N=8"; timeout 600 run 8 scripts/dist_check.py 2>&1 | grep -E "src|DIST_BFS" | tail -5
for N in 1 2 4 8; do
  echo "== chain N=$N"
  if [ $N = 1 ]; then timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/e.log | tail -1 > gpurun_out/scale_chain_n$N.json
  else timeout 900 run $N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/e.log | tail -1 > gpurun_out/scale_chain_n$N.json; fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_chain_n$N.json')); print({k: d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', round(d['e2e']['value']/1e9,1), d['clocks'])" || tail -3 gpurun_out/e.log
done
echo "== bfs N=8 scale 26"; timeout 1200 run 8 bench.py --gpus 8 --workload bfs --scale 26 --bfs-sources 16 --warmup 2 2> gpurun_out/e2.log | tail -1 | tee gpurun_out/bfs_n8_s26.json | cut -c1-700; tail -2 gpurun_out/e2.log | cut -c1-300
echo "== bfs N=1 scale 26"; timeout 1200 python bench.py --workload bfs --scale 26 --bfs-sources 16 --warmup 2 2> gpurun_out/e3.log | tail -1 | tee gpurun_out/bfs_n1_s26.json | cut -c1-500
echo "== triangles N=4 scale 24"; timeout 1200 run 4 bench.py --gpus 4 --workload triangles --scale 24 --steps 2 --warmup 1 2> gpurun_out/e4.log | tail -1 | tee gpurun_out/tri_n4_s24.json | cut -c1-700; tail -2 gpurun_out/e4.log | cut -c1-300
echo "== triangles N=1 scale 24"; timeout 1200 python bench.py --workload triangles --scale 24 --steps 2 --warmup 1 2> gpurun_out/e5.log | tail -1 | tee gpurun_out/tri_n1_s24.json | cut -c1-700
