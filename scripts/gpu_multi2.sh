#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== chain N=$N"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/e.log | tail -1 > gpurun_out/scale_chain_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/scale_chain_n$N.json')); print({k: d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', round(d['e2e']['value']/1e9,1), d['e2e']['result_format'], d['e2e']['csr_handoff'], d['cpu_baseline'], d['clocks'])" || tail -8 gpurun_out/e.log
echo "== reference arm N=$N"; timeout 900 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2> gpurun_out/e2.log | tail -1 | cut -c1-400
