#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (all gpu)"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_r1d.json')); print(d['value'], d['ms_per_step'], json.dumps(d['roofline']), json.dumps(d['cpu_baseline']))"; tail -2 gpurun_out/bench_r1d.err
