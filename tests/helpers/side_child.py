"""Stand-in for `bench.py --workload ...` in tests/test_bench_side.py: joins the children's own process group (gloo, env://),
all-reduces one number and lets rank 0 print the JSON line a side workload would print."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if "--fail" in sys.argv and rank == world - 1:
    sys.exit(3)
if world > 1:
    dist.init_process_group("gloo")
t = torch.tensor([rank + 1.0])
if world > 1:
    dist.all_reduce(t)
if rank == 0:
    print("some log line")
    print(json.dumps({"value": float(t.item()), "n_gpus": world, "argv": sys.argv[1:], "clocks": {"x": 1}}))
if world > 1:
    dist.destroy_process_group()
