"""Runs under torchrun in tests/test_bench_side.py: every rank calls bench.side_workload the way run_b200 does."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch.distributed as dist

import bench

dist.init_process_group("gloo")                      # the parents' own group stays up while the children run
a = types.SimpleNamespace(gpus=int(os.environ["WORLD_SIZE"]), edge_factor=16, seed=1, side_timeout=int(sys.argv[1]))
child = os.path.join(os.path.dirname(os.path.abspath(__file__)), "side_child.py")
out = {"ok": bench.side_workload(a, "bfs", 26, ["--bfs-sources", "8"], 1, script=child),
       "fail": bench.side_workload(a, "triangles", 24, ["--fail"], 2, script=child)}
dist.barrier()
if int(os.environ["RANK"]) == 0:
    print("PARENT " + json.dumps(out))
dist.destroy_process_group()
