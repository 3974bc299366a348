"""Run under torchrun (one rank per GPU): checks the 1-D row-partitioned BFS against the oracle, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import falkordb_b200 as fb          # noqa: E402
import oracle as orc                # noqa: E402
from falkordb_b200.dist_bfs import GpuBackend, bfs_gpu, partition  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fb.init()
scale = 15
A = orc.rmat_csr(scale, 16, 4)
be = GpuBackend(scale, 16, 4, rank, world)
ok = True
for src in [int(np.nonzero(np.diff(A.p))[0][k]) for k in (0, 11, 500)]:
    lv, par, edges, depth = bfs_gpu(be, src)
    wl, wp = orc.bfs(A, src)
    lo, hi = partition(A.nrows, rank, world)
    good = np.array_equal(lv.cpu().numpy(), wl[lo:hi].astype(np.int32)) and np.array_equal(par.cpu().numpy(), wp[lo:hi])
    e = torch.tensor([edges], device="cuda", dtype=torch.float64)
    dist.all_reduce(e)
    good = good and int(e.item()) == int(np.diff(A.p)[wl >= 0].sum())
    g = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(g, op=dist.ReduceOp.MIN)
    ok = ok and bool(g.item())
    if rank == 0:
        print(f"src {src}: depth {depth}, levels+parents+edges {'OK' if g.item() else 'MISMATCH'}")
be.close()
if rank == 0:
    print("DIST_BFS_PARITY", "PASS" if ok else "FAIL")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
