"""Run under torchrun (one rank per GPU): the 1-D row-partitioned BFS (B200_bfs_partitioned over NCCL) against the oracle,
level and parent bit for bit, with and without the sparse exchange."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import falkordb_b200 as fb          # noqa: E402
import oracle as orc                # noqa: E402
from falkordb_b200.dist_bfs import PartitionedBfs, partition  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fb.init()
scale = int(os.environ.get("DIST_CHECK_SCALE", "16"))


def bcast(b):
    box = [b]
    dist.broadcast_object_list(box, src=0)
    return box[0]


A = orc.rmat_csr(scale, 16, 4)
deg = np.diff(A.p)
pb = PartitionedBfs(scale, 16, 4, rank, world, bcast)
lo, hi = partition(A.nrows, rank, world)
ok = True
for sparse in (1, 0):
    fb.set_option("bfs_sparse_exchange", sparse)
    for src in [int(np.nonzero(deg)[0][k]) for k in (0, 11, 500)] + [int(np.argmax(deg))]:
        lv, par, info = pb.run(src)
        wl, wp = orc.bfs(A, src)
        good = np.array_equal(lv, wl[lo:hi]) and np.array_equal(par, wp[lo:hi]) and info["edges"] == int(deg[wl >= 0].sum())
        g = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(g, op=dist.ReduceOp.MIN)
        ok = ok and bool(g.item())
        if rank == 0:
            print(f"sparse={sparse} src {src}: {info} {'OK' if g.item() else 'MISMATCH'}")
pb.close()
if rank == 0:
    print("DIST_BFS_PARITY", "PASS" if ok else "FAIL")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
