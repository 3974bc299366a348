"""bench.py's side workloads (configs 5 and 4 riding the chain line) are child processes, one per rank, with their own process group
on MASTER_PORT + offset.  The rendezvous mechanics -- dropping torchrun's agent store for the children, the port offset, rank 0
keeping the child's JSON line, a failing child costing only its sub-object -- run here under a real torchrun with two CPU ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_side_workloads_form_their_own_group_under_torchrun():
    port = 23000 + os.getpid() % 4000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "helpers", "side_parent.py"), "25"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("PARENT ")]
    assert len(line) == 1, r.stdout[-2000:]
    out = json.loads(line[0][len("PARENT "):])
    ok = out["ok"]
    assert ok["value"] == 3.0 and ok["n_gpus"] == 2, ok            # the children all-reduced 1 + 2 in their own group
    assert "--workload" in ok["argv"] and "bfs" in ok["argv"] and "26" in ok["argv"] and "clocks" not in ok and "child_wall_s" in ok
    assert "error" in out["fail"], out["fail"]                    # a child that dies costs its sub-object only
