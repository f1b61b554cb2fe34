"""world_size-2 gloo test of bench.py's multi-rank logic (the only collectives on this path: max over ranks of the
elapsed time and the end-of-batch gather of the per-rank counters), plus the rank-0-only reference arm."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch.distributed as dist
import bench
dist.init_process_group("gloo")
r = dist.get_rank()
out = bench.reduce_over_ranks(10.0 + 5 * r, 20.0 - 3 * r, 1000 * (r + 1), 900 * (r + 1), 130, "cpu")
if r == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
"""


def test_reduce_over_ranks_gloo_world2(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER % ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(w)],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("[")][-1]
    ms, e2e, pairs, valid, launches = __import__("json").loads(line)
    assert (ms, e2e) == (15.0, 20.0)            # max over ranks
    assert (pairs, valid, launches) == (3000, 2700, 260)


def test_reduce_over_ranks_single_process():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reduce_over_ranks(1.0, 2.0, 3, 4, 5, "cpu") == (1.0, 2.0, 3, 4, 5)
    assert [bench.pingpong(t, 4) for t in range(8)] == [0, 1, 2, 3, 2, 1, 0, 1]
