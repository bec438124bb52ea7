"""N > 1 on real GPUs: two ranks (one per GPU) step their shards of ragged and block-aligned worlds
through both sharded paths (tiled: position gather only; pair-shared: gather + reduce-scatter of
partial accelerations) and rank 0 compares with the unsharded context AND with the CPU oracle (sampled force rows,
most-attracting links incl. near ties across ranks; tools/sharded_check.py).
Skipped on a single-GPU box; the CPU-side plumbing is covered by tests/test_dist_cpu.py (gloo)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharded_matches_unsharded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert out.count("-> PASS") == 18 and "FAIL" not in out, out[-3000:]
