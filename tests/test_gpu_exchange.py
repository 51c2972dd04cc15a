"""-m gpu: the fused feature gather (peer stores from the ray kernel into symmetric memory) against an NCCL all-gather, on as
many GPUs as the box has (1 or 2 ranks; with one rank the 'peer' is the rank's own symmetric buffer)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_feat_gather_matches_all_gather():
    n = min(2, torch.cuda.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_feat_exchange.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "feat exchange ok" in r.stdout
