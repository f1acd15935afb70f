"""Two-GPU check of the board-sharded solver (run with gpurun --gpus 2): launched through torch.distributed.run, every
rank must report the single-GPU exploitability trace (float32 sums are re-associated across shards: 1e-5 relative)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("which", ["fhp", "hulh"])
def test_two_rank_sharded_matches_single_gpu(which):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "sharded_check.py"), which], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    for a, b in zip(r["sharded"], r["single"]):
        assert abs(a - b) <= 1e-5 * abs(b), r


@pytest.mark.parametrize("algo", ["CFRPlus", "LinearCFR"])
def test_two_rank_board_engine_is_bit_identical_to_one_rank(algo):
    """board engine: the chance sums are int64 fixed point, so 2 ranks == 1 rank exactly (traces and trunk tables)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519",
                          os.path.join(ROOT, "tools", "sharded_board_check.py"), "999", algo], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["sharded"] == r["single"] and r["checksums"] == r["single_checksums"], r
