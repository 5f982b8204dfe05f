"""GPU, >= 2 devices: z-slab decomposition over NCCL vs the single-GPU run (bitwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("bc,overlap,problem", [
    ("periodic", "overlap", "ot3d"), ("open", "overlap", "ot3d"), ("periodic", "nooverlap", "ot3d"),
    ("periodic", "overlap", "mri"), ("periodic", "overlap", "implode"), ("periodic", "overlap", "kh32"),
    ("periodic", "overlap", "ot3d_diss"), ("periodic", "overlap", "mri_diss"), ("periodic", "overlap", "rt_mhd")])
def test_slabs_over_nccl_match_single_gpu(native, bc, overlap, problem):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_mhd3d_check.py"),
           "5", str(13 * world), bc, overlap, problem]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0 and "identical=True" in out, out[-3000:]
