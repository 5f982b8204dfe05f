"""GPU, >= 2 devices: z-slab decomposition vs the single-GPU run (bitwise); z halo by copy engines over peer-mapped
state arrays (default on one node) and by NCCL send/recv."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(bc, overlap, problem, halo, extra_planes=0):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_mhd3d_check.py"),
           "5", str(13 * world + extra_planes), bc, overlap, problem, halo]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0 and "identical=True" in out, out[-3000:]
    return out


@pytest.mark.parametrize("bc,overlap,problem", [
    ("periodic", "overlap", "ot3d"), ("open", "overlap", "ot3d"), ("periodic", "nooverlap", "ot3d"),
    ("periodic", "overlap", "mri"), ("periodic", "overlap", "implode"), ("periodic", "overlap", "kh32"),
    ("periodic", "overlap", "ot3d_diss"), ("periodic", "overlap", "mri_diss"), ("periodic", "overlap", "rt_mhd")])
def test_slabs_match_single_gpu(native, bc, overlap, problem):
    """Default halo path (peer copies on one node with peer access, NCCL otherwise)."""
    _run(bc, overlap, problem, "peer")


@pytest.mark.parametrize("bc,overlap,problem", [("periodic", "overlap", "ot3d"), ("open", "nooverlap", "ot3d"),
                                                ("periodic", "overlap", "mri"), ("periodic", "overlap", "implode"),
                                                ("periodic", "overlap", "ot3d_diss")])
def test_slabs_over_nccl_match_single_gpu(native, bc, overlap, problem):
    out = _run(bc, overlap, problem, "nccl")
    assert "peer_copies=False" in out, out[-2000:]


@pytest.mark.parametrize("bc,problem", [("periodic", "ot3d"), ("open", "ot3d"), ("periodic", "implode")])
def test_uneven_slabs_by_peer_copies(native, bc, problem):
    """One plane more than a multiple of the ranks: the first slab is one plane thicker, so a rank writes its neighbour's
    ghost planes at the NEIGHBOUR's offsets.  The GPU boxes have NVLink peer access: the peer-copy path must be the one used."""
    out = _run(bc, "overlap", problem, "peer", extra_planes=1)
    assert "peer_copies=True" in out, out[-2000:]
