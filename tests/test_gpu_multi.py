"""GPU tests that launch helper processes: the multi-GPU collection path (needs >= 2 GPUs) and
the drop-in ./sift binary (reference main.cpp linked against the shim)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_collection_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tests", "run_collection_nccl.py")],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "collection ok" in out.stdout


def test_drop_in_sift_binary(tmp_path):
    """the reference's main.cpp, unchanged, on top of sift_shim.cpp."""
    exe = os.path.join(ROOT, "oracle", "_ref", "sift_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sift_b200 is built only where /root/reference exists")
    g = os.path.join(ROOT, "tests", "golden")
    out = subprocess.run([exe, os.path.join(g, "image1.png"), os.path.join(g, "image2.png")], capture_output=True,
                         text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    finals = [int(l.split(":")[1]) for l in out.stdout.splitlines() if l.startswith("Final keypoints")]
    assert finals == [1286, 1430]          # SURVEY.md section 4 known answers
    assert (tmp_path / "matches.png").stat().st_size > 10000
    assert (tmp_path / "keypoints.png").stat().st_size > 10000
