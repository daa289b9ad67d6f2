"""The slab path as it really runs: one process per GPU under torchrun, NCCL bootstrap, halo by
peer puts over NVLink (and by NCCL send/recv) -- must reproduce the single-GPU run bit for bit.
Needs two GPUs; skipped on a one-GPU box (tests/test_gpu_slabs.py covers the same logic there
with virtual ranks)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
    except (OSError, subprocess.TimeoutExpired):
        return 0
    return sum(1 for line in out.splitlines() if line.startswith("GPU "))


@pytest.mark.parametrize("halo", ["put", "nccl"])
def test_two_ranks_reproduce_single_gpu_run(halo):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    if halo == "nccl":
        env["SPHB200_HALO"] = "nccl"
    else:
        env.pop("SPHB200_HALO", None)
    port = "29581" if halo == "put" else "29582"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", port,
                        os.path.join(ROOT, "tools", "check_multigpu.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MULTIGPU CHECK PASSED" in r.stdout, r.stdout[-2000:]
    if halo == "nccl":
        assert "halo mode: nccl" in r.stdout
    # "put" falls back to NCCL by itself on a box without peer access between the two GPUs
