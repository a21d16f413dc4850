"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU, blocks sharded,
halo exchange over NVLink peer-mapped arenas (par2d_peer_*) or over NCCL (external transport).  Each rank
checks its blocks against the oracle (tolerance) and BITWISE against a single-GPU run of the same case:
the number of GPUs must not change a bit of the state, the residual or the residual history."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(nproc, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "dist_worker.py"), "--backend", "nccl", "--impl", "gpu", "--tol", "1e-11"] + extra
    env = dict(os.environ, OMP_NUM_THREADS="1", PAR2D_PEER_TIMEOUT_MS="30000")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)


@pytest.mark.parametrize("transport", ["peer", "external"])
@pytest.mark.parametrize("extra", [
    ["--case", "Case_spikefine24", "--steps", "12", "--batches", "3"],   # 24 blocks, irregular neighbours, 3 batches
    ["--case", "Case_rktnoz9", "--steps", "8"],                          # 9 unequal blocks, subsonic far field
    ["--synthetic", "256x192:4x4", "--steps", "6", "--batches", "2"],    # regular grid -> patch partition
])
def test_two_gpus_match_oracle_and_single_gpu_bitwise(transport, extra):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    p = _launch(2, extra + ["--transport", transport])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "RANK0 OK" in p.stdout and "RANK1 OK" in p.stdout
    assert f"transport={transport}" in p.stdout


@pytest.mark.parametrize("transport", ["peer", "external"])
@pytest.mark.parametrize("extra", [
    ["--case", "Case_spikefine24", "--steps", "10", "--batches", "2"],
    ["--case", "Case_rktnoz9", "--steps", "8"],
])
def test_two_gpus_second_order_branch(transport, extra):
    """cfg.accur = 1 across ranks: the halo lines carry the reconstruction at two Gauss points per edge face
    (8 doubles per edge cell, SpatialDiscretization.cpp:243-256); means, slopes, right-hand sides and the residual
    history must equal the single-GPU run bit for bit and the oracle within tolerance."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    p = _launch(2, extra + ["--transport", transport, "--accur", "1"])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "RANK0 OK" in p.stdout and "RANK1 OK" in p.stdout


def test_all_gpus_peer_mapped():
    n = _ngpus()
    if n < 3:
        pytest.skip("needs more than 2 GPUs")
    p = _launch(n, ["--case", "Case_spikefine24", "--steps", "10", "--batches", "2", "--transport", "peer"])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert all(f"RANK{r} OK" in p.stdout for r in range(n))
