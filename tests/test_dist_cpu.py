"""world_size-2 gloo runs of the one-process-per-GPU driver (par2d_b200/dist.py) on CPU, with the
oracle as the per-rank backend: partition, exchange-buffer ordering, dt all-reduce and the
block-ordered residual gather must reproduce the single-process result bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import REPO


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(nproc, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "dist_worker.py"), "--backend", "gloo", "--impl", "oracle"] + extra
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


@pytest.mark.parametrize("extra", [
    ["--case", "Case_spikefine24", "--steps", "12"],          # 24 blocks, LPT partition, irregular neighbours
    ["--case", "Case_rktnoz9", "--steps", "8"],               # 9 unequal blocks
    ["--synthetic", "64x64:4x4", "--steps", "6"],             # regular grid -> patch partition
])
def test_two_ranks_match_single_process_bitwise(extra):
    p = _launch(2, extra)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "RANK0 OK" in p.stdout and "RANK1 OK" in p.stdout


def test_three_ranks_uneven(tmp_path):
    p = _launch(3, ["--case", "Case_ffstep8", "--steps", "5"])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert all(f"RANK{r} OK" in p.stdout for r in range(3))


@pytest.mark.parametrize("extra", [
    ["--case", "Case_spikefine24", "--steps", "8", "--accur", "1"],
    ["--case", "Case_ffstep8", "--steps", "10", "--accur", "1"],
])
def test_two_ranks_second_order_branch_bitwise(extra):
    """cfg.accur = 1: the exchange buffers carry two Gauss points per edge face (8 doubles per edge cell); the
    sharded run must still reproduce the single-process means and residual history bit for bit."""
    p = _launch(2, extra)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "RANK0 OK" in p.stdout and "RANK1 OK" in p.stdout
