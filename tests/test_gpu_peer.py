"""Cross-GPU epilogues over peer memory (csrc/peer.cu), checked against the CPU oracle.

The exchange kernels spin (bounded) on flags raised by other ranks, so each scenario runs in its own process with a
timeout: a broken build fails this test, not the CUDA context of the whole pytest session.

  * loopback       -- 4 ranks inside one process on one GPU, windows are plain device addresses (no IPC);
  * same device    -- 2 processes on cuda:0, windows travel as CUDA IPC handles over a gloo rendezvous;
  * two GPUs       -- 2 processes, one per GPU, NCCL and the peer-memory kernels side by side (skipped on a 1-GPU box).
"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(cmd, timeout):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, f"{' '.join(cmd)}\n--- stdout\n{r.stdout[-3000:]}\n--- stderr\n{r.stderr[-3000:]}"
    return r.stdout


def _torchrun(nproc, script, *args):
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
            "--master-port", str(_free_port()), script, *args]


def test_peer_exchange_loopback_four_ranks_one_process():
    out = _run([sys.executable, "tests/multigpu/peer_loopback.py", "--world", "4"], timeout=400)
    assert "peer_loopback ok: world=4" in out


def test_peer_exchange_ipc_two_processes_one_gpu():
    out = _run(_torchrun(2, "tests/multigpu/sharded_check.py", "--same-device"), timeout=600)
    assert "sharded_check ok: world=2" in out


def test_sharded_nccl_and_peer_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run(_torchrun(2, "tests/multigpu/sharded_check.py"), timeout=600)
    assert "sharded_check ok: world=2" in out
    # the one-process / G-devices model (ncclCommInitAll, grouped calls) through the C ABI alone
    out = _run([sys.executable, "tests/multigpu/nccl_single_process.py", "--gpus", "2"], timeout=600)
    assert "nccl_single_process ok: 2 devices" in out
