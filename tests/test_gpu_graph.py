"""The compute calls are capture-safe: a config-1-sized chain of six of them recorded into a CUDA graph on the context's
stream replays to bit-identical results (tools/graph_replay.py; run in its own process because it switches torch's
current stream)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_chain_captured_into_a_cuda_graph_replays_bit_identically():
    r = subprocess.run([sys.executable, "tools/graph_replay.py", "--iters", "20"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["bit_identical"] is True and line["calls"] == 6
