"""Sharded sensitivity sweep on the GPU path (BASELINE config 4): two ranks (gloo,
both on cuda:0 -- the box of the test run has one GPU) return the same per-candidate values as one
process.  The NCCL variant of the same script is tools/sweep_run.py under torchrun."""
import json
import os
import socket
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, extra=()):
    script = os.path.join(ROOT, "tools", "sweep_run.py")
    common = ["--arch", "resnet18", "--batches", "1", "--batch", "4", "--layers", "3", "--backend", "gloo"] + list(extra)
    if world == 1:
        cmd = [sys.executable, script] + common
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script] + common
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_sweep_two_ranks_equal_one_rank():
    one, two = _run(1), _run(2)
    assert one["world"] == 1 and two["world"] == 2
    assert one["values_sha256"] == two["values_sha256"]  # identical floats after the all_gather
    assert one["ranked_first8"] == two["ranked_first8"]
    assert one["flat_rows"] == two["flat_rows"]


def test_sweep_on_the_callers_net_equals_the_work_model_path():
    """functions.make_semilayers runs candidates 1.. on the caller's net when that net IS the pretrained model
    (no second model / engine); the general path builds a fresh pretrained work model (reference
    functions.py:258/393/528).  Same candidate values, and the caller's net leaves with the same weights
    (candidate 0's layer quantised, quirk Q3)."""
    fast, general = _run(1), _run(1, ["--force-work-model"])
    assert fast["values_sha256"] == general["values_sha256"]
    assert fast["net_after_sha256"] == general["net_after_sha256"]
    assert fast["ranked_first8"] == general["ranked_first8"]
