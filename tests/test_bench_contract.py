"""The bench contract on the CPU side: `bench.py --impl reference` (the reference's arithmetic on the host
cores) prints ONE JSON line with the keys the driver reads, also when torchrun pins OMP_NUM_THREADS=1; the
GPU arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    p = _run(["--impl", "reference", "--arch", "resnet18", "--ref-batch", "2", "--steps", "1", "--warmup", "1"],
             env={"OMP_NUM_THREADS": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("ResNet-18") and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    # "reference" when the unmodified reference modules are staged under oracle/_ref (oracle/make_ref.py),
    # else the oracle's restatement ("port")
    staged = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "resnet.py"))
    assert cb["kind"] == ("reference" if staged else "port")
    assert cb["value"] == d["value"] and "batch 2" in cb["sample"]
    # every core of the affinity mask, not torchrun's OMP_NUM_THREADS=1
    assert cb["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--arch", "resnet18", "--steps", "1", "--warmup", "0"], env={"RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    p = _run(["--steps", "1", "--warmup", "1"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
