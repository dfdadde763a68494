"""Integration test of the drop-in boundary (last-but-one row; BASELINE north_star:
"the functions.py / resnet.py entry points and the *_main.py flow stay as they are"):

the reference's UNMODIFIED ``resnet50_main.py`` (staged byte for byte under oracle/_ref by oracle/make_ref.py)
is executed with this repo's ``functions`` / ``resnet`` / ``imagenet`` modules first on sys.path, on the GPU,
and what it writes to ``output/resnet50_output.csv`` is compared with the fixture
``tests/golden/main_resnet50.npz`` = the same script run against the reference's own modules on the CPU
(oracle/run_main.py, 2 batches of 4 synthetic 64x64 images, seeded random labels and weights).

With random labels the baseline accuracy is 0, so ``acc >= preacc`` (resnet50_main.py:209) accepts every
semilayer in every phase: the run is the progressive 8 -> 6 -> 4 bit
re-quantisation of every channel.  What must then be IDENTICAL to the reference: the number of steps, the
bit / phase-flag / accuracy rows, the cumulative reduced-parameter total, per phase the multiset of
(layer number, channel count, parameter increment), and -- with the divide flavour pinned to the CPU one
 -- every final conv weight, bit for bit.  The ORDER of the semilayers inside a phase comes
from the KL/param ranking of the sweep, which this path perturbs by its u8 activation quantisation (the
reference never quantises activations, F2): it is reported (rank correlation), not asserted."""
import os

import numpy as np
import pytest

from helpers import GOLD

pytestmark = pytest.mark.gpu


def _phases(rows):
    """-> {(bit, flag): sorted list of (layer number, channels, param increment)} from the 7 CSV rows."""
    params = [float(v) for v in rows[0]]
    out = {}
    for i in range(1, len(params)):
        key = (int(rows[3][i]), int(rows[4][i]))
        out.setdefault(key, []).append((int(rows[5][i]), int(rows[6][i]), params[i] - params[i - 1]))
    return {k: sorted(v) for k, v in out.items()}


def test_unmodified_resnet50_main_against_the_package(monkeypatch):
    import run_main
    gold = np.load(os.path.join(GOLD, "main_resnet50.npz"))
    want = [r.split("\x1f") for r in gold["rows"]]
    b, n, hw, lseed, mseed = [int(v) for v in gold["config"]]
    monkeypatch.setenv("SLQ_DIV_MODE", "true")
    res = run_main.run_main("resnet50", "b200", batches=b, batch=n, hw=hw, loader_seed=lseed, model_seed=mseed)
    got = res["rows"]
    print("unmodified resnet50_main.py on the B200 path: %.1f s (reference on 6 CPU threads: %.1f s)" %
          (res["seconds"], float(gold["seconds"])))
    assert [len(r) for r in got] == [len(r) for r in want] == [280] * 7
    assert got[3] == want[3] and got[4] == want[4]          # bit-width and phase flag of every step
    assert [float(v) for v in got[1]] == [float(v) for v in want[1]]   # accuracy row (all 0.0)
    assert float(got[0][-1]) == float(want[0][-1]) == 18092032.0        # = 20,676,608 * 28 / 32
    assert _phases(got) == _phases(want)
    assert res["weights_sha256"] == str(gold["weights_sha256"])
    # order inside the 8-bit phase (sweep ranking): informative only
    g8 = [int(v) for v, bit in zip(got[5], got[3]) if bit == "8"]
    w8 = [int(v) for v, bit in zip(want[5], want[3]) if bit == "8"]
    same_pos = sum(1 for a, c in zip(g8, w8) if a == c)
    print("8-bit phase: %d of %d steps visit the same layer as the reference run" % (same_pos, len(w8)))
    kl = np.array([float(v) for v in got[2]])
    print("KL row, first steps: ours %s reference %s" % (kl[1:4], [float(v) for v in want[2][1:4]]))


def test_reference_files_are_staged_unmodified():
    """oracle/_ref holds byte-for-byte copies (digest written by oracle/make_ref.py)."""
    import hashlib
    import make_ref
    if not make_ref.staged():
        pytest.skip("oracle/_ref not staged")
    digest = hashlib.sha256()
    for f in make_ref.FILES:
        digest.update(f.encode() + b"\0" + open(os.path.join(make_ref.DST, f), "rb").read())
    assert digest.hexdigest() == open(os.path.join(make_ref.DST, "SHA256")).read().strip()
