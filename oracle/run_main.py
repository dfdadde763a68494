"""oracle/run_main.py -- TEST INFRASTRUCTURE ONLY: runs one of the reference's UNMODIFIED driver scripts
(``resnet18/34/50_main.py``, staged under oracle/_ref by oracle/make_ref.py) against a chosen pair of
``functions`` / ``resnet`` modules:

  * ``impl="reference"``: the reference's own functions.py / resnet.py (CPU) -- produces the fixture
    ``tests/golden/main_<arch>.npz`` (this file run as a script, in the build container);
  * ``impl="b200"``: this repo's drop-in modules (the package directory first on sys.path) on the GPU --
    what tests/test_reference_mains_gpu.py does.

Both get the two shims of SURVEY.md Appendix F: ``imagenet.val_loader`` is a seeded synthetic loader (the
reference's imagenet.py opens ./hogehoge at import time) and ``resnet.load_state_dict_from_url`` returns a
seeded random-init state_dict (no network).  The main runs in a scratch directory holding ``dataset/`` and
``output/``; its result CSV and final weights are returned.
"""
import csv
import hashlib
import importlib
import importlib.util
import io
import os
import runpy
import sys
import tempfile
import time
import types
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
REF = os.path.join(HERE, "_ref")


def _synthetic_loader():
    spec = importlib.util.spec_from_file_location("slq_imagenet_for_loader", os.path.join(PKG, "imagenet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.synthetic_loader


def conv_weights_sha256(net):
    h = hashlib.sha256()
    for name, p in net.state_dict().items():
        if p.dim() == 4:
            h.update(name.encode())
            h.update(p.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def run_main(arch, impl, batches=2, batch=4, hw=64, loader_seed=1, model_seed=0, quiet=True, strip_blank_csv=True):
    """Returns dict(rows=list of 7 lists of strings, weights_sha256, seconds, globals=the script's globals)."""
    assert impl in ("reference", "b200")
    script = os.path.join(REF, "%s_main.py" % arch)
    if not os.path.isfile(script):
        raise RuntimeError("oracle/_ref is not staged (python oracle/make_ref.py in the build container)")
    loader = _synthetic_loader()(batches, batch, hw, seed=loader_seed)
    saved_modules = {k: sys.modules.pop(k, None) for k in ("functions", "resnet", "imagenet")}
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    src_dir = REF if impl == "reference" else PKG
    tmp = tempfile.mkdtemp(prefix="slq_main_")
    try:
        sys.path.insert(0, src_dir)
        if impl == "reference":
            stub = types.ModuleType("imagenet")
            stub.val_loader, stub.train_loader = loader, None
            sys.modules["imagenet"] = stub
        else:
            imagenet = importlib.import_module("imagenet")
            imagenet.val_loader = loader
        resnet = importlib.import_module("resnet")
        functions = importlib.import_module("functions")
        assert os.path.dirname(os.path.abspath(functions.__file__)) == os.path.abspath(src_dir)
        torch.manual_seed(model_seed)
        sd = getattr(resnet, arch)(num_classes=1000).state_dict()
        resnet.load_state_dict_from_url = lambda url, progress=True: {k: v.clone() for k, v in sd.items()}
        os.makedirs(os.path.join(tmp, "dataset"))
        os.makedirs(os.path.join(tmp, "output"))
        for f in os.listdir(os.path.join(REF, "dataset")):
            raw = open(os.path.join(REF, "dataset", f), "rb").read()
            if strip_blank_csv and f.startswith("resnet18"):
                # quirk Q2: the shipped resnet18 table has two trailing empty cells per row, on which
                # resnet18_main.py:79 raises ValueError; they are stripped, nothing else is touched
                text = raw.decode("utf-8-sig")
                text = "\n".join(line.rstrip(",") for line in text.splitlines()) + "\n"
                raw = text.encode("utf-8-sig")
            with open(os.path.join(tmp, "dataset", f), "wb") as fh:
                fh.write(raw)
        os.chdir(tmp)
        t0 = time.time()
        sink = io.StringIO()
        if quiet:
            with redirect_stdout(sink):
                g = runpy.run_path(script, run_name="__main__")
        else:
            g = runpy.run_path(script, run_name="__main__")
        dt = time.time() - t0
        with open(os.path.join(tmp, "output", "%s_output.csv" % arch)) as fh:
            rows = [r for r in csv.reader(fh) if r]
        return dict(rows=rows, weights_sha256=conv_weights_sha256(g["net"]), seconds=dt, globals=g,
                    log_tail=sink.getvalue()[-2000:])
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k in ("functions", "resnet", "imagenet"):
            sys.modules.pop(k, None)
            if saved_modules[k] is not None:
                sys.modules[k] = saved_modules[k]


def main():
    import make_ref
    make_ref.make()
    arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
    torch.set_num_threads(int(os.environ.get("SLQ_FIXTURE_THREADS", "8")))
    res = run_main(arch, "reference", quiet=False)
    out = os.path.join(ROOT, "tests", "golden", "main_%s.npz" % arch)
    np.savez_compressed(out, rows=np.array(["\x1f".join(r) for r in res["rows"]]),
                        weights_sha256=res["weights_sha256"], seconds=res["seconds"],
                        config=np.array([2, 4, 64, 1, 0]))
    print("wrote", out, "rows", [len(r) for r in res["rows"]], "in %.1f s" % res["seconds"])


if __name__ == "__main__":
    main()
