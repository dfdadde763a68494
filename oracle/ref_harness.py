"""oracle/ref_harness.py -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference (``/root/reference/functions.py`` / ``resnet.py``) in this build
container so that oracle/gen_golden.py can produce golden vectors from it.
``/root/reference`` does not exist on the GPU box, so nothing under tests/ -m gpu, smoke() or
bench.py imports this module.

Two shims, no edits to reference files:
  1. ``imagenet`` is replaced by a stub module with a seeded synthetic ``val_loader``
     (reference imagenet.py:17-32 builds ImageFolders at import time and needs ./hogehoge).
  2. ``resnet.load_state_dict_from_url`` returns a seeded random-init state_dict
     (``pretrained='imagenet'`` is hard-coded at functions.py:258/393/528 and the mains).
"""
import csv
import importlib
import os
import sys
import types

import torch

REF_DIR = os.environ.get("SLQ_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "functions.py"))


def synthetic_loader(num_batches, batch, hw, seed=1):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_batches):
        x = torch.randn(batch, 3, hw, hw, generator=g)
        y = torch.randint(0, 1000, (batch,), generator=g)
        out.append((x, y))
    return out


def load_reference(loader=None):
    """Returns (functions, resnet) modules of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_DIR)
    for name in ("functions", "resnet", "imagenet"):
        sys.modules.pop(name, None)
    stub = types.ModuleType("imagenet")
    stub.val_loader = loader if loader is not None else synthetic_loader(1, 2, 64)
    stub.train_loader = None
    sys.modules["imagenet"] = stub
    sys.path.insert(0, REF_DIR)
    try:
        resnet = importlib.import_module("resnet")
        functions = importlib.import_module("functions")
    finally:
        sys.path.remove(REF_DIR)
    assert os.path.dirname(os.path.abspath(functions.__file__)) == os.path.abspath(REF_DIR)
    return functions, resnet


def unload_reference():
    for name in ("functions", "resnet", "imagenet"):
        sys.modules.pop(name, None)


def seeded_model(resnet_mod, arch, seed=0):
    torch.manual_seed(seed)
    return getattr(resnet_mod, arch)(num_classes=1000)


def install_seeded_pretrained(resnet_mod, arch, seed=0):
    sd = seeded_model(resnet_mod, arch, seed).state_dict()
    resnet_mod.load_state_dict_from_url = lambda url, progress=True: sd
    return sd


BLOCKS = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3], "resnet50": [3, 4, 6, 3]}
CONVS_PER_BLOCK = {"resnet18": 2, "resnet34": 2, "resnet50": 3}


def layer_map(arch, lnum):
    """CSV layer number (1-based) -> (layer_index, block_index, conv attribute name).

    Restates resnet50_main.py:81-136 / resnet18_main.py:86-115 with the INTENDED block map for
    ResNet-34 (resnet34_main.py reuses the ResNet-18 table)."""
    cpb = CONVS_PER_BLOCK[arch]
    flat_block = (lnum - 1) // cpb
    li = 0
    for li, nb in enumerate(BLOCKS[arch]):
        if flat_block < nb:
            break
        flat_block -= nb
    conv = "conv%d" % ((lnum - 1) % cpb + 1)
    return li, flat_block, conv


def read_deltaloss_csv(arch):
    """dataset/<arch>_deltaloss.csv -> (lnum[], cnum0[], rows of deltaloss floats).
    Blank trailing cells of the ResNet-18 file are stripped."""
    path = os.path.join(REF_DIR, "dataset", "%s_deltaloss.csv" % arch)
    with open(path, encoding="utf-8-sig") as f:
        rows = [[c for c in r if c != ""] for r in csv.reader(f)]
    rows = [r for r in rows if r]
    lnum = [int(v) for v in rows[0]]
    cnum = [int(v) - 1 for v in rows[1]]
    dl = [[float(v) for v in r] for r in rows[2:]]
    return lnum, cnum, dl


def p0_rows(functions, arch):
    """Policy P0 expressed with the reference's own split function:
    minus semilayer (deltaloss@4bit <= 0) -> 4 bit, plus -> 8 bit.
    Returns the reference-format 8-column rows, ascending global channel order."""
    lnum, cnum, dl = read_deltaloss_csv(arch)
    dl4 = dl[2]
    ds, rows = [], []
    for i in range(len(lnum)):
        li, bi, _ = layer_map(arch, lnum[i])
        ds.append([li, bi, lnum[i], cnum[i], dl[0][i], dl[1][i], dl4[i]])
        rows.append([li, bi, lnum[i], cnum[i], 8, 0, 32, i + 1])
    minus, plus = functions.make_divide_minusplusmodels(rows, ds, 6)
    for r in minus:
        r[4] = 4
    for r in plus:
        r[4] = 8
    allrows = sorted(minus + plus, key=lambda r: r[7])
    return allrows, minus, plus


def apply_rows(functions, net, arch, rows):
    layers = [net.layer1, net.layer2, net.layer3, net.layer4]
    for li, bi, lnum, cn, bit, _flag, _sel, _idx in rows:
        _, _, conv = layer_map(arch, lnum)
        m = getattr(layers[li][bi], conv)
        m.weight.data = functions.channel_wise_quantizationperchan(m.weight.data, bit, cn)
