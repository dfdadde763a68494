"""CPU suite for the host side that mirrors the reference's functions.py / resnet.py interface:
semilayer split, sweep, ranked list, evaluation and the module-tree contract, checked against
fixtures produced by the UNMODIFIED reference (oracle/gen_golden.py).  The CUDA entry points are
replaced by the oracle here (tests/cpu_standins.py)."""
import os

import numpy as np
import pytest
import torch

import cpu_standins
from helpers import GOLD


@pytest.fixture(scope="module")
def sw():
    return np.load(os.path.join(GOLD, "sweep_resnet18.npz"))


def _rows(a):
    return [[int(v) for v in r] for r in a]


def test_split_matches_reference(sw, capsys):
    import functions
    rows = _rows(sw["in_rows"])
    ds = [[r[0], r[1], r[2], r[3], float(d)] for r, d in zip(rows, sw["in_dl8"])]
    minus, plus = functions.make_divide_minusplusmodels(rows, ds, 4)
    assert minus == _rows(sw["minus_rows"]) and plus == _rows(sw["plus_rows"])
    assert "function debug:number of total channels= 3840" in capsys.readouterr().out
    # zero counts as minus; flags run 0,-1,-2.. / 1,2,3.. per layer
    assert minus[0][5] == 0 and plus[0][5] == 1
    assert min(r[5] for r in minus) == -15 and max(r[5] for r in plus) == 16


@pytest.mark.parametrize("force_work_model", [False, True])
def test_sweep_and_ranked_list_match_reference(sw, monkeypatch, force_work_model):
    """functions.make_semilayers_resnet18 + make_quantizedlists on the tiny synthetic loader the
    golden run used (2 batches of 4 images, 64x64): same semilayers, same sensitivities (KL/param),
    same ranked channel list; candidate 0 mutates the caller's net (reference quirk Q3).  Both routes of the
    candidates 1..: on the caller's net (it IS the pretrained model here) and on a fresh work model."""
    import functions
    import imagenet
    import resnet
    cpu_standins.install(monkeypatch)
    if force_work_model:
        monkeypatch.setattr(functions, "_is_pretrained", lambda *a, **k: False)
    loader = imagenet.synthetic_loader(2, 4, 64, seed=1)
    monkeypatch.setattr(imagenet, "val_loader", loader)
    torch.manual_seed(0)
    sd = resnet.resnet18(num_classes=1000).state_dict()
    monkeypatch.setattr(resnet, "load_state_dict_from_url", lambda url, progress=True: sd)
    net2 = resnet.resnet18(num_classes=1000, pretrained="imagenet")
    preacc, loss0, orig = functions.evaluate_acc_loss_softmax(net2, "cpu", loader)
    assert abs(loss0 - float(sw["loss0"])) < 1e-5 and preacc == float(sw["preacc"])
    minus, plus = _rows(sw["minus_rows"]), _rows(sw["plus_rows"])
    semilayers, orders = functions.make_semilayers_resnet18(net2, "cpu", orig, minus, plus)
    assert minus[-1] == [0, 0, 100, 0, 0, 0, 0, 0] and plus[-1][2] == 100  # caller's lists mutated
    assert [len(s) for s in semilayers] == sw["semilayer_sizes"].tolist()
    gold = sw["orders"]
    assert [o[0] for o in orders] == gold[:, 0].astype(int).tolist()
    got = np.array([o[1] for o in orders])
    assert np.allclose(got, gold[:, 1], rtol=2e-3, atol=1e-12), np.abs(got / gold[:, 1] - 1).max()
    c0 = int(sw["minus_rows"][0][3])
    assert np.array_equal(net2.layer1[0].conv1.weight.data[c0].reshape(-1).numpy(), sw["net2_row_after"])
    flat = functions.make_quantizedlists(semilayers, orders)
    assert [o[0] for o in orders] == sw["sorted_index"].tolist()
    assert np.array_equal(np.array(flat, np.int64), sw["flat"])


def test_kldiv_and_evaluate_conventions():
    import functions
    p = [torch.softmax(torch.randn(4, 10, generator=torch.Generator().manual_seed(i)), 1) for i in range(3)]
    q = [torch.softmax(torch.randn(4, 10, generator=torch.Generator().manual_seed(9 + i)), 1) for i in range(3)]
    want = sum((a[m] * (a[m] / b[m]).log()).sum() for a, b in zip(p, q) for m in range(4)) / 12
    assert abs(functions.KLdiv(p, q) - want.item()) < 1e-6
    assert functions.KLdiv(p, p) == 0.0


def test_quantizer_entry_points_inplace_contract(monkeypatch):
    import functions
    cpu_standins.install(monkeypatch)
    import slq_oracle as so
    t = torch.randn(4, 3, 3, 3, generator=torch.Generator().manual_seed(0))
    before = t.clone()
    out = functions.channel_wise_quantizationperchan(t, 4, 2)
    assert out is t
    assert torch.equal(t[0], before[0]) and not torch.equal(t[2], before[2])
    q, _, _, _, _ = so.quantize_row(before[2].reshape(-1).numpy(), 4)
    assert np.array_equal(t[2].reshape(-1).numpy(), q)
    fresh = functions.quantize_wgt(before[1], 8)
    assert fresh is not before[1] and fresh.shape == before[1].shape
    with pytest.raises(ZeroDivisionError):
        so.quantize_row(np.zeros(8, np.float32), 8)


def test_module_tree_contract():
    """Attribute names, state_dict keys and parameter shapes the mains index."""
    import resnet
    import torchvision
    for arch, counts in (("resnet18", 16), ("resnet34", 32), ("resnet50", 48)):
        net = getattr(resnet, arch)(num_classes=1000)
        tv = getattr(torchvision.models, arch)(weights=None)
        assert list(net.state_dict().keys()) == list(tv.state_dict().keys())
        for k, v in tv.state_dict().items():
            assert net.state_dict()[k].shape == v.shape
        blocks = [b for s in (net.layer1, net.layer2, net.layer3, net.layer4) for b in s]
        n = sum(1 for b in blocks for c in ("conv1", "conv2", "conv3") if hasattr(b, c))
        assert n == counts
        assert net.layers[0] is net.layer1
        w = net.layer2[0].conv2.weight
        assert w[3].data.numel() == w.shape[1] * 9  # resnet50_main.py:132 reads this
    net = resnet.resnet50(num_classes=1000)
    assert net.layer1[0].conv2.stride == (1, 1) and net.layer2[0].conv2.stride == (2, 2)  # v1.5
    assert net.layer2[0].conv1.stride == (1, 1) and net.layer2[0].downsample[0].stride == (2, 2)
    v0 = net.layer1[0].conv1.weight._version
    net.load_state_dict(net.state_dict())
    assert net.layer1[0].conv1.weight._version > v0  # how the engine sees a reload (slq_engine.Engine._sig)


def test_engine_invalidation_hooks(monkeypatch):
    """What tells a compiled engine that a layer must be re-packed: the per-storage write
    count kept by the quantizer entry points (writes through ``.data`` bump no version counter), the
    Parameter's version (load_state_dict), the storage pointer (.to / .data = ...).  A no-op ``net.to()``
    -- evaluate_acc_loss_softmax calls it before every evaluation -- must NOT invalidate anything."""
    import copy
    import functions
    import resnet
    import slq_engine
    cpu_standins.install(monkeypatch)
    net = resnet.resnet18(num_classes=10).eval()
    net._slq_dirty = False
    w = net.layer1[0].conv1.weight
    other = net.layer1[0].conv2.weight
    sig0, osig0 = slq_engine.Engine._sig(w), slq_engine.Engine._sig(other)
    real_quantize_rows = functions.quantize_rows

    def noting(tensor, rows, bits, **kw):  # the oracle stand-in + the product's own write note
        out = real_quantize_rows(tensor, rows, bits, **kw)
        resnet.note_weight_write(tensor)
        return out
    monkeypatch.setattr(functions, "quantize_rows", noting)
    w.data = functions.channel_wise_quantizationperchan(w.data, 8, 0)
    assert slq_engine.Engine._sig(w) != sig0 and slq_engine.Engine._sig(other) == osig0
    net.eval()
    net.to("cpu")
    assert not net._slq_dirty              # nothing moved: engines, scales and graphs are kept
    net.double()
    assert net._slq_dirty                  # dtype conversion re-allocates every parameter
    net.float()
    sig1 = slq_engine.Engine._sig(w)
    net.load_state_dict(net.state_dict())
    assert slq_engine.Engine._sig(w) != sig1
    net._slq_engines = {"x": object()}
    twin = copy.deepcopy(net)              # engines hold ctypes handles: they are not copied
    assert twin._slq_engines == {} and twin._slq_dirty
    assert torch.equal(twin.layer1[0].conv1.weight, net.layer1[0].conv1.weight)
    net.slq_invalidate()
    assert net._slq_dirty


def test_loader_side_formats():
    """imagenet.py mirror: raw u8 batches and the CPU normalisation the stem kernel reproduces
    (reference imagenet.py:14-15: ToTensor then Normalize(mean, std))."""
    import imagenet
    (x8, y), = imagenet.synthetic_loader(1, 3, 16, seed=5, dtype=torch.uint8)
    assert x8.dtype == torch.uint8 and x8.shape == (3, 3, 16, 16) and y.dtype == torch.int64
    got = imagenet.normalize_u8(x8)
    mean = torch.tensor(imagenet.MEAN).view(1, 3, 1, 1)
    std = torch.tensor(imagenet.STD).view(1, 3, 1, 1)
    want = (x8.float() / 255 - mean) / std
    assert got.dtype == torch.float32 and torch.equal(got, want)
    (xh, _), = imagenet.synthetic_loader(1, 2, 8, seed=5, dtype=torch.float16)
    (xf, _), = imagenet.synthetic_loader(1, 2, 8, seed=5)
    assert xh.dtype == torch.float16 and torch.equal(xh, xf.half())


def test_packed_row_sizes_match_the_library():
    """functions.quantize_rows sizes the packed-code blob with vectorised arithmetic; it must agree with
    slq_packed_row_bytes (a pure host function of the library) for every bit-width and row length."""
    import functions
    import slq_lib as L
    lib = L.lib()
    bits = np.arange(1, 9)
    for K in (1, 3, 64, 65, 576, 1152, 4608):
        want = [lib.slq_packed_row_bytes(K, int(b)) for b in bits]
        assert functions.packed_row_bytes(K, bits).tolist() == want, K


def test_packed_container_roundtrip():
    """SLQPACK1 (slq_store.dumps / loads) is plain host logic: header, 64-byte aligned arrays, dtypes."""
    import slq_store
    g = torch.Generator().manual_seed(0)
    packed = slq_store.Entry("layer1.0.conv1.weight", "packed", (4, 2, 3, 3), dict(
        bits=torch.tensor([8, 4, 4, 2], dtype=torch.int32), z=torch.tensor([-120, -7, -8, -1], dtype=torch.int32),
        s=torch.rand(4, generator=g), offsets=torch.tensor([0, 32, 48, 64], dtype=torch.int64),
        blob=torch.randint(0, 256, (80,), generator=g, dtype=torch.uint8)), K=18)
    dense = slq_store.Entry("bn1.running_var", "dense", (5,), dict(data=torch.rand(5, generator=g)))
    count = slq_store.Entry("bn1.num_batches_tracked", "dense", (), dict(data=torch.tensor(7)))
    raw = slq_store.dumps(slq_store.Snapshot([packed, dense, count], arch="BasicBlock"))
    assert raw[:8] == b"SLQPACK1"
    back = slq_store.loads(raw)
    assert back.arch == "BasicBlock" and [e.name for e in back.entries] == [packed.name, dense.name, count.name]
    for a, b in zip(back.entries, (packed, dense, count)):
        assert a.kind == b.kind and a.shape == b.shape and a.K == b.K
        for k in b.arrays:
            assert a.arrays[k].dtype == b.arrays[k].dtype and torch.equal(a.arrays[k], b.arrays[k])
    with pytest.raises(ValueError):
        slq_store.loads(b"NOTAPACK" + raw[8:])


def test_evaluate_cpu_route_is_the_reference_sequence():
    """Host tensors take the stock-torch route of functions._evaluate: identical to the reference's loop."""
    import functions
    g = torch.Generator().manual_seed(1)
    loader = [(torch.randn(5, 10, generator=g), torch.randint(0, 10, (5,), generator=g)) for _ in range(3)]
    net = torch.nn.Identity()
    acc, loss, outs = functions.evaluate_acc_loss_softmax(net, "cpu", loader)
    crit = torch.nn.CrossEntropyLoss()
    want_loss = (sum(crit(x, y) for x, y in loader) / 3).item()
    want_acc = (torch.cat([x.max(1)[1] for x, _ in loader]) == torch.cat([y for _, y in loader])).float().mean().item()
    assert loss == want_loss and abs(acc - want_acc) < 1e-7
    assert all(torch.equal(o, torch.softmax(x, 1)) for o, (x, _y) in zip(outs, loader))
    _, _, outs2, kl = functions._evaluate(net, "cpu", [(x * 1.1, y) for x, y in loader], ref_outputs=outs)
    assert abs(kl - functions.KLdiv(outs, outs2)) < 1e-9
