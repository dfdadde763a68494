"""oracle/gen_golden.py -- regenerates tests/golden/*.npz and the P0 bit-assignment table from the
UNMODIFIED reference (imported from /root/reference via oracle/ref_harness.py).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden.py            # everything (a few minutes of CPU)
    python oracle/gen_golden.py quant p0   # selected parts

Outputs (all small, committed):
  tests/golden/quant_rows.npz      rows -> functions.quantize_wgt outputs (+ z, s32, codes)
  tests/golden/model_<arch>.npz    seeded-init hashes, P0 fake-quant hashes, logits of the
                                   reference fp32 forward on 2 synthetic 224x224 images
  tests/golden/sweep_resnet18.npz  functions.make_semilayers_resnet18 + make_quantizedlists on a
                                   tiny synthetic loader
  <package>/data/p0_bits.npz       per-channel 4/8-bit assignment (policy P0) derived
                                   from dataset/*_deltaloss.csv with functions.make_divide_minusplusmodels
"""
import hashlib
import os
import struct
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
ARCHS = ("resnet18", "resnet34", "resnet50")


def _derive(w, q, bit):
    """z, s32 and codes from the reference's own python scalars (functions.py:35-40)."""
    t = torch.from_numpy(w)
    mn, mx = torch.min(t).item(), torch.max(t).item()
    scale = (mx - mn) / (2 ** bit - 1)
    z = round(mn / scale)
    s32 = np.float32(scale)
    k = np.rint(q.astype(np.float64) / float(s32))
    codes = (k - z).astype(np.int64)
    recon = ((codes + z).astype(np.float32) * s32).astype(np.float32)
    assert np.array_equal(recon, q), "reconstruction identity failed"
    return z, s32, codes.astype(np.int32)


def gen_quant():
    functions, resnet = rh.load_reference()
    rng = np.random.default_rng(1234)
    rows, bits = [], []
    Ks = [64, 128, 256, 512, 576, 1024, 1152, 2048, 2304, 4608, 7, 27, 147, 100]
    for K in Ks:
        std = np.sqrt(2.0 / max(K, 1))
        rows.append((rng.standard_normal(K) * std).astype(np.float32))
    # adversarial rows
    rows.append(np.arange(256, dtype=np.float32) / 2 - 63.75)        # exact .5 ties at 8 bit
    rows.append(np.linspace(-1, 1, 64).astype(np.float32))
    rows.append((np.abs(rng.standard_normal(128)) + 3).astype(np.float32))   # all positive, z > 0
    rows.append((-np.abs(rng.standard_normal(128)) - 0.1).astype(np.float32))  # all negative
    rows.append((rng.standard_normal(64) * 1e-6).astype(np.float32))   # tiny range
    rows.append((rng.standard_normal(64) * 1e4).astype(np.float32))    # large range
    rows.append(np.array([0.0, 1.0], np.float32))                      # 2 elements
    rows.append((1.0 + rng.standard_normal(64) * 1e-3).astype(np.float32))   # large positive z
    # real seeded-init rows from the reference model
    net = rh.seeded_model(resnet, "resnet50", 0)
    rows.append(net.layer1[0].conv1.weight.data[0].reshape(-1).numpy().copy())  # Appendix G KAT
    rows.append(net.layer3[1].conv2.weight.data[5].reshape(-1).numpy().copy())
    rows.append(net.layer4[2].conv2.weight.data[100].reshape(-1).numpy().copy())
    out = {}
    n = 0
    for w in rows:
        for bit in (8, 6, 4, 2):
            q = functions.quantize_wgt(torch.from_numpy(w.copy()), bit).numpy()
            z, s32, codes = _derive(w, q, bit)
            out["w%d" % n] = w
            out["bit%d" % n] = np.int32(bit)
            out["q%d" % n] = q
            out["z%d" % n] = np.int32(z)
            out["s%d" % n] = np.float32(s32)
            out["c%d" % n] = codes.astype(np.int32)  # may exceed 2^bit-1 on tie rows
            n += 1
    # progressive 8 -> 6 -> 4 on one row through channel_wise_quantizationperchan
    t = net.layer2[0].conv2.weight.data[:4].clone()
    chain = [t[1].reshape(-1).numpy().copy()]
    for bit in (8, 6, 4):
        t = functions.channel_wise_quantizationperchan(t, bit, 1)
        chain.append(t[1].reshape(-1).numpy().copy())
    out["chain"] = np.stack(chain)
    out["n"] = np.int32(n)
    # constant row -> ZeroDivisionError
    try:
        functions.quantize_wgt(torch.ones(16), 8)
        out["const_raises"] = np.int32(0)
    except ZeroDivisionError:
        out["const_raises"] = np.int32(1)
    np.savez_compressed(os.path.join(GOLD, "quant_rows.npz"), **out)
    print("quant_rows.npz:", n, "cases")
    rh.unload_reference()


def gen_p0():
    functions, resnet = rh.load_reference()
    out = {}
    for arch in ARCHS:
        rows, minus, plus = rh.p0_rows(functions, arch)
        arr = np.array([[r[2], r[3], r[4]] for r in rows], np.int32)  # lnum, cnum0, bit
        out[arch] = arr
        n4 = int((arr[:, 2] == 4).sum())
        print(arch, "channels", len(arr), "4-bit", n4, "minus", len(minus), "plus", len(plus))
    os.makedirs(os.path.join(PKG, "data"), exist_ok=True)
    np.savez_compressed(os.path.join(PKG, "data", "p0_bits.npz"), **out)
    rh.unload_reference()


def _sha_tensors(tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().numpy()).tobytes())
    return h.hexdigest()


def gen_models():
    functions, resnet = rh.load_reference()
    sys.path.insert(0, HERE)
    import slq_oracle as so
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 224, 224, generator=g)
    for arch in ARCHS:
        net = rh.seeded_model(resnet, arch, 0)
        sd = net.state_dict()
        init_hash = _sha_tensors([sd[k] for k in sd if k.endswith("weight") or k.endswith("bias")])
        conv_init_hash = _sha_tensors([m.weight for m in net.modules() if isinstance(m, torch.nn.Conv2d)])
        net.eval()
        with torch.no_grad():
            logits_fp32 = net(x).numpy()
        rows, _, _ = rh.p0_rows(functions, arch)
        # stream hash of (bit, z, s32, codes) in global channel order (SURVEY Appendix G format),
        # computed from the fp32 weights BEFORE they are overwritten
        layers = [net.layer1, net.layer2, net.layer3, net.layer4]
        h = hashlib.sha256()
        for li, bi, lnum, cn, bit, _f, _s, _i in rows:
            _, _, conv = rh.layer_map(arch, lnum)
            w = getattr(layers[li][bi], conv).weight.data[cn].reshape(-1).numpy()
            q = functions.quantize_wgt(torch.from_numpy(w.copy()), bit).numpy()
            z, s32, codes = _derive(w, q, bit)
            h.update(bytes([bit]))
            h.update(struct.pack("<i", z))
            h.update(struct.pack("<f", float(s32)))
            h.update(codes.astype(np.uint8).tobytes())
        rh.apply_rows(functions, net, arch, rows)
        fq_hash = _sha_tensors([getattr(b, c).weight for s in layers for b in s
                                for c in ("conv1", "conv2", "conv3") if hasattr(b, c)])
        with torch.no_grad():
            logits_p0 = net(x).numpy()
            logits_p0_oracle = so.torch_forward(net, x).numpy()
        assert np.array_equal(logits_p0, logits_p0_oracle), "oracle torch_forward != reference forward"
        np.savez_compressed(
            os.path.join(GOLD, "model_%s.npz" % arch),
            init_hash=np.array(init_hash), conv_init_hash=np.array(conv_init_hash),
            code_stream_hash=np.array(h.hexdigest()), fakequant_hash=np.array(fq_hash),
            logits_fp32=logits_fp32, logits_p0=logits_p0,
            first_row_q=net.layer1[0].conv1.weight.data[0].reshape(-1).numpy())
        print(arch, "codes", h.hexdigest()[:16], "fq", fq_hash[:16], "logits", logits_p0[0, :4],
              "argmax", logits_p0.argmax(1))
    rh.unload_reference()


def gen_sweep():
    loader = rh.synthetic_loader(2, 4, 64, seed=1)
    functions, resnet = rh.load_reference(loader)
    arch = "resnet18"
    rh.install_seeded_pretrained(resnet, arch, 0)
    net2 = resnet.resnet18(num_classes=1000, pretrained="imagenet")
    preacc, loss0, orig = functions.evaluate_acc_loss_softmax(net2, "cpu", loader)
    lnum, cnum, dl = rh.read_deltaloss_csv(arch)
    ds, rows = [], []
    for i in range(len(lnum)):
        li, bi, _ = rh.layer_map(arch, lnum[i])
        ds.append([li, bi, lnum[i], cnum[i], dl[0][i], dl[1][i], dl[2][i], dl[3][i]])
        rows.append([li, bi, lnum[i], cnum[i], 8, 0, 32, i + 1])
    minus, plus = functions.make_divide_minusplusmodels(rows, ds, 4)
    in_rows = np.array(rows, np.int64)
    in_dl8 = np.array([d[4] for d in ds], np.float64)
    minus_rows, plus_rows = np.array(minus, np.int64), np.array(plus, np.int64)
    semilayers, orders = functions.make_semilayers_resnet18(net2, "cpu", orig, minus, plus)
    orders_unsorted = [list(o) for o in orders]
    flat = functions.make_quantizedlists(semilayers, orders)
    np.savez_compressed(
        os.path.join(GOLD, "sweep_resnet18.npz"),
        preacc=np.float64(preacc), loss0=np.float64(loss0),
        orders=np.array(orders_unsorted, np.float64),
        sorted_index=np.array([o[0] for o in orders], np.int32),
        semilayer_sizes=np.array([len(s) for s in semilayers], np.int32),
        flat=np.array(flat, np.int64),
        in_rows=in_rows, in_dl8=in_dl8, minus_rows=minus_rows, plus_rows=plus_rows,
        net2_row_after=net2.layer1[0].conv1.weight.data[int(minus_rows[0][3])].reshape(-1).numpy(),
        minus_len=np.int32(len(minus) - 1), plus_len=np.int32(len(plus) - 1))
    print("sweep: orders", len(orders_unsorted), "flat", len(flat), "loss0", loss0)
    rh.unload_reference()


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    parts = sys.argv[1:] or ["quant", "p0", "models", "sweep"]
    torch.set_num_threads(8)
    if "quant" in parts:
        gen_quant()
    if "p0" in parts:
        gen_p0()
    if "models" in parts:
        gen_models()
    if "sweep" in parts:
        gen_sweep()
