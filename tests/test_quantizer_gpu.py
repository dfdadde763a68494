"""GPU parity of the quantizer kernels (through the C ABI) against the golden vectors of the
reference and against the oracle.  Bit-exact: codes, z, s32 and the fake-quantised fp32 rows."""
import hashlib
import os
import struct

import numpy as np
import pytest
import torch

import slq_oracle as so
from helpers import GOLD, p0_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qg():
    return np.load(os.path.join(GOLD, "quant_rows.npz"))


def _bits_eq(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def test_golden_rows_true_div(qg):
    """Every golden case (all K incl. unaligned/generic-path ones, bits 8/6/4/2, adversarial rows):
    CUDA tensor + div_mode TRUE must reproduce the reference's CPU result bit for bit."""
    import functions
    import slq_lib as L
    for i in range(int(qg["n"])):
        w, bit = qg["w%d" % i], int(qg["bit%d" % i])
        t = torch.from_numpy(w.copy()).reshape(1, -1).cuda()
        pr = functions.quantize_rows(t, [0], [bit], div_mode=L.DIV_TRUE)
        assert _bits_eq(t.cpu().numpy()[0], qg["q%d" % i]), "case %d K=%d bit=%d" % (i, w.size, bit)
        assert int(pr.z[0]) == int(qg["z%d" % i])
        assert _bits_eq(pr.s32.cpu().numpy()[0], qg["s%d" % i])
        gold = qg["c%d" % i].astype(np.int32)
        maxc = (1 << bit) - 1
        assert np.array_equal(pr.codes(0), np.clip(gold, 0, maxc)), "codes case %d" % i
        want = L.ROW_OK if (gold.min() >= 0 and gold.max() <= maxc) else L.ROW_CODE_RANGE
        assert int(pr.status[0]) == want


def test_host_buffer_entry_point_matches_golden(qg):
    """slq_quantize_rows_host (CPU tensors -> what the mains pass before net.to(device))."""
    import functions
    for i in range(0, int(qg["n"]), 3):
        w, bit = qg["w%d" % i], int(qg["bit%d" % i])
        t = torch.from_numpy(w.copy()).reshape(1, -1)
        out = functions.channel_wise_quantizationperchan(t, bit, 0)
        assert out is t
        assert _bits_eq(t.numpy()[0], qg["q%d" % i])
        q2 = functions.quantize_wgt(torch.from_numpy(w.copy()), bit)
        assert _bits_eq(q2.numpy(), qg["q%d" % i])


def test_recip_div_matches_aten_cuda_arithmetic(qg):
    """div_mode RECIP == the reference's own expression evaluated by ATen on CUDA tensors
    (functions.py:41 on a cuda tensor) and == the oracle's div_mode 1."""
    import functions
    rng = np.random.default_rng(7)
    mism_true = 0
    for K in (64, 576, 1152, 2304, 4608):
        w = (rng.standard_normal((64, K)) * 0.05).astype(np.float32)
        for bit in (8, 6, 4):
            t = torch.from_numpy(w.copy()).cuda()
            ref = t.clone()
            for r in range(64):  # the reference's arithmetic, by ATen, on the device
                row = ref[r]
                mn, mx = torch.min(row).item(), torch.max(row).item()
                scale = (mx - mn) / (2 ** bit - 1)
                z = round(mn / scale)
                ref[r] = (((row / scale) + z).round() - z) * scale
            functions.quantize_rows(t, np.arange(64), [bit] * 64)  # cuda tensor -> RECIP by default
            assert _bits_eq(t.cpu().numpy(), ref.cpu().numpy()), "K=%d bit=%d" % (K, bit)
            for r in (0, 17, 63):
                q1, _, _, _, _ = so.quantize_row(w[r], bit, so.DIV_RECIP)
                assert _bits_eq(t[r].cpu().numpy(), q1)
                q0, _, _, _, _ = so.quantize_row(w[r], bit, so.DIV_TRUE)
                mism_true += int((q0 != q1).sum())
    print("elements where the two divide flavours differ:", mism_true)


def test_constant_row_raises_and_leaves_row_untouched():
    import functions
    t = torch.ones(4, 64).cuda()
    t[1] = torch.linspace(-1, 1, 64)
    with pytest.raises(ZeroDivisionError):
        functions.channel_wise_quantizationperchan(t, 8, 0)
    assert bool((t[0] == 1).all())
    with pytest.raises(ZeroDivisionError):
        functions.quantize_wgt(torch.zeros(8), 4)
    with pytest.raises(IndexError):
        functions.quantize_rows(t, [9], [8])
    pr = functions.quantize_rows(t, [], [])
    assert pr.rows.size == 0


def test_progressive_chain_and_inplace_semantics(qg):
    import functions
    chain = qg["chain"]
    t = torch.from_numpy(np.stack([chain[0], chain[0]])).contiguous()
    for j, bit in enumerate((8, 6, 4), 1):  # CPU tensor: same divide flavour as the golden
        r = functions.channel_wise_quantizationperchan(t, bit, 1)
        assert r is t
        assert _bits_eq(t[1].numpy(), chain[j])
    assert _bits_eq(t[0].numpy(), chain[0])  # other rows untouched


@pytest.mark.parametrize("arch", ["resnet18", "resnet34", "resnet50"])
def test_whole_model_code_stream_hash(arch):
    """Full-size: every quantised conv of the model in one launch each; the (bit, z, s32, codes)
    stream and the fake-quant weights hash to the reference's values."""
    import functions
    import resnet
    import slq_lib as L
    g = np.load(os.path.join(GOLD, "model_%s.npz" % arch))
    torch.manual_seed(0)
    net = getattr(resnet, arch)(num_classes=1000).cuda()
    table = p0_table(arch)
    cpb = 3 if arch == "resnet50" else 2
    blocks = [b for s in (net.layer1, net.layer2, net.layer3, net.layer4) for b in s]
    h = hashlib.sha256()
    for lnum in np.unique(table[:, 0]):
        sel = table[table[:, 0] == lnum]
        conv = getattr(blocks[(lnum - 1) // cpb], "conv%d" % ((lnum - 1) % cpb + 1))
        assert np.array_equal(sel[:, 1], np.arange(conv.out_channels))
        pr = functions.quantize_rows(conv.weight.data, sel[:, 1], sel[:, 2], div_mode=L.DIV_TRUE)
        assert int(pr.status.max()) == 0
        z, s, blob = pr.z.cpu().numpy(), pr.s32.cpu().numpy(), pr.blob.cpu().numpy()
        K = pr.K
        for j in range(len(sel)):
            bit = int(sel[j, 2])
            nb = so.packed_row_bytes(K, bit)
            codes = so.unpack_codes(blob[pr.offsets[j]:pr.offsets[j] + nb], K, bit)
            h.update(bytes([bit]))
            h.update(struct.pack("<i", int(z[j])))
            h.update(struct.pack("<f", float(s[j])))
            h.update(codes.astype(np.uint8).tobytes())
    assert h.hexdigest() == str(g["code_stream_hash"])
    hf = hashlib.sha256()
    for b in blocks:
        for c in ("conv1", "conv2", "conv3"):
            if hasattr(b, c):
                hf.update(np.ascontiguousarray(getattr(b, c).weight.detach().cpu().numpy()).tobytes())
    assert hf.hexdigest() == str(g["fakequant_hash"])


def test_classify_then_encode_reproduces_quantizer_codes():
    """Content-derived path (slq_classify_rows + slq_encode_rows) on fake-quantised rows gives the
    same bit / z / codes that slq_quantize_rows emitted, and 16 bit for untouched fp32 rows."""
    import functions
    import slq_engine
    import slq_lib as L
    rng = np.random.default_rng(11)
    for K in (64, 576, 1152, 2304, 4608):
        w = torch.from_numpy((rng.standard_normal((96, K)) * 0.05).astype(np.float32)).cuda()
        bits = np.array([8, 4, 6, 2] * 16, np.int32)
        pr = functions.quantize_rows(w, np.arange(64), bits, div_mode=L.DIV_TRUE)
        (bit, z, s, bh, exact), = slq_engine.classify_weights([w])
        assert np.array_equal(bh[:64], bits)
        assert (bh[64:] == 16).all() and not exact   # untouched fp32 rows have no exact <= 8-bit grid
        assert np.array_equal(z.cpu().numpy()[:64], pr.z.cpu().numpy())
        # the refined scale IS the scale the rows were quantised with (bit for bit), so that
        # slq_decode_rows restores them exactly
        assert np.array_equal(s.cpu().numpy()[:64].view(np.uint32), pr.s32.cpu().numpy().view(np.uint32))
        (_b, _z, _s, _bh, exact64), = slq_engine.classify_weights([w[:64].contiguous()])
        assert exact64
        packed = slq_engine.encode_weight(w, bit, z, s, bh)
        blob = packed.blob.cpu().numpy()
        offs = packed.offsets.cpu().numpy()
        wh = w.cpu().numpy()
        for j in range(96):
            nb = so.packed_row_bytes(K, int(bh[j]))
            codes = so.unpack_codes(blob[offs[j]:offs[j] + nb], K, int(bh[j]))
            if j < 64:
                assert np.array_equal(codes, pr.codes(j)), "row %d K %d" % (j, K)
            ob, oc, oz, os_ = so.encode_row(wh[j])
            assert ob == bh[j] and oz == int(z[j]) and np.array_equal(oc, codes)
            assert np.float32(os_) == s.cpu().numpy()[j]
        # decode: packed rows -> fp32, bit-identical to the fake-quantised rows
        back = torch.empty_like(w)
        L.check(L.lib().slq_decode_rows(packed.blob.data_ptr(), packed.offsets.data_ptr(), packed.bits.data_ptr(),
                                        packed.z.data_ptr(), packed.s.data_ptr(), 96, K, back.data_ptr(),
                                        L.current_stream()))
        assert np.array_equal(back.cpu().numpy()[:64].view(np.uint32), wh[:64].view(np.uint32))
        assert np.allclose(back.cpu().numpy()[64:], wh[64:], atol=float(np.abs(wh[64:]).max()) / 60000)
