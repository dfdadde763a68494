"""Pins the CPU oracle (oracle/) to golden vectors generated from the UNMODIFIED reference
(oracle/gen_golden.py).  CPU only."""
import hashlib
import os
import struct

import numpy as np
import pytest
import torch

import slq_oracle as so
from helpers import GOLD, p0_table


@pytest.fixture(scope="module")
def qg():
    return np.load(os.path.join(GOLD, "quant_rows.npz"))


def test_quantizer_matches_reference_quantize_wgt(qg):
    n = int(qg["n"])
    assert n == 100
    for i in range(n):
        w, bit = qg["w%d" % i], int(qg["bit%d" % i])
        q, codes, z, s32, st = so.quantize_row(w, bit, so.DIV_TRUE)
        assert np.array_equal(q.view(np.uint32), qg["q%d" % i].view(np.uint32)), "case %d" % i
        assert z == int(qg["z%d" % i]) and s32 == qg["s%d" % i]
        gold = qg["c%d" % i].astype(np.int32)
        maxc = (1 << bit) - 1
        in_range = gold.min() >= 0 and gold.max() <= maxc
        # a tie at both ends can need 2^bit + 1 levels: the stored code is
        # clamped and the row is flagged; the fake-quantised values above are exact regardless
        assert st == (so.ST_OK if in_range else so.ST_CODE_RANGE), "case %d" % i
        assert np.array_equal(codes, np.clip(gold, 0, maxc))
        if in_range:  # reconstruction identity: real weight = (code + z) * s32
            assert np.array_equal(((codes + z).astype(np.float32) * s32).astype(np.float32), q)


def test_known_answer_row_survey_appendix_g(qg):
    # R50 layer1[0].conv1.weight[0], bit 8
    for i in range(int(qg["n"])):
        w = qg["w%d" % i]
        if w.size == 64 and abs(float(w.min()) + 0.3841761350631714) < 1e-12 and int(qg["bit%d" % i]) == 8:
            q, codes, z, s32, _ = so.quantize_row(w, 8)
            assert z == -123
            assert np.float32(s32).view(np.uint32) == 0x3B4CC52A
            assert codes[:8].tolist() == [129, 52, 96, 82, 175, 186, 162, 86]
            return
    pytest.fail("KAT row not found in golden file")


def test_progressive_requantisation_chain(qg):
    chain = qg["chain"]  # rows after fp32 -> 8 -> 6 -> 4 through channel_wise_quantizationperchan
    t = chain[0:1].copy()
    for j, bit in enumerate((8, 6, 4), 1):
        so.channel_wise(t, bit, 0)
        assert np.array_equal(t[0].view(np.uint32), chain[j].view(np.uint32))


def test_constant_row_raises_like_reference(qg):
    assert int(qg["const_raises"]) == 1
    with pytest.raises(ZeroDivisionError):
        so.quantize_row(np.ones(16, np.float32), 8)


def test_div_modes_differ_only_rarely():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(1 << 16) * 0.05).astype(np.float32)
    q0, c0, z0, s0, _ = so.quantize_row(w, 8, so.DIV_TRUE)
    q1, c1, z1, s1, _ = so.quantize_row(w, 8, so.DIV_RECIP)
    assert z0 == z1 and s0 == s1
    assert (c0 != c1).sum() < 16 and np.abs(c0 - c1).max() <= 1


@pytest.mark.parametrize("bit", [2, 4, 6, 8, 16])
def test_pack_roundtrip(bit):
    rng = np.random.default_rng(bit)
    for K in (64, 147, 7, 4608):
        codes = rng.integers(0, 1 << bit, K)
        p = so.pack_codes(codes, bit)
        assert p.size == so.packed_row_bytes(K, bit)
        assert np.array_equal(so.unpack_codes(p, K, bit), codes)


def test_encoder_recovers_bits_and_codes(qg):
    for i in range(int(qg["n"])):
        w, bit = qg["w%d" % i], int(qg["bit%d" % i])
        if w.size < 64:
            continue
        q = qg["q%d" % i]
        b2, codes, z, s = so.encode_row(q)
        recon = ((codes + z).astype(np.float32) * s).astype(np.float32)
        if b2 == 16:  # off-grid (257-level tie rows, huge |z|): stored on the 16-bit grid instead
            assert np.abs(recon.astype(np.float64) - q).max() <= float(s) * 0.51 + 1e-7 * np.abs(q).max()
            continue
        assert b2 <= bit  # an 8-bit row that only uses 4-bit levels may legally be stored narrower
        assert np.allclose(recon, q, rtol=3e-7, atol=0)
        gold = qg["c%d" % i].astype(np.int32)
        if b2 == bit and abs(z) < 1000 and gold.max() <= (1 << bit) - 1:
            assert np.array_equal(codes, gold)
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal(576) * 0.1).astype(np.float32)
    b, codes, z, s = so.encode_row(raw)
    assert b == 16
    assert np.abs(((codes + z).astype(np.float32) * s) - raw).max() <= float(s) * 0.51


@pytest.mark.parametrize("arch", ["resnet18", "resnet34", "resnet50"])
def test_model_code_stream_hash(arch):
    """Whole-model P0 quantisation through the oracle reproduces the reference's
    (bit, z, s32, codes) stream and fake-quant weights (SURVEY.md Appendix G hashes)."""
    import resnet
    g = np.load(os.path.join(GOLD, "model_%s.npz" % arch))
    torch.manual_seed(0)
    net = getattr(resnet, arch)(num_classes=1000)
    sd = net.state_dict()
    h0 = hashlib.sha256()
    for k in sd:
        if k.endswith("weight") or k.endswith("bias"):
            h0.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    assert h0.hexdigest() == str(g["init_hash"]), "seeded init differs from the reference's"
    cpb = 3 if arch == "resnet50" else 2
    blocks = [b for s in (net.layer1, net.layer2, net.layer3, net.layer4) for b in s]
    h = hashlib.sha256()
    for lnum, cn, bit in p0_table(arch):
        conv = getattr(blocks[(lnum - 1) // cpb], "conv%d" % ((lnum - 1) % cpb + 1))
        w2 = conv.weight.data.reshape(conv.out_channels, -1).numpy()
        q, codes, z, s32, _ = so.quantize_row(w2[cn], int(bit))
        h.update(bytes([int(bit)]))
        h.update(struct.pack("<i", z))
        h.update(struct.pack("<f", float(s32)))
        h.update(codes.astype(np.uint8).tobytes())
        w2[cn] = q
    assert h.hexdigest() == str(g["code_stream_hash"])
    hf = hashlib.sha256()
    for b in blocks:
        for c in ("conv1", "conv2", "conv3"):
            if hasattr(b, c):
                hf.update(np.ascontiguousarray(getattr(b, c).weight.detach().numpy()).tobytes())
    assert hf.hexdigest() == str(g["fakequant_hash"])
    assert np.array_equal(net.layer1[0].conv1.weight.data[0].reshape(-1).numpy(), g["first_row_q"])
    if arch == "resnet18":  # fp32 forward restatement vs the reference's own forward (2 images)
        gen = torch.Generator().manual_seed(1)
        x = torch.randn(2, 3, 224, 224, generator=gen)
        net.eval()
        logits = so.torch_forward(net, x).numpy()
        assert np.allclose(logits, g["logits_p0"], rtol=1e-4, atol=1e-4)
        assert (logits.argmax(1) == g["logits_p0"].argmax(1)).all()


def test_integer_conv_oracle_matches_float_conv():
    """conv_acc / epilogue (the pipeline restatement) against an independent float64 conv."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (2, 9, 9, 8), dtype=np.uint8)
    codes = rng.integers(0, 256, (16, 3, 3, 8))
    for stride in (1, 2):
        acc, S, Ho, Wo = so.conv_acc(x, codes, stride, 1)
        xt = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
        wt = torch.from_numpy(codes.astype(np.float64)).permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(xt, wt, None, stride, 1).permute(0, 2, 3, 1).reshape(-1, 16).numpy()
        assert np.array_equal(acc, ref.astype(np.int64))
        ones = torch.ones(1, 8, 3, 3, dtype=torch.float64)
        Sref = torch.nn.functional.conv2d(xt, ones, None, stride, 1).reshape(-1).numpy()
        assert np.array_equal(S, Sref.astype(np.int64))
