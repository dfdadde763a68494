"""GPU parity of the quantised convolution (tcgen05 implicit GEMM + fused epilogue, through the C
ABI) against the numpy restatement: integer accumulators and u8/s8/fp32 outputs are bit-exact."""
import numpy as np
import pytest
import torch

from helpers import (ConvCase, DOWNSAMPLE_SHAPES, R18_SHAPES, R50_SHAPES, expected_gemm_weights,
                     oracle_epilogue_case)

pytestmark = pytest.mark.gpu


def _mixed_bits(cout, seed, frac16=0.0):
    rng = np.random.default_rng(seed)
    bits = rng.choice([4, 8], cout).astype(np.int32)  # scattered 4/8-bit semilayers
    if frac16 > 0:
        bits[rng.random(cout) < frac16] = 32
    return bits


def _check_acc(case):
    lo, hi, S = case.oracle_acc()
    out, Sd = case.run_acc()
    assert np.array_equal(Sd.astype(np.int64), S), "window sums differ"
    assert np.array_equal(out[:, :case.cout].astype(np.int64), lo), "accumulators differ"
    if hi is not None:
        assert np.array_equal(out[:, case.cout:].astype(np.int64), hi), "high-limb accumulators differ"


def _check_epilogues(case, seed):
    import slq_lib as L
    rng = np.random.default_rng(seed)
    cout = case.cout
    zf = np.array([m[2] for m in case.meta], np.float32)
    s = np.array([m[3] for m in case.meta], np.float32)
    bn_a = (0.5 + rng.random(cout)).astype(np.float32)
    wscale = (s * bn_a).astype(np.float32)
    bias = (rng.standard_normal(cout) * 0.05).astype(np.float32)
    scales = np.array([0.02, 0.0, 0.013, 0.05], np.float32)
    # pick the output scale from the data so that the u8 range is actually used
    y = oracle_epilogue_case(case, wscale, zf, bias, scales, 0, 1, None, -1, 0, 1, L.OUT_F32)
    scales[1] = max(float(np.abs(y).max()), 1e-6) / 255.0
    res_u8 = rng.integers(0, 256, (case.M, cout), dtype=np.uint8)
    for (mode, res, rid, rs, relu) in [(L.OUT_F32, None, -1, 0, 1), (L.OUT_U8, None, -1, 0, 1),
                                       (L.OUT_U8, res_u8, 2, 0, 1), (L.OUT_U8, res_u8, 3, 1, 1),
                                       (L.OUT_S8, None, -1, 0, 0)]:
        want = oracle_epilogue_case(case, wscale, zf, bias, scales, 0, 1, res, rid, rs, relu, mode)
        got = case.run_epi(mode, wscale, zf, bias, scales, 0, 1, res, rid, rs, relu)
        if mode == L.OUT_F32:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "fp32 epilogue differs"
        else:
            assert np.array_equal(got, want), "mode %d res %s: %d bytes differ" % (mode, res is not None, (got != want).sum())


def test_gemm_ready_weights_layout():
    for (cin, cout, k, w16f) in [(64, 64, 3, 0.0), (128, 256, 1, 0.0), (64, 128, 1, 0.3), (256, 64, 3, 0.2)]:
        case = ConvCase(1, 8, cin, cout, k, 1, _mixed_bits(cout, 1, w16f), seed=3, impl=1)
        want = expected_gemm_weights(case.meta, cout, cin, k, case.w16)
        assert np.array_equal(case.wg.cpu().numpy(), want)
        case.close()


@pytest.mark.parametrize("shape", [(64, 64, 3, 1, 10), (128, 64, 1, 1, 9), (64, 128, 3, 2, 12), (128, 128, 1, 2, 8)])
@pytest.mark.parametrize("frac16", [0.0, 0.25])
def test_simt_checker_is_exact(shape, frac16):
    """The on-device dp4a checker itself is pinned to the numpy oracle."""
    cin, cout, k, stride, H = shape
    case = ConvCase(2, H, cin, cout, k, stride, _mixed_bits(cout, 2, frac16), seed=5, impl=1)
    _check_acc(case)
    _check_epilogues(case, 9)
    case.close()


@pytest.mark.parametrize("shape", R50_SHAPES, ids=lambda s: "c%d-%d_k%d_s%d_h%d" % s)
def test_umma_resnet50_shapes_exact(shape):
    cin, cout, k, stride, H = shape
    case = ConvCase(2, H, cin, cout, k, stride, _mixed_bits(cout, cin + cout), seed=cin)
    _check_acc(case)
    _check_epilogues(case, 13)
    case.close()


@pytest.mark.parametrize("shape", R18_SHAPES, ids=lambda s: "c%d-%d_k%d_s%d_h%d" % s)
def test_umma_resnet18_34_shapes_exact(shape):
    cin, cout, k, stride, H = shape
    case = ConvCase(2, H, cin, cout, k, stride, _mixed_bits(cout, cin * 3 + cout), seed=cin + 1)
    _check_acc(case)
    case.close()


@pytest.mark.parametrize("shape", DOWNSAMPLE_SHAPES, ids=lambda s: "c%d-%d_k%d_s%d_h%d" % s)
def test_umma_fp32_rows_two_limb_mode(shape):
    """Un-quantised (fp32) rows -> 16-bit codes in two u8 limbs (downsample convs)."""
    cin, cout, k, stride, H = shape
    case = ConvCase(2, H, cin, cout, k, stride, np.full(cout, 32, np.int32), seed=cout)
    assert case.w16 == 1
    _check_acc(case)
    _check_epilogues(case, 17)
    case.close()


def test_umma_mixed_fp32_and_quantised_rows():
    case = ConvCase(2, 14, 256, 256, 3, 1, _mixed_bits(256, 4, 0.5), seed=8)
    assert case.w16 == 1
    _check_acc(case)
    _check_epilogues(case, 19)
    case.close()


@pytest.mark.parametrize("N,H", [(1, 7), (3, 7), (1, 14), (5, 9), (1, 56)])
def test_umma_ragged_m_tails(N, H):
    """M = N*Ho*Wo not a multiple of the 128-pixel tile, single-tile and sub-tile problems."""
    for (cin, cout, k, stride) in [(128, 128, 3, 1), (256, 64, 1, 1), (128, 256, 3, 2), (64, 64, 1, 1)]:
        case = ConvCase(N, H, cin, cout, k, stride, _mixed_bits(cout, N + H), seed=N * 10 + H)
        _check_acc(case)
        case.close()


def test_umma_im2col_and_tiled_tma_agree():
    """1x1 stride-1: the A operand through the tiled map and through the im2col map."""
    import slq_lib as L
    for a_mode in (L.A_TILED, L.A_IM2COL):
        case = ConvCase(2, 28, 512, 128, 1, 1, _mixed_bits(128, 6), seed=21, a_mode=a_mode)
        _check_acc(case)
        case.close()


def test_umma_extreme_activations_no_overflow():
    """All-255 activations x all-255 codes at the largest K (4608): |acc| = 3.0e8 < 2^31."""
    case = ConvCase(1, 7, 512, 64, 3, 1, np.full(64, 8, np.int32), seed=1)
    case.xd.fill_(255)
    case.x[...] = 255
    case.rowsum.fill_(255 * 512)
    _check_acc(case)
    case.close()


def test_umma_equals_simt_at_full_batch():
    """Full BASELINE batch (256) on two layers: tcgen05 output bytes == dp4a checker bytes
    (size-independent cross-check where the numpy oracle is too slow)."""
    import ctypes
    import slq_lib as L
    for (cin, cout, k, stride, H) in [(256, 64, 1, 1, 56), (256, 256, 3, 1, 14)]:
        a = ConvCase(256, H, cin, cout, k, stride, _mixed_bits(cout, 3), seed=2)
        b = ConvCase(256, H, cin, cout, k, stride, _mixed_bits(cout, 3), seed=2, impl=L.IMPL_SIMT)
        oa, Sa = a.run_acc()
        ob, Sb = b.run_acc()
        assert np.array_equal(oa, ob) and np.array_equal(Sa, Sb)
        # checksum of checksums: column sums of the accumulators == (sum of A rows) . B
        A_sum = a.x.reshape(-1, cin).astype(np.int64).sum(0) if k == 1 else None
        if A_sum is not None:
            codes = np.stack([m[1] for m in a.meta]).astype(np.int64)
            assert np.array_equal(oa.astype(np.int64).sum(0), codes @ A_sum)
        a.close()
        b.close()


@pytest.mark.parametrize("cfg", [
    # (cin, cout, k, stride, H, fp32 rows, mode, residual: 0 none / 1 u8 / 2 s8)
    (64, 64, 1, 1, 56, False, "u8", 0),      # one K block per tile, many tiles per CTA
    (64, 256, 1, 1, 56, False, "u8", 2),     # one K block, two n-tiles, signed residual
    (64, 256, 1, 1, 56, True, "s8", 0),      # downsample: two-limb rows, signed output
    (256, 64, 1, 1, 56, False, "u8", 0),     # two K blocks per tile
    (64, 64, 3, 1, 56, False, "u8", 0),      # nine 64-byte K blocks
    (256, 512, 1, 2, 56, True, "s8", 0),     # strided two-limb downsample
    (128, 512, 1, 1, 28, False, "u8", 1),    # u8 residual, four n-tiles
    (512, 512, 3, 1, 14, False, "u8", 0),    # streamed weights, 36 K blocks
], ids=lambda c: "c%d-%d_k%d_s%d_h%d_%s_r%d" % (c[0], c[1], c[2], c[3], c[4], c[6], c[7]))
def test_umma_equals_simt_epilogue_modes_multi_tile(cfg):
    """Many tiles per CTA through every epilogue flavour the engine uses (u8 / s8 output, no / u8 / s8
    residual, two-limb rows): tcgen05 output bytes == dp4a checker bytes."""
    import slq_lib as L
    cin, cout, k, stride, H, fp32_rows, mode, res_kind = cfg
    N = 96  # >= 12 tiles per CTA: every barrier ring wraps more than once
    bits = np.full(cout, 32, np.int32) if fp32_rows else _mixed_bits(cout, cin + k)
    a = ConvCase(N, H, cin, cout, k, stride, bits, seed=4)
    b = ConvCase(N, H, cin, cout, k, stride, bits, seed=4, impl=L.IMPL_SIMT)
    rng = np.random.default_rng(9)
    K = cin * k * k
    wscale = (rng.uniform(0.5, 1.5, cout) / (K * 40.0 * (256.0 if fp32_rows else 1.0))).astype(np.float32)
    zf = -rng.integers(100, 156, cout).astype(np.float32) * (256.0 if fp32_rows else 1.0)
    bias = rng.normal(0, 0.5, cout).astype(np.float32)
    scales = np.array([1.0, 1.0 / 64, 1.0 / 32, 1.0], np.float32)
    res = None
    if res_kind:
        res = rng.integers(0, 256, (a.M, cout), dtype=np.uint8)
    out_mode = L.OUT_U8 if mode == "u8" else L.OUT_S8
    args = dict(res=res, res_id=2 if res_kind else -1, res_signed=1 if res_kind == 2 else 0, relu=1)
    oa = a.run_epi(out_mode, wscale, zf, bias, scales, 0, 1, **args)
    ob = b.run_epi(out_mode, wscale, zf, bias, scales, 0, 1, **args)
    assert len(np.unique(oa)) > 8  # the parameters exercise the quantiser's range
    assert np.array_equal(oa, ob)
    a.close()
    b.close()


def _expected_packed_gemm(case, bn_cols, n_tiles, k_block, num_kb):
    """numpy restatement of the packed GEMM-ready layout (include/slq.h, slq_conv_set_packed_weights)."""
    cout, cin, k = case.cout, case.cin, case.k
    codes = np.zeros((bn_cols * n_tiles, k * k * cin), np.int64)
    bits = np.full(bn_cols * n_tiles, 4, np.int64)
    for oc in range(cout):
        codes[oc] = case.meta[oc][1].reshape(cin, k, k).transpose(1, 2, 0).reshape(-1)  # (c,r,s) -> (r,s,c)
        bits[oc] = case.meta[oc][0]
    out = []
    for t in range(n_tiles):
        for kb in range(num_kb):
            for r in range(t * bn_cols, (t + 1) * bn_cols):
                seg = codes[r, kb * k_block:(kb + 1) * k_block]
                if bits[r] <= 4:
                    out.append((seg[0::2] | (seg[1::2] << 4)).astype(np.uint8))
                else:
                    out.append(seg.astype(np.uint8))
    return np.concatenate(out)


@pytest.mark.parametrize("shape", [(64, 256, 1, 1, 56), (256, 64, 1, 1, 56), (64, 64, 3, 1, 56), (512, 128, 1, 1, 28),
                                   (128, 512, 1, 1, 28), (64, 64, 1, 1, 56), (512, 2048, 1, 1, 7)],
                         ids=lambda s: "c%d-%d_k%d_s%d_h%d" % s)
def test_umma_packed_weights_unpacked_in_shared_memory(shape):
    """BASELINE north_star: 4-bit semilayer codes travel PACKED (two per byte) and are unpacked to int8 in shared
    memory.  Resident-weight layers with scattered 4/8-bit rows: the packed operand is what the layout says, it
    is smaller than the u8 matrix, and accumulators / epilogues stay bit-exact."""
    import slq_engine
    cin, cout, k, stride, H = shape
    bits = _mixed_bits(cout, cin + 3 * cout)
    case = ConvCase(2, H, cin, cout, k, stride, bits, seed=cout + 7, packed_b=True)
    assert case.packed_gemm is not None
    bn_cols, n_tiles, k_block, num_kb, resident = slq_engine.conv_tiling(case.desc)
    assert resident == 1
    want = _expected_packed_gemm(case, bn_cols, n_tiles, k_block, num_kb)
    got = case.packed_gemm.blob.cpu().numpy()[:want.size]
    assert np.array_equal(got, want)
    n4 = int((bits == 4).sum())
    assert want.size == (k * k * cin) * (cout - n4) + (k * k * cin // 2) * n4  # Cout is a multiple of the tile here
    assert want.size < case.wg.numel()
    _check_acc(case)
    _check_epilogues(case, 23)
    case.close()


def test_streamed_layers_keep_the_u8_operand():
    import slq_engine
    case = ConvCase(2, 14, 256, 256, 3, 1, _mixed_bits(256, 9), seed=3, packed_b=True)
    assert slq_engine.conv_tiling(case.desc)[4] == 0 and case.packed_gemm is None
    _check_acc(case)
    case.close()
