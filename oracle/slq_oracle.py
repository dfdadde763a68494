"""oracle/slq_oracle.py -- TEST INFRASTRUCTURE ONLY (the checker; never the thing shipped or timed,
except as bench.py's cpu_baseline / --impl reference leg).

CPU restatements used by tests/, __graft_entry__.smoke() and bench.py's CPU leg:

* quantizer  : ctypes wrapper over oracle/quant_oracle.c, a C restatement of the reference's
               functions.py:25-43 (quantize_wgt) / :9-23 (channel_wise_quantizationperchan).
* fp32 forward: ``torch_forward`` -- the reference's hot path IS stock torch.nn modules
               (resnet.py:204-220 -> aten::convolution / native_batch_norm / relu_ / add_), a
               third-party dependency that is not vendored under /root/reference; the installed
               torch (2.11.0) is that arithmetic, restated here call-for-call with
               torch.nn.functional on the module tree's tensors.
* integer pipeline: numpy restatement of THIS repo's u8 x u8 -> s32 implicit-GEMM convolution and
               fused epilogue so that the CUDA kernels' integer accumulators
               and u8 activations can be checked bit-for-bit at small sizes.

Parity pin: tests/test_oracle_pins.py checks the quantizer and the fp32 forward against
tests/golden/*.npz, produced by oracle/gen_golden.py from the unmodified reference run in the build
container.  The reference itself ships no tests or golden vectors.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libslq_oracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.slq_oracle_quantize_row.restype = ctypes.c_int
        lib.slq_oracle_quantize_row.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


DIV_TRUE = 0    # ATen CPU: true fp32 divide
DIV_RECIP = 1   # ATen CUDA: multiply by fp32 reciprocal of the (CPU-scalar) divisor

ST_OK, ST_ZERO_RANGE, ST_CODE_RANGE = 0, 1, 2


def quantize_row(w, bit, div_mode=DIV_TRUE):
    """functions.py:25-43.  Returns (q fp32[K], codes int32[K], z int, s32 np.float32, status)."""
    w = np.ascontiguousarray(w, dtype=np.float32).reshape(-1)
    K = w.size
    q = np.empty(K, np.float32)
    codes = np.empty(K, np.int32)
    z = ctypes.c_int32(0)
    s32 = ctypes.c_float(0)
    s64 = ctypes.c_double(0)
    st = _lib().slq_oracle_quantize_row(
        w.ctypes.data, K, int(bit), int(div_mode), q.ctypes.data, codes.ctypes.data,
        ctypes.byref(z), ctypes.byref(s32), ctypes.byref(s64))
    if st == ST_ZERO_RANGE:
        raise ZeroDivisionError("float division by zero")  # functions.py:40 behaviour
    return q, codes, int(z.value), np.float32(s32.value), st


def channel_wise(tensor2d, bit, i, div_mode=DIV_TRUE):
    """functions.py:9-23 on a numpy [rows, K] fp32 array, in place."""
    q, _, _, _, _ = quantize_row(tensor2d[i], bit, div_mode)
    tensor2d[i] = q
    return tensor2d


# ----------------------------------------------------------------------------------------------
# storage format of the codes: little-endian bit packing inside a byte
# ----------------------------------------------------------------------------------------------
def packed_row_bytes(K, bit):
    if bit == 16:
        return 2 * K
    if bit in (8, 6):
        return K
    if bit == 4:
        return (K + 1) // 2
    if bit == 2:
        return (K + 3) // 4
    raise ValueError(bit)


def pack_codes(codes, bit):
    c = np.asarray(codes).astype(np.int64).reshape(-1)
    K = c.size
    if bit in (8, 6):
        return c.astype(np.uint8)
    if bit == 16:  # low-limb plane then high-limb plane
        return np.concatenate([(c & 255).astype(np.uint8), (c >> 8).astype(np.uint8)])
    per = 8 // bit
    pad = (-K) % per
    c = np.concatenate([c, np.zeros(pad, np.int64)]).reshape(-1, per)
    out = np.zeros(c.shape[0], np.int64)
    for j in range(per):
        out |= c[:, j] << (bit * j)
    return out.astype(np.uint8)


def unpack_codes(packed, K, bit):
    p = np.asarray(packed, dtype=np.uint8).astype(np.int64)
    if bit in (8, 6):
        return p[:K].astype(np.int32)
    if bit == 16:
        return (p[:K] | (p[K:2 * K] << 8)).astype(np.int32)
    per = 8 // bit
    cols = [(p >> (bit * j)) & ((1 << bit) - 1) for j in range(per)]
    return np.stack(cols, 1).reshape(-1)[:K].astype(np.int32)


# ----------------------------------------------------------------------------------------------
# content-derived encoder: smallest b in {2,4,6,8} for which the fp32 row
# already lies on the b-bit affine grid spanned by its own min/max; otherwise a 16-bit grid.
# ----------------------------------------------------------------------------------------------
ENCODE_TOL = np.float32(0.02)


def encode_row(w):
    """Returns (bit, codes int32[K], z int, s np.float32).  real weight ~= (codes + z) * s."""
    w = np.ascontiguousarray(w, dtype=np.float32).reshape(-1)
    mn, mx = np.float32(w.min()), np.float32(w.max())
    if mx == mn:
        if mn == 0:
            return 16, np.zeros(w.size, np.int32), 0, np.float32(1.0)
        return 16, np.zeros(w.size, np.int32), 1, mn
    for bit in (2, 4, 6, 8, 16):
        levels = (1 << bit) - 1
        scale = (float(mx) - float(mn)) / levels
        s32 = np.float32(scale)
        if s32 == 0 or not np.isfinite(np.float32(1.0) / s32):
            continue
        zd = float(np.rint(float(mn) / scale))
        if abs(zd) > (8.0e6 if bit == 16 else 1.0e6):
            continue
        zf = np.float32(zd)
        t = (w / s32).astype(np.float32)
        r = np.rint(t).astype(np.float32)
        if bit != 16:
            if np.max(np.abs(t - r)) > ENCODE_TOL:
                continue
            # refinement (csrc/quantizer.cu classify_rows_kernel): the neighbouring scale, within 2 ulp, that
            # reproduces every element exactly, fp32(rint(w / s) * s) == w
            for delta in (0, -1, 1, -2, 2):
                sc = (s32.view(np.int32) + np.int32(delta)).view(np.float32)
                kmin = np.float32(np.rint(np.float32(mn / sc)))
                k = np.rint((w / sc).astype(np.float32)).astype(np.float32)
                c = (k - kmin).astype(np.float32)
                if np.array_equal((k * sc).astype(np.float32), w) and c.min() >= 0 and c.max() <= levels \
                        and abs(float(kmin)) <= 1.0e6:
                    s32, zd, zf, r = sc, float(kmin), kmin, k
                    break
        u = (r - zf).astype(np.float32)
        u = np.minimum(np.maximum(u, np.float32(0)), np.float32(levels))
        return bit, u.astype(np.int32), int(zd), s32
    # degenerate range (denormal scale): represent exactly nothing better than a constant row
    return 16, np.zeros(w.size, np.int32), 1, mn if mn != 0 else np.float32(1.0)


# ----------------------------------------------------------------------------------------------
# integer implicit-GEMM convolution + fused epilogue of this repo's pipeline (numpy, exact)
# ----------------------------------------------------------------------------------------------
def im2col_nhwc(x, kh, kw, stride, pad):
    """x [N,H,W,C] -> A [N*Ho*Wo, kh*kw*C] with K ordered (r, s, c); zero padding."""
    N, H, W, C = x.shape
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    xp = np.zeros((N, H + 2 * pad, W + 2 * pad, C), x.dtype)
    xp[:, pad:pad + H, pad:pad + W, :] = x
    cols = []
    for r in range(kh):
        for s in range(kw):
            cols.append(xp[:, r:r + stride * Ho:stride, s:s + stride * Wo:stride, :])
    A = np.concatenate(cols, axis=3).reshape(N * Ho * Wo, kh * kw * C)
    return A, Ho, Wo


def conv_acc(x_u8, codes_ohwi, stride, pad):
    """Exact integer accumulators.  x_u8 [N,H,W,C] uint8; codes_ohwi [Cout,kh,kw,C] ints >= 0.
    Returns (acc int64 [M,Cout], S int64 [M], Ho, Wo): acc = sum x*u, S = window sum of x."""
    Cout, kh, kw, C = codes_ohwi.shape
    A, Ho, Wo = im2col_nhwc(x_u8.astype(np.int64), kh, kw, stride, pad)
    B = codes_ohwi.reshape(Cout, kh * kw * C).astype(np.int64)
    return A @ B.T, A.sum(1), Ho, Wo


def fma32(a, b, c):
    """Exactly-rounded float32 fused multiply-add on numpy arrays (what __fmaf_rn computes).
    a*b is exact in float64; the float64 sum is then corrected for double rounding with the exact
    TwoSum error term, so the result equals round_to_float32(a*b + c) computed exactly."""
    a64, b64, c64 = (np.asarray(v, np.float32).astype(np.float64) for v in (a, b, c))
    p = a64 * b64
    s = p + c64
    bb = s - p
    err = (p - (s - bb)) + (c64 - bb)          # p + c == s + err exactly
    r = s.astype(np.float32)
    rd = r.astype(np.float64)
    up = np.nextafter(r, np.float32(np.inf))
    dn = np.nextafter(r, np.float32(-np.inf))
    other = np.where(s > rd, up, dn)
    od = other.astype(np.float64)
    tie = (s == (rd + od) / 2) & (s != rd)
    toward = np.sign(err) == np.sign(od - rd)
    return np.where(tie & (err != 0) & toward, other, r).astype(np.float32)


def _epilogue_core(acc, S, zf, wscale, bias, s_in, inv_out, res_u8, s_res, acc_hi, res_signed):
    f = np.float32
    wsc = (wscale.astype(f) * f(s_in)).astype(f)
    zw = (zf.astype(f) * wsc).astype(f)
    b = bias.astype(f)
    if inv_out is not None:  # quantised output: the re-quantisation multiply is folded per channel
        inv = f(inv_out)
        wsc, zw, b = (wsc * inv).astype(f), (zw * inv).astype(f), (b * inv).astype(f)
    accf = acc.astype(f)
    if acc_hi is not None:
        accf = fma32(acc_hi.astype(f), f(256.0), accf)
    c2 = fma32(S.astype(f)[:, None], zw[None, :], b[None, :])
    y = fma32(accf, wsc[None, :], c2)
    if res_u8 is not None:
        r = res_u8.view(np.int8).astype(f) if res_signed else res_u8.astype(f)
        sr = f(s_res) if inv_out is None else (f(s_res) * f(inv_out)).astype(f)
        y = fma32(r, sr, y)
    return y


def epilogue(acc, S, zf, wscale, bias, s_in, res_u8=None, s_res=None, relu=True, acc_hi=None,
             res_signed=False):
    """fp32-output epilogue exactly as csrc/epilogue.cuh computes it (fma = one rounding):
         wsc = wscale*s_in ; zw = zf*wsc
         accf = f32(acc) [two limbs: fma(f32(hi), 256, f32(lo))]
         y = fma(accf, wsc, fma(f32(S), zw, bias)) ; y = fma(f32(res), s_res, y)
       Returns the fp32 output (ReLU applied when relu)."""
    y = _epilogue_core(acc, S, zf, wscale, bias, s_in, None, res_u8, s_res, acc_hi, res_signed)
    return np.maximum(y, np.float32(0)) if relu else y


def epilogue_q(acc, S, zf, wscale, bias, s_in, s_out, res_u8=None, s_res=None, acc_hi=None,
               res_signed=False, signed_out=False):
    """u8 / s8-output epilogue of csrc/epilogue.cuh: same chain with inv = 1/s_out folded into the
    per-channel constants (A = wsc*inv, Z = zw*inv, B = bias*inv, sr = s_res*inv), then
    sat(rint(y)).  Saturation at 0 is the ReLU of u8 tensors.  Returns raw bytes (uint8)."""
    inv = np.float32(1.0) / np.float32(s_out)
    y = _epilogue_core(acc, S, zf, wscale, bias, s_in, inv, res_u8, s_res, acc_hi, res_signed)
    q = np.nan_to_num(np.rint(y), nan=0.0)
    if signed_out:
        return np.clip(q, -128, 127).astype(np.int8).view(np.uint8)
    return np.clip(q, 0, 255).astype(np.uint8)


def _block_tail_core(acc3, S3, zf3, wscale3, bias3, s_y2, lo, hi, Sd, zfd, wscaled, biasd, s_x, inv_out):
    f = np.float32
    a3 = (wscale3.astype(f) * f(s_y2)).astype(f)
    ad = (wscaled.astype(f) * f(s_x)).astype(f)
    b = (bias3.astype(f) + biasd.astype(f)).astype(f)
    if inv_out is not None:
        inv = f(inv_out)
        a3, ad, b = (a3 * inv).astype(f), (ad * inv).astype(f), (b * inv).astype(f)
    z3 = (zf3.astype(f) * a3).astype(f)
    zd = (zfd.astype(f) * ad).astype(f)
    accd = fma32(hi.astype(f), f(256.0), lo.astype(f))
    y = fma32(S3.astype(f)[:, None], z3[None, :], b[None, :])
    y = fma32(Sd.astype(f)[:, None], zd[None, :], y)
    y = fma32(acc3.astype(f), a3[None, :], y)
    return fma32(accd, ad[None, :], y)


def block_tail(acc3, S3, zf3, wscale3, bias3, s_y2, lo, hi, Sd, zfd, wscaled, biasd, s_x):
    """fp32 output of the fused block tail (csrc/block_tail.cu; include/slq.h section 2b), i.e. reference
    resnet.py:107-114  relu(bn3(conv3(y2)) + bn_d(conv_d(x)))  on integer accumulators:
         A3 = wscale3*s_y2 ; Ad = wscaled*s_x ; B = bias3 + biasd ; Z3 = zf3*A3 ; Zd = zfd*Ad
         y = fma(fma(hi, 256, lo), Ad, fma(acc3, A3, fma(Sd, Zd, fma(S3, Z3, B)))) ; max(y, 0)
    acc3 [M, Cout], S3 [M]: conv3 accumulators / window sums; lo, hi [M, Cout], Sd [M]: the two limbs of the
    downsample conv's 16-bit codes and its window sums."""
    y = _block_tail_core(acc3, S3, zf3, wscale3, bias3, s_y2, lo, hi, Sd, zfd, wscaled, biasd, s_x, None)
    return np.maximum(y, np.float32(0))


def block_tail_q(acc3, S3, zf3, wscale3, bias3, s_y2, lo, hi, Sd, zfd, wscaled, biasd, s_x, s_out):
    """u8 output of the fused block tail: the same chain with inv = 1/s_out folded into A3, Ad and B BEFORE the
    zero-point constants are formed, then sat_u8(rint(y))."""
    inv = np.float32(1.0) / np.float32(s_out)
    y = _block_tail_core(acc3, S3, zf3, wscale3, bias3, s_y2, lo, hi, Sd, zfd, wscaled, biasd, s_x, inv)
    return np.clip(np.nan_to_num(np.rint(y), nan=0.0), 0, 255).astype(np.uint8)


def requant_u8(y, s_out):
    """cvt.rni.sat.u8.f32(y * (1/s_out))"""
    inv = np.float32(1.0) / np.float32(s_out)
    q = np.rint((np.asarray(y, np.float32) * inv).astype(np.float32))
    return np.clip(np.nan_to_num(q, nan=0.0), 0, 255).astype(np.uint8)


def requant_s8(y, s_out):
    """cvt.rni.sat.s8.f32(y * (1/s_out)), returned as the raw byte"""
    inv = np.float32(1.0) / np.float32(s_out)
    q = np.rint((np.asarray(y, np.float32) * inv).astype(np.float32))
    return np.clip(np.nan_to_num(q, nan=0.0), -128, 127).astype(np.int8).view(np.uint8)


def act_scale_from_absmax(amax):
    amax = np.float32(amax)
    return np.float32(1.0) if amax == 0 else np.float32(amax / np.float32(255.0))


# ----------------------------------------------------------------------------------------------
# fp32 forward of the reference model (resnet.py:204-220) restated on a module tree's tensors
# ----------------------------------------------------------------------------------------------
def torch_forward(net, x):
    """Stock-torch fp32 forward of a ResNet-shaped module tree (reference resnet.py:204-220,
    BasicBlock :55-68, Bottleneck :97-116).  Works for the reference's modules and for this
    repo's resnet.py mirror alike (same attribute names).  eval-mode BatchNorm."""
    import torch
    import torch.nn.functional as F

    def bn(m, t):
        return F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)

    def conv(m, t):
        return F.conv2d(t, m.weight, None, m.stride, m.padding)

    with torch.no_grad():
        t = F.relu(bn(net.bn1, conv(net.conv1, x)))
        t = F.max_pool2d(t, 3, 2, 1)
        for stage in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in stage:
                idt = t
                o = F.relu(bn(blk.bn1, conv(blk.conv1, t)))
                o = bn(blk.bn2, conv(blk.conv2, o))
                if hasattr(blk, "conv3"):
                    o = bn(blk.bn3, conv(blk.conv3, F.relu(o)))
                if blk.downsample is not None:
                    idt = bn(blk.downsample[1], conv(blk.downsample[0], t))
                t = F.relu(o + idt)
        t = torch.flatten(F.adaptive_avg_pool2d(t, (1, 1)), 1)
        return F.linear(t, net.fc.weight, net.fc.bias)
