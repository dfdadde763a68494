"""tools/diag_div.py -- which fp32 arithmetic does ATen use on CUDA tensors for each op of
functions.py:41  (((tensor/scale) + z).round() - z) * scale  ?   (to be confirmed on
the GPU box).  Compares torch-on-CUDA results bitwise with numpy float32/float64 hypotheses."""
import numpy as np
import torch

rng = np.random.default_rng(0)
tot = {}
n_el = 0
for trial in range(40):
    K = 1 << 19
    w = (rng.standard_normal(K) * rng.uniform(0.01, 0.2)).astype(np.float32)
    bit = [8, 6, 4, 2][trial % 4]
    mn, mx = float(w.min()), float(w.max())
    scale = (mx - mn) / (2 ** bit - 1)          # python double
    z = round(mn / scale)
    s32 = np.float32(scale)
    t = torch.from_numpy(w).cuda()
    d_gpu = (t / scale).cpu().numpy()
    hyp = {
        "div: w / s32 (true fp32 divide)": (w / s32).astype(np.float32),
        "div: w * f32(1f / s32)": (w * (np.float32(1.0) / s32)).astype(np.float32),
        "div: w * f32(1.0 / scale64)": (w * np.float32(1.0 / scale)).astype(np.float32),
        "div: w * f32(1.0 / f64(s32))": (w * np.float32(1.0 / float(s32))).astype(np.float32),
        "div: f32(f64(w) / scale64)": (w.astype(np.float64) / scale).astype(np.float32),
        "div: f32(f64(w) * (1.0/scale64))": (w.astype(np.float64) * (1.0 / scale)).astype(np.float32),
    }
    for k, v in hyp.items():
        tot[k] = tot.get(k, 0) + int((v.view(np.uint32) != d_gpu.view(np.uint32)).sum())
    # downstream ops, each fed with the GPU's own previous result
    a_gpu = (torch.from_numpy(d_gpu).cuda() + z).cpu().numpy()
    tot["add: f32(t1 + f32(z))"] = tot.get("add: f32(t1 + f32(z))", 0) + int(((d_gpu + np.float32(z)).astype(np.float32).view(np.uint32) != a_gpu.view(np.uint32)).sum())
    r_gpu = torch.from_numpy(a_gpu).cuda().round().cpu().numpy()
    tot["round: rint"] = tot.get("round: rint", 0) + int((np.rint(a_gpu).view(np.uint32) != r_gpu.view(np.uint32)).sum())
    s_gpu = (torch.from_numpy(r_gpu).cuda() - z).cpu().numpy()
    tot["sub: f32(t3 - f32(z))"] = tot.get("sub: f32(t3 - f32(z))", 0) + int(((r_gpu - np.float32(z)).astype(np.float32).view(np.uint32) != s_gpu.view(np.uint32)).sum())
    m_gpu = (torch.from_numpy(s_gpu).cuda() * scale).cpu().numpy()
    tot["mul: f32(k * s32)"] = tot.get("mul: f32(k * s32)", 0) + int(((s_gpu * s32).astype(np.float32).view(np.uint32) != m_gpu.view(np.uint32)).sum())
    tot["mul: f32(f64(k) * scale64)"] = tot.get("mul: f32(f64(k) * scale64)", 0) + int(((s_gpu.astype(np.float64) * scale).astype(np.float32).view(np.uint32) != m_gpu.view(np.uint32)).sum())
    # the whole chain as the reference writes it, on the device
    full_gpu = ((((t / scale) + z).round() - z) * scale).cpu().numpy()
    chain = ((w * np.float32(1.0 / scale)).astype(np.float32) + np.float32(z)).astype(np.float32)
    k = (np.rint(chain) - np.float32(z)).astype(np.float32)
    tot["chain: H(1.0/scale64)"] = tot.get("chain: H(1.0/scale64)", 0) + int(((k * s32).astype(np.float32).view(np.uint32) != full_gpu.view(np.uint32)).sum())
    n_el += K
print("elements per hypothesis:", n_el, "torch", torch.__version__)
for k, v in tot.items():
    print("%-40s mismatches %d" % (k, v))
