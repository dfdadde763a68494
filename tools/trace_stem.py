"""tools/trace_stem.py -- CTA 0's builder / MMA / epilogue timeline of the fused stem kernel."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200"))
os.environ.setdefault("SLQ_DEBUG_LIB", "1")  # the tracing build of the library
import slq_lib as L
lib = L.lib()
N, H, W = 256, 224, 224
x = torch.randn(N, 3, H, W, device="cuda")
w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
a = torch.ones(64, device="cuda"); b = torch.zeros(64, device="cuda")
sc = torch.full((4,), 0.02, device="cuda")
ws = torch.empty(lib.slq_stem_workspace_bytes(N, H, W), dtype=torch.uint8, device="cuda")
h = ctypes.c_void_p()
L.check(lib.slq_stem_create(N, H, W, ws.data_ptr(), ctypes.byref(h)))
L.check(lib.slq_stem_set_weights(h, w.data_ptr(), L.current_stream()))
out = torch.empty(N * 56 * 56 * 64, dtype=torch.uint8, device="cuda")
scratch = torch.empty(16, device="cuda")
cap = 24 * 2400
buf = torch.zeros(3 * cap, dtype=torch.int64, device="cuda")
for rep in range(2):
    buf.zero_(); torch.cuda.synchronize()
    lib.slq_debug_set_trace(buf.data_ptr() if rep == 1 else None, cap)
    L.check(lib.slq_stem_launch(h, x.data_ptr(), a.data_ptr(), b.data_ptr(), sc.data_ptr(), 0, out.data_ptr(), L.OUT_U8, scratch.data_ptr(), None, L.current_stream()))
    torch.cuda.synchronize()
lib.slq_debug_set_trace(None, 0)
hh = buf.cpu().numpy().reshape(cap, 3)
issuer = np.repeat(np.arange(24), cap // 24)
keep = hh[:, 0] > 0
ev = np.concatenate([hh[keep], issuer[keep, None]], 1)
ev[:, 0] -= 1
ev = ev[np.argsort(ev[:, 2], kind="stable")]
if os.environ.get("STEM_OLD"):
    names = {0: "BLD row start", 1: "BLD aempty ok", 2: "BLD built", 3: "BLD arrived", 8: "BLD row end", 4: "MMA go", 9: "MMA issued",
             5: "EPI tfull ok", 10: "EPI stored", 11: "EPI bar ok", 6: "EPI row end"}
else:  # stem_ts_kernel: issuer 0 producer, 1 builder thread 0, 2-5 MMA warps, 6 epilogue thread 0, 7 pool thread 0
    names = {0: "PRD top", 1: "PRD sempty ok", 2: "BLD top", 3: "BLD sfull ok", 4: "BLD converted", 5: "BLD pfree ok",
             6: "BLD stored", 7: "MMA top", 8: "MMA tempty ok", 9: "MMA pfull ok", 10: "MMA issued", 11: "EPI top",
             12: "EPI tfull ok", 13: "EPI loaded", 14: "EPI emit", 15: "EPI vfree ok", 16: "POOL top", 17: "POOL vfull ok",
             18: "POOL done"}
t0 = ev[0, 2]
lo, hi = int(os.environ.get("TRACE_FROM", "600")), int(os.environ.get("TRACE_TO", "760"))
for e, idx, t, who in ev[lo:hi]:
    print("%8d  %-14s idx %4d (issuer %d)" % (t - t0, names.get(int(e), str(e)), idx, who))
# per role: where the time of one loop trip goes (mean clocks from each event to the next one of the same issuer)
for who in sorted(set(ev[:, 3].tolist())):
    sub = ev[ev[:, 3] == who]
    sub = sub[len(sub) // 4:]          # steady state
    if len(sub) < 20:
        continue
    d = np.diff(sub[:, 2])
    print("issuer %d: %d events, span %d clk" % (who, len(sub), sub[-1, 2] - sub[0, 2]))
    for e in sorted(set(sub[:-1, 0].tolist())):
        m = sub[:-1, 0] == e
        print("    after %-14s mean %7.0f clk  (n=%d, total %d)" % (names.get(int(e), str(e)), d[m].mean(), m.sum(), d[m].sum()))
