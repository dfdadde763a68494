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
cap = 24 * 800
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
names = {0: "BLD row start", 1: "BLD aempty ok", 2: "BLD built", 3: "BLD arrived", 8: "BLD row end", 4: "MMA go", 9: "MMA issued",
         5: "EPI tfull ok", 10: "EPI stored", 11: "EPI bar ok", 6: "EPI row end"}
t0 = ev[0, 2]
lo, hi = int(os.environ.get("TRACE_FROM", "300")), int(os.environ.get("TRACE_TO", "420"))
for e, idx, t, who in ev[lo:hi]:
    print("%8d  %-14s row %4d (issuer %d)" % (t - t0, names.get(int(e), str(e)), idx, who))
for e in (0, 4, 5):
    tt = ev[ev[:, 0] == e][:, 2]
    if len(tt) > 10:
        print(names[e], "period median %.0f clk" % np.median(np.diff(tt)))
