"""tools/trace_conv.py -- prints CTA 0's producer / MMA / epilogue timeline of one conv layer
(slq_debug_set_trace).  usage: python tools/trace_conv.py cin cout k stride H [N [w16]]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("SLQ_DEBUG_LIB", "1")  # the tracing build of the library
import slq_lib as L  # noqa: E402
from helpers import ConvCase  # noqa: E402

cin, cout, k, stride, H = [int(v) for v in sys.argv[1:6]]
N = int(sys.argv[6]) if len(sys.argv) > 6 else 256
bits = np.full(cout, 32 if (len(sys.argv) > 7 and sys.argv[7] == 'w16') else 8, np.int32)
case = ConvCase(N, H, cin, cout, k, stride, bits, seed=1)
lib = L.lib()
cap = 24 * 600
buf = torch.zeros(3 * cap, dtype=torch.int64, device="cuda")
out = torch.empty((case.M, cout), dtype=torch.uint8, device="cuda")
ws = torch.ones(cout, device="cuda")
zz = torch.zeros(cout, device="cuda")
sc = torch.ones(4, device="cuda")
e = L.Epilogue(ws.data_ptr(), zz.data_ptr(), zz.data_ptr(), sc.data_ptr(), 0, 1, -1, None, 0, out.data_ptr(), None,
               L.OUT_U8, 1, case.rowsum.data_ptr(), None, 1, case.rowsum.numel())
for rep in range(2):
    buf.zero_()
    torch.cuda.synchronize()
    L.check(lib.slq_debug_set_trace(buf.data_ptr() if rep == 1 else None, cap))
    L.check(lib.slq_conv_launch(case.handle, ctypes.byref(e), L.current_stream()))
    torch.cuda.synchronize()
lib.slq_debug_set_trace(None, 0)
h = buf.cpu().numpy().reshape(cap, 3)
issuer = np.repeat(np.arange(24), cap // 24)
keep = h[:, 0] > 0
ev = np.concatenate([h[keep], issuer[keep, None]], 1)
ev[:, 0] -= 1
n = len(ev)
ev = ev[np.argsort(ev[:, 2], kind="stable")]
t0 = ev[0, 2]
names = {0: "A issue>", 1: "A issue<", 2: "B issue>", 3: "B issue<", 4: "MMA kb", 5: "EPI begin", 6: "EPI end", 7: "MMA wait"}
print("events", n)
for evn, idx, t, who in ev[:int(os.environ.get("TRACE_ROWS", "160"))]:
    print("%8d  %-10s %4d  (issuer %d)" % (t - t0, names[int(evn)], idx, who))
mma = ev[ev[:, 0] == 4]
if len(mma) > 10:
    d = np.diff(mma[:, 2])
    print("MMA k-block period: median %.0f mean %.0f clk over %d" % (np.median(d), d.mean(), len(d)))
for a_, b_, nm in ((0, 1, "A"), (2, 3, "B"), (7, 4, "MMA full-wait")):
    s_ = {int(i): t for e_, i, t, _w in ev if e_ == a_}
    f_ = {int(i): t for e_, i, t, _w in ev if e_ == b_}
    dd = [f_[i] - s_[i] for i in s_ if i in f_]
    if dd:
        print("%s issue duration: median %.0f clk" % (nm, np.median(dd)))
