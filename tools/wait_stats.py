"""tools/wait_stats.py -- where the roles of conv_umma_kernel's CTA 0 spend their time: cycles blocked in
mbarrier waits per role over one launch (slq_debug_set_trace with a negative capacity = statistics only,
no per-step overhead).  usage: python tools/wait_stats.py cin cout k stride H [N [w16]]"""
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
mode = sys.argv[7] if len(sys.argv) > 7 else ""
bits = np.full(cout, 32 if mode == "w16" else 8, np.int32)
case = ConvCase(N, H, cin, cout, k, stride, bits, seed=1)
lib = L.lib()
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
out = torch.empty((case.M, cout), dtype=torch.uint8, device="cuda")
ws = torch.ones(cout, device="cuda")
zz = torch.zeros(cout, device="cuda")
sc = torch.ones(4, device="cuda")
res = torch.randint(0, 256, (case.M, cout), dtype=torch.uint8, device="cuda") if mode == "res" else None
rs_out = torch.zeros((max(cout // 64, 1), case.M), dtype=torch.int32, device="cuda")
e = L.Epilogue(ws.data_ptr(), zz.data_ptr(), zz.data_ptr(), sc.data_ptr(), 0, 1, 2 if res is not None else -1, L.ptr(res), 0,
               out.data_ptr(), None, L.OUT_U8, 1, case.rowsum.data_ptr(), rs_out.data_ptr() if mode == "res" else None, 1,
               case.rowsum.numel())
for rep in range(3):
    buf.zero_()
    L.check(lib.slq_debug_set_trace(buf.data_ptr() if rep == 2 else None, -1))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(lib.slq_conv_launch(case.handle, ctypes.byref(e), L.current_stream()))
    b.record()
    torch.cuda.synchronize()
lib.slq_debug_set_trace(None, 0)
h = buf.cpu().numpy()
print("layer %d->%d k%d s%d H%d N%d %s: %.1f us" % (cin, cout, k, stride, H, N, mode, 1e3 * a.elapsed_time(b)))
names = ["producer0 wait(empty)", "producer1 wait(empty)", "mma0 wait(full)", "mma1 wait(full)", "mma0 wait(acc/tstart)",
         "mma1 wait(acc/tstart)", "epi team0 wait(tfull)", "epi team1 wait(tfull)"]
life = [8, 9, 10, 11, 10, 11, 12, 13]
print("  roles start %d clk after kernel entry; CTA 0 lives %d clk; first CTAs %s; last CTAs %s" % (
    h[14], h[15], h[16:24].tolist(), h[48:56].tolist()))
for i, n in enumerate(names):
    tot = h[life[i]]
    print("  %-24s %9d clk of %9d (%.0f%%)" % (n, h[i], tot, 100.0 * h[i] / max(tot, 1)))
if h[62] > 0:
    segs = ["loop top (barriers, constants, window-sum gather)", "wait accumulator (tfull)", "wait residual tile / S from TMEM",
            "unit loop (TMEM loads, arithmetic, staging stores)", "hand-offs (release, fence, team barrier, TMA store)", "rowsum tail"]
    tot = float(sum(h[56:62]))
    print("  epilogue warp 4, %d tiles, %.0f clk per tile:" % (h[62], tot / h[62]))
    for i, n in enumerate(segs):
        print("    %-52s %7.0f clk per tile (%.0f%%)" % (n, h[56 + i] / h[62], 100.0 * h[56 + i] / max(tot, 1)))
