"""tools/layer_time.py -- device time of ONE conv layer launched like the engine launches it (u8 output with
rowsum side tensor, optional u8 / s8 residual, or the two-limb downsample mode), CUDA events, best of 5 after
warm-up.  $SLQ_LIB_VARIANT=<name> loads libslq_b200_<name>.so (an A/B build of the same sources).

    python tools/layer_time.py 64 256 1 1 56 256 res      # Cin Cout k stride H N [res|sres|w16]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("semilayer-wise-mixed-precision-quantization_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import slq_lib as L  # noqa: E402
from helpers import ConvCase  # noqa: E402

cin, cout, k, stride, H, N = [int(v) for v in sys.argv[1:7]]
mode = sys.argv[7] if len(sys.argv) > 7 else ""
bits = np.full(cout, 32 if mode == "w16" else 8, np.int32)
bits[::2] = 4 if mode != "w16" else 32
case = ConvCase(N, H, cin, cout, k, stride, bits, seed=1)
lib = L.lib()
dev = "cuda"
out = torch.empty((case.M, cout), dtype=torch.uint8, device=dev)
ws = torch.full((cout,), 1e-3, device=dev)
zz = torch.full((cout,), -120.0, device=dev)
bb = torch.zeros(cout, device=dev)
sc = torch.tensor([1.0, 0.05, 0.5, 1.0], device=dev)
res = torch.randint(0, 256, (case.M, cout), dtype=torch.uint8, device=dev) if mode in ("res", "sres") else None
rs_out = torch.zeros((max(cout // 64, 1), case.M), dtype=torch.int32, device=dev)
out_mode = L.OUT_S8 if mode == "w16" else L.OUT_U8
e = L.Epilogue(ws.data_ptr(), zz.data_ptr(), bb.data_ptr(), sc.data_ptr(), 0, 1, 2 if res is not None else -1, L.ptr(res),
               1 if mode == "sres" else 0, out.data_ptr(), None, out_mode, 0 if mode == "w16" else 1,
               case.rowsum.data_ptr(),
               rs_out.data_ptr() if out_mode == L.OUT_U8 and not os.environ.get("LT_NO_RS") else None,  # $LT_NO_RS: no consumer gathers the output's channel sums
               1, case.rowsum.numel())
best = 1e9
for rep in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(lib.slq_conv_launch(case.handle, ctypes.byref(e), L.current_stream()))
    b.record()
    torch.cuda.synchronize()
    if rep >= 3:
        best = min(best, a.elapsed_time(b))
print("%s layer %d->%d k%d s%d H%d N%d %s: %.1f us" % (os.environ.get("SLQ_LIB_VARIANT", "prod"), cin, cout, k, stride, H, N, mode, 1e3 * best))
