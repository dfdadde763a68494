"""tools/launch_gap.py -- per-launch cost of one conv layer launched back to back (stream and CUDA graph)
against the lifetime of its CTAs (wait-statistics mode): what a launch costs beyond the CTAs' own work.
usage: python tools/launch_gap.py cin cout k stride H [N]"""
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
case = ConvCase(N, H, cin, cout, k, stride, np.full(cout, 8, np.int32), seed=1)
lib = L.lib()
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
out = torch.empty((case.M, cout), dtype=torch.uint8, device="cuda")
ws = torch.ones(cout, device="cuda")
zz = torch.zeros(cout, device="cuda")
sc = torch.ones(4, device="cuda")
e = L.Epilogue(ws.data_ptr(), zz.data_ptr(), zz.data_ptr(), sc.data_ptr(), 0, 1, -1, None, 0, out.data_ptr(), None,
               L.OUT_U8, 1, case.rowsum.data_ptr(), None, 1, case.rowsum.numel())
st = torch.cuda.current_stream()


def launch(n):
    for _ in range(n):
        L.check(lib.slq_conv_launch(case.handle, ctypes.byref(e), st.cuda_stream))


launch(3)
torch.cuda.synchronize()
L.check(lib.slq_debug_set_trace(buf.data_ptr(), -1))
launch(1)
torch.cuda.synchronize()
lib.slq_debug_set_trace(None, 0)
life = int(buf.cpu().numpy()[15])
R = 30
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
launch(R)
b.record()
torch.cuda.synchronize()
t_stream = 1e3 * a.elapsed_time(b) / R
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(st)
with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
    for _ in range(R):
        L.check(lib.slq_conv_launch(case.handle, ctypes.byref(e), torch.cuda.current_stream().cuda_stream))
g.replay()
torch.cuda.synchronize()
a.record()
g.replay()
b.record()
torch.cuda.synchronize()
t_graph = 1e3 * a.elapsed_time(b) / R
print("%d->%d k%d s%d H%d: CTA 0 lifetime %.1f us | per launch: stream %.1f us, graph %.1f us  (pdl=%s)" % (
    cin, cout, k, stride, H, life / 1965.0, t_stream, t_graph, os.environ.get("SLQ_NO_PDL") is None))
