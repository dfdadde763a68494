"""tools/time_stem.py -- times slq_stem_launch (u8 mode) at batch 256 under $SLQ_STEM_DBG variants."""
import ctypes, os, sys
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
for dbg in [int(v) for v in os.environ.get("STEM_DBG_LIST", "0").split(",")]:
    os.environ["SLQ_STEM_DBG"] = str(dbg)
    for _ in range(2):
        L.check(lib.slq_stem_launch(h, x.data_ptr(), a.data_ptr(), b.data_ptr(), sc.data_ptr(), 0, out.data_ptr(), L.OUT_U8, scratch.data_ptr(), None, L.current_stream()))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        L.check(lib.slq_stem_launch(h, x.data_ptr(), a.data_ptr(), b.data_ptr(), sc.data_ptr(), 0, out.data_ptr(), L.OUT_U8, scratch.data_ptr(), None, L.current_stream()))
    e1.record(); torch.cuda.synchronize()
    print("dbg=%d  %.1f us per launch" % (dbg, 1e3 * e0.elapsed_time(e1) / 5))
