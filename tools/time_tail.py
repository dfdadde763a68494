"""tools/time_tail.py -- times slq_blocktail_launch (u8 mode) of the three ResNet-50 stages at batch 256 under the
$SLQ_BT_DBG variants of the debug build (1 no constant loads, 2 no limb TMEM loads, 4 no staging / store, 8 no
downsample MMAs, 32 no epilogue arithmetic): which part of the kernel the time goes to."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200"))
os.environ.setdefault("SLQ_DEBUG_LIB", "1")
import slq_lib as L
lib = L.lib()
N = int(os.environ.get("BT_N", "256"))
for cin, cmid, cout, stride, H in [(64, 64, 256, 1, 56), (256, 128, 512, 2, 56), (512, 256, 1024, 2, 28)]:
    Ho = (H - 1) // stride + 1
    M = N * Ho * Ho
    y2 = torch.randint(0, 256, (M, cmid), dtype=torch.uint8, device="cuda")
    x = torch.randint(0, 256, (N * H * H, cin), dtype=torch.uint8, device="cuda")
    wg3 = torch.randint(0, 256, (cout, cmid), dtype=torch.uint8, device="cuda")
    wgd = torch.randint(0, 256, (2 * cout, cin), dtype=torch.uint8, device="cuda")
    out = torch.empty((M, cout), dtype=torch.uint8, device="cuda")
    vec = [torch.rand(cout, device="cuda") * 1e-4 for _ in range(6)]
    sc = torch.full((4,), 0.02, device="cuda")
    desc = L.BlockTailDesc(N, H, H, cin, stride, cmid, cout, L.IMPL_UMMA)
    h = ctypes.c_void_p()
    L.check(lib.slq_blocktail_create(ctypes.byref(desc), y2.data_ptr(), x.data_ptr(), wg3.data_ptr(), wgd.data_ptr(), ctypes.byref(h)))
    e = L.BlockTailEpilogue(*[v.data_ptr() for v in vec], sc.data_ptr(), 0, 1, 2, out.data_ptr(), L.OUT_U8, None)
    res = []
    for dbg in [int(v) for v in os.environ.get("BT_DBG_LIST", "0,1,2,4,8,32,3,35,39,47").split(",")]:
        os.environ["SLQ_BT_DBG"] = str(dbg)
        for _ in range(2):
            L.check(lib.slq_blocktail_launch(h, ctypes.byref(e), L.current_stream()))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.check(lib.slq_blocktail_launch(h, ctypes.byref(e), L.current_stream()))
        e1.record(); torch.cuda.synchronize()
        res.append("dbg=%d %.1f" % (dbg, 1e3 * e0.elapsed_time(e1) / 5))
    print("Cin %d Cmid %d Cout %d s%d H%d: " % (cin, cmid, cout, stride, H) + " | ".join(res) + "  (us per launch)")
    # where CTA 0's roles wait (cycles blocked in mbarrier waits / the crew barrier, out of each role's loop time)
    os.environ["SLQ_BT_DBG"] = "0"
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    lib.slq_debug_set_trace(buf.data_ptr(), -64)
    L.check(lib.slq_blocktail_launch(h, ctypes.byref(e), L.current_stream()))
    torch.cuda.synchronize()
    lib.slq_debug_set_trace(None, 0)
    st = buf.cpu().numpy()
    tiles = max(int(st[13]), 1)
    print("   CTA 0: %d tiles; per tile: producers wait(empty) %d / %d of %d / %d clk | MMA wait(full) %d / %d, wait(tempty) %d / %d of %d / %d | "
          "crew wait(tfull) %d, barrier %d of %d" % (tiles, st[0] * 2 // tiles, st[1] * 2 // tiles, st[2] * 2 // tiles, st[3] * 2 // tiles,
                                                    st[4] * 2 // tiles, st[5] * 2 // tiles, st[6] * 2 // tiles, st[7] * 2 // tiles,
                                                    st[8] * 2 // tiles, st[9] * 2 // tiles, st[10] // tiles, st[11] // tiles, st[12] // tiles))
    print("   crew thread 0, per tile: TMEM loads %d | fence+arrive %d | arithmetic+pack+STS %d | proxy fence %d | crew barrier %d | store issue + wait_group %d clk"
          % tuple(int(v) // tiles for v in st[16:22]))
    lib.slq_blocktail_destroy(h)
