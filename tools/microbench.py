"""tools/microbench.py -- SURVEY.md 8(d) "Microbench": every quantised conv shape of an architecture x
batch N in {1 .. 512}, and the weight quantizer per K x bit in {4, 8}; BASELINE.json configs[4].

Per conv shape and batch: device time of ONE launch (CUDA events on the launching stream, best of
`--reps` after a warm-up, a 256 MB L2 flush between repetitions), TOPS = 2*MAC / t, % of the INT8 tensor
peak, algorithmic GB/s (u8 NHWC in + out + residual + weights once) and which roof is nearer.
Per (K, bit): one slq_quantize_rows launch over 128 MB of fp32 rows, GB/s of algorithmic bytes (fp32 row read +
write-back + packed codes + 8 B metadata).

usage (GPU box):  python tools/microbench.py [--arch resnet50] [--batches 1,8,32,128,256,512] [--out FILE.json]
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
sys.path.insert(0, PKG)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--batches", default="1,8,32,128,256,512")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch
    import slq_build
    slq_build.build()
    import functions
    import resnet
    import slq_lib as L
    sys.path.insert(0, ROOT)
    import bench  # load_peaks: the measured roofs of this pool's B200s
    peaks = bench.load_peaks()
    int8_peak = 2.0 * peaks["bf16"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    results = {"arch": args.arch, "int8_peak_tops": int8_peak, "hbm_peak_gbs": peaks["hbm"], "conv": [], "quantizer": []}

    # ---- model with the P0 8/4-bit assignment (same as bench.py) --------------------------------
    torch.manual_seed(0)
    net = getattr(resnet, args.arch)(num_classes=1000).to(dev).eval()
    table = np.load(os.path.join(PKG, "data", "p0_bits.npz"))[args.arch]
    cpb = 3 if args.arch == "resnet50" else 2
    blocks = [b for s in (net.layer1, net.layer2, net.layer3, net.layer4) for b in s]
    for lnum in np.unique(table[:, 0]):
        sel = table[table[:, 0] == lnum]
        conv = getattr(blocks[(lnum - 1) // cpb], "conv%d" % ((lnum - 1) % cpb + 1))
        functions.quantize_rows(conv.weight.data, sel[:, 1], sel[:, 2], div_mode=L.DIV_TRUE, want_codes=False)

    print("%-28s %5s %9s %9s %7s %9s %6s" % ("shape (Cin Cout k s H res w16)", "N", "us", "TOPS", "%int8", "GB/s", "near"))
    for N in [int(v) for v in args.batches.split(",")]:
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randn(N, 3, 224, 224, generator=g, device=dev)
        eng = net.slq_engine(x)
        eng.refresh_weights()
        eng.calibrate(x)
        eng.forward(x)
        torch.cuda.synchronize()
        seen = {}
        for op in eng.ops:
            key = (op.Cin, op.Cout, op.k, op.stride, op.H, 1 if op.res_id >= 0 else 0, op.w16)
            if key in seen:
                continue
            mode = L.OUT_S8 if op.signed else L.OUT_U8
            e = eng._epilogue(op, mode, eng.act[op.out_id].data_ptr())
            best = 1e30
            for rep in range(args.reps + 1):
                flush.fill_(rep)  # evict the layer's tensors from L2 (inputs larger than L2 at N >= 128 anyway)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                L.check(eng.lib.slq_conv_launch(op.handle, ctypes.byref(e), st.cuda_stream))
                b.record(st)
                torch.cuda.synchronize()
                if rep:
                    best = min(best, a.elapsed_time(b))
            ops = 2.0 * op.M * op.Cout * op.k * op.k * op.Cin
            byts = (eng.act[op.in_id].numel() + eng.act[op.out_id].numel() + op.wg.numel() +
                    (eng.act[op.res_id].numel() if op.res_id >= 0 else 0))
            tops = ops / (best * 1e-3) / 1e12
            gbs = byts / (best * 1e-3) / 1e9
            near = "tensor" if tops / int8_peak > gbs / peaks["hbm"] else "hbm"
            seen[key] = 1
            results["conv"].append({"Cin": op.Cin, "Cout": op.Cout, "k": op.k, "stride": op.stride, "H": op.H,
                                    "res": key[5], "w16": op.w16, "N": N, "us": 1e3 * best, "tops": tops,
                                    "frac_int8": tops / int8_peak, "gbs": gbs, "frac_hbm": gbs / peaks["hbm"], "near": near})
            print("%4d %4d %d %d %3d %d %d        %5d %9.1f %9.1f %6.1f%% %9.1f %6s" % (
                key + (N, 1e3 * best, tops, 100 * tops / int8_peak, gbs, near)))
        del eng
        net._slq_engines = {}
        torch.cuda.empty_cache()

    # ---- quantizer: one launch over 128 MB of rows per (K, bit) ----------------------------------
    print("\n%-6s %4s %9s %9s %7s" % ("K", "bit", "us", "GB/s", "%hbm"))
    lib = L.lib()
    for K in (64, 128, 256, 512, 576, 1024, 1152, 2048, 2304, 4608):
        rows_n = (1 << 25) // K  # 128 MB of fp32 rows per launch: larger than L2, long enough to time
        for bit in (4, 8):
            w = torch.randn(rows_n, K, device=dev)
            rows = torch.arange(rows_n, dtype=torch.int32, device=dev)
            bits = torch.full((rows_n,), bit, dtype=torch.int32, device=dev)
            row_bytes = (lib.slq_packed_row_bytes(K, bit) + 15) // 16 * 16
            offs = (torch.arange(rows_n, dtype=torch.int64, device=dev) * row_bytes)
            blob = torch.empty(rows_n * row_bytes, dtype=torch.uint8, device=dev)
            z = torch.empty(rows_n, dtype=torch.int32, device=dev)
            s32 = torch.empty(rows_n, dtype=torch.float32, device=dev)
            status = torch.empty(rows_n, dtype=torch.int32, device=dev)
            best = 1e30
            for rep in range(args.reps + 1):
                flush.fill_(rep)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                L.check(lib.slq_quantize_rows(w.data_ptr(), rows_n, K, rows.data_ptr(), bits.data_ptr(), rows_n,
                                              L.DIV_RECIP, 1, blob.data_ptr(), offs.data_ptr(), z.data_ptr(),
                                              s32.data_ptr(), status.data_ptr(), st.cuda_stream))
                b.record(st)
                torch.cuda.synchronize()
                if rep:
                    best = min(best, a.elapsed_time(b))
            byts = rows_n * (8.0 * K + lib.slq_packed_row_bytes(K, bit) + 8)
            gbs = byts / (best * 1e-3) / 1e9
            results["quantizer"].append({"K": K, "bit": bit, "rows": rows_n, "us": 1e3 * best, "gbs": gbs,
                                         "frac_hbm": gbs / peaks["hbm"]})
            print("%-6d %4d %9.1f %9.1f %6.1f%%" % (K, bit, 1e3 * best, gbs, 100 * gbs / peaks["hbm"]))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)
        print("wrote", args.out)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("microbench: %.0f s" % (time.time() - t0))
