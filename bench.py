#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json metric):
ResNet-50 semilayer-wise mixed 8/4-bit inference, batch 256 per GPU, synthetic 224x224 images,
random-init weights, policy-P0 bit assignment derived from the reference's delta-loss table.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own path on the host CPU cores

A "step" is one forward pass of one batch (256 images per GPU) through the whole hot path:
stem -> 52 quantised/fp32-weight convs on tcgen05 (fused dequant+BN+residual+ReLU epilogues)
-> avgpool+fc.  `value` has the inputs resident in HBM; `e2e` goes through the public call
`net(x)` with pinned-host inputs copied H2D and logits copied D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
sys.path.insert(0, PKG)

METRIC = "ResNet-50 mixed 8/4-bit images/s"


def metric_name(arch):  # BASELINE.json's metric; the other architectures are parity / side configurations
    return METRIC.replace("ResNet-50", {"resnet18": "ResNet-18", "resnet34": "ResNet-34"}.get(arch, "ResNet-50"))
UNIT = "images/s"
GOP_PER_IMG = {"resnet50": 8.178, "resnet34": 7.328, "resnet18": 3.628}  # 2*MAC, convs + fc
ALGO_MB_PER_IMG = {"resnet50": 27.84, "resnet34": 9.17, "resnet18": 6.22}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Timed region starts here: samples taken before this call (warm-up) are dropped, except
        the last one, so that even a region shorter than the sampling period has a reading."""
        self.rows = self.rows[-1:]

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_forward(arch, batch, reps, threads=None):
    """The reference's path on the host cores: fp32 fake-quant forward of the same P0 model, stock
    torch modules on CPU (oracle.torch_forward restates resnet.py:204-220 call for call)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import slq_oracle as so
    import resnet
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = getattr(resnet, arch)(num_classes=1000).eval()
    table = np.load(os.path.join(PKG, "data", "p0_bits.npz"))[arch]
    cpb = 3 if arch == "resnet50" else 2
    blocks = [b for s in (net.layer1, net.layer2, net.layer3, net.layer4) for b in s]
    for lnum, cn, bit in table:  # oracle quantizer: the CPU restatement of functions.py:9-43
        conv = getattr(blocks[(lnum - 1) // cpb], "conv%d" % ((lnum - 1) % cpb + 1))
        so.channel_wise(conv.weight.data.reshape(conv.out_channels, -1).numpy(), int(bit), int(cn))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    so.torch_forward(net, x)  # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        so.torch_forward(net, x)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    # all the host threads this process may use: torchrun pins OMP_NUM_THREADS=1 per rank, and only rank 0
    # runs this arm, so take the cores of its affinity mask instead
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(threads)  # before torch is imported
    times, threads = cpu_reference_forward(args.arch, batch, args.steps + args.warmup, threads=threads)
    times = times[args.warmup:] if len(times) > args.warmup else times
    total = sum(times)
    val = batch * len(times) / total
    line = {
        "metric": metric_name(args.arch), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "%s semilayer 8/4-bit (P0) inference, synthetic 224x224, random-init" % args.arch,
                   "sample": "batch %d per step on the host CPU" % batch, "batch_per_gpu": args.batch},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d fp32 forwards of batch %d (torch %s CPU, stock nn.functional ops)" %
                                   (len(times), batch, __import__("torch").__version__)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=16, help="images per step of the CPU reference arm")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-agree", action="store_true", help="skip the fp32 agreement check (torch kernels)")
    ap.add_argument("--simt", action="store_true", help="run the dp4a checker kernels instead (debug)")
    ap.add_argument("--layers", default="", help="write the per-layer CUDA-event table to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import slq_build
    slq_build.build()
    import functions
    import resnet
    import slq_engine
    import slq_lib as L

    # ---- model: seeded random init + P0 8/4-bit assignment through the product quantizer --------
    torch.manual_seed(0)
    net = getattr(resnet, args.arch)(num_classes=1000).to(dev).eval()
    table = np.load(os.path.join(PKG, "data", "p0_bits.npz"))[args.arch]
    cpb = 3 if args.arch == "resnet50" else 2
    blocks = [b for s in (net.layer1, net.layer2, net.layer3, net.layer4) for b in s]
    torch.cuda.synchronize()
    tq0 = time.perf_counter()
    packed_bytes = 0
    for lnum in np.unique(table[:, 0]):
        sel = table[table[:, 0] == lnum]
        conv = getattr(blocks[(lnum - 1) // cpb], "conv%d" % ((lnum - 1) % cpb + 1))
        pr = functions.quantize_rows(conv.weight.data, sel[:, 1], sel[:, 2], div_mode=L.DIV_TRUE)
        packed_bytes += int(pr.blob.numel())
    torch.cuda.synchronize()
    quant_ms = 1e3 * (time.perf_counter() - tq0)

    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(B, 3, 224, 224, generator=g, device=dev)
    x_host = x.cpu().pin_memory()
    impl = L.IMPL_SIMT if args.simt else L.IMPL_UMMA
    eng = net.slq_engine(x, impl=impl)
    eng.refresh_weights()
    eng.calibrate(x)
    eng.epoch = resnet.WEIGHT_EPOCH[0]
    net._slq_dirty = False
    logits = eng.forward(x).clone()
    torch.cuda.synchronize()

    # ---- accuracy side-channel: agreement with the fp32 torch restatement (not timed) -----------
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    agree = rel = None
    try:
        if args.no_agree:
            raise RuntimeError("skipped")
        import slq_oracle as so
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        nb = min(B, 64)
        ref = so.torch_forward(net, x[:nb])
        agree = float((ref.argmax(1) == logits[:nb].argmax(1)).float().mean())
        rel = float((ref - logits[:nb]).norm() / ref.norm())
    except Exception as e:  # the checker is optional here
        agree, rel = None, "unavailable: %s" % type(e).__name__

    use_graph = not args.no_graph
    if use_graph:
        try:
            eng.capture_graph(x)
        except Exception as e:
            print("graph capture failed (%s); timing direct launches" % e, file=sys.stderr)
            use_graph = False
    st = torch.cuda.current_stream(dev)

    def step():
        if use_graph:
            eng.graph.replay()
        else:
            eng.launch_all(x.data_ptr(), st.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(args.steps):
        if not use_graph and i == args.steps - 1:  # profiling hook: ncu --nvtx --nvtx-include "slq_step/"
            torch.cuda.nvtx.range_push("slq_step")
            step()
            torch.cuda.nvtx.range_pop()
        else:
            step()
    e1.record(st)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: the public call net(x) with host buffers: H2D + forward + D2H every step ---------
    # The loader-side pattern of the reference (functions.py:110-113: inputs.to(device); net(inputs))
    # with pinned memory and a copy stream: batch i+1 crosses PCIe while batch i computes; every
    # step's logits come back to the host (a blocking read), so nothing is deferred past the region.
    e2e_steps = max(3, min(args.steps, 20))
    copy_stream = torch.cuda.Stream(dev)
    NB = 3  # input buffers: the copy engine always has the next batch queued behind the one in flight
    dbuf = [torch.empty_like(x) for _ in range(NB)]
    landed = [torch.cuda.Event() for _ in range(NB)]

    def e2e_loop(n, src=None, bufs=None):
        src = x_host if src is None else src
        bufs = dbuf if bufs is None else bufs

        def fetch(i):  # host -> device copy of step i's input on the copy stream
            with torch.cuda.stream(copy_stream):
                bufs[i % NB].copy_(src, non_blocking=True)
                landed[i % NB].record(copy_stream)

        fetch(0)
        if n > 1:
            fetch(1)
        out = None
        for i in range(n):
            st.wait_event(landed[i % NB])     # step i's input has landed
            if i + 2 < n:
                fetch(i + 2)                  # bufs[(i+2)%3] was last read by step i-1, which has finished
            out = net(bufs[i % NB]).cpu()     # forward + D2H of the logits (synchronises)
        return out

    e2e_loop(2 * NB)  # every input buffer seen twice: its forward is a captured graph from here on
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(st)
    out_host = e2e_loop(e2e_steps)
    t1.record(st)
    barrier()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * B * e2e_steps / (float(e2e_ms.item()) * 1e-3)

    # the same loop with the loader-side formats the stem also accepts (not the headline: the reference's
    # loader hands over fp32): the batch as fp16 (bit-identical logits) and as raw u8 pixels (normalised in
    # the stem kernel); same engine, same static scales
    import imagenet
    e2e_alt = {}
    net.input_norm = (imagenet.MEAN, imagenet.STD)
    for name, src in (("fp16", x_host.half().pin_memory()),
                      ("u8", torch.randint(0, 256, tuple(x.shape), dtype=torch.uint8).pin_memory())):
        bufs = [torch.empty(tuple(x.shape), dtype=src.dtype, device=dev) for _ in range(NB)]
        e2e_loop(2 * NB, src, bufs)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(st)
        e2e_loop(e2e_steps, src, bufs)
        a1.record(st)
        barrier()
        tm = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_alt[name] = {"value": world * B * e2e_steps / (float(tm.item()) * 1e-3), "unit": UNIT,
                         "h2d_bytes_per_step": int(src.numel() * src.element_size())}
        del bufs

    # ---- roofline of the dominant kernel (conv_umma_kernel), timed live per launch -------------
    peaks = load_peaks()
    conv_ms, conv_ops, conv_bytes = 0.0, 0.0, 0.0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in eng.ops]
    import ctypes
    torch.cuda.synchronize()
    for rep in range(3):
        for (a, b), op in zip(ev, eng.ops):
            mode = L.OUT_S8 if op.signed else L.OUT_U8
            e = eng._epilogue(op, mode, eng.act[op.out_id].data_ptr())
            a.record(st)
            L.check(eng.lib.slq_conv_launch(op.handle, ctypes.byref(e), st.cuda_stream))
            b.record(st)
        torch.cuda.synchronize()
    per_layer = []
    for (a, b), op in zip(ev, eng.ops):
        t = a.elapsed_time(b)
        ops = 2.0 * op.M * op.Cout * op.k * op.k * op.Cin
        byts = op.N_in_bytes if hasattr(op, "N_in_bytes") else (
            eng.act[op.in_id].numel() + eng.act[op.out_id].numel() + op.wg.numel() +
            (eng.act[op.res_id].numel() if op.res_id >= 0 else 0))
        conv_ms += t
        conv_ops += ops
        conv_bytes += byts
        per_layer.append((t, ops, byts))
    if args.layers and rank == 0:
        with open(args.layers, "w") as f:
            f.write("idx Cin Cout k s H M w16 res  us  TOPS  GB/s\n")
            for i, ((t, ops, byts), op) in enumerate(zip(per_layer, eng.ops)):
                f.write("%2d %4d %4d %d %d %3d %7d %d %d  %7.1f %7.1f %7.1f\n" % (
                    i, op.Cin, op.Cout, op.k, op.stride, op.H, op.M, op.w16, 1 if op.res_id >= 0 else 0,
                    1e3 * t, ops / (t * 1e-3) / 1e12, byts / (t * 1e-3) / 1e9))
    achieved_tops = conv_ops / (conv_ms * 1e-3) / 1e12
    int8_peak = 2.0 * peaks["bf16"]  # dense INT8 = 2x the measured dense bf16 tensor throughput
    # DRAM traffic of the same launches from the committed `ncu --set full` capture of one step
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r1_ncu_full_step_summary.csv")
    if args.arch == "resnet50" and B == 256 and os.path.exists(prof):
        import csv
        rows = [r for r in csv.DictReader(open(prof)) if "conv_umma" in r["kernel"]]
        if len(rows) == len(eng.ops):
            traffic = 1e6 * sum(float(r["dram_rd_MB"]) + float(r["dram_wr_MB"]) for r in rows) / len(rows)
    roofline = {"bound": "tensor", "achieved": achieved_tops, "peak": int8_peak, "unit": "TFLOP/s",
                "frac": achieved_tops / int8_peak, "traffic": traffic,
                "traffic_note": "mean DRAM bytes (read+write) per conv launch, profiles/r1_ncu_full_step_summary.csv; "
                                "algorithmic mean %.1f MB" % (conv_bytes / len(eng.ops) / 1e6),
                "kernel": "conv_umma_kernel: the %d conv launches of a step, each timed with CUDA events on the "
                          "launching stream; achieved = sum(2*MAC) / sum(duration)" % len(eng.ops),
                "peak_source": "2 x %s bf16 dense (MEASURED_PEAKS.json sustained): INT8 tensor peak" % peaks["src"],
                "hbm_achieved_gbs": conv_bytes / (conv_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm"],
                "hbm_frac": conv_bytes / (conv_ms * 1e-3) / 1e9 / peaks["hbm"],
                "conv_ms_per_step_serialised": conv_ms, "step_ms": ms / args.steps}

    line = {
        "metric": metric_name(args.arch), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "ours",
        "config": {"workload": "%s semilayer 8/4-bit (P0: %d of %d channels 4-bit) inference, batch %d per GPU, "
                               "synthetic 224x224, random-init" % (args.arch, int((table[:, 2] == 4).sum()), len(table), B),
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "cuda_graph": use_graph, "kernels": "simt-checker" if args.simt else "tcgen05",
                   "l2": "per-step working set (u8 activations ~%.1f GB) >> 126 MB L2; no flush needed" %
                         (sum(a.numel() for a in eng.act) / 1e9),
                   "activation_quant": "static per-tensor u8 (calibrated on one batch)"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": e2e_steps,
                "overlap": "pinned fp32 H2D two batches ahead on a copy stream; blocking D2H of the logits each step"},
        "e2e_other_input_formats": e2e_alt,
        "gpu_launches": int(eng.kernel_launches * args.steps),
        "roofline": roofline,
        "pct_int8_peak_whole_net": 100.0 * (value / world) * GOP_PER_IMG[args.arch] * 1e9 / (int8_peak * 1e12),
        "top1_agreement_vs_fp32": agree, "logits_rel_l2_vs_fp32": rel,
        "quantizer": {"ms_all_layers": quant_ms, "packed_bytes": packed_bytes},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            times, threads = cpu_reference_forward(args.arch, 32, 16)
            v = 32 * len(times) / sum(times)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d fp32 forwards of batch 32 (%.1f s) on the host CPU after one warm-up "
                                              "(stock torch ops = the reference's own arithmetic)" % (len(times), sum(times))}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": "failed: %s" % e}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
