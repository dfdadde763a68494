"""tools/sweep_run.py -- BASELINE config 4: the per-semilayer sensitivity sweep
(reference functions.py:456-588 make_semilayers_resnet50) sharded over the ranks of a torchrun job.

    python tools/sweep_run.py --arch resnet50 --batches 2 --batch 64
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29512 tools/sweep_run.py --arch resnet50

Every rank builds the same seeded model and synthetic loader, evaluates its share of the candidate
semilayers on its own GPU (quantizer kernel + tcgen05 forward) and the per-candidate values are
exchanged with ONE all_gather (NCCL).  Rank 0 prints a JSON line with the ranked order, a digest of
the gathered values and the wall time.  The metric is the reference's KL/param for resnet18 and
delta-loss for the deeper nets (KL is NaN on random-init R34/R50)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
sys.path.insert(0, PKG)


def rows_from_p0(arch):
    """listminus / listplus in the reference's 8-column row format (resnet50_main.py:152) from the
    committed P0 table: 4-bit channels are the 'minus' semilayer of their layer, 8-bit the 'plus'."""
    table = np.load(os.path.join(PKG, "data", "p0_bits.npz"))[arch]
    cpb = 3 if arch == "resnet50" else 2
    depth = [2, 2, 2, 2] if arch == "resnet18" else [3, 4, 6, 3]
    starts = np.cumsum([0] + depth)
    minus, plus = [], []
    for gi, (lnum, cn, bit) in enumerate(table, 1):
        blk = (int(lnum) - 1) // cpb
        li = int(np.searchsorted(starts, blk, side="right") - 1)
        bi = blk - int(starts[li])
        row = [li, bi, int(lnum), int(cn), 8, 0, 32, gi]
        (minus if bit == 4 else plus).append(row)
    for lst, sign in ((minus, -1), (plus, 1)):  # per-layer flags like make_divide_minusplusmodels
        for r in lst:
            r[5] = -(r[2] - 1) if sign < 0 else r[2]
    return minus, plus


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--layers", type=int, default=0, help="only the first N layers (0 = all)")
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--trace-evals", action="store_true", help="rank 0 logs the wall time of every evaluation to stderr")
    ap.add_argument("--force-work-model", action="store_true",
                    help="candidates 1.. on a second, freshly built pretrained model even when the caller's net is the "
                         "pretrained one (the general path; tests compare it with the shortcut)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if args.backend == "nccl":
        torch.cuda.set_device(local)
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        if args.backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group(args.backend)
    import slq_build
    slq_build.build()
    import functions
    import imagenet
    import resnet
    imagenet.val_loader = imagenet.synthetic_loader(args.batches, args.batch, args.hw, seed=1)
    torch.manual_seed(0)
    sd = getattr(resnet, args.arch)(num_classes=1000).state_dict()
    resnet.load_state_dict_from_url = lambda url, progress=True: sd  # 'pretrained' = seeded random init
    net2 = getattr(resnet, args.arch)(num_classes=1000, pretrained="imagenet")
    metric = "kl" if args.arch == "resnet18" else "dloss"
    minus, plus = rows_from_p0(args.arch)
    if args.layers:
        minus = [r for r in minus if r[2] <= args.layers]
        plus = [r for r in plus if r[2] <= args.layers]
    if args.force_work_model:
        functions._is_pretrained = lambda *a, **k: False
    if args.trace_evals and rank == 0:  # wall time of every evaluation (and of the gaps between them) on rank 0
        inner = functions._evaluate
        last = [time.perf_counter()]

        def timed(*a, **k):
            t_in = time.perf_counter()
            out = inner(*a, **k)
            torch.cuda.synchronize()
            t_out = time.perf_counter()
            sys.stderr.write("eval %.1f ms (gap before it %.1f ms)\n" % (1e3 * (t_out - t_in), 1e3 * (t_in - last[0])))
            last[0] = t_out
            return out
        functions._evaluate = timed
    if world > 1:  # communicator set-up (lazy, on the first collective) is not part of the sweep
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, orig = functions.evaluate_acc_loss_softmax(net2, dev, imagenet.val_loader)
    devnull = open(os.devnull, "w")
    stdout = sys.stdout
    sys.stdout = devnull  # the reference prints one line per candidate
    try:
        semilayers, orders = functions.make_semilayers(args.arch, net2, dev, orig, minus, plus, metric=metric)
        flat = functions.make_quantizedlists(semilayers, [list(o) for o in orders])
    finally:
        sys.stdout = stdout
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    vals = np.array([o[1] for o in orders], np.float64)
    # what the sweep left in the caller's net (quirk Q3: candidate 0's layer quantised, everything else untouched)
    h = hashlib.sha256()
    for k, v in sorted(net2.state_dict().items()):
        if k.endswith("weight") and v.dim() == 4:
            h.update(v.detach().cpu().numpy().tobytes())
    if rank == 0:
        ranked = [int(i) for i in np.argsort(vals, kind="stable")]
        print(json.dumps({
            "workload": "%s semilayer sensitivity sweep, %d candidates, %d x %d synthetic %dx%d images" % (
                args.arch, len(semilayers), args.batches, args.batch, args.hw, args.hw),
            "metric": metric, "world": world, "backend": args.backend if world > 1 else "none",
            "seconds": dt, "candidates_per_s": len(semilayers) / dt,
            "values_sha256": hashlib.sha256(vals.tobytes()).hexdigest(), "net_after_sha256": h.hexdigest(),
            "ranked_first8": ranked[:8],
            "flat_rows": len(flat), "values_first4": [float(v) for v in vals[:4]]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
