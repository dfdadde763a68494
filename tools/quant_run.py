"""tools/quant_run.py -- the whole-model quantizer pass alone (for ncu): ResNet-50, P0 8/4-bit assignment,
ONE slq_quantize_jobs launch per repetition, L2 flushed in between.  Prints device time per launch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
sys.path.insert(0, PKG)
import slq_build  # noqa: E402

slq_build.build()
import functions  # noqa: E402
import resnet  # noqa: E402
import slq_lib as L  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
net = getattr(resnet, arch)(num_classes=1000).cuda().eval()
table = np.load(os.path.join(PKG, "data", "p0_bits.npz"))[arch]
convs = dict(functions.quantized_convs(arch, net))
items = [(convs[int(l)].weight.data, table[table[:, 0] == l][:, 1], table[table[:, 0] == l][:, 2]) for l in np.unique(table[:, 0])]
fresh = [t.clone() for t, _r, _b in items]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
plan = functions.QuantPlan(items, div_mode=L.DIV_TRUE)
for rep in range(reps):
    for (t, _r, _b), f in zip(items, fresh):
        t.copy_(f)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pm = plan.run()
    e1.record()
    torch.cuda.synchronize()
    n_w = sum(len(r) * t[0].numel() for t, r, _b in items)
    byts = 8 * n_w + pm.nbytes + 12 * len(table)
    print("rep %d: %.1f us, %.0f GB/s algorithmic (%.1f MB)" % (rep, 1e3 * e0.elapsed_time(e1), byts / e0.elapsed_time(e1) / 1e6, byts / 1e6))
