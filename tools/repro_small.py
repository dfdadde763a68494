"""tools/repro_small.py -- smallest end-to-end case for compute-sanitizer: ResNet-50 (P0 8/4-bit), batch 4, 64x64
images: calibrate + forward, repeated forwards must be bit-identical, snapshot -> wreck -> restore -> same logits."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("semilayer-wise-mixed-precision-quantization_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import functions  # noqa: E402
from helpers import build_p0_model  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
net = build_p0_model(arch, "cuda")
x = torch.randn(4, 3, hw, hw, generator=torch.Generator().manual_seed(1)).cuda()
with torch.no_grad():
    l0 = net(x).clone()
    for i in range(3):
        li = net(x).clone()
        print("repeat", i, "identical:", bool(torch.equal(li, l0)), float((li - l0).abs().max()))
    snap = functions.snapshot(net)
    w = net.layer3[1].conv2.weight
    functions.quantize_rows(w.data, np.arange(w.shape[0]), np.full(w.shape[0], 2))
    net.layer1[0].bn1.weight.mul_(1.5)
    l1 = net(x).clone()
    functions.restore(net, snap)
    l2 = net(x).clone()
    print("after restore identical:", bool(torch.equal(l2, l0)), float((l2 - l0).abs().max()), "wrecked differs:", not torch.equal(l1, l0))
    eng = next(iter(net._slq_engines.values()))
    acts = [a.clone() for a in eng.act]
    net(x)
    bad = [i for i, (a, b) in enumerate(zip(acts, eng.act)) if not torch.equal(a, b)]
    print("activation tensors that differ between two identical forwards:", bad[:10])
torch.cuda.synchronize()
print("done")
