"""Logits parity on BASELINE.json's own configurations, with HELD-OUT calibration:

  * configs[2]  ResNet-50 semilayer 8/4-bit (P0), batch 256          (the headline configuration)
  * configs[1]  ResNet-34 semilayer 8/4-bit (P0), batch 128
  * both again with RANDOMISED BatchNorm (gamma, beta, running_mean, running_var), so that the BN fold of
    the epilogue (slq_engine.Engine._fold_bn; reference resnet.py:58,61,100,104,108 eval-mode BN) is
    exercised end to end with non-identity statistics.

The static u8 activation scales are calibrated on seeded batch A through the explicit API
(``net.slq_calibrate``); batches B and C, which the engine has never seen, are then compared with the
reference's fp32 fake-quant forward (oracle ``torch_forward`` = reference resnet.py:204-220 on stock torch
ops, cuDNN fp32 with TF32 off).  Stated tolerance: rel-L2 of the logits <= 1e-2.
Top-1 agreement is reported and bounded; every disagreement must be a near-tie of the reference itself.
"""
import numpy as np
import pytest
import torch

import slq_oracle as so
from helpers import build_p0_model, rel_l2

pytestmark = pytest.mark.gpu

LOGITS_REL_L2_TOL = 1e-2
TOP1_AGREEMENT_MIN = 0.90   # random-init logits are nearly flat: flips are near-ties (checked below)


def _batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, 224, 224, generator=g)


def _randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                c = m.num_features
                m.weight.copy_(0.5 + torch.rand(c, generator=g))
                m.bias.copy_(0.2 * torch.randn(c, generator=g))
                m.running_mean.copy_(0.2 * torch.randn(c, generator=g))
                m.running_var.copy_(0.5 + torch.rand(c, generator=g))


def _check_heldout(net, batch, seeds, tag):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net.slq_calibrate(_batch(batch, seeds[0]).cuda())
    eng = next(iter(net._slq_engines.values()))
    scales = eng.act_scales.clone()
    worst, agree, total = 0.0, 0, 0
    for s in seeds[1:]:
        x = _batch(batch, s).cuda()
        with torch.no_grad():
            got = net(x)
        ref = so.torch_forward(net, x)
        err = rel_l2(got.cpu().numpy(), ref.cpu().numpy())
        worst = max(worst, err)
        a, r = got.argmax(1), ref.argmax(1)
        agree += int((a == r).sum())
        total += batch
        # a flip needs the reference's own top-1 margin to be below twice the largest logit error of that image
        flipped = (a != r).nonzero().flatten()
        if len(flipped):
            top2 = ref[flipped].topk(2, dim=1).values
            margin = (top2[:, 0] - top2[:, 1])
            dmax = (got[flipped] - ref[flipped]).abs().max(1).values
            assert bool((margin <= 2 * dmax + 1e-6).all())
            rel_margin = (margin / ref[flipped].abs().max(1).values).max().item()
            print("%s seed %d: %d flips, largest relative top-1 margin among them %.3e" % (tag, s, len(flipped), rel_margin))
            assert rel_margin < 2e-2
        print("%s seed %d: rel-L2 %.3e, top-1 agreement %d/%d" % (tag, s, err, int((a == r).sum()), batch))
    assert torch.equal(scales, eng.act_scales)  # held-out batches did not re-calibrate
    assert worst <= LOGITS_REL_L2_TOL, worst
    assert agree >= TOP1_AGREEMENT_MIN * total, (agree, total)
    return worst, agree / total


def test_resnet50_batch256_heldout_calibration():
    net = build_p0_model("resnet50", "cuda")
    _check_heldout(net, 256, (11, 12, 13), "R50@256")


def test_resnet34_batch128_heldout_calibration():
    net = build_p0_model("resnet34", "cuda")
    _check_heldout(net, 128, (21, 22, 23), "R34@128")


@pytest.mark.parametrize("arch,batch", [("resnet18", 32), ("resnet50", 32)])
def test_randomised_batchnorm_heldout(arch, batch):
    net = build_p0_model(arch, "cuda")
    _randomise_bn(net, 5)
    _check_heldout(net, batch, (31, 32), "%s random-BN" % arch)
    # changing only BN statistics afterwards is seen by the next forward (no weight re-pack needed)
    x = _batch(batch, 33).cuda()
    with torch.no_grad():
        before = net(x).clone()
        net.layer1[0].bn1.running_mean.add_(0.5)
        after = net(x).clone()
    assert not torch.equal(before, after)
    ref = so.torch_forward(net, x)
    assert rel_l2(after.cpu().numpy(), ref.cpu().numpy()) <= LOGITS_REL_L2_TOL


def test_calibration_api_semantics():
    """Scales are set by slq_calibrate (running abs-max over the batches given, optional head-room) and are
    NOT touched by weight changes, state_dict reloads or later batches; the first forward of an
    un-calibrated net calibrates on its own batch."""
    import functions
    net = build_p0_model("resnet18", "cuda")
    a, b = _batch(4, 41).cuda(), _batch(4, 42).cuda()
    with torch.no_grad():
        net(a)                                   # implicit calibration on a
    eng = next(iter(net._slq_engines.values()))
    s_a = eng.act_scales.clone()
    net.slq_calibrate([a, b])
    s_ab = eng.act_scales.clone()
    assert bool((s_ab >= s_a * (1 - 1e-6)).all()) and not torch.equal(s_ab, s_a)
    net.slq_calibrate([(a, None)], headroom=1.25)  # loader-style items
    assert torch.allclose(eng.act_scales[:1], s_a[:1] * 1.25, rtol=1e-6)  # stem output: same input, same weights
    net.slq_calibrate(a)
    assert torch.equal(eng.act_scales, s_a)
    w = net.layer2[1].conv1.weight
    w.data = functions.channel_wise_quantizationperchan(w.data, 2, 3)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        l1 = net(b).clone()
        net.load_state_dict(sd)
        l2 = net(b).clone()
    assert torch.equal(eng.act_scales, s_a) and torch.equal(l1, l2)
    assert eng is next(iter(net._slq_engines.values()))
