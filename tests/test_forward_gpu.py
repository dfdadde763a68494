"""Whole-network parity on the GPU: logits of the B200 engine vs the reference's fp32 fake-quant
forward (golden fixtures generated from the unmodified reference; torch fp32 restatement for
larger batches).  Tolerance (stated): relative L2 error of the logits <= 1e-2."""
import os

import numpy as np
import pytest
import torch

import slq_oracle as so
from helpers import GOLD, build_p0_model, rel_l2

pytestmark = pytest.mark.gpu

LOGITS_REL_L2_TOL = 1e-2


def _golden_x():
    g = torch.Generator().manual_seed(1)
    return torch.randn(2, 3, 224, 224, generator=g)


@pytest.mark.parametrize("arch", ["resnet18", "resnet34", "resnet50"])
def test_logits_vs_reference_golden(arch):
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % arch))
    net = build_p0_model(arch, "cuda")
    x = _golden_x().cuda()
    with torch.no_grad():
        logits = net(x).cpu().numpy()
    err = rel_l2(logits, gold["logits_p0"])
    print(arch, "rel-L2 vs reference fp32 fake-quant logits:", err, "argmax", logits.argmax(1), gold["logits_p0"].argmax(1))
    assert err <= LOGITS_REL_L2_TOL
    assert (logits.argmax(1) == gold["logits_p0"].argmax(1)).all()
    eng = next(iter(net._slq_engines.values()))
    # every quantised block conv is in one-limb mode with only 4/8-bit rows; downsample convs fp32
    for op in eng.ops:
        bits = set(np.unique(op.bits_host).tolist())
        if op.signed:
            assert bits == {16} and op.w16 == 1
        else:
            assert bits <= {4, 8} and op.w16 == 0


def test_simt_and_umma_engines_produce_identical_bytes():
    """Same model through the tcgen05 path and through the dp4a checker: every activation tensor
    and the logits are byte-identical (same integer accumulators, same epilogue arithmetic)."""
    import slq_engine
    import slq_lib as L
    net = build_p0_model("resnet18", "cuda")
    x = _golden_x().cuda()
    outs = []
    for impl in (L.IMPL_UMMA, L.IMPL_SIMT):
        eng = slq_engine.Engine(net, 2, 224, 224, x.device, impl=impl)
        eng.refresh_weights()
        eng.calibrate(x)
        calib_bytes = [a.clone() for a in eng.act]
        logits = eng.forward(x).clone()
        # the static-scale pass reproduces the bytes the calibration pass left behind
        for a, b in zip(calib_bytes, eng.act):
            assert torch.equal(a, b)
        outs.append((logits, [a.clone() for a in eng.act], eng.act_scales.clone()))
    assert torch.equal(outs[0][2], outs[1][2])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)
    assert torch.equal(outs[0][0], outs[1][0])


def test_activation_pipeline_matches_numpy_oracle_first_block():
    """First bottleneck conv of ResNet-50 checked against the numpy restatement end to end
    (stem output bytes -> conv1 u8 output bytes)."""
    import slq_engine
    net = build_p0_model("resnet50", "cuda")
    x = _golden_x().cuda()
    eng = slq_engine.Engine(net, 2, 224, 224, x.device)
    eng.refresh_weights()
    eng.calibrate(x)
    eng.forward(x)
    op = eng.ops[0]
    xin = eng.act[op.in_id].cpu().numpy()
    scales = eng.act_scales.cpu().numpy()
    w = op.conv.weight.detach().cpu().numpy().reshape(op.Cout, -1)
    meta = [so.encode_row(w[oc]) for oc in range(op.Cout)]
    codes = np.stack([m[1].reshape(op.Cin, op.k, op.k).transpose(1, 2, 0) for m in meta]).astype(np.int64)
    acc, S, _, _ = so.conv_acc(xin, codes, op.stride, op.pad)
    want = so.epilogue_q(acc, S, op.zf.cpu().numpy(), op.wscale.cpu().numpy(), op.bias.cpu().numpy(),
                         scales[op.in_id], scales[op.out_id]).reshape(eng.act[op.out_id].shape)
    assert np.array_equal(eng.act[op.out_id].cpu().numpy(), want)
    # stem: fp32 conv + bn + relu + maxpool against torch, then the same quantiser
    import torch.nn.functional as F
    with torch.no_grad():
        t = F.conv2d(x, net.conv1.weight, None, 2, 3)
        t = F.relu(F.batch_norm(t, net.bn1.running_mean, net.bn1.running_var, net.bn1.weight, net.bn1.bias, False, 0.0, net.bn1.eps))
        t = F.max_pool2d(t, 3, 2, 1).permute(0, 2, 3, 1).contiguous().cpu().numpy()
    s0 = scales[0]
    # the stem runs fp16 operands on the tensor core (fp32 accumulation): ~3e-4 relative
    assert abs(float(t.max()) / 255.0 - float(s0)) <= 1e-3 * float(s0) + 1e-9
    q = so.requant_u8(t, s0)
    got = eng.act[0].cpu().numpy().astype(np.int32)
    assert np.abs(got - q.astype(np.int32)).max() <= 1  # fp16 tensor-core operands / summation order
    assert (got != q).mean() < 0.05


def test_batch_32_top1_agreement_and_dropin_evaluate():
    """functions.evaluate_acc_loss_softmax drives the engine exactly like the mains do; the
    predictions agree with the fp32 torch restatement of the reference forward."""
    import functions
    import imagenet
    net = build_p0_model("resnet18", "cuda")
    loader = imagenet.synthetic_loader(2, 32, 224, seed=3) + imagenet.synthetic_loader(1, 5, 224, seed=4)
    acc, loss, outs = functions.evaluate_acc_loss_softmax(net, "cuda", loader)
    assert len(outs) == 3 and outs[2].shape == (5, 1000)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    agree, total, errs = 0, 0, []
    for (x, _y), p in zip(loader, outs):
        ref = so.torch_forward(net, x.cuda())
        agree += int((ref.argmax(1) == p.argmax(1)).sum())
        total += x.shape[0]
        errs.append(rel_l2(torch.log(p).cpu().numpy() - torch.log(p).cpu().numpy().mean(1, keepdims=True),
                           torch.log_softmax(ref, 1).cpu().numpy() - torch.log_softmax(ref, 1).cpu().numpy().mean(1, keepdims=True)))
    print("top-1 agreement %d/%d, centred log-prob rel-L2 %s" % (agree, total, errs))
    assert agree >= total - 1
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0


def test_weight_mutation_is_seen_by_the_next_forward():
    """the mains mutate conv.weight.data in place and reload state_dicts."""
    import functions
    import resnet
    torch.manual_seed(0)
    net = resnet.resnet18(num_classes=1000).cuda().eval()
    x = _golden_x().cuda()
    with torch.no_grad():
        l0 = net(x).clone()
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        w = net.layer1[0].conv1.weight
        for c in range(w.shape[0]):
            w.data = functions.channel_wise_quantizationperchan(w.data, 2, c)
        l1 = net(x).clone()
        assert not torch.equal(l0, l1)
        eng = next(iter(net._slq_engines.values()))
        assert set(np.unique(eng.ops[0].bits_host).tolist()) <= {2}
        net.load_state_dict(sd)
        l2 = net(x).clone()
        assert torch.equal(l0, l2)


def test_tensor_core_stem_matches_exact_fp32_stem():
    """The tcgen05 stem (fp16 operands, fp32 accumulate) against the exact-fp32 CUDA-core stem:
    same static scale to 1e-3, u8 activations within one level, fp32 pre-quantisation values to 2e-3."""
    import slq_engine
    import slq_lib as L
    net = build_p0_model("resnet18", "cuda")
    for n in (2, 5):
        g = torch.Generator().manual_seed(n)
        x = torch.randn(n, 3, 224, 224, generator=g).cuda()
        res = {}
        for kind in ("umma", "simt"):
            eng = slq_engine.Engine(net, n, 224, 224, x.device, stem=kind)
            eng.refresh_weights()
            assert (eng.stem is not None) == (kind == "umma")
            eng._stem(x.data_ptr(), eng.f32_scratch.data_ptr(), L.OUT_F32, L.current_stream())
            torch.cuda.synchronize()
            f32 = eng.f32_scratch[:eng.act[0].numel()].clone()
            eng.calibrate(x)
            eng.forward(x)
            res[kind] = (f32, eng.act[0].clone(), float(eng.act_scales[0]))
        fa, fb = res["umma"][0], res["simt"][0]
        assert float((fa - fb).abs().max()) <= 2e-3 * float(fb.abs().max())
        assert abs(res["umma"][2] - res["simt"][2]) <= 1e-3 * res["simt"][2]
        d = (res["umma"][1].int() - res["simt"][1].int()).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.05


def test_u8_and_fp16_inputs_give_the_fp32_logits_bit_for_bit():
    """The loader-side data formats (reference imagenet.py:14-40): raw u8 pixels normalised inside the stem
    kernel == the reference's CPU ToTensor + Normalize followed by the fp32 call, and the fp32 batch rounded
    to fp16 == the fp32 call (the stem rounds its operands to fp16 anyway): identical logits, bit for bit."""
    import imagenet
    net = build_p0_model("resnet18", "cuda")
    net.input_norm = (imagenet.MEAN, imagenet.STD)
    (x_u8, _y), = imagenet.synthetic_loader(1, 5, 224, seed=7, dtype=torch.uint8)
    x32 = imagenet.normalize_u8(x_u8)  # CPU, like the reference's DataLoader workers
    ref = net(x32.cuda()).cpu().numpy()  # calibrates the engine on this batch
    got_u8 = net(x_u8.cuda()).cpu().numpy()
    got_f16 = net(x32.half().cuda()).cpu().numpy()
    assert np.isfinite(ref).all() and np.abs(ref).max() > 0
    assert np.array_equal(got_u8.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(got_f16.view(np.uint32), ref.view(np.uint32))
    # and a u8 batch without the normalisation constants is an error, not a guess
    del net.input_norm
    with pytest.raises(ValueError):
        net(x_u8.cuda())


@pytest.mark.parametrize("N,HW,C,O", [(256, 49, 2048, 1000), (5, 49, 512, 1000), (130, 4, 512, 10)])
def test_tensor_core_tail_matches_fp32(N, HW, C, O):
    """slq_tail_forward (avg-pool + split-K TF32 GEMM on tcgen05 + fixed-order reduction + bias) against the fp32
    restatement of resnet.py:216-218; both operands enter as two TF32 terms (hi + lo, three products): fp32-class
    accuracy, rel-L2 <= 2e-5, deterministic."""
    import slq_lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(N + C)
    x = torch.randint(0, 256, (N, HW, C), generator=g, dtype=torch.uint8).cuda()
    w = (torch.randn(O, C, generator=g) * 0.02).cuda()
    b = torch.randn(O, generator=g).cuda()
    scales = torch.tensor([0.5, 0.0123], device="cuda")
    ws = torch.empty(lib.slq_tail_workspace_bytes(N, C, O) // 4, dtype=torch.float32, device="cuda")
    w2 = torch.empty((2, O, C), dtype=torch.float32, device="cuda")
    L.check(lib.slq_tail_split_weights(w.data_ptr(), O, C, w2.data_ptr(), L.current_stream()))
    assert torch.equal(w2[0] + w2[1], w) and torch.equal(w2[0].view(torch.int32) & 0x1fff, torch.zeros_like(w2[0], dtype=torch.int32))
    outs = []
    for _ in range(2):
        logits = torch.full((N, O), float("nan"), device="cuda")
        L.check(lib.slq_tail_forward(x.data_ptr(), N, HW, C, scales.data_ptr(), 1, w2.data_ptr(), b.data_ptr(), O,
                                     ws.data_ptr(), logits.data_ptr(), L.current_stream()))
        torch.cuda.synchronize()
        outs.append(logits)
    assert torch.equal(outs[0], outs[1])
    torch.backends.cuda.matmul.allow_tf32 = False
    pooled = x.float().mean(1) * 0.0123
    ref = pooled @ w.t() + b
    err = rel_l2(outs[0].cpu().numpy(), ref.cpu().numpy())
    print("tail N=%d C=%d O=%d: rel-L2 %.3e" % (N, C, O, err))
    assert err <= 2e-5
