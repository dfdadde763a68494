"""GPU tests of the rows SURVEY.md 8(f) ranks after the conv path:
  N1  fused evaluation tail (slq_eval_tail / slq_kl_rows) vs the reference's stock-torch sequence
      (functions.py:109-129 and the per-sample loop of KLdiv, functions.py:142-146)
  N2  device-resident snapshot / restore instead of torch.save / torch.load (resnet50_main.py:212, :233)
  N3  delta-loss table generator (the reference's dead evaluate_loss, functions.py:45-82)
  N4  packed on-disk model format
and the one-launch whole-model quantizer (slq_quantize_jobs) against the per-layer calls."""
import csv
import os

import numpy as np
import pytest
import torch

import slq_oracle as so
from helpers import build_p0_model, p0_table

pytestmark = pytest.mark.gpu


class _Passthrough(torch.nn.Module):
    def forward(self, x):
        return x


def _reference_eval(loader):
    """functions.py:84-129 + :131-149 restated with stock torch ops on the CPU (the reference's own sequence)."""
    crit = torch.nn.CrossEntropyLoss()
    ys, preds, outs, loss_sum = [], [], [], 0
    for x, y in loader:
        preds.append(x.max(1)[1])
        loss_sum = loss_sum + crit(x, y)
        outs.append(torch.softmax(x, dim=1))
        ys.append(y)
    ys, preds = torch.cat(ys), torch.cat(preds)
    acc = ((ys == preds).float().sum() / len(ys)).item()
    return acc, (loss_sum / len(loader)).item(), outs


def _reference_kl(n_out, out):
    kls = []
    for l in range(len(out)):
        for m in range(out[l].size()[0]):
            kls.append((n_out[l][m] * (n_out[l][m] / out[l][m]).log()).sum())
    return (sum(kls) / len(kls)).item()


def test_fused_eval_tail_matches_reference_sequence():
    import functions
    g = torch.Generator().manual_seed(3)
    loader = [(4 * torch.randn(b, 1000, generator=g), torch.randint(0, 1000, (b,), generator=g)) for b in (64, 64, 37)]
    for x, y in loader:  # make some predictions correct, and one exact tie (first index must win)
        x[torch.arange(0, x.shape[0], 3), y[::3]] += 30.0
    loader[0][0][1, 5] = loader[0][0][1, 900] = 99.0
    acc_r, loss_r, outs_r = _reference_eval(loader)
    net = _Passthrough()
    acc, loss, outs = functions.evaluate_acc_loss_softmax(net, "cuda", loader)
    assert acc == acc_r
    assert abs(loss - loss_r) <= 2e-6 * abs(loss_r)
    for a, b in zip(outs, outs_r):
        assert a.is_cuda and torch.allclose(a.cpu(), b, rtol=2e-6, atol=1e-12)
    # KL of a perturbed model against the stored outputs: fused in the evaluation and via KLdiv
    # (moderate logits: every softmax entry stays positive, so the reference's KL is finite)
    mild = [(1.5 * torch.randn(b, 1000, generator=g), torch.randint(0, 1000, (b,), generator=g)) for b in (64, 37)]
    mild2 = [(x + 0.05 * torch.randn(x.shape, generator=g), y) for x, y in mild]
    _, _, m_r = _reference_eval(mild)
    _, _, m2_r = _reference_eval(mild2)
    kl_r = _reference_kl(m_r, m2_r)
    assert np.isfinite(kl_r) and kl_r > 0
    _, _, m_outs = functions.evaluate_acc_loss_softmax(net, "cuda", mild)
    _, _, m2_outs, kl_fused = functions._evaluate(net, "cuda", mild2, ref_outputs=m_outs)
    kl_fn = functions.KLdiv(m_outs, m2_outs)
    assert abs(kl_fused - kl_r) <= 1e-4 * abs(kl_r) and abs(kl_fn - kl_r) <= 1e-4 * abs(kl_r)
    assert functions.KLdiv(m_outs, m_outs) == 0.0
    # quirk Q10: softmax underflow -> 0 * log(0 / 0) = NaN, p * log(p / 0) = inf, exactly like the reference
    big = [(300 * torch.randn(8, 1000, generator=g), torch.randint(0, 1000, (8,), generator=g))]
    _, _, ob_r = _reference_eval(big)
    _, _, ob = functions.evaluate_acc_loss_softmax(net, "cuda", big)
    big2 = [(big[0][0].flip(1).contiguous(), big[0][1])]
    _, _, ob2_r = _reference_eval(big2)
    _, _, ob2 = functions.evaluate_acc_loss_softmax(net, "cuda", big2)
    r, f = _reference_kl(ob_r, ob2_r), functions.KLdiv(ob, ob2)
    assert (np.isnan(r) and np.isnan(f)) or (np.isinf(r) and np.isinf(f)) or abs(r - f) <= 1e-4 * abs(r)


def test_resident_loader_cache_follows_the_host_tensors():
    """An in-memory loader is copied to the device once and reused by later evaluations (functions._resident_batches):
    the copies must follow in-place changes of the host tensors, a streaming loader must pass through untouched, and a
    module that already lives on the device must not be walked by net.to() again (same object, same parameters)."""
    import functions
    g = torch.Generator().manual_seed(11)
    loader = [(2 * torch.randn(16, 1000, generator=g), torch.randint(0, 1000, (16,), generator=g)) for _ in range(2)]
    net = _Passthrough()
    _, loss0, outs0 = functions.evaluate_acc_loss_softmax(net, "cuda", loader)
    first = functions._RESIDENT["batches"]
    _, loss1, _ = functions.evaluate_acc_loss_softmax(net, "cuda", loader)
    assert functions._RESIDENT["batches"] is first and loss1 == loss0  # reused, same result
    loader[0][0].mul_(0.5)  # in-place change of a host batch: the version counter moves, the copy is refreshed
    _, loss2, _ = functions.evaluate_acc_loss_softmax(net, "cuda", loader)
    _, loss2_r, _ = _reference_eval(loader)
    assert functions._RESIDENT["batches"] is not first
    assert loss2 != loss0 and abs(loss2 - loss2_r) <= 2e-6 * abs(loss2_r)
    gen = ((x, y) for x, y in loader)  # a streaming loader (anything that is not a list / tuple) is not cached
    assert functions._resident_batches(gen, "cuda") is gen
    lin = torch.nn.Linear(8, 8).cuda()
    w_ptr = lin.weight.data_ptr()
    assert functions._to_device(lin, "cuda") is lin and lin.weight.data_ptr() == w_ptr
    cpu_lin = torch.nn.Linear(8, 8)
    assert functions._to_device(cpu_lin, "cuda").weight.is_cuda


def test_whole_model_quantizer_one_launch_equals_per_layer_calls():
    import functions
    import resnet
    import slq_lib as L
    table = p0_table("resnet50")
    nets = []
    for mode in ("per_layer", "one_launch"):
        torch.manual_seed(0)
        net = resnet.resnet50(num_classes=1000).cuda().eval()
        convs = dict(functions.quantized_convs("resnet50", net))
        items = [(convs[int(l)].weight.data, table[table[:, 0] == l][:, 1], table[table[:, 0] == l][:, 2])
                 for l in np.unique(table[:, 0])]
        if mode == "per_layer":
            prs = [functions.quantize_rows(t, r, b, div_mode=L.DIV_TRUE) for t, r, b in items]
            z = np.concatenate([p.z.cpu().numpy() for p in prs])
            s = np.concatenate([p.s32.cpu().numpy() for p in prs])
            nbytes = sum(int(p.blob.numel()) for p in prs)
        else:
            pm = functions.quantize_model(items, div_mode=L.DIV_TRUE)
            inv = np.empty_like(pm.perm)
            inv[pm.perm] = np.arange(pm.perm.size)       # caller order -> launch order
            zm, sm = pm.z.cpu().numpy()[inv], pm.s32.cpu().numpy()[inv]
            assert int((pm.status.cpu() != 0).sum()) == 0
            assert pm.nbytes == nbytes
        nets.append(net)
    assert np.array_equal(z, zm) and np.array_equal(s.view(np.uint32), sm.view(np.uint32))
    for (ka, a), (_kb, b) in zip(nets[0].state_dict().items(), nets[1].state_dict().items()):
        assert torch.equal(a, b), ka
    # a constant row is reported once, at the end, like the reference's ZeroDivisionError
    w = torch.ones(8, 64, device="cuda")
    with pytest.raises(ZeroDivisionError):
        functions.quantize_model([(w, [0, 1], [8, 4])])


def test_snapshot_restore_and_packed_file(tmp_path):
    import functions
    import resnet
    import slq_store
    net = build_p0_model("resnet18", "cuda")
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        l0 = net(x).clone()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    snap = functions.snapshot(net)
    assert len(snap.packed_names()) == 16             # every quantised block conv is held as packed codes
    assert snap.nbytes < 0.45 * snap.fp32_bytes
    # wreck the model: 2-bit a layer, rescale a BatchNorm
    w = net.layer3[1].conv2.weight
    functions.quantize_rows(w.data, np.arange(w.shape[0]), np.full(w.shape[0], 2))
    with torch.no_grad():
        net.layer1[0].bn1.weight.mul_(1.5)
        l1 = net(x).clone()
    assert not torch.equal(l0, l1)
    functions.restore(net, snap)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k
    with torch.no_grad():
        assert torch.equal(net(x), l0)
    # packed file -> a different model object
    path = str(tmp_path / "r18.slqpack")
    nbytes = functions.save_packed(net, path)
    assert nbytes == os.path.getsize(path) and nbytes < 0.45 * snap.fp32_bytes
    torch.manual_seed(123)
    other = resnet.resnet18(num_classes=1000).cuda().eval()
    functions.load_packed(path, other)
    for k, v in other.state_dict().items():
        assert torch.equal(v, before[k]), k
    with torch.no_grad():
        other.slq_share_calibration(net)
        assert torch.equal(other(x), l0)
    raw = open(path, "rb").read()
    assert raw[:8] == slq_store.MAGIC
    # a model that still has fp32 (never-quantised) block convs keeps those layers dense
    torch.manual_seed(0)
    fresh = resnet.resnet18(num_classes=1000).cuda().eval()
    with torch.no_grad():
        fresh(x)
    assert functions.snapshot(fresh).packed_names() == []


def test_deltaloss_table_generator(tmp_path):
    """Per-channel delta-loss (N3) against the reference's arithmetic: oracle quantizer + fp32 torch forward."""
    import functions
    import imagenet
    import resnet
    torch.manual_seed(0)
    net = resnet.resnet18(num_classes=1000).cuda().eval()
    loader = imagenet.synthetic_loader(1, 8, 64, seed=2)
    chans = lambda lnum, cout: (0, 7, 33)
    before = net.layer1[0].conv1.weight.detach().clone()
    lnums, cnums, table = functions.make_deltaloss_table("resnet18", net, "cuda", loader, bits=(8, 2), layers={1, 4},
                                                         channels=chans)
    assert lnums == [1, 1, 1, 4, 4, 4] and cnums == [1, 8, 34] * 2
    assert torch.equal(net.layer1[0].conv1.weight.detach(), before)          # every row restored
    # reference arithmetic on the CPU
    crit = torch.nn.CrossEntropyLoss()
    x, y = loader[0]
    cpu = resnet.resnet18(num_classes=1000).eval()
    cpu.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    base = crit(so.torch_forward(cpu, x), y).item()
    convs = dict(functions.quantized_convs("resnet18", cpu))
    for j, (lnum, c1) in enumerate(zip(lnums, cnums)):
        for b in (8, 2):
            w2 = convs[lnum].weight.data.reshape(convs[lnum].out_channels, -1)
            keep = w2[c1 - 1].clone()
            so.channel_wise(w2.numpy(), b, c1 - 1)
            ref = crit(so.torch_forward(cpu, x), y).item() - base
            w2[c1 - 1].copy_(keep)
            got = table[b][j]
            print("layer %d channel %d %d bit: delta-loss %.4e (reference %.4e)" % (lnum, c1, b, got, ref))
            assert abs(got - ref) <= 0.35 * abs(ref) + 2e-3 * abs(base)
    path = str(tmp_path / "dl.csv")
    functions.write_deltaloss_csv(path, lnums, cnums, table, bits=(8, 2))
    with open(path, encoding="utf-8-sig") as f:  # parsed the way resnet50_main.py:59-79 does
        rows = list(csv.reader(f))
    assert [int(v) for v in rows[0]] == lnums and [int(v) for v in rows[1]] == cnums
    assert [float(v) for v in rows[2]] == table[8] and [float(v) for v in rows[3]] == table[2]
