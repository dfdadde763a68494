"""functions.py -- drop-in mirror of the reference's ``functions.py`` on top of the B200 path.

Same ten module-level entry points, same argument meaning, return values and error behaviour
(SURVEY.md 8b), so the unmodified ``resnet18/34/50_main.py`` run against this module:

  channel_wise_quantizationperchan / quantize_wgt   reference functions.py:9-43   -> slq_quantize_rows
  evaluate_loss / evaluate_acc_loss_softmax / KLdiv reference functions.py:45-149 (callers of net(x))
  make_divide_minusplusmodels                       reference functions.py:151-184
  make_semilayers_resnet18/34/50                    reference functions.py:186-588 (one shared body,
                                                    candidates sharded over torch.distributed ranks)
  make_quantizedlists                               reference functions.py:590-612

Extensions (not in the reference): ``quantize_rows`` (a whole semilayer / layer in one launch,
returns the packed codes) and ``make_semilayers`` (arch-generic sweep, optional delta-loss metric).
"""
import os

import numpy as np
import torch
import torch.nn

import imagenet
import resnet
import slq_lib as L

SENTINEL = [0, 0, 100, 0, 0, 0, 0, 0]


# ==============================================================================================
# quantizer
# ==============================================================================================
class PackedRows:
    """Result of quantize_rows: job j -> codes blob[offsets[j]:...], z[j], s32[j], status[j]."""

    def __init__(self, rows, bits, offsets, blob, z, s32, status, K):
        self.rows, self.bits, self.offsets, self.blob = rows, bits, offsets, blob
        self.z, self.s32, self.status, self.K = z, s32, status, K

    def codes(self, j):
        """Unpacked codes of job j as a uint8/int32 numpy array (host copy; for inspection/tests)."""
        bit = int(self.bits[j])
        nbytes = L.lib().slq_packed_row_bytes(self.K, bit)
        off = int(self.offsets[j])
        raw = self.blob[off:off + nbytes].cpu().numpy().astype(np.int64)
        if bit == 4:
            return np.stack([raw & 15, raw >> 4], 1).reshape(-1)[:self.K].astype(np.int32)
        if bit == 2:
            return np.stack([(raw >> (2 * i)) & 3 for i in range(4)], 1).reshape(-1)[:self.K].astype(np.int32)
        return raw[:self.K].astype(np.int32)


def packed_row_bytes(K, bits):
    """slq_packed_row_bytes (include/slq.h) for an array of bit-widths <= 8 at once -- a ctypes call per
    row would dominate a whole-model pass: 4-bit rows pack two codes per byte, 2-bit rows four, every other
    width one."""
    b64 = np.asarray(bits, dtype=np.int64)
    return np.where(b64 == 4, (K + 1) // 2, np.where(b64 == 2, (K + 3) // 4, K)).astype(np.int64)


def _div_mode_for(tensor, div_mode):
    if div_mode is not None:
        return div_mode
    forced = os.environ.get("SLQ_DIV_MODE")  # "true" | "recip": pin one flavour for a whole run (tests)
    if forced:
        return {"true": L.DIV_TRUE, "recip": L.DIV_RECIP}[forced]
    # what the reference's own ATen call does on that device
    return L.DIV_RECIP if tensor.is_cuda else L.DIV_TRUE


def quantize_rows(tensor, rows, bits, write_back=True, want_codes=True, div_mode=None):
    """Quantises rows ``rows[j]`` of a contiguous fp32 weight tensor ``[Cout, ...]`` to ``bits[j]``
    bits in ONE kernel launch (reference: one quantize_wgt call = 7 launches + 2 syncs per row).
    In place when write_back (the contract of functions.py:22).  Raises ZeroDivisionError for a
    constant row like the reference (functions.py:40).  Returns PackedRows."""
    lib = L.lib()
    if tensor.dtype != torch.float32 or not tensor.is_contiguous():
        raise TypeError("quantize_rows needs a contiguous float32 tensor")
    n_rows = tensor.shape[0]
    K = tensor[0].numel()
    rows_h = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
    bits_h = np.ascontiguousarray(bits, dtype=np.int32).reshape(-1)
    if rows_h.size != bits_h.size:
        raise ValueError("rows and bits must have the same length")
    if rows_h.size and (rows_h.min() < 0 or rows_h.max() >= n_rows):
        raise IndexError("row index out of range")
    if bits_h.size and (bits_h.min() < 1 or bits_h.max() > 8):
        raise ValueError("bit-width must be in 1..8")
    n_jobs = rows_h.size
    sizes = (packed_row_bytes(K, bits_h) + 15) // 16 * 16
    offs_h = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if n_jobs else np.zeros(0, np.int64)
    total = int(sizes.sum())
    dm = _div_mode_for(tensor, div_mode)
    if n_jobs == 0:
        e = torch.empty(0)
        return PackedRows(rows_h, bits_h, offs_h, e.to(torch.uint8), e.to(torch.int32), e, e.to(torch.int32), K)
    if tensor.is_cuda:
        dev = tensor.device
        with torch.cuda.device(dev):
            # job table in ONE host->device copy: offsets (int64) | rows | bits (int32)
            table = np.concatenate([offs_h.view(np.int32), rows_h, bits_h])
            table_d = torch.from_numpy(table).to(dev)
            offs_d = table_d[:2 * n_jobs].view(torch.int64)
            rows_d, bits_d = table_d[2 * n_jobs:3 * n_jobs], table_d[3 * n_jobs:]
            blob = torch.empty(max(total, 16), dtype=torch.uint8, device=dev) if want_codes else None
            z = torch.empty(n_jobs, dtype=torch.int32, device=dev)
            s32 = torch.empty(n_jobs, dtype=torch.float32, device=dev)
            status = torch.empty(n_jobs, dtype=torch.int32, device=dev)
            L.check(lib.slq_quantize_rows(tensor.data_ptr(), n_rows, K, rows_d.data_ptr(), bits_d.data_ptr(),
                                          n_jobs, dm, 1 if write_back else 0, L.ptr(blob), offs_d.data_ptr(),
                                          z.data_ptr(), s32.data_ptr(), status.data_ptr(), L.current_stream(dev)))
            st_h = status.cpu()  # the reference is synchronous too (.item())
    else:
        blob = torch.empty(max(total, 16), dtype=torch.uint8) if want_codes else None
        z = torch.empty(n_jobs, dtype=torch.int32)
        s32 = torch.empty(n_jobs, dtype=torch.float32)
        status = torch.zeros(n_jobs, dtype=torch.int32)
        rc = lib.slq_quantize_rows_host(tensor.data_ptr(), n_rows, K, rows_h.ctypes.data, bits_h.ctypes.data,
                                        n_jobs, dm, 1 if write_back else 0, L.ptr(blob), offs_h.ctypes.data,
                                        z.data_ptr(), s32.data_ptr(), status.data_ptr())
        if rc not in (L.SLQ_OK, L.SLQ_ERR_ZERO_RANGE):
            L.check(rc)
        st_h = status
    if write_back:
        resnet.note_weight_write(tensor)  # engines re-pack exactly this tensor's layer on their next forward
    if bool((st_h & L.ROW_ZERO_RANGE).any()):
        raise ZeroDivisionError("float division by zero")
    return PackedRows(rows_h, bits_h, offs_h, blob, z, s32, status, K)


def quantize_wgt(tensor, bit):
    """Per-tensor min/max affine fake-quantiser; returns a NEW tensor (reference functions.py:25-43)."""
    flat = tensor.detach().to(torch.float32).contiguous().reshape(1, -1).clone()
    quantize_rows(flat, [0], [bit], write_back=True, want_codes=False)
    return flat.reshape(tensor.shape)


def channel_wise_quantizationperchan(tensor, bit, i):
    """Overwrites output channel ``i`` of ``tensor`` with its ``bit``-bit fake-quantised value, in
    place, and returns the same tensor object (reference functions.py:9-23)."""
    if tensor.is_contiguous() and tensor.dtype == torch.float32:
        quantize_rows(tensor, [i], [bit], write_back=True, want_codes=False)
    else:
        tensor[i] = quantize_wgt(tensor[i], bit)
        resnet.note_weight_write(tensor)
    return tensor


# ==============================================================================================
# evaluation (callers of the forward)
# ==============================================================================================
def evaluate_loss(net, device, data_loader):
    """Mean-of-batch-means cross-entropy (reference functions.py:45-82; unused by the mains)."""
    net.eval()
    criterion = torch.nn.CrossEntropyLoss()
    total, count = 0, 0
    for x, y in data_loader:
        x, y = x.to(device), y.to(device)
        with torch.no_grad():
            total = total + criterion(net(x), y)
        count += 1
    return (total / count).item()


def evaluate_acc_loss_softmax(net, device, data_loader):
    """-> (accuracy, mean-of-batch-means CE loss, [softmax per batch])  (reference functions.py:84-129)."""
    net.to(device)
    net.eval()
    criterion = torch.nn.CrossEntropyLoss()
    labels, preds, outputs = [], [], []
    loss_sum, count = 0, 0
    for x, y in data_loader:
        x, y = x.to(device), y.to(device)
        with torch.no_grad():
            out = net(x)
            preds.append(out.max(1)[1])
            loss_sum = loss_sum + criterion(out, y)
            outputs.append(torch.softmax(out, dim=1))
        labels.append(y)
        count += 1
    labels, preds = torch.cat(labels), torch.cat(preds)
    acc = (labels == preds).float().sum() / len(labels)
    return acc.item(), (loss_sum / count).item(), outputs


def KLdiv(n_out, out):
    """Mean over samples of sum_c p*log(p/q), p = before, q = after (reference functions.py:131-149)."""
    total, n = None, 0
    for p, q in zip(n_out, out):
        kl = (p * (p / q).log()).sum(dim=1)
        total = kl.sum() if total is None else total + kl.sum()
        n += kl.numel()
    return (total / n).item()


# ==============================================================================================
# semilayer bookkeeping
# ==============================================================================================
def make_divide_minusplusmodels(paramlists, dlists, index):
    """Splits channel rows into the 'minus' (deltaloss <= 0) and 'plus' semilayer lists; column 5
    becomes the per-layer sign flag 0,-1,-2,.. / 1,2,3,..  (reference functions.py:151-184)."""
    listminus, listplus = [], []
    mflag, pflag = 0, 1
    last = len(paramlists) - 1
    for i, p in enumerate(paramlists):
        if dlists[i][index] <= 0:
            listminus.append([p[0], p[1], p[2], p[3], p[4], mflag, p[6], p[7]])
        else:
            listplus.append([p[0], p[1], p[2], p[3], p[4], pflag, p[6], p[7]])
        if i == last:
            print('function debug:number of total channels=', len(listminus) + len(listplus),
                  'No.1:', len(listminus), 'No.2:', len(listplus))
            break
        if p[2] != paramlists[i + 1][2]:
            mflag -= 1
            pflag += 1
    return listminus, listplus


_CONVS_PER_BLOCK = {"resnet18": 2, "resnet34": 2, "resnet50": 3}


def _conv_of(arch, block, lnum):
    """Layer number -> conv module inside its block (functions.py:236-242 / :504-512)."""
    cpb = _CONVS_PER_BLOCK[arch]
    return getattr(block, "conv%d" % ((lnum - 1) % cpb + 1))


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def make_semilayers(arch, net, device, originaloutputs, listminus, listplus, metric=None, shard=True):
    """Sensitivity sweep (reference functions.py:186-588, three copies): one candidate per semilayer
    = quantise its channels on fresh weights, evaluate the whole loader, sensitivity =
    KL(before, after) / sum(numel * (32-bit)/32).

    Candidates are independent, so with torch.distributed initialised they are dealt round-robin
    to the ranks and the per-candidate values are exchanged with ONE all_gather;
    every rank returns the full ``(semilayers, orders)``.
    metric: 'kl' (reference) or 'dloss' (loss_candidate - loss_base; KL is NaN for random-init
    R34/R50).  Default from $SLQ_SWEEP_METRIC, else 'kl'."""
    metric = metric or os.environ.get("SLQ_SWEEP_METRIC", "kl")
    dist, rank, world = _dist() if shard else (None, 0, 1)
    listminus.append(list(SENTINEL))  # the reference mutates the caller's lists the same way
    listplus.append(list(SENTINEL))
    # ---- pure bookkeeping: candidates = runs of equal layer number inside each list ------------
    semilayers = []
    for lst in (listminus, listplus):
        cur = []
        for i, row in enumerate(lst):
            if row[2] == 100:
                break
            cur.append(list(row[:8]))
            if row[2] != lst[i + 1][2]:
                semilayers.append(cur)
                cur = []
    factory = getattr(resnet, arch)
    # delta-loss needs the un-quantised loss from logits (the stored softmax underflows to exact zeros on
    # random-init R34/R50, which is what makes KL NaN there): one more evaluation of the caller's net,
    # before candidate 0 mutates it
    base = evaluate_acc_loss_softmax(net, device, imagenet.val_loader)[1] if metric == "dloss" else None
    values = torch.zeros(len(semilayers), dtype=torch.float64)
    for index, rows in enumerate(semilayers):
        mine = (index % world) == rank
        if not mine and index != 0:
            continue
        # candidate 0 runs on the CALLER's net object and mutates it (reference quirk Q3); every
        # other candidate gets a fresh 'pretrained' model (functions.py:258/393/528)
        cand = net if index == 0 else factory(num_classes=1000, pretrained='imagenet')
        layers = [cand.layer1, cand.layer2, cand.layer3, cand.layer4]
        param = 0.0
        by_conv = {}
        for li, bi, lnum, cnum, w_bit, _flag, _sel, _gi in rows:
            conv = _conv_of(arch, layers[li][bi], lnum)
            by_conv.setdefault(id(conv), (conv, [], []))
            by_conv[id(conv)][1].append(cnum)
            by_conv[id(conv)][2].append(w_bit)
            param += conv.weight[cnum].data.numel() * ((32 - w_bit) / 32)
        for conv, chans, bits in by_conv.values():
            # a semilayer's rows are distinct channels of one conv: one launch, same result as the
            # reference's per-channel loop
            conv.weight.data = _quantize_channels(conv.weight.data, chans, bits)
        if not mine:
            continue
        acc, loss, after = evaluate_acc_loss_softmax(cand, device, imagenet.val_loader)
        val = (loss - base) if metric == "dloss" else KLdiv(originaloutputs, after) / param
        values[index] = val
        print(rows[0][4], 'bit', 'semilayer-No.', index, 'layernumber=', rows[0][2], 'channels=', len(rows),
              'total KL divergence=' if metric == "kl" else 'delta loss=', val)
    if world > 1:
        gathered = [torch.zeros_like(values) for _ in range(world)]
        pad = values.clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        bufs = [g.to(dev) for g in gathered]
        dist.all_gather(bufs, pad.to(dev))
        for index in range(len(semilayers)):
            values[index] = bufs[index % world][index].cpu()
    orders = [[i, float(values[i])] for i in range(len(semilayers))]
    return semilayers, orders


def _quantize_channels(weight, chans, bits):
    if weight.is_contiguous() and weight.dtype == torch.float32:
        quantize_rows(weight, chans, bits, write_back=True, want_codes=False)
        return weight
    for c, b in zip(chans, bits):
        weight = channel_wise_quantizationperchan(weight, b, c)
    return weight


def make_semilayers_resnet18(net, device, originaloutputs, listminus, listplus):
    return make_semilayers("resnet18", net, device, originaloutputs, listminus, listplus)


def make_semilayers_resnet34(net, device, originaloutputs, listminus, listplus):
    return make_semilayers("resnet34", net, device, originaloutputs, listminus, listplus)


def make_semilayers_resnet50(net, device, originaloutputs, listminus, listplus):
    return make_semilayers("resnet50", net, device, originaloutputs, listminus, listplus)


def make_quantizedlists(semilayers, orders):
    """Least-sensitive semilayer first: sorts ``orders`` in place by sensitivity and flattens the
    semilayers in that order, sentinel-terminated (reference functions.py:590-612)."""
    orders.sort(key=lambda o: o[1])
    flat = []
    for idx, _sens in orders:
        flat.extend(list(r[:8]) for r in semilayers[idx])
        print('debug layernum=', semilayers[idx][-1][2], 'number of channels=', len(semilayers[idx]))
    print('number of valuationfirsts list=', len(flat))
    flat.append(list(SENTINEL))
    return flat
