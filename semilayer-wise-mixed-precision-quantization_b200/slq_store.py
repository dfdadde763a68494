"""slq_store.py -- the mixed-precision model as PACKED CODES: device-resident snapshots
and the on-disk / wire format (N4).

The reference only ever stores fp32 fake-quantised ``state_dict``s: its greedy search writes one to disk at
every accepted step and reads it back at every rejected one (resnet50_main.py:212 ``torch.save``,
:233-234 ``torch.load`` + ``load_state_dict``; ~100 MB each way for ResNet-50), and the "reduced parameters"
it reports (resnet50_main.py:190) are never materialised.

Here a snapshot keeps, for every quantised conv layer whose rows all lie on a <= 8-bit grid, the engine's
packed store (per output channel: bit-width, zero point z, float32 scale s, little-endian packed codes --
15.3 MB for the P0 ResNet-50) and restores it with ONE decode kernel per layer (slq_decode_rows:
w = fp32((code + z) * s), bit-identical to what the quantizer wrote back).  Layers that still hold
never-quantised fp32 rows, and everything the reference keeps in fp32 (stem, downsample convs, BatchNorm,
fc), are kept as fp32 tensors.  Nothing leaves the device.

File format ``SLQPACK1`` (little endian):
    bytes 0..7    magic  b"SLQPACK1"
    bytes 8..15   uint64 header length H
    bytes 16..    H bytes of UTF-8 JSON: {"format": 1, "arch": ..., "entries": [ {"name", "kind": "packed" |
                  "dense", "shape", "K", "arrays": {array name: [dtype, element count, byte offset]}} ]}
    then          the arrays, each starting at a multiple of 64 bytes counted from the end of the header
                  (packed: bits int32[Cout], z int32[Cout], s float32[Cout], offsets int64[Cout], blob uint8[..];
                   dense: data <dtype>[prod(shape)])
"""
import json
import struct

import numpy as np
import torch

import slq_lib as L

MAGIC = b"SLQPACK1"


class Entry:
    def __init__(self, name, kind, shape, arrays, K=0):
        self.name, self.kind, self.shape, self.arrays, self.K = name, kind, tuple(shape), arrays, K

    @property
    def nbytes(self):
        return sum(int(a.numel()) * a.element_size() for a in self.arrays.values())


class Snapshot:
    """Ordered ``state_dict`` key -> Entry.  ``nbytes``: what it holds; ``fp32_bytes``: the state_dict it replaces."""

    def __init__(self, entries, arch=None):
        self.entries, self.arch = entries, arch

    @property
    def nbytes(self):
        return sum(e.nbytes for e in self.entries)

    @property
    def fp32_bytes(self):
        return sum(int(np.prod(e.shape)) * 4 if e.kind == "packed" else e.nbytes for e in self.entries)

    def packed_names(self):
        return [e.name for e in self.entries if e.kind == "packed"]


def _engine_of(net):
    engines = getattr(net, "_slq_engines", None) or {}
    return next(iter(engines.values()), None)


def snapshot(net):
    """Device-resident snapshot of ``net``'s parameters and buffers (replaces ``torch.save(net.state_dict(),
    pthname)``, resnet50_main.py:212).  Zero-copy for packed layers: the engine's packed store is immutable
    (a re-pack allocates a new one), so the snapshot just keeps a reference."""
    eng = _engine_of(net)
    packed_of = {}
    if eng is not None:
        eng.sync_weights(force=getattr(net, "_slq_dirty", False))
        net._slq_dirty = False
        for op in eng.ops:
            if int(op.bits_host.max()) <= 8 and op.packed_exact:  # every row decodes bit for bit
                packed_of[id(op.conv.weight)] = op.packed
    params = dict(net.named_parameters())
    entries = []
    for name, t in net.state_dict().items():
        p = params.get(name)
        pk = packed_of.get(id(p)) if p is not None else None
        if pk is not None:
            entries.append(Entry(name, "packed", t.shape,
                                 dict(bits=pk.bits, z=pk.z, s=pk.s, offsets=pk.offsets, blob=pk.blob), K=pk.K))
        else:
            entries.append(Entry(name, "dense", t.shape, dict(data=t.detach().clone())))
    return Snapshot(entries, arch=getattr(net, "block_name", None))


def restore(net, snap):
    """Writes a snapshot back into ``net`` IN PLACE (replaces ``net.load_state_dict(torch.load(pthname))``,
    resnet50_main.py:233-234): packed layers through slq_decode_rows on the parameter's device."""
    import resnet
    lib = L.lib()
    sd = net.state_dict()
    names = [e.name for e in snap.entries]
    if names != list(sd.keys()):
        raise KeyError("snapshot does not match this module tree")
    with torch.no_grad():
        for e in snap.entries:
            dst = sd[e.name]
            if tuple(dst.shape) != e.shape:
                raise ValueError("shape mismatch for %s" % e.name)
            if e.kind == "dense":
                dst.copy_(e.arrays["data"])
                if dst.dim() == 4:
                    resnet.note_weight_write(dst)
                continue
            if not dst.is_cuda:
                raise RuntimeError("restoring packed layers needs the model on the CUDA device (no CPU fallback)")
            if dst.dtype != torch.float32 or not dst.is_contiguous():
                raise TypeError("packed layers restore into contiguous float32 parameters")
            a = {k: v.to(dst.device) for k, v in e.arrays.items()}
            with torch.cuda.device(dst.device):
                L.check(lib.slq_decode_rows(a["blob"].data_ptr(), a["offsets"].data_ptr(), a["bits"].data_ptr(),
                                            a["z"].data_ptr(), a["s"].data_ptr(), e.shape[0], e.K, dst.data_ptr(),
                                            L.current_stream(dst.device)))
            resnet.note_weight_write(dst)


# ----------------------------------------------------------------------------------------------
# on-disk / wire format
# ----------------------------------------------------------------------------------------------
_DTYPES = {"int32": np.int32, "int64": np.int64, "float32": np.float32, "uint8": np.uint8, "float64": np.float64,
           "float16": np.float16, "int8": np.int8, "bool": np.bool_}


def dumps(snap):
    """Snapshot -> bytes in the SLQPACK1 container."""
    index, blobs, pos = [], [], 0
    for e in snap.entries:
        arrays = {}
        for k, t in e.arrays.items():
            a = t.detach().cpu().contiguous().numpy()
            pos = (pos + 63) // 64 * 64
            arrays[k] = [str(a.dtype), int(a.size), pos]
            blobs.append((pos, a.tobytes()))
            pos += a.nbytes
        index.append(dict(name=e.name, kind=e.kind, shape=list(e.shape), K=int(e.K), arrays=arrays))
    header = json.dumps(dict(format=1, arch=snap.arch, entries=index)).encode()
    out = bytearray(MAGIC + struct.pack("<Q", len(header)) + header)
    base = len(out)
    out.extend(b"\0" * pos)
    for off, raw in blobs:
        out[base + off:base + off + len(raw)] = raw
    return bytes(out)


def loads(raw, device="cpu"):
    """bytes -> Snapshot with its arrays on ``device``."""
    if raw[:8] != MAGIC:
        raise ValueError("not an SLQPACK1 file")
    (hlen,) = struct.unpack("<Q", raw[8:16])
    meta = json.loads(raw[16:16 + hlen].decode())
    if meta.get("format") != 1:
        raise ValueError("unsupported SLQPACK format %r" % meta.get("format"))
    base = 16 + hlen
    entries = []
    for e in meta["entries"]:
        arrays = {}
        for k, (dt, count, off) in e["arrays"].items():
            a = np.frombuffer(raw, dtype=_DTYPES[dt], count=count, offset=base + off)
            t = torch.from_numpy(a.copy())
            if k == "data":
                t = t.reshape(e["shape"])
            arrays[k] = t.to(device)
        entries.append(Entry(e["name"], e["kind"], e["shape"], arrays, K=e["K"]))
    return Snapshot(entries, arch=meta.get("arch"))


def save_packed(net, path):
    """Writes the mixed-precision model as packed codes (the format the reference never materialises)."""
    raw = dumps(snapshot(net))
    with open(path, "wb") as f:
        f.write(raw)
    return len(raw)


def load_packed(path, net):
    """Reads an SLQPACK1 file into ``net`` (which must be on the CUDA device if the file has packed layers)."""
    with open(path, "rb") as f:
        raw = f.read()
    dev = next(net.parameters()).device
    restore(net, loads(raw, device=dev))
    return net
