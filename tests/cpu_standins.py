"""tests/cpu_standins.py -- lets the CPU-only suite exercise the host logic of functions.py /
resnet.py without a GPU: the CUDA entry points are replaced by the ORACLE (tests are allowed to do
that; the product never does)."""
import numpy as np
import torch

import slq_oracle as so


class _Packed:
    def __init__(self, n):
        self.status = torch.zeros(n, dtype=torch.int32)


def oracle_quantize_rows(tensor, rows, bits, write_back=True, want_codes=True, div_mode=None):
    import resnet
    assert not tensor.is_cuda and tensor.is_contiguous() and tensor.dtype == torch.float32
    w2 = tensor.detach().reshape(tensor.shape[0], -1).numpy()
    for r, b in zip(np.asarray(rows).reshape(-1), np.asarray(bits).reshape(-1)):
        q, _, _, _, _ = so.quantize_row(w2[int(r)], int(b), so.DIV_TRUE if div_mode is None else div_mode)
        if write_back:
            w2[int(r)] = q
    if write_back:
        resnet.bump_weight_epoch()
    return _Packed(len(rows))


def install(monkeypatch):
    import functions
    import resnet
    monkeypatch.setattr(functions, "quantize_rows", oracle_quantize_rows)
    monkeypatch.setattr(resnet.ResNet, "cpu_checker", staticmethod(so.torch_forward))
