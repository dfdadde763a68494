"""imagenet.py -- synthetic stand-in for the reference's ``imagenet.py`` (which builds
``ImageFolder('./hogehoge/{train,val}')`` at import time, reference imagenet.py:17-40).

ImageNet is not available offline, so ``val_loader`` is a list of seeded synthetic batches shaped
like the reference's loader output (fp32 NCHW, normalised statistics ~ N(0,1); int64 labels).
Size knobs: $SLQ_SYNTH_BATCHES (default 2), $SLQ_SYNTH_BATCH (default 8), $SLQ_SYNTH_HW (default 224).
Replace this module (or set ``imagenet.val_loader``) to evaluate on real data.

B200-side data format: the forward also accepts the batch as raw u8 pixels (set
``net.input_norm = (MEAN, STD)``; ToTensor + Normalize of reference imagenet.py:14-15 then run inside the
stem kernel, bit-exactly) or as fp16 -- 4x / 2x fewer bytes over PCIe than the reference's fp32 batches.
``synthetic_loader(..., dtype=torch.uint8)`` yields such batches.
"""
import os

import torch


MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)  # reference imagenet.py:14-15


def normalize_u8(x_u8):
    """What the reference's transforms make of raw pixels: ToTensor (u8 -> fp32 / 255) + Normalize, computed
    on the CPU like the reference's DataLoader workers do (ATen's CUDA kernel for tensor / python-scalar
    multiplies by a reciprocal instead of dividing).  Returns a CPU fp32 tensor."""
    t = x_u8.cpu().to(torch.float32).div(255)
    mean = torch.tensor(MEAN, dtype=torch.float32)[:, None, None]
    std = torch.tensor(STD, dtype=torch.float32)[:, None, None]
    return t.sub_(mean).div_(std)


def synthetic_loader(num_batches, batch, hw, seed=1, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    if dtype == torch.uint8:
        return [(torch.randint(0, 256, (batch, 3, hw, hw), generator=g, dtype=torch.uint8),
                 torch.randint(0, 1000, (batch,), generator=g)) for _ in range(num_batches)]
    return [(torch.randn(batch, 3, hw, hw, generator=g).to(dtype), torch.randint(0, 1000, (batch,), generator=g))
            for _ in range(num_batches)]


val_loader = synthetic_loader(int(os.environ.get("SLQ_SYNTH_BATCHES", "2")),
                              int(os.environ.get("SLQ_SYNTH_BATCH", "8")),
                              int(os.environ.get("SLQ_SYNTH_HW", "224")))
train_loader = None
