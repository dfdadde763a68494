"""imagenet.py -- synthetic stand-in for the reference's ``imagenet.py`` (which builds
``ImageFolder('./hogehoge/{train,val}')`` at import time, reference imagenet.py:17-40).

ImageNet is not available offline, so ``val_loader`` is a list of seeded synthetic batches shaped
like the reference's loader output (fp32 NCHW, normalised statistics ~ N(0,1); int64 labels).
Size knobs: $SLQ_SYNTH_BATCHES (default 2), $SLQ_SYNTH_BATCH (default 8), $SLQ_SYNTH_HW (default 224).
Replace this module (or set ``imagenet.val_loader``) to evaluate on real data.
"""
import os

import torch


def synthetic_loader(num_batches, batch, hw, seed=1):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, hw, hw, generator=g), torch.randint(0, 1000, (batch,), generator=g))
            for _ in range(num_batches)]


val_loader = synthetic_loader(int(os.environ.get("SLQ_SYNTH_BATCHES", "2")),
                              int(os.environ.get("SLQ_SYNTH_BATCH", "8")),
                              int(os.environ.get("SLQ_SYNTH_HW", "224")))
train_loader = None
