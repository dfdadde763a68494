"""Multi-GPU host logic on CPU: the sharded sensitivity sweep with world_size 2 over gloo returns,
on every rank, the same (semilayers, orders) as the single-process sweep.
The CUDA entry points are replaced by the oracle (tests/cpu_standins.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import GOLD, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, n_layers):
    for p in ("semilayer-wise-mixed-precision-quantization_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(ROOT, p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import cpu_standins
    import functions
    import imagenet
    import resnet
    import slq_oracle as so
    torch.set_num_threads(2)
    functions.quantize_rows = cpu_standins.oracle_quantize_rows
    resnet.ResNet.cpu_checker = staticmethod(so.torch_forward)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    loader = imagenet.synthetic_loader(1, 4, 32, seed=1)
    imagenet.val_loader = loader
    torch.manual_seed(0)
    sd = resnet.resnet18(num_classes=1000).state_dict()
    resnet.load_state_dict_from_url = lambda url, progress=True: sd
    net2 = resnet.resnet18(num_classes=1000, pretrained="imagenet")
    _, _, orig = functions.evaluate_acc_loss_softmax(net2, "cpu", loader)
    sw = np.load(os.path.join(GOLD, "sweep_resnet18.npz"))
    keep = lambda a: [[int(v) for v in r] for r in a if r[2] <= n_layers]
    minus, plus = keep(sw["minus_rows"]), keep(sw["plus_rows"])
    semilayers, orders = functions.make_semilayers_resnet18(net2, "cpu", orig, minus, plus)
    flat = functions.make_quantizedlists(semilayers, [list(o) for o in orders])
    np.savez(os.path.join(out_dir, "r%d_w%d.npz" % (rank, world)), orders=np.array(orders, np.float64),
             sizes=np.array([len(s) for s in semilayers]), flat=np.array(flat, np.int64),
             net2=net2.layer1[0].conv1.weight.detach().numpy())
    if world > 1:
        dist.destroy_process_group()


def test_sharded_sweep_world2_equals_single_process(tmp_path):
    n_layers = 3  # first 3 conv layers -> 6 candidate semilayers
    _worker(0, 1, 0, str(tmp_path), n_layers)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), n_layers), nprocs=2, join=True)
    one = np.load(tmp_path / "r0_w1.npz")
    assert one["orders"].shape == (6, 2)
    for r in range(2):
        two = np.load(tmp_path / ("r%d_w2.npz" % r))
        assert np.array_equal(two["sizes"], one["sizes"])
        assert np.array_equal(two["orders"], one["orders"])  # identical floats, not just close
        assert np.array_equal(two["flat"], one["flat"])
        assert np.array_equal(two["net2"], one["net2"])       # quirk Q3 applied on every rank
