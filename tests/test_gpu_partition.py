"""GPU: the vertex-partitioned layer through the C ABI.  world_size 1 exercises the row-sub-range plumbing
(interior / boundary / halo calls); the 2-rank NCCL case (needs 2 GPUs, `gpurun --gpus 2`) exercises the halo
exchange overlapped with the interior rows and compares against the single-GPU layer on the whole mesh."""
import os
import socket
import sys

import pytest
import torch

from conftest import ROOT, assert_close_normwise

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _layer_and_data(n_side, c, b, r, seed=0):
    import fieldconv_b200 as fcb
    from fieldconv_b200.synthetic import random_features, torus_mesh
    mesh = torus_mesh(n_side, deg=30.0, seed=seed, device="cpu")
    torch.manual_seed(seed)
    layer = fcb.FCResNetBlock(c, c, b, r, 1, precision="fp32")
    x = random_features(mesh.num_nodes, c, seed=1)
    gy = random_features(mesh.num_nodes, c, seed=2, zero_frac=0)
    return mesh, layer, x, gy


def _to(mesh, dev):
    import types
    out = types.SimpleNamespace(**vars(mesh))
    for k in ("supp_edges", "logMag", "logAng", "xp", "w"):
        setattr(out, k, getattr(mesh, k).to(dev))
    return out


def _reference(mesh, layer, x, gy, dev):
    import fieldconv_b200 as fcb
    m = _to(mesh, dev)
    layer = layer.to(dev)
    plan = fcb.build_plan(m.supp_edges, m.logMag, m.logAng, m.xp, m.w, layer.conv1.R, m.epsilon)
    xr = x.to(dev).requires_grad_(True)
    for p in layer.parameters():
        p.grad = None
    y = layer(xr, plan)
    y.backward(gy.to(dev))
    return y.detach().cpu(), xr.grad.cpu(), {k: p.grad.detach().cpu().clone() for k, p in layer.named_parameters()}


def test_partition_world1_matches_plain_layer():
    import fieldconv_b200 as fcb
    mesh, layer, x, gy = _layer_and_data(24, 8, 1, 4)
    y_ref, gx_ref, gp_ref = _reference(mesh, layer, x, gy, DEV)
    part = fcb.partition_mesh(_to(mesh, DEV), 1, 0)
    xo = x.to(DEV)[part.own_global].requires_grad_(True)
    for p in layer.parameters():
        p.grad = None
    y = layer(xo, part)
    y.backward(gy.to(DEV)[part.own_global])
    assert_close_normwise(y, y_ref[part.own_global.cpu()], 1e-5, "y")
    assert_close_normwise(xo.grad, gx_ref[part.own_global.cpu()], 1e-5, "gx")
    for k, p in layer.named_parameters():
        assert_close_normwise(p.grad, gp_ref[k], 1e-5, k)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import fieldconv_b200 as fcb
        mesh, layer, x, gy = _layer_and_data(32, 8, 2, 6)
        y_ref, gx_ref, gp_ref = _reference(mesh, layer, x, gy, dev)
        part = fcb.partition_mesh(_to(mesh, dev), world, rank)
        assert part.n_halo > 0 and 0 < part.n_interior < part.n_own
        own = part.own_global
        xo = x.to(dev)[own].requires_grad_(True)
        for p in layer.parameters():
            p.grad = None
        y = layer(xo, part)
        y.backward(gy.to(dev)[own])
        fcb.allreduce_gradients(list(layer.parameters()))
        torch.cuda.synchronize()
        assert_close_normwise(y, y_ref[own.cpu()], 1e-5, "y")
        assert_close_normwise(xo.grad, gx_ref[own.cpu()], 1e-5, "gx")
        for k, p in layer.named_parameters():
            assert_close_normwise(p.grad, gp_ref[k], 2e-5, k)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_partition_two_gpus_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok0")) and os.path.exists(os.path.join(str(tmp_path), "ok1"))
