"""CPU (gloo, world_size 2 and 3): host logic of the vertex partition — local renumbering, halo lists, the
forward halo exchange and its adjoint.  The arithmetic of the layer is the oracle's (tests may use it): a FieldConv
evaluated per rank on [owned | halo] rows must reproduce the single-process result, outputs and gradients."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fieldconv_b200.partition import HaloExchange, equal_bounds, partition_mesh
        from fieldconv_b200.synthetic import random_features, torus_mesh
        from oracle import restate
        torch.set_num_threads(1)
        B, R, C = 1, 3, 4
        mesh = torus_mesh(12, deg=14.0, seed=3)
        n = mesh.num_nodes
        part = partition_mesh(mesh, world, rank)
        # --- bookkeeping invariants
        b = equal_bounds(n, world)
        assert part.n_own == b[rank + 1] - b[rank]
        assert sorted(part.own_global.tolist()) == list(range(b[rank], b[rank + 1]))
        assert sum(part.recv_counts) == part.n_halo and part.recv_counts[rank] == 0
        assert sum(part.send_counts) == part.send_idx.numel() and part.send_counts[rank] == 0
        assert bool(((part.halo_global < b[rank]) | (part.halo_global >= b[rank + 1])).all())
        assert part.supp_edges.min() >= 0 and part.supp_edges.max() < part.n_ext
        assert bool((part.supp_edges[:, 1] < part.n_own).all())
        # interior targets have no foreign source
        keep = part.logMag <= part.epsilon
        src, tgt = part.supp_edges[keep, 0], part.supp_edges[keep, 1]
        assert not bool((src[tgt < part.n_interior] >= part.n_own).any())
        if world > 1:
            assert bool((src[tgt >= part.n_interior] >= part.n_own).any())

        # --- layer on the partition == layer on the whole mesh (oracle arithmetic), forward and backward
        torch.manual_seed(0)
        W = torch.randn(C, C, R, 2 * B + 1, dtype=torch.complex64)
        x = random_features(n, C, seed=1)
        gy = random_features(n, C, seed=2, zero_frac=0)
        e_g, sten_g, _, _, _ = restate.fc_precomp(mesh.logMag, mesh.logAng, mesh.w, mesh.supp_edges, mesh.xp, B, R, mesh.epsilon)
        xg = x.clone().requires_grad_(True)
        y_ref = restate.field_conv_lean(xg, e_g, sten_g, W, B)
        (y_ref.real * gy.real + y_ref.imag * gy.imag).sum().backward()

        x_own = x[part.own_global].clone().requires_grad_(True)
        x_ext = HaloExchange.apply(x_own, part)
        assert torch.equal(x_ext.detach(), part.to_local(x))
        e_l, sten_l, _, _, _ = restate.fc_precomp(part.logMag, part.logAng, part.w, part.supp_edges, part.xp, B, R, part.epsilon)
        y_loc = restate.field_conv_lean(x_ext, e_l, sten_l, W, B)
        # pad: targets without kept edges at the end of the range shrink fc_precomp's dim_size
        y_own = torch.zeros(part.n_own, C, dtype=y_loc.dtype)
        y_own[:min(part.n_own, y_loc.shape[0])] = y_loc[:part.n_own]
        gy_own = gy[part.own_global]
        (y_own.real * gy_own.real + y_own.imag * gy_own.imag).sum().backward()
        err_y = float((y_own.detach() - y_ref.detach()[part.own_global]).abs().max() / y_ref.detach().abs().max())
        err_g = float((x_own.grad - xg.grad[part.own_global]).abs().max() / xg.grad.abs().max())
        assert err_y < 1e-6 and err_g < 1e-6, (err_y, err_g)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("%g %g %d %d" % (err_y, err_g, part.n_interior, part.n_halo))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_exchange_gloo(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(os.path.join(str(tmp_path), "ok%d" % r))


def test_single_rank_partition_is_identity():
    from fieldconv_b200.partition import partition_mesh
    from fieldconv_b200.synthetic import torus_mesh
    mesh = torus_mesh(8, deg=10.0, seed=1)
    part = partition_mesh(mesh, 1, 0)
    assert part.n_halo == 0 and part.n_interior == part.n_own == mesh.num_nodes
    assert torch.equal(part.supp_edges, mesh.supp_edges)


def _dp_worker(rank, world, port, out_dir):
    """Data-parallel mode (batches of small meshes): every rank holds its own mesh; the only collective
    is the all-reduce of the parameter gradients (fieldconv_b200.allreduce_gradients)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fieldconv_b200 as fcb
        from fieldconv_b200.synthetic import random_features, torus_mesh
        from oracle import restate
        torch.set_num_threads(1)
        B, R, C = 1, 3, 4

        def local_grads(seed, params):
            mesh = torus_mesh(10, deg=12.0, seed=seed)
            e, sten, _, _, _ = restate.fc_precomp(mesh.logMag, mesh.logAng, mesh.w, mesh.supp_edges, mesh.xp, B, R, mesh.epsilon)
            x = random_features(mesh.num_nodes, C, seed=10 + seed)
            gy = random_features(mesh.num_nodes, C, seed=20 + seed, zero_frac=0)
            W = fcb.fold_weights(params[0], params[1], params[2], 1, B)
            y = restate.field_conv_lean(x, e, sten, W, B)
            return torch.autograd.grad((y.real * gy.real + y.imag * gy.imag).sum(), params)

        torch.manual_seed(0)                      # identical parameters on every rank
        layer = fcb.FieldConv(C, C, B, R, 1)
        params = [layer.zonal, layer.spherical, layer.phase]
        for p, g in zip(params, local_grads(rank, params)):
            p.grad = g.clone()
        fcb.allreduce_gradients(params)
        total = [sum(gs) for gs in zip(*[local_grads(r, params) for r in range(world)])]
        for p, t in zip(params, total):
            assert float((p.grad - t).abs().max()) <= 1e-6 * float(t.abs().max()) + 1e-12
        open(os.path.join(out_dir, "dp%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(os.path.join(str(tmp_path), "dp%d" % r))
