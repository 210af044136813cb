"""Seeded synthetic support graphs with the exact attribute set the reference's offline
transforms produce (transforms/support_graph.py:56-59, transforms/compute_log_xport.py:36-50):

    supp_edges (E,2) int64  rows (j source, i target), grouped by source
    logMag, logAng (E,) f32 polar log-map log_j(i) in j's frame (self loop: r=0, theta=0)
    xp (E,) complex64       parallel transport of j's frame into i's frame
    w (N,1) f32             lumped vertex mass

The surface is a flat torus sampled on a jittered n_side x n_side grid, so every vertex has the
same expected degree and the geometry has a closed form (log_j(i) = p_i - p_j, zero holonomy);
per-vertex random frame angles alpha_v make theta and xp non-trivial.  Vertices are numbered in
row-major tiles (spatially coherent like a real mesh) or, with ``permute=True``, randomly
(cache-hostile case).  Pure torch, runs on CPU or CUDA; used by tests and bench only.
"""
import math
import types

import torch


def torus_mesh(n_side, deg=40.0, seed=0, device="cpu", jitter=0.25, tile=8, permute=False,
               chunk=1 << 18):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = n_side * n_side
    h = 1.0 / n_side
    eps = h * math.sqrt(deg / math.pi)
    dev = torch.device(device)

    # --- vertex numbering: new id -> grid cell (cx, cy)
    cy, cx = torch.meshgrid(torch.arange(n_side), torch.arange(n_side), indexing="ij")
    cx, cy = cx.reshape(-1), cy.reshape(-1)
    if permute:
        order = torch.randperm(n, generator=g)
    else:
        tiles_per_row = (n_side + tile - 1) // tile
        key = ((cy // tile) * tiles_per_row + (cx // tile)) * (tile * tile) + (cy % tile) * tile + (cx % tile)
        order = torch.argsort(key, stable=True)
    cell_of_id = (cy[order] * n_side + cx[order])          # new id -> linear cell
    id_of_cell = torch.empty(n, dtype=torch.long)
    id_of_cell[cell_of_id] = torch.arange(n)

    jit = (torch.rand(n, 2, generator=g) * 2 - 1) * jitter  # per cell
    alpha = (torch.rand(n, generator=g) * 2 - 1) * math.pi   # per cell
    w = (torch.rand(n, 1, generator=g) + 0.5)                # per new id
    px = ((cx.float() + 0.5 + jit[:, 0]) * h).to(dev)        # per cell
    py = ((cy.float() + 0.5 + jit[:, 1]) * h).to(dev)
    alpha = alpha.to(dev)
    cell_of_id = cell_of_id.to(dev)
    id_of_cell = id_of_cell.to(dev)

    reach = int(math.ceil(eps / h + 2 * jitter))
    offs = torch.arange(-reach, reach + 1, device=dev)
    oy, ox = torch.meshgrid(offs, offs, indexing="ij")
    ox, oy = ox.reshape(-1), oy.reshape(-1)

    out_e, out_r, out_t, out_xp = [], [], [], []
    for lo in range(0, n, chunk):
        src_id = torch.arange(lo, min(n, lo + chunk), device=dev)
        src_cell = cell_of_id[src_id]
        sx, sy = src_cell % n_side, src_cell // n_side
        tx = (sx[:, None] + ox[None, :]) % n_side
        ty = (sy[:, None] + oy[None, :]) % n_side
        tgt_cell = ty * n_side + tx
        dx = px[tgt_cell] - px[src_cell][:, None]
        dy = py[tgt_cell] - py[src_cell][:, None]
        dx = dx - torch.round(dx)                           # minimum image on the unit torus
        dy = dy - torch.round(dy)
        dist = torch.sqrt(dx * dx + dy * dy)
        keep = dist <= eps
        s_idx, o_idx = torch.nonzero(keep, as_tuple=True)   # row-major => grouped by source
        tc = tgt_cell[s_idx, o_idx]
        a_src = alpha[src_cell][s_idx]
        theta = torch.atan2(dy[s_idx, o_idx], dx[s_idx, o_idx]) - a_src
        theta = torch.remainder(theta + math.pi, 2 * math.pi) - math.pi
        rr = dist[s_idx, o_idx]
        theta = torch.where(rr == 0, torch.zeros_like(theta), theta)   # angle(0) = 0
        ang = a_src - alpha[tc]
        out_e.append(torch.stack((src_id[s_idx], id_of_cell[tc]), dim=1))
        out_r.append(rr.float())
        out_t.append(theta.float())
        out_xp.append(torch.polar(torch.ones_like(ang), ang).to(torch.complex64))

    data = types.SimpleNamespace()
    data.num_nodes = n
    data.supp_edges = torch.cat(out_e)
    data.logMag = torch.cat(out_r)
    data.logAng = torch.cat(out_t)
    data.xp = torch.cat(out_xp)
    data.w = w.to(dev)
    data.epsilon = eps
    return data


def merge_meshes(meshes):
    """Block-diagonal union of independent meshes (vertex offsets applied to the edges) —
    how a batch of small meshes is presented to FieldConv."""
    out = types.SimpleNamespace()
    off, edges = 0, []
    for m in meshes:
        edges.append(m.supp_edges + off)
        off += m.num_nodes
    out.num_nodes = off
    out.supp_edges = torch.cat(edges)
    out.logMag = torch.cat([m.logMag for m in meshes])
    out.logAng = torch.cat([m.logAng for m in meshes])
    out.xp = torch.cat([m.xp for m in meshes])
    out.w = torch.cat([m.w for m in meshes])
    out.epsilon = meshes[0].epsilon
    return out


def random_features(n, channels, seed=0, zero_frac=0.01, device="cpu"):
    """complex64 features with a fraction of exact zeros (exercises the origin branch,
    utils/field.py:14-16)."""
    g = torch.Generator(device="cpu").manual_seed(seed + 7919)
    x = torch.complex(torch.randn(n, channels, generator=g), torch.randn(n, channels, generator=g))
    if zero_frac > 0:
        x = torch.where(torch.rand(n, channels, generator=g) < zero_frac, torch.zeros_like(x), x)
    return x.to(device)


def with_edge_cases(mesh, isolated=(), duplicate=0, shuffle_seed=None, far=0):
    """The reference's edge cases applied to a synthetic mesh: `isolated` targets lose every
    incoming edge (y = 0 there, field_conv.py:134 with dim_size = N), the first `duplicate` edges appear twice (summed,
    no coalescing), `far` extra edges get r > epsilon (dropped before the weight normalisation, fc_precomp.py:69-74,87),
    and the edge list is shuffled (any edge order is allowed)."""
    out = types.SimpleNamespace(**vars(mesh))
    e, r, t, xp = mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp
    if len(isolated):
        iso = torch.as_tensor(list(isolated), device=e.device)
        keep = ~torch.isin(e[:, 1], iso)
        e, r, t, xp = e[keep], r[keep], t[keep], xp[keep]
    if duplicate:
        e, r, t, xp = (torch.cat((a, a[:duplicate])) for a in (e, r, t, xp))
    if far:
        e = torch.cat((e, e[:far].flip(1)))
        r = torch.cat((r, torch.full((far,), 1.5 * mesh.epsilon, device=r.device, dtype=r.dtype)))
        t, xp = torch.cat((t, t[:far])), torch.cat((xp, xp[:far]))
    if shuffle_seed is not None:
        perm = torch.randperm(e.shape[0], generator=torch.Generator().manual_seed(shuffle_seed)).to(e.device)
        e, r, t, xp = e[perm], r[perm], t[perm], xp[perm]
    out.supp_edges, out.logMag, out.logAng, out.xp = e.contiguous(), r.contiguous(), t.contiguous(), xp.contiguous()
    return out


def skewed_degree(mesh, seed=0, isolated_frac=0.1, exponent=3.0):
    """A degree-skewed variant of a synthetic mesh: every target keeps each of its incoming edges with its own probability
    p_i = u_i^exponent (u_i uniform), and `isolated_frac` of the targets lose all of them — power-law-like in-degrees with
    many empty rows and a few full ones.  Self loops go through the same lottery."""
    out = types.SimpleNamespace(**vars(mesh))
    g = torch.Generator().manual_seed(seed)
    n = mesh.num_nodes
    p = torch.rand(n, generator=g) ** exponent
    p[torch.rand(n, generator=g) < isolated_frac] = 0.0
    p[torch.rand(n, generator=g) < 0.02] = 1.0
    keep = (torch.rand(mesh.supp_edges.shape[0], generator=g) < p[mesh.supp_edges[:, 1].cpu()]).to(mesh.supp_edges.device)
    for k in ("supp_edges", "logMag", "logAng", "xp"):
        setattr(out, k, getattr(mesh, k)[keep].contiguous())
    return out
