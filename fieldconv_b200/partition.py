"""Vertex-partitioned FieldConv for ONE large mesh across the ranks of a torch.distributed group
(BASELINE.json configs[3]).

FieldConv is one-hop (nn/field_conv.py:130-134): target i needs x[j] only for its in-neighbours j.  Each rank owns a
contiguous range of the (spatially ordered) global vertex numbering and every edge whose TARGET it owns; the sources
it does not own are its halo.  Per layer and direction there is exactly one exchange step:

  forward   owners send their boundary rows of x to the ranks that list them as halo           (HaloExchange.forward)
  backward  ranks send the partial grad-x rows they computed for halo sources back to the
            owners, who add them in rank order (deterministic)                                  (HaloExchange.backward)

and the parameter gradients are all-reduced by the caller like in data parallel.  Local vertex numbering is
[interior targets | boundary targets | halo], applied ONCE per mesh (features, labels and outputs of every layer
live in it), so the layer can run the interior rows — whose sources are all local — while the halo rows are still in
flight on a side stream, and the boundary rows after.

Exchange = batched point-to-point (NCCL send/recv between neighbouring ranks only; gloo on CPU for the host-logic
tests).  The arithmetic is the ordinary single-GPU kernels called on row sub-ranges through the C ABI.
"""
import torch
import torch.distributed as dist

from . import _lib, ops


def equal_bounds(n, world):
    """Contiguous, near-equal vertex ranges: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(n, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


class MeshPartition:
    """This rank's share of the mesh, in local numbering (see module docstring).

    own_global   (n_own,)  global id of local vertex v < n_own      (interior first, then boundary)
    halo_global  (n_halo,) global id of local vertex n_own + h       (ascending => grouped by owner rank)
    supp_edges / logMag / logAng / xp   the edges whose target is owned, endpoints in LOCAL ids
    w            (n_ext, 1) lumped masses of the local vertices
    send_idx     local ids (< n_own) of the rows peers need, concatenated by destination rank
    send_counts / recv_counts   rows per peer (recv rows land in halo order)
    """

    def __init__(self):
        self.plan = None

    @property
    def n_ext(self):
        return self.n_own + self.n_halo

    def to_local(self, t_global):
        """Rows of a global per-vertex tensor for [owned | halo] in local order."""
        return t_global[torch.cat((self.own_global, self.halo_global))]

    def build_plan(self, n_rings):
        from .plan import build_plan
        self.plan = build_plan(self.supp_edges, self.logMag, self.logAng, self.xp, self.w, n_rings, self.epsilon,
                               num_nodes=self.n_ext)
        return self.plan


def partition_mesh(mesh, world, rank, group=None, bounds=None):
    """Split `mesh` (the reference's per-mesh attributes: supp_edges (E,2) (j,i), logMag, logAng, xp, w, epsilon —
    transforms/compute_log_xport.py:36-50) for `rank` of `world`.  Every rank calls this with the same global mesh;
    the only communication is the exchange of halo id lists (who needs which of my rows)."""
    n = int(mesh.num_nodes)
    dev = mesh.supp_edges.device
    b = bounds if bounds is not None else equal_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    e = mesh.supp_edges
    tgt_owned = (e[:, 1] >= lo) & (e[:, 1] < hi)
    e_loc = e[tgt_owned]
    lm = mesh.logMag[tgt_owned]
    # halo from the edges FCPrecomp keeps (r <= epsilon, fc_precomp.py:69-74); the margin only ever adds a harmless
    # extra halo row, the plan builder applies the exact filter
    kept = lm <= mesh.epsilon * (1.0 + 1e-5)
    src = e_loc[kept, 0]
    foreign = (src < lo) | (src >= hi)
    halo_global = torch.unique(src[foreign])                                  # sorted
    is_boundary = torch.zeros(hi - lo, dtype=torch.bool, device=dev)
    is_boundary[e_loc[kept, 1][foreign] - lo] = True
    owned = torch.arange(lo, hi, device=dev)
    own_global = torch.cat((owned[~is_boundary], owned[is_boundary]))

    p = MeshPartition()
    p.rank, p.world, p.bounds, p.group = rank, world, b, group
    p.num_global = n
    p.n_own, p.n_halo = hi - lo, int(halo_global.numel())
    p.n_interior = int((~is_boundary).sum().item())
    p.own_global, p.halo_global = own_global, halo_global
    loc = torch.full((n,), -1, dtype=torch.long, device=dev)
    loc[own_global] = torch.arange(p.n_own, device=dev)
    loc[halo_global] = p.n_own + torch.arange(p.n_halo, device=dev)
    # edges from sources outside own+halo are exactly the dropped (r > eps) ones: point them at the target itself
    # (any valid id) — the plan builder discards them by their logMag
    ls = loc[e_loc[:, 0]]
    lt = loc[e_loc[:, 1]]
    ls = torch.where(ls < 0, lt, ls)
    p.supp_edges = torch.stack((ls, lt), 1).contiguous()
    p.logMag, p.logAng, p.xp = lm.contiguous(), mesh.logAng[tgt_owned].contiguous(), mesh.xp[tgt_owned].contiguous()
    p.w = mesh.w[torch.cat((own_global, halo_global))].contiguous()
    p.epsilon = mesh.epsilon

    # who owns my halo rows -> recv_counts; tell the owners which rows I need -> their send lists
    bt = torch.tensor(b, device=dev)
    owner = torch.bucketize(halo_global, bt[1:], right=True)
    p.recv_counts = torch.bincount(owner, minlength=world).tolist()
    if world == 1:
        p.send_counts, p.send_idx = [0], torch.zeros(0, dtype=torch.long, device=dev)
        return p
    counts = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
    mine = [torch.tensor([c], dtype=torch.long, device=dev) for c in p.recv_counts]
    _all_to_all(counts, mine, group)
    p.send_counts = [int(c.item()) for c in counts]
    want = list(torch.split(halo_global, p.recv_counts))
    asked = [torch.empty(c, dtype=torch.long, device=dev) for c in p.send_counts]
    _all_to_all(asked, want, group)
    p.send_idx = loc[torch.cat(asked)] if sum(p.send_counts) else torch.zeros(0, dtype=torch.long, device=dev)
    assert bool((p.send_idx >= 0).all()) and bool((p.send_idx < p.n_own).all())
    return p


def _all_to_all(out_list, in_list, group):
    """Neighbour exchange as batched point-to-point ops (works on NCCL and on gloo); empty messages are skipped."""
    rank = dist.get_rank(group)
    reqs = []
    for q, (o, i) in enumerate(zip(out_list, in_list)):
        if q == rank:
            if o.numel():
                o.copy_(i)
            continue
        peer = q if group is None else dist.get_global_rank(group, q)
        if i.numel():
            reqs.append(dist.P2POp(dist.isend, i, peer, group))
        if o.numel():
            reqs.append(dist.P2POp(dist.irecv, o, peer, group))
    if reqs:
        for w in dist.batch_isend_irecv(reqs):
            w.wait()


def _exchange_rows(rows_by_peer_out, rows_by_peer_in, group):
    _all_to_all(rows_by_peer_out, rows_by_peer_in, group)


class HaloExchange(torch.autograd.Function):
    """x_own (n_own, C) complex -> x_ext (n_own + n_halo, C); backward returns the owners' summed gradient."""

    @staticmethod
    def forward(ctx, x_own, part):
        ctx.part = part
        halo = halo_forward(x_own, part)
        return torch.cat((x_own, halo), 0)

    @staticmethod
    def backward(ctx, g_ext):
        part = ctx.part
        g_own = g_ext[:part.n_own].clone()
        halo_backward_add(g_own, g_ext[part.n_own:], part)
        return g_own, None


def halo_forward(x_own, part):
    """Send my boundary rows, receive my halo rows (n_halo, C)."""
    xr = torch.view_as_real(x_own) if x_own.is_complex() else x_own
    tail = xr.shape[1:]
    halo = torch.empty((part.n_halo,) + tuple(tail), dtype=xr.dtype, device=xr.device)
    if part.world > 1:
        send = xr[part.send_idx].contiguous()
        _exchange_rows(list(torch.split(halo, part.recv_counts)), list(torch.split(send, part.send_counts)), part.group)
    return torch.view_as_complex(halo) if x_own.is_complex() else halo


def halo_backward_add(g_own, g_halo, part):
    """Reverse exchange: halo-row gradients go back to their owners and are added peer by peer, in rank order."""
    if part.world == 1:
        return
    gr = torch.view_as_real(g_halo.contiguous()) if g_halo.is_complex() else g_halo.contiguous()
    back = torch.empty((sum(part.send_counts),) + tuple(gr.shape[1:]), dtype=gr.dtype, device=gr.device)
    _exchange_rows(list(torch.split(back, part.send_counts)), list(torch.split(gr, part.recv_counts)), part.group)
    tgt = torch.view_as_real(g_own) if g_own.is_complex() else g_own
    off = 0
    for c in part.send_counts:                      # one peer at a time: indices are unique within a peer
        if c:
            tgt.index_add_(0, part.send_idx[off:off + c], back[off:off + c])
        off += c


# ----------------------------------------------------------------------------- the partitioned layer (CUDA)
# bench.py sets HALO_TIMING = [] for one step: every exchange then records (comm start, comm end, compute-stream wait start,
# wait end) CUDA events; collect_halo_timing() turns them into the communication time and its exposed (non-overlapped) part
HALO_TIMING = None


def _timing_events():
    return [torch.cuda.Event(enable_timing=True) for _ in range(4)] if HALO_TIMING is not None else None


def collect_halo_timing():
    torch.cuda.synchronize()
    comm = sum(e[0].elapsed_time(e[1]) for e in (HALO_TIMING or []))
    exposed = sum(e[2].elapsed_time(e[3]) for e in (HALO_TIMING or []))
    return {"comm_ms": comm, "exposed_ms": exposed, "exchanges": len(HALO_TIMING or [])}


class _PartitionedFieldConv(torch.autograd.Function):
    """FieldConv over the owned rows with the halo exchange overlapped:
       forward : [comm stream] halo rows of x      || [main] interior rows;  then boundary rows
       backward: [main] grad-x of the halo sources -> [comm] send them back || [main] grad-x + grad-W of owned rows"""

    @staticmethod
    def forward(ctx, x_own, W, part, band_limit, flags):
        plan = part.plan
        dev = x_own.device
        n_own, n_int, n_ext = part.n_own, part.n_interior, part.n_ext
        ci, co = x_own.shape[1], W.shape[0]
        k = plan.n_rings * ci * (2 * band_limit + 1)
        x_own, W = x_own.contiguous(), W.contiguous()
        x_ext = torch.empty(n_ext, ci, dtype=torch.complex64, device=dev)
        main = torch.cuda.current_stream(dev)
        comm = _comm_stream(dev)
        x_ext[:n_own].copy_(x_own)
        ready = torch.cuda.Event()
        ready.record(main)
        done = torch.cuda.Event()
        tev = _timing_events()
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            if tev:
                tev[0].record(comm)
            halo = halo_forward(x_own, part)
            x_ext[n_own:].copy_(halo)
            if tev:
                tev[1].record(comm)
            done.record(comm)
        y = torch.empty(n_own, co, dtype=torch.complex64, device=dev)
        fused = bool(flags & _lib.FLAG_FUSED)
        flags &= ~_lib.FLAG_FUSED
        keep = ops.keep_contrib_default(n_own * k * 8, dev) and not fused
        contrib = torch.empty((0 if fused else n_own), k, dtype=torch.complex64, device=dev)
        cmax = torch.zeros(1, dtype=torch.float32, device=dev)      # max|contrib| over both row ranges (atomic max)

        def rows(a, b):
            if b <= a:
                return
            if fused:      # band_limit <= 1: gather -> shared-memory tile -> tcgen05 in one kernel, no contrib buffer at all
                nbytes = _lib.query_bytes("fcb_fwd_fused_workspace_bytes", ci, co, band_limit, plan.n_rings)
                ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
                _lib.call("fcb_fwd_fused_f32", torch.view_as_real(x_ext).data_ptr(), torch.view_as_real(W).data_ptr(),
                          plan.rowptr_tgt[a:].data_ptr(), plan.rec_tgt.data_ptr(), plan.rot_tgt.data_ptr(),
                          plan.norms.data_ptr(), torch.view_as_real(y)[a:].data_ptr(), b - a, n_ext, ci, co, band_limit,
                          plan.n_rings, ws.data_ptr(), nbytes, _lib.stream_ptr())
                return
            nbytes = _lib.query_bytes("fcb_fwd_workspace_bytes", b - a, ci, co, band_limit, plan.n_rings, flags)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            _lib.call("fcb_fwd_f32", torch.view_as_real(x_ext).data_ptr(), torch.view_as_real(W).data_ptr(),
                      plan.rowptr_tgt[a:].data_ptr(), plan.rec_tgt.data_ptr(), plan.rot_tgt.data_ptr(),
                      torch.view_as_real(y)[a:].data_ptr(), torch.view_as_real(contrib)[a:].data_ptr(), cmax.data_ptr(),
                      b - a, ci, co, band_limit, plan.n_rings, flags, ws.data_ptr(), nbytes, _lib.stream_ptr())

        with torch.cuda.device(dev):
            rows(0, n_int)                       # interior: every source is local
            if tev:
                tev[2].record(main)
            main.wait_event(done)
            if tev:
                tev[3].record(main)
                HALO_TIMING.append(tev)
            rows(n_int, n_own)                   # boundary: needs the halo rows
        halo.record_stream(main)
        ctx.part, ctx.cfg = part, (band_limit, flags)
        ctx.save_for_backward(x_ext, W, contrib if keep else torch.empty(0, dtype=torch.complex64, device=dev), cmax)
        return y

    @staticmethod
    def backward(ctx, gy):
        part = ctx.part
        plan = part.plan
        band_limit, flags = ctx.cfg
        x_ext, W, contrib, cmax = ctx.saved_tensors
        dev = gy.device
        n_own, n_ext = part.n_own, part.n_ext
        ci, co = x_ext.shape[1], W.shape[0]
        gy = gy.contiguous()
        need_gx, need_gw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx_ext = torch.empty(n_ext, ci, dtype=torch.complex64, device=dev) if need_gx else None
        gw = torch.empty_like(W) if need_gw else None
        have = contrib.numel() > 0
        # no contrib kept: gW = sum over the SOURCE rows of conj(xhat) G, so the halo sources contribute their share too
        gw_halo = torch.zeros_like(W) if (need_gw and not have and n_ext > n_own) else None

        def rows(a, b, with_gw, gw_out=None):
            gw_out = gw if gw_out is None else gw_out
            if b <= a or not (need_gx or with_gw):
                return
            fl = flags | (0x100 if (have or not with_gw) else 0)
            nbytes = _lib.query_bytes("fcb_bwd_workspace_bytes", b - a, ci, co, band_limit, plan.n_rings, fl)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            _lib.call("fcb_bwd_f32", torch.view_as_real(x_ext)[a:].data_ptr(), torch.view_as_real(W).data_ptr(),
                      torch.view_as_real(gy).data_ptr(),
                      torch.view_as_real(contrib)[a:].data_ptr() if (have and with_gw) else 0,
                      cmax.data_ptr() if (have and with_gw) else 0,
                      plan.rowptr_tgt[a:].data_ptr(), plan.rec_tgt.data_ptr(), plan.rot_tgt.data_ptr(),
                      plan.rowptr_src[a:].data_ptr(), plan.rec_src.data_ptr(), plan.rot_src.data_ptr(),
                      torch.view_as_real(gx_ext)[a:].data_ptr() if need_gx else 0,
                      torch.view_as_real(gw_out).data_ptr() if with_gw else 0, None,
                      b - a, ci, co, band_limit, plan.n_rings, flags, ws.data_ptr(), nbytes, _lib.stream_ptr())

        main = torch.cuda.current_stream(dev)
        comm = _comm_stream(dev)
        with torch.cuda.device(dev):
            if need_gx or gw_halo is not None:
                rows(n_own, n_ext, gw_halo is not None, gw_halo)   # partial grad-x (and grad-W share) of the halo sources
            if need_gx:
                halo_ready = torch.cuda.Event()
                halo_ready.record(main)
            rows(0, n_own, need_gw)              # owned rows: grad-x and grad-W, overlaps the send below
            if gw_halo is not None:
                gw = gw + gw_halo
        gx_own = None
        if need_gx:
            gx_own = gx_ext[:n_own]
            back_done = torch.cuda.Event()
            own_ready = torch.cuda.Event()
            own_ready.record(main)
            tev = _timing_events()
            with torch.cuda.stream(comm):
                comm.wait_event(halo_ready)
                comm.wait_event(own_ready)       # the adds below touch gx_own
                if tev:
                    tev[0].record(comm)
                halo_backward_add(gx_own, gx_ext[n_own:], part)
                if tev:
                    tev[1].record(comm)
                back_done.record(comm)
            if tev:
                tev[2].record(main)
            main.wait_event(back_done)
            if tev:
                tev[3].record(main)
                HALO_TIMING.append(tev)
        return gx_own, gw, None, None, None


_COMM_STREAMS = {}


def _comm_stream(dev):
    key = torch.device(dev).index
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COMM_STREAMS[key]


def partitioned_field_conv(layer, x_own, part):
    """`layer` is a fieldconv_b200.FieldConv; x_own (n_own, Ci) complex64 in the partition's local order."""
    from .nn import _resolve_precision, fused_flags
    if part.plan is None or part.plan.n_rings != layer.R:
        part.build_plan(layer.R)
    ci, co = layer.in_channels, layer.out_channels
    if ci % 2 or co % 2:
        raise ValueError("partitioned FieldConv needs even channel counts")
    if x_own.shape[0] != part.n_own or x_own.shape[1] != ci:
        raise ValueError("x_own must be (%d owned rows, %d channels), got %s" % (part.n_own, ci, tuple(x_own.shape)))
    flags = _resolve_precision(layer.precision, ci, co, layer.R, layer.B) & _lib.GEMM_MASK   # row sub-ranges: fp32 operand layout
    if layer.precision in ("auto", "2xf16", "2xf16p"):
        flags = fused_flags(flags, part.plan, ci, co, layer.B, layer.R)
    return _PartitionedFieldConv.apply(x_own, layer.weight(), part, layer.B, flags)


def allreduce_gradients(params, group=None):
    """Sum the parameter gradients of all ranks (each holds the contribution of its owned targets)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])
