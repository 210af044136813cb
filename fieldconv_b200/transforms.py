"""Device-side drop-in for the reference's edge-organising transform.

    FCPrecomp(band_limit, n_rings, epsilon)(data) -> (supp_edges, supp_sten, ln, wxp)      transforms/fc_precomp.py:30-97

Same constructor, same call signature, same four outputs (kept edges in input order, dense stencil (E',R,2B+1),
ln = polar(r/eps, theta), wxp) — computed by libfieldconv_b200 on the GPU.  The compact plan built on the way is
attached to the returned ``supp_sten`` so that ``FieldConv.forward(x, supp_edges, supp_sten)`` — the reference's exact
call — takes the compact fast path without any change to the calling network (nn/fc_resnet_block.py:84-88 style callers).
"""
import torch

from . import _lib
from .plan import build_plan


class FCPrecomp(object):
    def __init__(self, band_limit, n_rings, epsilon):
        self.B = band_limit
        self.R = n_rings
        self.max_r = epsilon

    def __call__(self, data):
        r, theta, w, supp_edges, xp = data.logMag, data.logAng, data.w, data.supp_edges, data.xp
        if not supp_edges.is_cuda:
            raise RuntimeError("fieldconv_b200.FCPrecomp runs on CUDA tensors only (no CPU path)")
        dev = supp_edges.device
        plan = build_plan(supp_edges, r, theta, xp, w, self.R, self.max_r)
        e_kept = plan.num_edges                      # one host sync, like torch.nonzero in fc_precomp.py:69
        e_in = int(supp_edges.shape[0])
        m = 2 * self.B + 1
        edges_out = torch.empty(e_kept, 2, dtype=torch.int64, device=dev)
        sten = torch.empty(e_kept, self.R, m, dtype=torch.complex64, device=dev)
        ln = torch.empty(e_kept, dtype=torch.complex64, device=dev)
        wxp = torch.empty(e_kept, dtype=torch.complex64, device=dev)
        if e_kept > 0:
            nbytes = _lib.query_bytes("fcb_precomp_workspace_bytes", e_in)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            edges = supp_edges.to(torch.int64).contiguous()
            with torch.cuda.device(dev):
                _lib.call("fcb_precomp_expand_f32", edges.data_ptr(), r.contiguous().data_ptr(), theta.contiguous().data_ptr(),
                          float(self.max_r), e_in, plan.num_nodes, self.R, self.B, plan.rowptr_tgt.data_ptr(),
                          plan.rec_tgt.data_ptr(), plan.perm_tgt.data_ptr(), e_kept, edges_out.data_ptr(),
                          torch.view_as_real(sten).data_ptr(), torch.view_as_real(ln).data_ptr(),
                          torch.view_as_real(wxp).data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr())
        attach_plan(sten, edges_out, plan)
        return edges_out, sten, ln, wxp

    def __repr__(self):
        return '{}(n_rings={}, epsilon={})'.format(self.__class__.__name__, self.R, self.max_r)


def radius_graph(pos, epsilon, max_num_neighbors=512):
    """(E, 2) int64 rows (j, i) with |pos_i - pos_j| <= epsilon, self loops included, grouped by j, at most
    max_num_neighbors per j — the radius query of transforms/support_graph.py:56-59 on the device (csrc/radius.cu)."""
    if not pos.is_cuda:
        raise RuntimeError("fieldconv_b200.radius_graph runs on CUDA tensors only (no CPU path)")
    p = pos.detach().to(torch.float32).contiguous()
    if p.dim() != 2 or p.shape[1] != 3:
        raise ValueError("pos must be (N, 3)")
    n = int(p.shape[0])
    dev = p.device
    if n == 0:
        return torch.zeros(0, 2, dtype=torch.int64, device=dev)
    lo = p.min(0).values.tolist()
    nbytes = _lib.query_bytes("fcb_radius_workspace_bytes", n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fcb_radius_count", p.data_ptr(), n, float(epsilon), int(max_num_neighbors), lo[0], lo[1], lo[2],
                  counts.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr())
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=offsets[1:])
        e = int(offsets[-1].item())                      # one host sync: the edge count sizes the output
        edges = torch.empty(e, 2, dtype=torch.int64, device=dev)
        if e:
            _lib.call("fcb_radius_fill", p.data_ptr(), n, float(epsilon), int(max_num_neighbors), lo[0], lo[1], lo[2],
                      offsets.data_ptr(), edges.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr())
    return edges


def farthest_point_sample(pos, n_samples, start=0):
    """Deterministic farthest-point sampling (the reference uses torch_geometric.nn.fps with a random start,
    transforms/support_graph.py:46): returns sorted indices like `.sort()[0]` there."""
    n = pos.shape[0]
    idx = torch.empty(n_samples, dtype=torch.long, device=pos.device)
    d = torch.full((n,), float("inf"), device=pos.device)
    cur = torch.tensor(start, device=pos.device)
    for k in range(n_samples):
        idx[k] = cur
        d = torch.minimum(d, ((pos - pos[cur]) ** 2).sum(1))
        cur = torch.argmax(d)
    return idx.sort()[0]


class SupportGraph(object):
    """transforms/support_graph.py:11-64 — computes the filter-support edges: optional FPS subsampling to `sample_n`
    points (stored as data.sample_idx), then the Euclidean radius graph of the samples with at most 512 neighbours and
    self loops; data.supp_edges is (E, 2) int64, rows (j, i) grouped by column 0.  Same constructor and call signature;
    runs on the device (data.pos must be a CUDA tensor)."""

    def __init__(self, epsilon, sample_n=None):
        self.epsilon = epsilon
        self.sample_n = sample_n

    def __call__(self, data):
        pos = data.pos
        if hasattr(data, "sample_idx"):
            sample_idx = data.sample_idx
        else:
            if self.sample_n is not None and not self.sample_n > pos.size(0):
                sample_idx = farthest_point_sample(pos, int(self.sample_n))
            else:
                sample_idx = torch.arange(pos.size(0), device=pos.device)
            data.sample_idx = sample_idx
        data.supp_edges = radius_graph(pos[sample_idx], self.epsilon, 512)      # indices into the samples, like original_idx[...]
        return data

    def __repr__(self):
        return '{}(epsilon={}, sample_n={})'.format(self.__class__.__name__, self.epsilon, self.sample_n)


def attach_plan(supp_sten, supp_edges, plan):
    """Remember that (supp_edges, supp_sten) are the dense form of `plan` (valid while neither is modified in place)."""
    supp_sten._fcb_plan = (plan, supp_edges.data_ptr(), supp_edges._version, supp_sten._version)


def attached_plan(supp_edges, supp_sten, n_rings, num_nodes):
    """The compact plan FCPrecomp attached to this very (supp_edges, supp_sten) pair, else None."""
    rec = getattr(supp_sten, "_fcb_plan", None)
    if rec is None or supp_edges is None:
        return None
    plan, ptr, v_edges, v_sten = rec
    if (supp_edges.data_ptr() != ptr or supp_edges._version != v_edges or supp_sten._version != v_sten or
            plan.n_rings != n_rings or plan.num_nodes != num_nodes):
        return None
    return plan
