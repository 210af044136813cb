"""Edge plans (K0): the device-side replacement of FCPrecomp's arithmetic and of the grouping
that scatter_add / autograd perform implicitly in the reference (transforms/fc_precomp.py:53-97,
nn/field_conv.py:134).  Built once per mesh, shared by every FieldConv layer, forward and backward."""
import os

import torch

from . import _lib


def _check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("fieldconv_b200: %s must be a CUDA tensor (no CPU path)" % name)


class Plan:
    """Compact plan: both CSR orders with 16-byte edge records + (cos,sin) of the log-map angle.

    Attributes are device tensors sized for the E edges given to build_plan; only the first
    ``rowptr_tgt[N]`` entries are meaningful (edges with r > epsilon are dropped on the device
    without a host sync, fc_precomp.py:67-74)."""

    dense = False

    def __init__(self, n, n_rings, epsilon, e_cap):
        self.num_nodes, self.n_rings, self.epsilon, self.e_cap = n, n_rings, epsilon, e_cap

    @property
    def num_edges(self):  # host sync
        return int(self.rowptr_tgt[-1].item())

    def tensors(self):
        """Every device tensor of the plan (e.g. to record_stream() them when the plan is built on a side stream)."""
        names = ("rowptr_tgt", "rowptr_src", "rec_tgt", "rec_src", "rot_tgt", "rot_src", "perm_tgt", "perm_src", "norms")
        return tuple(getattr(self, a) for a in names if getattr(self, a, None) is not None)

    def edges_by_target(self):
        """(j, i) of the kept edges in by-(target, ring) order — for index parity checks."""
        e = self.num_edges
        src = (self.rec_tgt[:e, 0] & ((1 << 27) - 1)).long()
        counts = (self.rowptr_tgt[1:] - self.rowptr_tgt[:-1]).long()
        tgt = torch.repeat_interleave(torch.arange(self.num_nodes, device=src.device), counts)
        return torch.stack((src, tgt), 1)

    def edges_by_source(self):
        e = self.num_edges
        tgt = (self.rec_src[:e, 0] & ((1 << 27) - 1)).long()
        counts = (self.rowptr_src[1:] - self.rowptr_src[:-1]).long()
        src = torch.repeat_interleave(torch.arange(self.num_nodes, device=tgt.device), counts)
        return torch.stack((src, tgt), 1)


class DensePlan:
    """CSR orders only; the stencil stays the caller's dense supp_sten (E,R,M)."""

    dense = True

    def __init__(self, n, e):
        self.num_nodes, self.e_cap = n, e

    @property
    def num_edges(self):
        return int(self.rowptr_tgt[-1].item())


def ring_radii(n_rings, device):
    # transforms/fc_precomp.py:12, evaluated with the same torch ops so the float32 values match
    return torch.sqrt(torch.div(torch.arange(n_rings, device=device), n_rings - 1)).float().contiguous()


def _validate_edges(supp_edges, n):
    """The reference fails with an index error on an edge endpoint outside [0, N) (nn/field_conv.py:130-134); the plan
    kernels drop such edges instead, so a corrupted or mis-offset batched supp_edges would silently become a smaller stencil.
    One reduction + host sync: opt-in (validate=True or FIELDCONV_B200_VALIDATE=1)."""
    if supp_edges.numel() and bool(((supp_edges < 0) | (supp_edges >= n)).any()):
        bad = int(((supp_edges < 0) | (supp_edges >= n)).any(dim=1).sum())
        raise IndexError("fieldconv_b200: %d of %d supp_edges rows have an endpoint outside [0, %d)" % (bad, supp_edges.shape[0], n))


def build_plan(supp_edges, logMag, logAng, xp, w, n_rings, epsilon, num_nodes=None, validate=None):
    """Compact plan from the attributes the reference's offline transforms store on `data`
    (supp_edges, logMag, logAng, xp, w — transforms/compute_log_xport.py:36-50) and FCPrecomp's
    (n_rings, epsilon).  Reproduces fc_precomp.py:67-95 on the device."""
    for t, name in ((supp_edges, "supp_edges"), (logMag, "logMag"), (logAng, "logAng"), (xp, "xp"), (w, "w")):
        _check_cuda(t, name)
    if logMag.dtype != torch.float32 or logAng.dtype != torch.float32 or w.dtype != torch.float32:
        raise TypeError("fieldconv_b200: logMag/logAng/w must be float32 (the reference's FCPrecomp is float32-only)")
    if n_rings < 2:
        raise ValueError("n_rings must be >= 2 (the reference divides by n_rings-1, fc_precomp.py:12)")
    dev = supp_edges.device
    n = int(w.shape[0]) if num_nodes is None else int(num_nodes)
    e = int(supp_edges.shape[0])
    edges = supp_edges.to(torch.int64).contiguous()
    if validate or (validate is None and os.environ.get("FIELDCONV_B200_VALIDATE", "0") == "1"):
        _validate_edges(edges, n)
    lm, la = logMag.contiguous(), logAng.contiguous()
    xpc = xp.to(torch.complex64).contiguous()
    wv = w.reshape(-1).contiguous()
    radii = ring_radii(n_rings, dev)
    p = Plan(n, n_rings, float(epsilon), e)
    cap = max(e, 1)
    p.rowptr_tgt = torch.empty(n + 1, dtype=torch.int32, device=dev)
    p.rowptr_src = torch.empty(n + 1, dtype=torch.int32, device=dev)
    p.rec_tgt = torch.zeros(cap, 4, dtype=torch.int32, device=dev)
    p.rec_src = torch.zeros(cap, 4, dtype=torch.int32, device=dev)
    p.rot_tgt = torch.zeros(cap, 2, dtype=torch.float32, device=dev)
    p.rot_src = torch.zeros(cap, 2, dtype=torch.float32, device=dev)
    p.perm_tgt = torch.full((cap,), -1, dtype=torch.int32, device=dev)
    p.perm_src = torch.full((cap,), -1, dtype=torch.int32, device=dev)
    nbytes = _lib.query_bytes("fcb_plan_workspace_bytes", e, n, n_rings)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fcb_plan_build", edges.data_ptr(), lm.data_ptr(), la.data_ptr(), torch.view_as_real(xpc).data_ptr(),
                  wv.data_ptr(), radii.data_ptr(), float(epsilon), e, n, n_rings,
                  p.rowptr_tgt.data_ptr(), p.rec_tgt.data_ptr(), p.rot_tgt.data_ptr(), p.perm_tgt.data_ptr(),
                  p.rowptr_src.data_ptr(), p.rec_src.data_ptr(), p.rot_src.data_ptr(), p.perm_src.data_ptr(),
                  ws.data_ptr(), nbytes, _lib.stream_ptr())
        # [0]: max over targets of sum_e |wxp_e| (by-target order), [1]: the same over sources — the a-priori bounds
        # max|contrib| <= norms[0] max|x| and max|G| <= norms[1] max|gy| that scale the packed fp16 operand planes
        p.norms = torch.zeros(2, dtype=torch.float32, device=dev)
        _lib.call("fcb_plan_norm", p.rowptr_tgt.data_ptr(), p.rec_tgt.data_ptr(), n, p.norms[0:].data_ptr(), _lib.stream_ptr())
        _lib.call("fcb_plan_norm", p.rowptr_src.data_ptr(), p.rec_src.data_ptr(), n, p.norms[1:].data_ptr(), _lib.stream_ptr())
    return p


def build_dense_plan(supp_edges, num_nodes):
    _check_cuda(supp_edges, "supp_edges")
    dev = supp_edges.device
    n, e = int(num_nodes), int(supp_edges.shape[0])
    edges = supp_edges.to(torch.int64).contiguous()
    p = DensePlan(n, e)
    cap = max(e, 1)
    for name in ("nbr_tgt", "perm_tgt", "nbr_src", "perm_src"):
        setattr(p, name, torch.zeros(cap, dtype=torch.int32, device=dev))
    p.rowptr_tgt = torch.empty(n + 1, dtype=torch.int32, device=dev)
    p.rowptr_src = torch.empty(n + 1, dtype=torch.int32, device=dev)
    nbytes = _lib.query_bytes("fcb_plan_dense_workspace_bytes", e, n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fcb_plan_build_dense", edges.data_ptr(), e, n, p.rowptr_tgt.data_ptr(), p.nbr_tgt.data_ptr(),
                  p.perm_tgt.data_ptr(), p.rowptr_src.data_ptr(), p.nbr_src.data_ptr(), p.perm_src.data_ptr(),
                  ws.data_ptr(), nbytes, _lib.stream_ptr())
    return p
