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
