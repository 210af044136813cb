"""TransField / LiftBlock behind the reference's module API:

    TransField(in_channels, out_channels, n_rings=6, ftype=1).forward(x, supp_edges, lift_sten)      nn/trans_field.py:27-113
    LiftBlock(in_channels, out_channels, n_rings=6, ftype=1).forward(x, supp_edges, lift_sten)       nn/lift_block.py:6-55

Parameter names, shapes and initialisation are the reference's (zonalAng, zonalMag, phase).  The two scatter_adds over the
support edges (nn/trans_field.py:104-110) run as deterministic CSR segmented reductions in CUDA (csrc/lift.cu, custom op
``fieldconv_b200::lift_aggregate`` with its adjoint for grad x); the per-vertex weighting (nn/trans_field.py:10-25) is a few
tiny torch ops on (N, Co, Ci) tensors in the trig-free form  y = sum_c |m| (a / |a|) e^{i phase}  (a / |a| := 1 at origin
entries — softAngle, utils/field.py:40-48), so parameter gradients come from autograd exactly as in the reference.
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import Parameter

from . import _lib
from .plan import DensePlan


@torch.library.custom_op("fieldconv_b200::lift_aggregate", mutates_args=())
def lift_aggregate(x: Tensor, lift_sten: Tensor, rowptr_tgt: Tensor, nbr_tgt: Tensor, perm_tgt: Tensor, rowptr_src: Tensor,
                   nbr_src: Tensor, perm_src: Tensor) -> Tuple[Tensor, Tensor]:
    if not x.is_cuda:
        raise RuntimeError("fieldconv_b200: TransField runs on CUDA tensors only — there is no CPU path")
    if x.dtype != torch.float32 or lift_sten.dtype != torch.complex64:
        raise TypeError("fieldconv_b200: TransField needs float32 features and a complex64 stencil")
    x, lift_sten = x.contiguous(), lift_sten.contiguous()
    n, ci = x.shape
    r = lift_sten.shape[1]
    agg = torch.empty(n, ci + 1, r, dtype=torch.complex64, device=x.device)
    mag = torch.empty(n, ci, r, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("fcb_lift_aggregate_f32", x.data_ptr(), torch.view_as_real(lift_sten).data_ptr(), rowptr_tgt.data_ptr(),
                  nbr_tgt.data_ptr(), perm_tgt.data_ptr(), torch.view_as_real(agg).data_ptr(), mag.data_ptr(), n, ci, r,
                  _lib.stream_ptr())
    return agg, mag


@lift_aggregate.register_fake
def _(x, lift_sten, rowptr_tgt, nbr_tgt, perm_tgt, rowptr_src, nbr_src, perm_src):
    n, ci = x.shape
    r = lift_sten.shape[1]
    return lift_sten.new_empty(n, ci + 1, r), x.new_empty(n, ci, r)


@torch.library.custom_op("fieldconv_b200::lift_aggregate_bwd", mutates_args=())
def lift_aggregate_bwd(g_agg: Tensor, g_mag: Tensor, agg: Tensor, lift_sten: Tensor, rowptr_src: Tensor, nbr_src: Tensor,
                       perm_src: Tensor) -> Tensor:
    g_agg, g_mag, agg, lift_sten = g_agg.contiguous(), g_mag.contiguous(), agg.contiguous(), lift_sten.contiguous()
    n, c1, r = g_agg.shape
    gx = torch.empty(n, c1 - 1, dtype=torch.float32, device=g_agg.device)
    with torch.cuda.device(g_agg.device):
        _lib.call("fcb_lift_aggregate_bwd_f32", torch.view_as_real(g_agg).data_ptr(), g_mag.data_ptr(),
                  torch.view_as_real(agg).data_ptr(), torch.view_as_real(lift_sten).data_ptr(), rowptr_src.data_ptr(), nbr_src.data_ptr(), perm_src.data_ptr(),
                  gx.data_ptr(), n, c1 - 1, r, _lib.stream_ptr())
    return gx


@lift_aggregate_bwd.register_fake
def _(g_agg, g_mag, agg, lift_sten, rowptr_src, nbr_src, perm_src):
    return g_mag.new_empty(g_agg.shape[0], g_agg.shape[1] - 1)


def _la_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1], inputs[5], inputs[6], inputs[7], output[0])
    ctx.shape = (inputs[0].shape[0], inputs[0].shape[1], inputs[1].shape[1])
    ctx.set_materialize_grads(False)


def _la_backward(ctx, g_agg, g_mag):
    lift_sten, rowptr_src, nbr_src, perm_src, agg = ctx.saved_tensors
    if not ctx.needs_input_grad[0] or (g_agg is None and g_mag is None):
        return (None,) * 8
    n, ci, r = ctx.shape
    if g_agg is None:
        g_agg = torch.zeros(n, ci + 1, r, dtype=torch.complex64, device=lift_sten.device)
    if g_mag is None:
        g_mag = torch.zeros(n, ci, r, dtype=torch.float32, device=lift_sten.device)
    return (lift_aggregate_bwd(g_agg, g_mag, agg, lift_sten, rowptr_src, nbr_src, perm_src),) + (None,) * 7


lift_aggregate.register_autograd(_la_backward, setup_context=_la_setup)


class TransField(nn.Module):
    def __init__(self, in_channels, out_channels, n_rings=6, ftype=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.R, self.ftype = n_rings, ftype
        self.zonalAng = Parameter(torch.empty(out_channels, in_channels, n_rings))
        self.zonalMag = Parameter(torch.empty(out_channels, in_channels, n_rings))
        if ftype == 0:
            self.register_buffer("phase", torch.zeros(out_channels, in_channels))
        else:
            self.phase = Parameter(torch.empty(out_channels, in_channels))
            nn.init.xavier_uniform_(self.phase)
        nn.init.xavier_uniform_(self.zonalAng)
        nn.init.xavier_uniform_(self.zonalMag)

    def forward(self, x, supp_edges, lift_sten, *, plan=None):
        """x (N, Ci) float32, supp_edges (E, 2) int64 rows (j, i), lift_sten (E, R, 2) complex64 -> (N, Co) complex64.
        `plan`: an optional DensePlan of supp_edges to reuse its CSR orders (else the shared per-tensor cache)."""
        from .nn import shared_dense_plan
        if not x.is_cuda:
            raise RuntimeError("fieldconv_b200.TransField runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.shape[1] != self.in_channels:
            raise ValueError("expected %d input channels, got %d" % (self.in_channels, x.shape[1]))
        if tuple(lift_sten.shape[1:]) != (self.R, 2) or lift_sten.shape[0] != supp_edges.shape[0]:
            raise ValueError("lift_sten must be (E=%d, %d, 2), got %s" % (supp_edges.shape[0], self.R, tuple(lift_sten.shape)))
        dp = plan if isinstance(plan, DensePlan) else shared_dense_plan(supp_edges, x.shape[0])
        if dp.num_nodes != x.shape[0] or dp.e_cap != lift_sten.shape[0]:
            raise ValueError("the dense plan does not match x / lift_sten")
        agg, mag = lift_aggregate(x.float(), lift_sten, dp.rowptr_tgt, dp.nbr_tgt, dp.perm_tgt, dp.rowptr_src, dp.nbr_src,
                                  dp.perm_src)
        ci = self.in_channels
        a_ring = -agg[:, :ci, :]                 # contribAng = -sum (x_j - x_i) s1  (nn/trans_field.py:104-106), (N, Ci, R) complex
        # the reference's own broadcast-multiply-and-sum over the rings (elementwise kernels + a reduction: plain fp32
        # adds, no library GEMM whose reduction order / precision mode could differ from the reference's)
        a = (a_ring[:, None] * self.zonalAng[None]).sum(dim=3)                                       # :12 / :19 before softAngle
        m = (mag[:, None] * self.zonalMag[None]).sum(dim=3).abs()                                    # softAbsolute(:14 / :21)
        origin = (a.real.abs() < 1e-7) & (a.imag.abs() < 1e-7)                                       # utils/field.py:14-16
        mod = a.abs()
        unit = torch.where(origin, torch.ones_like(a), a / torch.where(origin, torch.ones_like(mod), mod))
        if self.ftype == 1:
            unit = unit * torch.polar(torch.ones_like(self.phase), self.phase)[None]                 # :19
        return (m * unit).sum(dim=-1)                                                                # :16 / :23


class LiftBlock(nn.Module):
    """nn/lift_block.py:6-55 — TransField followed by the modReLU non-linearity."""

    def __init__(self, in_channels, out_channels, n_rings=6, ftype=1):
        super().__init__()
        from .nn import TangentNonLin
        self.field = TransField(in_channels, out_channels, n_rings=n_rings, ftype=ftype)
        self.nonlin = TangentNonLin(out_channels)

    def forward(self, x, supp_edges, lift_sten, *, plan=None):
        return self.nonlin(self.field(x, supp_edges, lift_sten, plan=plan))
