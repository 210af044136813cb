"""ECHO descriptors / ECHOBlock behind the reference's module API:

    ECHO(channels, n_bins=2).forward(x, supp_edges, ln, wxp) -> (N, channels, hdim) float32          nn/echo.py:64-148
    ECHOBlock(in_channels, out_channels, n_des=None, n_bins=3, band_limit=1, n_rings=6, ftype=1)
        .forward(x, supp_edges, supp_sten, ln, wxp)                                                   nn/echo_block.py:21-103

The per-(edge, channel) bilinear histogram votes (four scatter_adds with float atomics in the reference, after a nonzero()
compaction that syncs the host) run as one deterministic CUDA kernel per direction (csrc/echo.cu, custom op
``fieldconv_b200::echo`` with its backward).  ECHOBlock's MLP / residual are the reference's plain ``nn.Linear`` layers.
"""
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib
from .plan import DensePlan


def disk_map(n_bins):
    """nn/echo.py:11-27 — raster cell -> histogram bin (cells outside the disk keep the fill value 0) and the bin count."""
    w = 2 * n_bins + 1
    ind = [w * i + j for i in range(w) for j in range(w) if (i - n_bins) ** 2 + (j - n_bins) ** 2 <= (n_bins + 0.25) ** 2]
    d_map = torch.zeros(w * w, dtype=torch.long)
    d_map[torch.tensor(ind)] = torch.arange(len(ind))
    return d_map, len(ind)


def hist_dim(n_bins):
    """nn/echo_block.py:9-19."""
    return disk_map(n_bins)[1]


@torch.library.custom_op("fieldconv_b200::echo", mutates_args=())
def echo_fwd(x: Tensor, ln: Tensor, wxp: Tensor, dmap: Tensor, rowptr_tgt: Tensor, nbr_tgt: Tensor, perm_tgt: Tensor,
             rowptr_src: Tensor, nbr_src: Tensor, perm_src: Tensor, n_bins: int, hdim: int) -> Tuple[Tensor, Tensor]:
    if not x.is_cuda:
        raise RuntimeError("fieldconv_b200: ECHO runs on CUDA tensors only — there is no CPU path")
    for t, name in ((x, "x"), (ln, "ln"), (wxp, "wxp")):
        if t.dtype != torch.complex64:
            raise TypeError("fieldconv_b200: %s must be complex64" % name)
    x, ln, wxp = x.contiguous(), ln.contiguous(), wxp.contiguous()
    n, c = x.shape
    hist = torch.empty(n, c, hdim, dtype=torch.complex64, device=x.device)
    out = torch.empty(n, c, hdim, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("fcb_echo_fwd_f32", torch.view_as_real(x).data_ptr(), torch.view_as_real(ln).data_ptr(),
                  torch.view_as_real(wxp).data_ptr(), rowptr_tgt.data_ptr(), nbr_tgt.data_ptr(), perm_tgt.data_ptr(),
                  dmap.data_ptr(), torch.view_as_real(hist).data_ptr(), out.data_ptr(), n, c, n_bins, hdim, _lib.stream_ptr())
    return out, hist


@echo_fwd.register_fake
def _(x, ln, wxp, dmap, rowptr_tgt, nbr_tgt, perm_tgt, rowptr_src, nbr_src, perm_src, n_bins, hdim):
    n, c = x.shape
    return x.new_empty(n, c, hdim, dtype=torch.float32), x.new_empty(n, c, hdim)


@torch.library.custom_op("fieldconv_b200::echo_bwd", mutates_args=())
def echo_bwd(x: Tensor, ln: Tensor, wxp: Tensor, dmap: Tensor, hist: Tensor, g_out: Tensor, rowptr_src: Tensor, nbr_src: Tensor,
             perm_src: Tensor, n_bins: int, hdim: int) -> Tensor:
    x, ln, wxp, hist, g_out = x.contiguous(), ln.contiguous(), wxp.contiguous(), hist.contiguous(), g_out.contiguous().float()
    n, c = x.shape
    gx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.call("fcb_echo_bwd_f32", torch.view_as_real(x).data_ptr(), torch.view_as_real(ln).data_ptr(),
                  torch.view_as_real(wxp).data_ptr(), rowptr_src.data_ptr(), nbr_src.data_ptr(), perm_src.data_ptr(),
                  dmap.data_ptr(), torch.view_as_real(hist).data_ptr(), g_out.data_ptr(), torch.view_as_real(gx).data_ptr(),
                  n, c, n_bins, hdim, _lib.stream_ptr())
    return gx


@echo_bwd.register_fake
def _(x, ln, wxp, dmap, hist, g_out, rowptr_src, nbr_src, perm_src, n_bins, hdim):
    return torch.empty_like(x)


def _echo_setup(ctx, inputs, output):
    x, ln, wxp, dmap, _, _, _, rowptr_src, nbr_src, perm_src, n_bins, hdim = inputs
    ctx.save_for_backward(x, ln, wxp, dmap, output[1], rowptr_src, nbr_src, perm_src)
    ctx.cfg = (n_bins, hdim)
    ctx.set_materialize_grads(False)


def _echo_backward(ctx, g_out, _g_hist):
    if g_out is None or not ctx.needs_input_grad[0]:
        return (None,) * 12
    x, ln, wxp, dmap, hist, rowptr_src, nbr_src, perm_src = ctx.saved_tensors
    return (echo_bwd(x, ln, wxp, dmap, hist, g_out, rowptr_src, nbr_src, perm_src, *ctx.cfg),) + (None,) * 11


echo_fwd.register_autograd(_echo_backward, setup_context=_echo_setup)


class ECHO(nn.Module):
    def __init__(self, channels, n_bins=2):
        super().__init__()
        if not 1 <= n_bins <= 3:
            raise ValueError("fieldconv_b200.ECHO supports n_bins 1..3 (the reference's nets use 2 and 3)")
        self.channels, self.n_bins = channels, n_bins
        d_map, dim = disk_map(n_bins)
        self.register_buffer("dMap", d_map)
        self.hdim = dim

    def forward(self, x, supp_edges, ln, wxp, *, plan=None):
        """x (N, C) complex64, supp_edges (E, 2) int64 rows (j, i), ln / wxp (E,) complex64 -> (N, C, hdim) float32."""
        from .nn import shared_dense_plan
        if not x.is_cuda:
            raise RuntimeError("fieldconv_b200.ECHO runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.shape[1] != self.channels:
            raise ValueError("expected %d channels, got %d" % (self.channels, x.shape[1]))
        if ln.shape[0] != supp_edges.shape[0] or wxp.shape[0] != supp_edges.shape[0]:
            raise ValueError("ln / wxp must have one entry per support edge")
        dp = plan if isinstance(plan, DensePlan) else shared_dense_plan(supp_edges, x.shape[0])
        if dp.num_nodes != x.shape[0] or dp.e_cap != supp_edges.shape[0]:
            raise ValueError("the dense plan does not match x / supp_edges")
        dmap = self.dMap.to(device=x.device, dtype=torch.int32)
        out, _ = echo_fwd(x, ln, wxp, dmap, dp.rowptr_tgt, dp.nbr_tgt, dp.perm_tgt, dp.rowptr_src, dp.nbr_src, dp.perm_src,
                          self.n_bins, self.hdim)
        return out


class ECHOBlock(nn.Module):
    """nn/echo_block.py:21-103 — FieldConv -> TangentNonLin -> ECHO -> 3-layer MLP, plus a linear residual of |x|."""

    def __init__(self, in_channels, out_channels, n_des=None, n_bins=3, band_limit=1, n_rings=6, ftype=1):
        super().__init__()
        from .nn import FieldConv, TangentNonLin
        if n_des is None:
            n_des = in_channels
        self.conv = FieldConv(in_channels, n_des, band_limit, n_rings, ftype)
        self.nonlin = TangentNonLin(in_channels)
        self.echo = ECHO(n_des, n_bins)
        mid_channels = n_des * hist_dim(n_bins)
        self.lin1 = nn.Linear(mid_channels, 128)
        self.lin2 = nn.Linear(128, 64)
        self.lin3 = nn.Linear(64, out_channels)
        self.res = nn.Linear(in_channels, out_channels)

    def forward(self, x, supp_edges, supp_sten, ln, wxp):
        x_e = self.nonlin(self.conv(x, supp_edges, supp_sten))
        x_e = self.echo(x_e, supp_edges, ln, wxp)
        x_e = torch.reshape(x_e, (x_e.size(0), -1))
        x_e = F.relu(self.lin1(x_e))
        x_e = F.relu(self.lin2(x_e))
        mag = x.abs()
        mag = torch.where((x.real.abs() < 1e-7) & (x.imag.abs() < 1e-7), torch.zeros_like(mag), mag)      # softAbs, utils/field.py:29-37
        return self.lin3(x_e) + self.res(mag)
