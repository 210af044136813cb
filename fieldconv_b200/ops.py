"""PyTorch custom ops over the C ABI (torch is plumbing: device memory, streams, autograd graph).

``fieldconv_b200::fc_fwd / fc_bwd``          compact plan (fast path)
``fieldconv_b200::fc_fwd_dense / _bwd_dense`` dense supp_sten (E,R,M), exact drop-in for
                                             nn/field_conv.py:104 of the reference
``fieldconv_b200::modrelu(_bwd)``            TangentNonLin (nn/tangent_nonlin.py:24-35)
``fieldconv_b200::gemm``                     real fp32 GEMM used by TangentLin (nn/tangent_lin.py:27-29)

The folded filter W (Co,Ci,R,M) complex is an op INPUT: it is built from (zonal, spherical, phase)
with ordinary differentiable torch ops (tiny tensors), so parameter gradients flow from the op's
gW through autograd exactly as in the reference (nn/field_conv.py:10-33).
"""
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

# The backward takes the weight gradient from G and xhat (gW[o,c,r,m] = sum_j conj(xhat[j,c,m]) G[j,m,r,o], csrc/api.cu
# backward_common), so the forward keeps NOTHING of size N x K by default: no saved contrib, no recompute.
# FIELDCONV_B200_SAVE_CONTRIB=1 restores the older behaviour (contrib saved by the forward, gW = contrib^H gy) for A/B runs.
SAVE_CONTRIB = os.environ.get("FIELDCONV_B200_SAVE_CONTRIB", "0") == "1"


def keep_contrib_default(nbytes=0, device=None):
    return SAVE_CONTRIB


# ---------------------------------------------------------------------------- operand bounds (struct fcb_bounds)
# The 2xFP16 tensor-core paths scale an operand by a power of two taken from a bound of its largest magnitude.  The kernels
# that WRITE a tensor can report that bound for free (the fused block epilogue for the activation, the modReLU backward for
# the gradient it returns); it travels with the tensor object as an attribute, tagged with the tensor's version counter,
# and the consumers hand it to the library instead of letting every call take its own pass over the operand.  A tensor
# without a bound gets ONE pass (bound_of) shared by all its consumers.  FIELDCONV_B200_BOUNDS=0: every call computes its own.
BOUNDS = os.environ.get("FIELDCONV_B200_BOUNDS", "1") == "1"


def set_bound(t, b):
    """Attach to t the device scalar b >= max_i |t_i| (complex modulus / largest |component|) its producer reported."""
    if BOUNDS and b is not None:
        t._fcb_bound = (b, t._version)
    return t


def peek_bound(t):
    rec = getattr(t, "_fcb_bound", None) if BOUNDS else None
    if rec is not None and rec[1] == t._version and rec[0].device == t.device:
        return rec[0]
    return None


def bound_of(t):
    """The bound attached to the complex64 tensor t, computed (one pass, fcb_bound_f32) and attached when there is none."""
    if not BOUNDS or not t.is_cuda or t.dtype != torch.complex64 or not t.is_contiguous():
        return None
    b = peek_bound(t)
    if b is None:
        b = torch.empty(1, dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            _lib.call("fcb_bound_f32", _real(t).data_ptr(), t.numel(), b.data_ptr(), _lib.stream_ptr())
        set_bound(t, b)
    return b


def _uses_bounds(flags):
    return BOUNDS and (flags & _lib.GEMM_MASK) == _lib.GEMM_TC_2XF16


def _real(t):
    return torch.view_as_real(t)


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


def _check(x, name, dtype=torch.complex64):
    if not x.is_cuda:
        raise RuntimeError("fieldconv_b200: %s must be a CUDA tensor — there is no CPU path" % name)
    if x.dtype != dtype:
        raise TypeError("fieldconv_b200: %s must be %s, got %s" % (name, dtype, x.dtype))


# --------------------------------------------------------------------------- compact plan ops
def _padded_rows(n):
    return (n + 127) // 128 * 128


@torch.library.custom_op("fieldconv_b200::fc_fwd", mutates_args=())
def fc_fwd(x: Tensor, W: Tensor, rowptr_tgt: Tensor, rec_tgt: Tensor, rot_tgt: Tensor, rowptr_src: Tensor,
           rec_src: Tensor, rot_src: Tensor, norms: Tensor, band_limit: int, n_rings: int, flags: int,
           keep_contrib: bool, x_bound: Optional[Tensor] = None, w_bound: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    # the by-source plan tensors are unused here; they are inputs so autograd can hand them to fc_bwd
    _check(x, "x")
    _check(W, "W")
    x, W = x.contiguous(), W.contiguous()
    n, ci = x.shape
    co = W.shape[0]
    k = n_rings * ci * (2 * band_limit + 1)
    packed = bool(flags & _lib.FLAG_PACKED)
    fused = bool(flags & _lib.FLAG_FUSED) and not keep_contrib
    cflags = flags & ~(_lib.FLAG_PACKED | _lib.FLAG_FUSED | _lib.FLAG_PACKED_G)
    y = torch.empty(n, co, dtype=torch.complex64, device=x.device)
    if fused:      # band_limit <= 1: one kernel, contrib never exists in device memory (csrc/fused_fwd.cu)
        nbytes = _lib.query_bytes("fcb_fwd_fused_workspace_bytes", ci, co, band_limit, n_rings)
        ws = _ws(nbytes, x.device)
        with torch.cuda.device(x.device):
            _lib.call("fcb_fwd_fused_f32", _real(x).data_ptr(), _real(W).data_ptr(), rowptr_tgt.data_ptr(), rec_tgt.data_ptr(),
                      rot_tgt.data_ptr(), norms.data_ptr(), _real(y).data_ptr(), n, n, ci, co, band_limit, n_rings,
                      ws.data_ptr(), nbytes, _lib.stream_ptr())
        return y, torch.empty(0, dtype=torch.complex64, device=x.device), torch.zeros(1, dtype=torch.float32, device=x.device)
    # rows padded to whole 128-row tiles: the packed (PK) layout needs them, the fp32 layout ignores the tail
    contrib = torch.empty(_padded_rows(n), k, dtype=torch.complex64, device=x.device)
    cmax = torch.zeros(1, dtype=torch.float32, device=x.device)     # max|contrib| (or its bound): operand scale of the 2xFP16 contraction
    nbytes = _lib.query_bytes("fcb_fwd_workspace_bytes", n, ci, co, band_limit, n_rings, cflags)
    ws = _ws(nbytes, x.device)
    with torch.cuda.device(x.device):
        if packed:
            _lib.call("fcb_fwd_pk_f32", _real(x).data_ptr(), _real(W).data_ptr(), rowptr_tgt.data_ptr(), rec_tgt.data_ptr(),
                      rot_tgt.data_ptr(), norms.data_ptr(), _real(y).data_ptr(), _real(contrib).data_ptr(), cmax.data_ptr(),
                      _lib.bounds(x=x_bound, w=w_bound), n, ci, co, band_limit, n_rings, cflags, ws.data_ptr(), nbytes,
                      _lib.stream_ptr())
        elif w_bound is not None:     # fcb_fwd_f32 with the filter bound: the epilogue-less form of fcb_fwd_act_f32
            _lib.call("fcb_fwd_act_f32", _real(x).data_ptr(), _real(W).data_ptr(), rowptr_tgt.data_ptr(), rec_tgt.data_ptr(),
                      rot_tgt.data_ptr(), _real(y).data_ptr(), _real(contrib).data_ptr(), cmax.data_ptr(), None, None, None,
                      _lib.bounds(w=w_bound), n, ci, co, band_limit, n_rings, cflags, ws.data_ptr(), nbytes, _lib.stream_ptr())
        else:
            _lib.call("fcb_fwd_f32", _real(x).data_ptr(), _real(W).data_ptr(), rowptr_tgt.data_ptr(), rec_tgt.data_ptr(),
                      rot_tgt.data_ptr(), _real(y).data_ptr(), _real(contrib).data_ptr(), cmax.data_ptr(), n, ci, co,
                      band_limit, n_rings, cflags, ws.data_ptr(), nbytes, _lib.stream_ptr())
    if not keep_contrib:
        contrib = torch.empty(0, dtype=torch.complex64, device=x.device)
    return y, contrib, cmax


@fc_fwd.register_fake
def _(x, W, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit, n_rings, flags, keep_contrib,
      x_bound=None, w_bound=None):
    n, ci = x.shape
    k = n_rings * ci * (2 * band_limit + 1)
    return (x.new_empty(n, W.shape[0]), x.new_empty((_padded_rows(n), k) if keep_contrib else (0,)),
            x.new_empty(1, dtype=torch.float32))


@torch.library.custom_op("fieldconv_b200::fc_bwd", mutates_args=())
def fc_bwd(x: Tensor, W: Tensor, gy: Tensor, contrib: Tensor, cmax: Tensor, rowptr_tgt: Tensor, rec_tgt: Tensor, rot_tgt: Tensor,
           rowptr_src: Tensor, rec_src: Tensor, rot_src: Tensor, norms: Tensor, band_limit: int, n_rings: int, flags: int,
           need_gx: bool, need_gw: bool, x_bound: Optional[Tensor] = None, gy_bound: Optional[Tensor] = None,
           w_bound: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    _check(gy, "grad_output")
    x, W, gy = x.contiguous(), W.contiguous(), gy.contiguous()
    n, ci = x.shape
    co = W.shape[0]
    have_contrib = contrib.numel() > 0
    # FLAG_PACKED_G: the forward ran the fp32 operand layout, the backward packs G (only meaningful when no fp32 contrib
    # was kept: the packed entry point cannot read one)
    packed = bool(flags & _lib.FLAG_PACKED) or (bool(flags & _lib.FLAG_PACKED_G) and not have_contrib)
    cflags = flags & ~(_lib.FLAG_PACKED | _lib.FLAG_FUSED | _lib.FLAG_PACKED_G)
    gx = torch.empty_like(x) if need_gx else torch.empty(0, dtype=x.dtype, device=x.device)
    gw = torch.empty_like(W) if need_gw else torch.empty(0, dtype=W.dtype, device=x.device)
    nbytes = _lib.query_bytes("fcb_bwd_workspace_bytes", n, ci, co, band_limit, n_rings,
                              cflags | (_lib.FLAG_HAVE_CONTRIB if have_contrib else 0))
    ws = _ws(nbytes, x.device)
    bnd = _lib.bounds(x=x_bound, gy=gy_bound, w=w_bound)
    with torch.cuda.device(x.device):
        if packed:
            _lib.call("fcb_bwd_pk_f32", _real(x).data_ptr(), _real(W).data_ptr(), _real(gy).data_ptr(),
                      _real(contrib).data_ptr() if have_contrib else 0, cmax.data_ptr() if have_contrib else 0,
                      rowptr_tgt.data_ptr(), rec_tgt.data_ptr(), rot_tgt.data_ptr(), norms[0:].data_ptr(),
                      rowptr_src.data_ptr(), rec_src.data_ptr(), rot_src.data_ptr(), norms[1:].data_ptr(),
                      _real(gx).data_ptr() if need_gx else 0, _real(gw).data_ptr() if need_gw else 0, bnd,
                      n, ci, co, band_limit, n_rings, cflags, ws.data_ptr(), nbytes, _lib.stream_ptr())
        else:
            _lib.call("fcb_bwd_f32", _real(x).data_ptr(), _real(W).data_ptr(), _real(gy).data_ptr(),
                      _real(contrib).data_ptr() if have_contrib else 0, cmax.data_ptr() if have_contrib else 0,
                      rowptr_tgt.data_ptr(), rec_tgt.data_ptr(), rot_tgt.data_ptr(),
                      rowptr_src.data_ptr(), rec_src.data_ptr(), rot_src.data_ptr(),
                      _real(gx).data_ptr() if need_gx else 0, _real(gw).data_ptr() if need_gw else 0, bnd,
                      n, ci, co, band_limit, n_rings, cflags, ws.data_ptr(), nbytes, _lib.stream_ptr())
    return gx, gw


@fc_bwd.register_fake
def _(x, W, gy, contrib, cmax, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit, n_rings, flags,
      need_gx, need_gw, x_bound=None, gy_bound=None, w_bound=None):
    return (torch.empty_like(x) if need_gx else x.new_empty(0)), (torch.empty_like(W) if need_gw else W.new_empty(0))


def _fc_setup(ctx, inputs, output):
    (x, W, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit, n_rings, flags, keep, x_bound,
     w_bound) = inputs
    _, contrib, cmax = output
    ctx.save_for_backward(x, W, contrib, cmax, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms)
    ctx.x_bound, ctx.w_bound = x_bound, w_bound      # 1-element device scalars, not part of the graph
    ctx.cfg = (band_limit, n_rings, flags)
    ctx.set_materialize_grads(False)      # no N*K zero tensor for the unused contrib output


def _fc_backward(ctx, gy, _gcontrib, _gcmax):
    if gy is None:
        return (None,) * 15
    x, W, contrib, cmax, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms = ctx.saved_tensors
    band_limit, n_rings, flags = ctx.cfg
    gy = gy.contiguous()
    gx, gw = fc_bwd(x, W, gy, contrib, cmax, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit,
                    n_rings, flags, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.x_bound,
                    peek_bound(gy) if _uses_bounds(flags) else None, ctx.w_bound)
    return (gx if ctx.needs_input_grad[0] else None, gw if ctx.needs_input_grad[1] else None) + (None,) * 13


fc_fwd.register_autograd(_fc_backward, setup_context=_fc_setup)


# --------------------------------------------------------------------------- layer + block epilogue in one op
@torch.library.custom_op("fieldconv_b200::fc_fwd_act", mutates_args=())
def fc_fwd_act(x: Tensor, W: Tensor, res: Tensor, bias: Tensor, rowptr_tgt: Tensor, rec_tgt: Tensor, rot_tgt: Tensor,
               rowptr_src: Tensor, rec_src: Tensor, rot_src: Tensor, norms: Tensor, band_limit: int, n_rings: int, flags: int,
               has_res: bool, x_bound: Optional[Tensor] = None, w_bound: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(act, z, bound) with z = FieldConv(x) (+ res) and act = modReLU(z, bias): nn/fc_resnet_block.py:84-88 with the
    TangentNonLin (and the residual add) applied in the contraction kernel's epilogue (fcb_fwd_act_f32 / fcb_fwd_act_pk_f32);
    bound = max_i |act_i|, reported by that epilogue.  Nothing of size N x K is kept: the backward is modrelu_bwd + fc_bwd with
    gW from G and xhat."""
    _check(x, "x")
    _check(W, "W")
    x, W = x.contiguous(), W.contiguous()
    n, ci = x.shape
    co = W.shape[0]
    k = n_rings * ci * (2 * band_limit + 1)
    packed = bool(flags & _lib.FLAG_PACKED)
    cflags = flags & ~(_lib.FLAG_PACKED | _lib.FLAG_FUSED | _lib.FLAG_PACKED_G)
    z = torch.empty(n, co, dtype=torch.complex64, device=x.device)
    act = torch.empty(n, co, dtype=torch.complex64, device=x.device)
    b = bias.reshape(-1).contiguous().float()
    r_ptr = 0
    if has_res:
        _check(res, "res")
        res = res.contiguous()
        r_ptr = _real(res).data_ptr()
    contrib = torch.empty(_padded_rows(n), k, dtype=torch.complex64, device=x.device)      # transient: freed on return
    cmax = torch.zeros(1, dtype=torch.float32, device=x.device)
    act_bound = torch.empty(1, dtype=torch.float32, device=x.device)
    nbytes = _lib.query_bytes("fcb_fwd_workspace_bytes", n, ci, co, band_limit, n_rings, cflags)
    ws = _ws(nbytes, x.device)
    bnd = _lib.bounds(x=x_bound, act=act_bound, w=w_bound)
    with torch.cuda.device(x.device):
        if packed:
            _lib.call("fcb_fwd_act_pk_f32", _real(x).data_ptr(), _real(W).data_ptr(), rowptr_tgt.data_ptr(), rec_tgt.data_ptr(),
                      rot_tgt.data_ptr(), norms.data_ptr(), _real(z).data_ptr(), _real(contrib).data_ptr(), cmax.data_ptr(),
                      r_ptr, b.data_ptr(), _real(act).data_ptr(), bnd, n, ci, co, band_limit, n_rings, cflags, ws.data_ptr(), nbytes,
                      _lib.stream_ptr())
        else:
            _lib.call("fcb_fwd_act_f32", _real(x).data_ptr(), _real(W).data_ptr(), rowptr_tgt.data_ptr(), rec_tgt.data_ptr(),
                      rot_tgt.data_ptr(), _real(z).data_ptr(), _real(contrib).data_ptr(), cmax.data_ptr(), r_ptr, b.data_ptr(),
                      _real(act).data_ptr(), bnd, n, ci, co, band_limit, n_rings, cflags, ws.data_ptr(), nbytes, _lib.stream_ptr())
    return act, z, act_bound


@fc_fwd_act.register_fake
def _(x, W, res, bias, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit, n_rings, flags, has_res,
      x_bound=None, w_bound=None):
    return x.new_empty(x.shape[0], W.shape[0]), x.new_empty(x.shape[0], W.shape[0]), x.new_empty(1, dtype=torch.float32)


def _fa_setup(ctx, inputs, output):
    (x, W, res, bias, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit, n_rings, flags, has_res,
     x_bound, w_bound) = inputs
    ctx.save_for_backward(x, W, bias, output[1], rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms)
    ctx.x_bound, ctx.w_bound = x_bound, w_bound
    ctx.cfg = (band_limit, n_rings, flags, has_res)
    ctx.set_materialize_grads(False)


def _fa_backward(ctx, g_act, g_z, _g_bound):
    x, W, bias, z, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms = ctx.saved_tensors
    band_limit, n_rings, flags, has_res = ctx.cfg
    if g_act is None and g_z is None:
        return (None,) * 17
    gb = None
    gz = g_z
    gz_bound = None
    if g_act is not None:
        gz_a, gbv, gzb = modrelu_bwd(z, bias, g_act.contiguous())
        if gz is None:
            gz, gz_bound = gz_a, gzb
            set_bound(gz, gzb)        # the residual branch's backward receives this very tensor
        else:
            gz = gz + gz_a
        gb = gbv.reshape(bias.shape)
    empty = torch.empty(0, dtype=torch.complex64, device=x.device)
    cm = torch.zeros(1, dtype=torch.float32, device=x.device)
    gx, gw = fc_bwd(x, W, gz, empty, cm, rowptr_tgt, rec_tgt, rot_tgt, rowptr_src, rec_src, rot_src, norms, band_limit, n_rings,
                    flags & ~_lib.FLAG_FUSED, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.x_bound,
                    gz_bound if _uses_bounds(flags) else None, ctx.w_bound)
    return (gx if ctx.needs_input_grad[0] else None, gw if ctx.needs_input_grad[1] else None,
            gz if (has_res and ctx.needs_input_grad[2]) else None, gb if ctx.needs_input_grad[3] else None) + (None,) * 13


fc_fwd_act.register_autograd(_fa_backward, setup_context=_fa_setup)


def field_conv_act(x, W, plan, band_limit, bias, res=None, flags=0):
    """modReLU(FieldConv(x) + res, bias) for the compact plan, the block epilogue fused into the layer; differentiable w.r.t.
    x, W, res and bias.  Returns the activated output."""
    norms = getattr(plan, "norms", None)
    if norms is None:
        if flags & (_lib.FLAG_PACKED | _lib.FLAG_PACKED_G):
            raise RuntimeError("fieldconv_b200: the packed path needs a plan built by build_plan (plan.norms)")
        norms = torch.zeros(2, dtype=torch.float32, device=x.device)
    has_res = res is not None
    if not has_res:
        res = torch.empty(0, dtype=torch.complex64, device=x.device)
    x = x.contiguous()
    act, _, act_bound = fc_fwd_act(x, W, res, bias, plan.rowptr_tgt, plan.rec_tgt, plan.rot_tgt, plan.rowptr_src, plan.rec_src,
                                   plan.rot_src, norms, band_limit, plan.n_rings, flags, has_res, _x_bound_for(x, W, flags),
                                   _w_bound_for(W, flags))
    return set_bound(act, act_bound)


def _w_bound_for(W, flags):
    """Bound of max|W| of the folded filter: attached by prefold() for a whole network at once, else one pass shared by the
    forward and the backward of this layer."""
    if not _uses_bounds(flags):
        return None
    return bound_of(W)


def _x_bound_for(x, W, flags):
    """Bound of the layer input when some kernel of this layer will need it: the packed forward (operand scale of contrib)
    or the weight gradient (operand scale of xhat)."""
    if not _uses_bounds(flags):
        return None
    if (flags & _lib.FLAG_PACKED) or (torch.is_grad_enabled() and W.requires_grad):
        return bound_of(x)
    return peek_bound(x)


def field_conv(x, W, plan, band_limit, flags=0, keep_contrib=None):
    """y = FieldConv(x) for the compact plan; differentiable w.r.t. x and W.  flags may carry _lib.FLAG_PACKED."""
    n, ci = x.shape
    if keep_contrib is None:
        keep_contrib = keep_contrib_default(n * plan.n_rings * ci * (2 * band_limit + 1) * 8, x.device)
    norms = getattr(plan, "norms", None)
    if norms is None:
        if flags & (_lib.FLAG_PACKED | _lib.FLAG_FUSED | _lib.FLAG_PACKED_G):
            raise RuntimeError("fieldconv_b200: the packed / fused paths need a plan built by build_plan (plan.norms)")
        norms = torch.zeros(2, dtype=torch.float32, device=x.device)
    x = x.contiguous()
    y, _, _ = fc_fwd(x, W, plan.rowptr_tgt, plan.rec_tgt, plan.rot_tgt, plan.rowptr_src, plan.rec_src, plan.rot_src, norms,
                     band_limit, plan.n_rings, flags, bool(keep_contrib), _x_bound_for(x, W, flags), _w_bound_for(W, flags))
    return y


# --------------------------------------------------------------------------- dense-stencil ops
@torch.library.custom_op("fieldconv_b200::fc_fwd_dense", mutates_args=())
def fc_fwd_dense(x: Tensor, W: Tensor, sten: Tensor, rowptr_tgt: Tensor, nbr_tgt: Tensor, perm_tgt: Tensor,
                 rowptr_src: Tensor, nbr_src: Tensor, perm_src: Tensor, flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    _check(x, "x")
    _check(W, "W")
    _check(sten, "supp_sten")
    x, W, sten = x.contiguous(), W.contiguous(), sten.contiguous()
    n, ci = x.shape
    co, _, r, m = W.shape
    b = (m - 1) // 2
    y = torch.empty(n, co, dtype=torch.complex64, device=x.device)
    contrib = torch.empty(n, r * ci * m, dtype=torch.complex64, device=x.device)
    cmax = torch.zeros(1, dtype=torch.float32, device=x.device)
    nbytes = _lib.query_bytes("fcb_fwd_workspace_bytes", n, ci, co, b, r, flags)
    ws = _ws(nbytes, x.device)
    with torch.cuda.device(x.device):
        _lib.call("fcb_fwd_dense_f32", _real(x).data_ptr(), _real(W).data_ptr(), _real(sten).data_ptr(),
                  rowptr_tgt.data_ptr(), nbr_tgt.data_ptr(), perm_tgt.data_ptr(), _real(y).data_ptr(),
                  _real(contrib).data_ptr(), cmax.data_ptr(), n, ci, co, b, r, flags, ws.data_ptr(), nbytes,
                  _lib.stream_ptr())
    return y, contrib, cmax


@fc_fwd_dense.register_fake
def _(x, W, sten, rowptr_tgt, nbr_tgt, perm_tgt, rowptr_src, nbr_src, perm_src, flags):
    n, ci = x.shape
    co, _, r, m = W.shape
    return x.new_empty(n, co), x.new_empty(n, r * ci * m), x.new_empty(1, dtype=torch.float32)


@torch.library.custom_op("fieldconv_b200::fc_bwd_dense", mutates_args=())
def fc_bwd_dense(x: Tensor, W: Tensor, gy: Tensor, contrib: Tensor, cmax: Tensor, sten: Tensor, rowptr_src: Tensor, nbr_src: Tensor,
                 perm_src: Tensor, flags: int, need_gx: bool, need_gw: bool) -> Tuple[Tensor, Tensor]:
    x, W, gy, sten = x.contiguous(), W.contiguous(), gy.contiguous(), sten.contiguous()
    n, ci = x.shape
    co, _, r, m = W.shape
    b = (m - 1) // 2
    gx = torch.empty_like(x) if need_gx else torch.empty(0, dtype=x.dtype, device=x.device)
    gw = torch.empty_like(W) if need_gw else torch.empty(0, dtype=W.dtype, device=x.device)
    nbytes = _lib.query_bytes("fcb_bwd_workspace_bytes", n, ci, co, b, r, flags | 0x100)
    ws = _ws(nbytes, x.device)
    with torch.cuda.device(x.device):
        _lib.call("fcb_bwd_dense_f32", _real(x).data_ptr(), _real(W).data_ptr(), _real(gy).data_ptr(),
                  _real(contrib).data_ptr(), cmax.data_ptr(), _real(sten).data_ptr(), rowptr_src.data_ptr(), nbr_src.data_ptr(),
                  perm_src.data_ptr(), _real(gx).data_ptr() if need_gx else 0, _real(gw).data_ptr() if need_gw else 0,
                  n, ci, co, b, r, flags, ws.data_ptr(), nbytes, _lib.stream_ptr())
    return gx, gw


@fc_bwd_dense.register_fake
def _(x, W, gy, contrib, cmax, sten, rowptr_src, nbr_src, perm_src, flags, need_gx, need_gw):
    return (torch.empty_like(x) if need_gx else x.new_empty(0)), (torch.empty_like(W) if need_gw else W.new_empty(0))


def _fcd_setup(ctx, inputs, output):
    x, W, sten, rowptr_tgt, nbr_tgt, perm_tgt, rowptr_src, nbr_src, perm_src, flags = inputs
    ctx.save_for_backward(x, W, output[1], output[2], sten, rowptr_src, nbr_src, perm_src)
    ctx.flags = flags
    ctx.set_materialize_grads(False)


def _fcd_backward(ctx, gy, _gc, _gm):
    if gy is None:
        return (None,) * 10
    x, W, contrib, cmax, sten, rowptr_src, nbr_src, perm_src = ctx.saved_tensors
    gx, gw = fc_bwd_dense(x, W, gy, contrib, cmax, sten, rowptr_src, nbr_src, perm_src, ctx.flags,
                          ctx.needs_input_grad[0], ctx.needs_input_grad[1])
    return (gx if ctx.needs_input_grad[0] else None, gw if ctx.needs_input_grad[1] else None) + (None,) * 8


fc_fwd_dense.register_autograd(_fcd_backward, setup_context=_fcd_setup)


def field_conv_dense(x, W, supp_sten, plan, flags=0):
    y, _, _ = fc_fwd_dense(x, W, supp_sten, plan.rowptr_tgt, plan.nbr_tgt, plan.perm_tgt, plan.rowptr_src, plan.nbr_src,
                        plan.perm_src, flags)
    return y


# --------------------------------------------------------------------------- modReLU
@torch.library.custom_op("fieldconv_b200::modrelu", mutates_args=())
def modrelu(x: Tensor, bias: Tensor) -> Tensor:
    _check(x, "x")
    x = x.contiguous()
    b = bias.reshape(-1).contiguous().float()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.call("fcb_modrelu_fwd_f32", _real(x).data_ptr(), b.data_ptr(), _real(y).data_ptr(), x.shape[0], x.shape[1],
                  _lib.stream_ptr())
    return y


@modrelu.register_fake
def _(x, bias):
    return torch.empty_like(x)


@torch.library.custom_op("fieldconv_b200::modrelu_bwd", mutates_args=())
def modrelu_bwd(x: Tensor, bias: Tensor, gy: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (gx, gb, bound of gx: its largest |component|, reported by the kernel)"""
    x, gy = x.contiguous(), gy.contiguous()
    b = bias.reshape(-1).contiguous().float()
    n, c = x.shape
    gx = torch.empty_like(x)
    gb = torch.empty(c, dtype=torch.float32, device=x.device)
    gx_bound = torch.empty(1, dtype=torch.float32, device=x.device)
    nbytes = _lib.query_bytes("fcb_modrelu_bwd_workspace_bytes", n, c)
    ws = _ws(nbytes, x.device)
    with torch.cuda.device(x.device):
        _lib.call("fcb_modrelu_bwd_f32", _real(x).data_ptr(), b.data_ptr(), _real(gy).data_ptr(), _real(gx).data_ptr(),
                  gb.data_ptr(), gx_bound.data_ptr(), n, c, ws.data_ptr(), nbytes, _lib.stream_ptr())
    return gx, gb, gx_bound


@modrelu_bwd.register_fake
def _(x, bias, gy):
    return torch.empty_like(x), x.new_empty(x.shape[1], dtype=torch.float32), x.new_empty(1, dtype=torch.float32)


def _mr_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _mr_backward(ctx, gy):
    x, bias = ctx.saved_tensors
    gx, gb, gxb = modrelu_bwd(x, bias, gy)
    return set_bound(gx, gxb), gb.reshape(bias.shape)


modrelu.register_autograd(_mr_backward, setup_context=_mr_setup)


# --------------------------------------------------------------------------- real GEMM (TangentLin)
def _gemm_call(a, b, trans_a, flags, a_bound=None, b_bound=None):
    a, b = a.contiguous(), b.contiguous()
    if trans_a:
        k, m = a.shape
    else:
        m, k = a.shape
    n = b.shape[1]
    c = torch.empty(m, n, dtype=torch.float32, device=a.device)
    split = 1
    if trans_a:
        tiles = (m + 127) // 128
        split = max(1, min(k // 512, (4 * 148) // tiles))      # at most four full waves of CTAs on 148 SMs
    nbytes = _lib.query_bytes("fcb_gemm_workspace_bytes", m, n, k, 1 if trans_a else 0, 1, split, flags)
    ws = _ws(nbytes, a.device)
    with torch.cuda.device(a.device):
        _lib.call("fcb_gemm_f32", a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, a.shape[1], n, n,
                  1 if trans_a else 0, 1, 0, 0, 0, split, _lib.ptr(a_bound) or None, _lib.ptr(b_bound) or None, ws.data_ptr(), nbytes,
                  flags, _lib.stream_ptr())
    return c


@torch.library.custom_op("fieldconv_b200::gemm", mutates_args=())
def gemm(a: Tensor, b: Tensor, trans_a: bool, flags: int = 0, a_bound: Optional[Tensor] = None,
         b_bound: Optional[Tensor] = None) -> Tensor:
    """C = A @ B (trans_a False, A is MxK) or A^T @ B (trans_a True, A is KxM); fp32, row-major.  a_bound / b_bound: device
    scalars >= max|A|, max|B| when their producer knows them (struct fcb_bounds of the header)."""
    _check(a, "a", torch.float32)
    _check(b, "b", torch.float32)
    return _gemm_call(a, b, trans_a, flags, a_bound, b_bound)


@gemm.register_fake
def _(a, b, trans_a, flags=0, a_bound=None, b_bound=None):
    return a.new_empty(a.shape[1] if trans_a else a.shape[0], b.shape[1])


def _gemm_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.trans_a = inputs[2]
    ctx.flags = inputs[3]


def _gemm_backward(ctx, gc):
    a, b = ctx.saved_tensors
    if ctx.trans_a:
        raise RuntimeError("fieldconv_b200::gemm: backward of the transposed form is not needed")
    ga = gemm(gc, b.t().contiguous(), False, ctx.flags) if ctx.needs_input_grad[0] else None
    gb = gemm(a, gc, True, ctx.flags) if ctx.needs_input_grad[1] else None
    return ga, gb, None, None, None, None


gemm.register_autograd(_gemm_backward, setup_context=_gemm_setup)


# --------------------------------------------------------------------------- TangentLin on complex tensors
@torch.library.custom_op("fieldconv_b200::tangent_lin", mutates_args=())
def tangent_lin(x: Tensor, emb: Tensor, flags: int, x_bound: Optional[Tensor] = None,
                emb_bound: Optional[Tensor] = None) -> Tensor:
    """y = x @ (Re + i Im)^T (nn/tangent_lin.py:27-29) as ONE real GEMM on the interleaved storage: [x_re, x_im] @ emb with emb
    the (2Ci, 2Co) real embedding of the weight.  Ci and Co even (16-byte rows).  Taking and returning the complex tensors
    themselves lets the backward see the very gradient tensor its producer attached a bound to (set_bound)."""
    _check(x, "x")
    _check(emb, "emb", torch.float32)
    x = x.contiguous()
    n, ci = x.shape
    yr = _gemm_call(_real(x).reshape(n, 2 * ci), emb, False, flags, x_bound, emb_bound)
    return torch.view_as_complex(yr.reshape(n, emb.shape[1] // 2, 2))


@tangent_lin.register_fake
def _(x, emb, flags, x_bound=None, emb_bound=None):
    return x.new_empty(x.shape[0], emb.shape[1] // 2)


def _tl_setup(ctx, inputs, output):
    x, emb, flags, x_bound, emb_bound = inputs
    ctx.save_for_backward(x, emb)
    ctx.flags, ctx.x_bound, ctx.emb_bound = flags, x_bound, emb_bound


def _tl_backward(ctx, gy):
    x, emb = ctx.saved_tensors
    gy = gy.contiguous()
    n, ci = x.shape
    co = emb.shape[1] // 2
    gyb = bound_of(gy) if _uses_bounds(ctx.flags) else None      # attached by the producer of gy, else ONE pass for both products
    gyr = _real(gy).reshape(n, 2 * co)
    gx = gemb = None
    if ctx.needs_input_grad[0]:
        gxr = _gemm_call(gyr, emb.t().contiguous(), False, ctx.flags, gyb, ctx.emb_bound)     # the transpose has the same entries
        gx = torch.view_as_complex(gxr.reshape(n, ci, 2))
    if ctx.needs_input_grad[1]:
        gemb = _gemm_call(_real(x.contiguous()).reshape(n, 2 * ci), gyr, True, ctx.flags, ctx.x_bound, gyb)
    return gx, gemb, None, None, None


tangent_lin.register_autograd(_tl_backward, setup_context=_tl_setup)
