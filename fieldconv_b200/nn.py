"""Drop-in modules mirroring the reference's operator interface for the hot path:

    FieldConv(in_channels, out_channels, band_limit=1, n_rings=6, ftype=1)     nn/field_conv.py:62
        .forward(x, supp_edges, supp_sten)                                       nn/field_conv.py:104
    TangentLin, TangentNonLin, FCResNetBlock                                     nn/tangent_lin.py, nn/tangent_nonlin.py,
                                                                                 nn/fc_resnet_block.py:43-88

Parameter / buffer names, shapes and initialisation are the reference's (state_dicts interchange).
Extra, keyword-only: ``precision=`` selects the contraction arithmetic, and ``forward(x, plan)`` accepts
a compact ``fieldconv_b200.Plan`` instead of (supp_edges, supp_sten) — the fast path.
"""
import os
import weakref

import torch
import torch.nn as nn
from torch.nn import Parameter

from . import _lib, ops
from .partition import MeshPartition, partitioned_field_conv
from .plan import DensePlan, Plan, build_dense_plan
from .transforms import attached_plan

_PRECISIONS = {"fp32": _lib.GEMM_SIMT_FP32, "3xtf32": _lib.GEMM_TC_3XTF32, "tf32": _lib.GEMM_TC_TF32,
               "2xf16": _lib.GEMM_TC_2XF16, "2xf16p": _lib.GEMM_TC_2XF16 | _lib.FLAG_PACKED, "auto": -1}

# precision="auto" takes the packed-operand variant of the 2xFP16 path ("2xf16p": same arithmetic, the aggregation kernels
# write the fp16 operand planes themselves) where the library supports the layer shape and the B200 measurements favour it
# (profiles/r01f: band_limit <= 1 gains 7-8 % per layer — 1M vertices C=32 11.2 -> 10.3 ms, C=128 57.0 -> 52.8 ms; at
# band_limit 2 the extra conversion work in the aggregation cancels the contraction's gain).
# FIELDCONV_B200_PACKED=1: wherever supported; =0: never; unset/"auto": band_limit <= 1.
PACKED_POLICY = os.environ.get("FIELDCONV_B200_PACKED", "auto")


# FIELDCONV_B200_FUSED=1: the fused forward (gather -> shared-memory tile -> tcgen05, csrc/fused_fwd.cu) wherever the shape
# is supported (band_limit <= 1, Ci % 32 == 0, Co <= 128) and nothing of size N x K has to be kept; =0: never.
FUSED_POLICY = os.environ.get("FIELDCONV_B200_FUSED", "0")


def fused_flags(flags, plan, ci, co, band_limit, n_rings):
    if FUSED_POLICY != "1" or (flags & _lib.GEMM_MASK) != _lib.GEMM_TC_2XF16 or ops.SAVE_CONTRIB:
        return flags
    if getattr(plan, "norms", None) is None or not _lib.fused_supported(ci, co, band_limit, n_rings):
        return flags
    return flags | _lib.FLAG_FUSED


LIN_POLICY = os.environ.get("FIELDCONV_B200_LIN", "2xf16")
# FIELDCONV_B200_PACKED_G=0: band_limit 2 keeps the fp32 G layout in the backward too (A/B switch)
PACKED_G_POLICY = os.environ.get("FIELDCONV_B200_PACKED_G", "1")
# FIELDCONV_B200_BLOCK_EPILOGUE=0: FCResNetBlock runs TangentNonLin / the residual add as separate kernels (A/B switch)
BLOCK_EPILOGUE = os.environ.get("FIELDCONV_B200_BLOCK_EPILOGUE", "1")


def _packed_by_default(band_limit):
    if PACKED_POLICY == "0":
        return False
    return True if PACKED_POLICY == "1" else band_limit <= 1


def packed_flags(flags, plan, n, ci, co, band_limit, n_rings, explicit, auto=False):
    """Resolve FLAG_PACKED for one call: keep it only for a compact plan with norms and a supported shape.  An
    explicit precision="2xf16p" raises when the shape is not supported instead of silently changing kernels;
    precision="auto" follows the measured policy above."""
    want = bool(flags & _lib.FLAG_PACKED) or (auto and _packed_by_default(band_limit) and (flags & _lib.GEMM_MASK) == _lib.GEMM_TC_2XF16)
    flags &= ~_lib.FLAG_PACKED
    ok = getattr(plan, "norms", None) is not None and _lib.pk_supported(n, ci, co, band_limit, n_rings)
    if not want:
        # band_limit 2 under "auto": fp32 contrib in the forward (the packing store costs the aggregation more than the
        # contraction gains), packed G in the backward, where TWO contractions read it (measured r02c at the cfg-2 layer:
        # transposed aggregation +0.056 ms, grouped grad-x -0.053 ms, weight gradient -0.039 ms)
        if auto and ok and band_limit == 2 and PACKED_POLICY != "0" and PACKED_G_POLICY == "1" and \
                (flags & _lib.GEMM_MASK) == _lib.GEMM_TC_2XF16 and not ops.SAVE_CONTRIB:
            flags |= _lib.FLAG_PACKED_G
        return flags
    if ok:
        return flags | _lib.FLAG_PACKED
    if explicit:
        raise RuntimeError("fieldconv_b200: precision='2xf16p' does not support this layer shape / plan "
                           "(N=%d, Ci=%d, Co=%d, B=%d, R=%d)" % (n, ci, co, band_limit, n_rings))
    return flags


import functools


@functools.lru_cache(maxsize=None)
def _resolve_precision(precision, ci, co, n_rings, band_limit):
    """"auto": error-compensated tensor cores — operands as scaled fp16 (hi, lo) pairs ("2xf16", fastest), else
    3xTF32 — when the library's accumulation plan (column chunks x TMEM accumulators, at most 400 accumulating MMAs
    each — DESIGN.md §4) keeps the forward and the grad-x contractions inside the fp32 path's 1e-5 parity budget,
    else the FP32-FMA kernels."""
    if precision != "auto":
        return _PRECISIONS[precision]
    m = 2 * band_limit + 1
    for mode in (_lib.GEMM_TC_2XF16, _lib.GEMM_TC_3XTF32):
        fwd = _lib.tc_feasible(2 * co, 2 * n_rings * ci * m, flags=mode)
        bwd = _lib.tc_feasible(2 * ci, 2 * n_rings * co, flags=mode)
        if fwd and bwd:
            return mode
    return _lib.GEMM_SIMT_FP32


# Dense-stencil plans (CSR orders of the caller's supp_edges), ONE per (supp_edges tensor, N), shared by every layer of the
# network.  Keyed on the tensor OBJECT (id + in-place version) and validated through a weak reference, so an address the
# caching allocator hands to the next batch can never match a stale entry; the entry dies with the tensor.
_DENSE_PLANS = {}


def shared_dense_plan(supp_edges, n):
    key = (id(supp_edges), supp_edges._version, int(supp_edges.shape[0]), int(n))
    hit = _DENSE_PLANS.get(key)
    if hit is not None and hit[0]() is supp_edges:
        return hit[1]
    plan = build_dense_plan(supp_edges, n)
    for k in [k for k in _DENSE_PLANS if k[0] == key[0]]:        # older versions / sizes of the same object
        del _DENSE_PLANS[k]
    _DENSE_PLANS[key] = (weakref.ref(supp_edges, lambda _r, k=key: _DENSE_PLANS.pop(k, None)), plan)
    return plan


def _param_versions(m):
    return tuple(p._version for p in (m.zonal, m.spherical, m.phase))


def fold_weights(zonal, spherical, phase, ftype, band_limit):
    """(zonal, spherical, phase) -> W (Co,Ci,R,2B+1) complex with y = sum contrib * W, i.e. the
    coefficient tensors of weightContribReal / Offset / Complex divided by 2B+1
    (nn/field_conv.py:12-14, :18-25, :31-33).  Differentiable torch code on tiny tensors."""
    B = band_limit
    sph = torch.view_as_complex(spherical.contiguous())
    if ftype == 2:
        zc = torch.view_as_complex(zonal.contiguous())
        w = torch.cat((sph[..., :B], zc.unsqueeze(-1), sph[..., B:]), dim=-1)
    else:
        zc = torch.complex(zonal, torch.zeros_like(zonal)).unsqueeze(-1)
        w = torch.cat((sph.conj().flip(-1), zc, sph), dim=-1)
        if ftype == 1:
            ph = torch.cat((phase[..., 1:].flip(-1), phase), dim=-1)          # |m| = B..1, 0, 1..B
            w = w * torch.polar(torch.ones_like(ph), ph).unsqueeze(-2)
    return (w / (2 * B + 1)).resolve_conj().contiguous()


def prefold(module):
    """Fold the filters of every FieldConv inside `module` (and build the real embedding of every TangentLin) in a few
    batched torch ops and hand each layer its operand for its next forward.  fold_weights is ~30 tiny kernels per layer per step (forward + autograd); a 10-layer network spends
    0.7 ms of a 20 ms step on them.  Layers with identical (ftype, shapes) are stacked, folded once, and unbound — the
    same arithmetic on the same values, so outputs and gradients are unchanged (tests/test_host_logic.py).  Optional:
    a layer that was not prefolded folds its own filter as before.  Call once per forward pass, before the layers run."""
    groups = {}
    for m in module.modules():
        if isinstance(m, FieldConv):
            key = (m.ftype, m.B, tuple(m.zonal.shape), tuple(m.spherical.shape), m.zonal.device)
            groups.setdefault(key, []).append(m)
    for (ftype, band_limit, _, _, _), layers in groups.items():
        if len(layers) == 1:
            continue
        w = fold_weights(torch.stack([m.zonal for m in layers]), torch.stack([m.spherical for m in layers]),
                         torch.stack([m.phase for m in layers]), ftype, band_limit)
        # max|W| of every layer from ONE reduction: the operand scale of the packed filter (struct fcb_bounds, field w)
        wb = w.detach().abs().flatten(1).amax(dim=1) if (ops.BOUNDS and w.is_cuda) else None
        for i, (m, wi) in enumerate(zip(layers, w.unbind(0))):
            if wb is not None:
                ops.set_bound(wi, wb[i:i + 1])
            m._prefolded = (wi, _param_versions(m))       # valid only while the parameters are unchanged
    lins = {}
    for m in module.modules():
        if isinstance(m, TangentLin):
            lins.setdefault((m.in_channels, m.out_channels, m.Re.device), []).append(m)
    for layers in lins.values():
        if len(layers) > 1:
            emb = _lin_embedding(torch.stack([m.Re for m in layers]), torch.stack([m.Im for m in layers]))
            eb = emb.detach().abs().flatten(1).amax(dim=1) if (ops.BOUNDS and emb.is_cuda) else None
            for i, (m, e) in enumerate(zip(layers, emb.unbind(0))):
                if eb is not None:
                    ops.set_bound(e, eb[i:i + 1])
                m._preemb = (e, (m.Re._version, m.Im._version))


def _lin_embedding(Re, Im):
    """(…, Co, Ci) real and imaginary parts -> (…, 2Ci, 2Co) real matrix E with  [x_re, x_im] @ E = [y_re, y_im]  for
    y = x @ (Re + i Im)^T  (nn/tangent_lin.py:27-29 on the interleaved storage)."""
    re, im = Re.transpose(-1, -2), Im.transpose(-1, -2)              # (…, Ci, Co)
    emb = torch.stack((torch.stack((re, im), -1), torch.stack((-im, re), -1)), -3)   # (…, Ci, 2, Co, 2)
    return emb.reshape(*re.shape[:-2], 2 * re.shape[-2], 2 * re.shape[-1])


class FieldConv(nn.Module):
    def __init__(self, in_channels, out_channels, band_limit=1, n_rings=6, ftype=1, *, precision="auto"):
        super().__init__()
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        self.in_channels, self.out_channels = in_channels, out_channels
        self.R, self.B, self.ftype = n_rings, band_limit, ftype
        self.precision = precision
        if ftype == 0 or ftype == 1:
            self.zonal = Parameter(torch.empty(out_channels, in_channels, n_rings))
            self.spherical = Parameter(torch.empty(out_channels, in_channels, n_rings, band_limit, 2))
            if ftype == 1:
                self.phase = Parameter(torch.empty(out_channels, in_channels, band_limit + 1))
                nn.init.xavier_uniform_(self.phase)
            else:
                self.register_buffer("phase", torch.zeros(out_channels, in_channels, band_limit + 1))
        else:
            self.zonal = Parameter(torch.empty(out_channels, in_channels, n_rings, 2))
            self.spherical = Parameter(torch.empty(out_channels, in_channels, n_rings, 2 * band_limit, 2))
            self.register_buffer("phase", torch.zeros(out_channels, in_channels, band_limit + 1))
        nn.init.xavier_uniform_(self.zonal)
        nn.init.xavier_uniform_(self.spherical)
        self._prefolded = None

    def weight(self):
        pre, self._prefolded = self._prefolded, None
        # handed over by prefold() for exactly one forward, and only while the parameters it was folded from are
        # unchanged (a layer skipped in that forward must not run on pre-optimizer-step weights later)
        if pre is not None and pre[1] == _param_versions(self):
            return pre[0]
        return fold_weights(self.zonal, self.spherical, self.phase, self.ftype, self.B)

    def _dense_plan(self, supp_edges, n):
        return shared_dense_plan(supp_edges, n)

    def forward(self, x, supp_edges=None, supp_sten=None, *, plan=None):
        if isinstance(supp_edges, (Plan, DensePlan, MeshPartition)):
            plan, supp_edges = supp_edges, None
        if not x.is_cuda:
            raise RuntimeError("fieldconv_b200.FieldConv runs on CUDA (sm_100a) only; there is no CPU fallback")
        if isinstance(plan, MeshPartition):      # one large mesh split across ranks: x holds this rank's owned rows
            return partitioned_field_conv(self, x, plan)
        ci, co = self.in_channels, self.out_channels
        flags = _resolve_precision(self.precision, ci + ci % 2, co + co % 2, self.R, self.B)
        w = self.weight()
        if x.shape[1] != ci:
            raise ValueError("expected %d input channels, got %d" % (ci, x.shape[1]))
        if ci % 2:  # 16-byte feature rows: pad a zero channel (contributes nothing)
            x = torch.cat((x, torch.zeros_like(x[:, :1])), dim=1)
            w = torch.cat((w, torch.zeros_like(w[:, :1])), dim=1)
        if co % 2:
            w = torch.cat((w, torch.zeros_like(w[:1])), dim=0)
        if supp_sten is not None:
            if tuple(supp_sten.shape[1:]) != (self.R, 2 * self.B + 1):
                raise ValueError("supp_sten must be (E, %d, %d), got %s" % (self.R, 2 * self.B + 1, tuple(supp_sten.shape)))
            if supp_edges is not None and supp_sten.shape[0] != supp_edges.shape[0]:
                raise ValueError("supp_sten has %d edges, supp_edges %d" % (supp_sten.shape[0], supp_edges.shape[0]))
        if plan is None and supp_sten is not None:
            # (supp_edges, supp_sten) straight from fieldconv_b200.FCPrecomp: use the compact plan it was expanded from
            plan = attached_plan(supp_edges, supp_sten, self.R, x.shape[0])
        if plan is not None and x.shape[0] != plan.num_nodes:
            # the reference raises an index error on a mismatched (features, mesh) pair; here it would be an out-of-bounds
            # gather on the device
            raise ValueError("x has %d rows, the plan was built for %d vertices" % (x.shape[0], plan.num_nodes))
        if plan is not None and not plan.dense:
            if plan.n_rings != self.R:
                raise ValueError("plan was built for n_rings=%d, layer has %d" % (plan.n_rings, self.R))
            flags = packed_flags(flags, plan, x.shape[0], x.shape[1], w.shape[0], self.B, self.R, self.precision == "2xf16p",
                                 auto=self.precision == "auto")
            if self.precision in ("auto", "2xf16", "2xf16p"):
                flags = fused_flags(flags, plan, x.shape[1], w.shape[0], self.B, self.R)
            y = ops.field_conv(x, w, plan, self.B, flags)
        else:
            flags &= ~_lib.FLAG_PACKED
            if supp_sten is None:
                raise ValueError("forward needs (supp_edges, supp_sten) or a compact plan")
            dplan = plan if plan is not None else self._dense_plan(supp_edges, x.shape[0])
            if dplan.e_cap != supp_sten.shape[0]:
                raise ValueError("the dense plan was built for %d edges, supp_sten has %d" % (dplan.e_cap, supp_sten.shape[0]))
            y = ops.field_conv_dense(x, w, supp_sten, dplan, flags)
        return y[:, :co] if co % 2 else y


def _fused_block_ok(plan, supp_sten):
    """The block epilogue (modReLU, residual) rides in the layer's contraction kernel on the compact-plan path."""
    return BLOCK_EPILOGUE == "1" and isinstance(plan, Plan) and not plan.dense and not ops.SAVE_CONTRIB


def _conv_act(layer, x, plan, bias, res):
    """modReLU(layer(x) + res, bias) through the fused-epilogue op; shapes / flags resolved as in FieldConv.forward."""
    ci, co = layer.in_channels, layer.out_channels
    if ci % 2 or co % 2 or x.shape[1] != ci:
        return None                                   # odd channel counts take the padded, unfused route
    if x.shape[0] != plan.num_nodes:
        raise ValueError("x has %d rows, the plan was built for %d vertices" % (x.shape[0], plan.num_nodes))
    if plan.n_rings != layer.R:
        raise ValueError("plan was built for n_rings=%d, layer has %d" % (plan.n_rings, layer.R))
    flags = _resolve_precision(layer.precision, ci, co, layer.R, layer.B)
    flags = packed_flags(flags, plan, x.shape[0], ci, co, layer.B, layer.R, layer.precision == "2xf16p", auto=layer.precision == "auto")
    return ops.field_conv_act(x, layer.weight(), plan, layer.B, bias, res, flags)


class TangentLin(nn.Module):
    """nn/tangent_lin.py:12-29 — y = x @ (Re + i Im)^T, carried as one real GEMM on the interleaved storage."""

    def __init__(self, in_channels, out_channels, *, precision="auto"):
        super().__init__()
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        self.in_channels, self.out_channels = in_channels, out_channels
        # "auto": FIELDCONV_B200_LIN selects "2xf16" (scaled fp16-pair tensor cores, fp32-grade 5e-6; the library falls back to
        # FP32 FMA where no plan fits) or "fp32" (FP32-FMA kernel: one launch instead of absmax + pack + GEMM)
        self.gemm_flags = _PRECISIONS[LIN_POLICY] if precision == "auto" else (_PRECISIONS[precision] & _lib.GEMM_MASK)
        self._preemb = None
        self.Re = Parameter(torch.empty(out_channels, in_channels))
        self.Im = Parameter(torch.empty(out_channels, in_channels))
        nn.init.xavier_uniform_(self.Re)
        nn.init.xavier_uniform_(self.Im, gain=0.1)

    def forward(self, x):
        ci, co = self.in_channels, self.out_channels
        pre, self._preemb = self._preemb, None                  # handed over by prefold() for exactly one forward
        emb = pre[0] if (pre is not None and pre[1] == (self.Re._version, self.Im._version)) else None
        if emb is None:
            emb = _lin_embedding(self.Re, self.Im)              # (2Ci, 2Co)
        if ci % 2 == 0 and co % 2 == 0 and x.is_cuda and x.dtype == torch.complex64:
            x = x.contiguous()
            xb = ops.bound_of(x) if ops._uses_bounds(self.gemm_flags) else None     # the A-operand scale; shared with the conv
            emb = emb.contiguous()
            return ops.tangent_lin(x, emb, self.gemm_flags, xb, ops.peek_bound(emb) if xb is not None else None)
        xr = torch.view_as_real(x.contiguous()).reshape(x.shape[0], 2 * ci)
        if (2 * ci) % 4 or (2 * co) % 4:                         # GEMM wants 16-byte rows
            pad_i, pad_o = (2 * ci) % 4, (2 * co) % 4
            xr = torch.nn.functional.pad(xr, (0, pad_i))
            emb = torch.nn.functional.pad(emb, (0, pad_o, 0, pad_i))
        y = ops.gemm(xr, emb.contiguous(), False, self.gemm_flags)[:, :2 * co]
        return torch.view_as_complex(y.reshape(x.shape[0], co, 2).contiguous())


class TangentNonLin(nn.Module):
    """nn/tangent_nonlin.py:12-35 — modReLU."""

    def __init__(self, in_channels):
        super().__init__()
        self.bias = Parameter(torch.zeros(1, in_channels))

    def forward(self, x):
        return ops.modrelu(x, self.bias)


class TangentPerceptron(nn.Module):
    """nn/tangent_perceptron.py:7-25 — nonlin(lin(x)): the complex fully-connected layer + modReLU of the reference nets'
    'meta' residual connections (correspondence.ipynb / feature_matching.ipynb `Net.res*`)."""

    def __init__(self, in_channels, out_channels, *, precision="auto"):
        super().__init__()
        self.lin = TangentLin(in_channels, out_channels, precision=precision)
        self.nonlin = TangentNonLin(out_channels)

    def forward(self, x):
        return self.nonlin(self.lin(x))


class FCResNetBlock(nn.Module):
    """nn/fc_resnet_block.py:43-88 — nonlin2(res(x) + conv2(nonlin1(conv1(x))))."""

    def __init__(self, in_channels, out_channels, band_limit=1, n_rings=6, ftype=1, frontload=False, *,
                 precision="auto"):
        super().__init__()
        mid = in_channels if frontload else out_channels
        self.conv1 = FieldConv(in_channels, mid, band_limit=band_limit, n_rings=n_rings, ftype=ftype, precision=precision)
        self.conv2 = FieldConv(mid, out_channels, band_limit=band_limit, n_rings=n_rings, ftype=ftype, precision=precision)
        self.nonlin1 = TangentNonLin(mid)
        self.nonlin2 = TangentNonLin(out_channels)
        self.res = TangentLin(in_channels, out_channels, precision=precision)

    def forward(self, x, supp_edges=None, supp_sten=None, *, plan=None):
        if isinstance(supp_edges, (Plan, DensePlan, MeshPartition)):
            plan, supp_edges = supp_edges, None
        if supp_sten is not None and plan is None:
            plan = attached_plan(supp_edges, supp_sten, self.conv1.R, x.shape[0])
        if _fused_block_ok(plan, supp_sten) and x.is_cuda and FUSED_POLICY != "1":
            # nonlin1 and (residual add + nonlin2) run in the epilogue of conv1's / conv2's contraction kernel
            h = _conv_act(self.conv1, x, plan, self.nonlin1.bias, None)
            if h is not None:
                out = _conv_act(self.conv2, h, plan, self.nonlin2.bias, self.res(x))
                if out is not None:
                    return out
        h = self.nonlin1(self.conv1(x, supp_edges, supp_sten, plan=plan))
        h = self.conv2(h, supp_edges, supp_sten, plan=plan)
        return self.nonlin2(self.res(x) + h)
