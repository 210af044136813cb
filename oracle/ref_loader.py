"""Import the UNMODIFIED reference (when /root/reference exists) behind shims.

Test infrastructure only.  The reference imports torch_scatter / torch_geometric /
torch_sparse / fcutils at module import time (nn/field_conv.py:6, utils/field.py:5-6,
nn/tangent_nonlin.py:5, transforms/support_graph.py:4-8); none are installed here.
`scatter_add` is an index-sum along dim 0 (order independent up to rounding), which
`Tensor.index_add` restates exactly; everything else is imported but unused on the
path, so empty stand-ins suffice.  This never runs on the GPU box (no
/root/reference there); it exists to generate `tests/golden/*.npz`.
"""
import importlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("FIELDCONV_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "nn", "field_conv.py"))


def _scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and out is None
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add(0, index, src)


def _install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def _missing(*a, **k):
        raise RuntimeError("shimmed third-party function is not on the FieldConv path")

    mod("torch_scatter", scatter_add=_scatter_add, scatter_min=_missing)
    tg = mod("torch_geometric")
    tg.data = mod("torch_geometric.data", Data=object)
    tg.nn = mod("torch_geometric.nn", radius=_missing, fps=_missing)
    tg.nn.inits = mod("torch_geometric.nn.inits", zeros=_missing)
    tg.utils = mod("torch_geometric.utils", degree=_missing, to_undirected=_missing)
    tg.io = mod("torch_geometric.io", read_ply=_missing, read_off=_missing, read_obj=_missing)
    tg.transforms = mod("torch_geometric.transforms")
    mod("torch_sparse", coalesce=_missing)
    mod("fcutils")
    mod("progressbar")


_cache = {}


def load():
    """Returns a namespace with the reference FieldConv, FCResNetBlock, TangentLin,
    TangentNonLin and FCPrecomp classes (unmodified reference code)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    # the reference's top-level packages are called `nn`, `utils`, `transforms`
    for name in ("nn", "utils", "transforms"):
        if name in sys.modules:
            raise RuntimeError("module name clash importing the reference: " + name)
    sys.path.insert(0, REF_ROOT)
    try:
        ref_nn = importlib.import_module("nn")
        ref_fc = importlib.import_module("nn.field_conv")
        ref_pre = importlib.import_module("transforms.fc_precomp")
        ref_field = importlib.import_module("utils.field")
    finally:
        sys.path.remove(REF_ROOT)
    ns = types.SimpleNamespace(
        FieldConv=ref_fc.FieldConv,
        FCResNetBlock=ref_nn.FCResNetBlock,
        TangentLin=ref_nn.TangentLin,
        TangentNonLin=ref_nn.TangentNonLin,
        FCPrecomp=ref_pre.FCPrecomp,
        radialInterpolant=ref_pre.radialInterpolant,
        softAngle=ref_field.softAngle,
        TransField=ref_nn.TransField,
        LiftBlock=ref_nn.LiftBlock,
        ECHO=ref_nn.ECHO,
        ECHOBlock=ref_nn.ECHOBlock,
        TangentPerceptron=ref_nn.TangentPerceptron,
    )
    _cache["ns"] = ns
    return ns
