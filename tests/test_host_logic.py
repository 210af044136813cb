"""CPU: host-side mirror of the reference interface (module surface, weight folding, input checks)."""
import pytest
import torch

import fieldconv_b200 as fcb
from conftest import assert_close_normwise, golden_names, load_golden
from oracle import restate


@pytest.mark.parametrize("name", golden_names("fc_"))
def test_fold_weights_matches_oracle(name):
    g = load_golden(name)
    w = fcb.fold_weights(g["zonal"], g["spherical"], g["phase"], g["ftype"], g["B"])
    w_ref = restate.fold_weights(g["zonal"], g["spherical"], g["phase"], g["ftype"], g["B"])
    assert w.shape == (g["co"], g["ci"], g["R"], 2 * g["B"] + 1)
    assert_close_normwise(w, w_ref, 1e-7, "fold_weights")


@pytest.mark.parametrize("ftype", [0, 1, 2])
def test_state_dict_contract(ftype):
    """Parameter/buffer names, shapes and Parameter-vs-buffer status follow nn/field_conv.py:71-98."""
    m = fcb.FieldConv(6, 4, 2, 5, ftype)
    sd = m.state_dict()
    assert set(sd) == {"zonal", "spherical", "phase"}
    if ftype == 2:
        assert sd["zonal"].shape == (4, 6, 5, 2) and sd["spherical"].shape == (4, 6, 5, 4, 2)
    else:
        assert sd["zonal"].shape == (4, 6, 5) and sd["spherical"].shape == (4, 6, 5, 2, 2)
    assert sd["phase"].shape == (4, 6, 3)
    params = dict(m.named_parameters())
    assert ("phase" in params) == (ftype == 1)
    assert (m.in_channels, m.out_channels, m.R, m.B, m.ftype) == (6, 4, 5, 2, ftype)


def test_reference_state_dict_loads():
    g = load_golden("block_b2r6")
    blk = fcb.FCResNetBlock(g["ci"], g["co"], g["B"], g["R"], 1)
    sd = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    missing, unexpected = blk.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_reference_network_state_dict_loads():
    """The whole notebook-style network (LiftBlock, FCResNetBlock incl. frontload, TangentPerceptron, ECHOBlock) built from
    this package's modules takes the reference network's state_dict as is: same keys, same shapes (net_b2r6 golden)."""
    from oracle.make_golden import build_net
    g = load_golden("net_b2r6")
    net = build_net(fcb, g["B"], g["R"], g["ftype"], g["n_classes"], g["n_des"], g["n_bins"])
    sd = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    assert set(sd) == set(net.state_dict())
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert {k for k, _ in net.named_parameters()} == {k[2:] for k in g if k.startswith("g.")}
    with pytest.raises(RuntimeError):          # no CPU path anywhere in the network
        net(g["pos"], g["supp_edges"], g["supp_sten"], g["ln"], g["wxp"])


def test_no_cpu_fallback():
    m = fcb.FieldConv(4, 4)
    x = torch.zeros(5, 4, dtype=torch.complex64)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(x, torch.zeros(3, 2, dtype=torch.long), torch.zeros(3, 6, 3, dtype=torch.complex64))
    with pytest.raises(RuntimeError, match="CUDA"):
        fcb.build_plan(torch.zeros(3, 2, dtype=torch.long), torch.zeros(3), torch.zeros(3),
                       torch.zeros(3, dtype=torch.complex64), torch.ones(5, 1), 6, 1.0)


def test_product_does_not_import_oracle():
    import os
    import re
    from conftest import ROOT
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fieldconv_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_packed_path_selection_policy(monkeypatch):
    """precision="auto" takes the packed-operand path for band_limit <= 1 where the library supports the shape;
    an explicit "2xf16p" raises on an unsupported shape instead of silently switching kernels."""
    from fieldconv_b200 import _lib
    from fieldconv_b200 import nn as fnn

    class FakePlan:
        norms = object()

    class NoNorms:
        pass

    f16 = _lib.GEMM_TC_2XF16
    monkeypatch.setattr(fnn, "PACKED_POLICY", "auto")
    assert fnn.packed_flags(f16, FakePlan(), 5041, 32, 32, 1, 6, False, auto=True) == f16 | _lib.FLAG_PACKED
    # band_limit 2: fp32 contrib in the forward, packed G in the backward (two contractions read it)
    assert fnn.packed_flags(f16, FakePlan(), 80656, 48, 48, 2, 6, False, auto=True) == f16 | _lib.FLAG_PACKED_G
    monkeypatch.setattr(fnn, "PACKED_G_POLICY", "0")
    assert fnn.packed_flags(f16, FakePlan(), 80656, 48, 48, 2, 6, False, auto=True) == f16
    monkeypatch.setattr(fnn, "PACKED_G_POLICY", "1")
    assert fnn.packed_flags(f16, FakePlan(), 80656, 48, 48, 3, 6, False, auto=True) == f16          # band_limit 3: not measured, unpacked
    assert fnn.packed_flags(f16, NoNorms(), 5041, 32, 32, 1, 6, False, auto=True) == f16            # plan without norms
    assert fnn.packed_flags(f16, FakePlan(), 144, 6, 6, 1, 3, False, auto=True) == f16              # 108 columns: unsupported
    assert fnn.packed_flags(f16, FakePlan(), 5041, 32, 32, 1, 6, False, auto=False) == f16          # explicit "2xf16" stays unpacked
    assert fnn.packed_flags(_lib.GEMM_TC_3XTF32, FakePlan(), 5041, 32, 32, 1, 6, False, auto=True) == _lib.GEMM_TC_3XTF32
    assert fnn.packed_flags(f16 | _lib.FLAG_PACKED, FakePlan(), 80656, 48, 48, 2, 6, True) == f16 | _lib.FLAG_PACKED
    with pytest.raises(RuntimeError, match="2xf16p"):
        fnn.packed_flags(f16 | _lib.FLAG_PACKED, FakePlan(), 144, 6, 6, 1, 3, True)
    monkeypatch.setattr(fnn, "PACKED_POLICY", "0")
    assert fnn.packed_flags(f16, FakePlan(), 5041, 32, 32, 1, 6, False, auto=True) == f16
    monkeypatch.setattr(fnn, "PACKED_POLICY", "1")
    assert fnn.packed_flags(f16, FakePlan(), 80656, 48, 48, 2, 6, False, auto=True) == f16 | _lib.FLAG_PACKED


def test_packed_format_mirror_roundtrip():
    """fieldconv_b200.packed: encode a small matrix on the host exactly as store_ring_packed addresses it (byte_offset),
    decode with unpack(): values come back to 2^-21 relative, the tail rows of the last tile are zero."""
    import numpy as np
    from fieldconv_b200 import packed
    rows, cols = 130, 128
    rng = np.random.default_rng(0)
    a = rng.standard_normal((rows, cols)).astype(np.float32) * np.float32(3.0)
    bound = float(np.abs(a).max()) * 1.5
    s = np.float32(packed.scale_of(bound))
    assert 2.0 ** 14 <= bound * float(s) < 2.0 ** 15
    raw = np.zeros(packed.pk_bytes(rows, cols), dtype=np.uint8)
    hi = (a * s).astype(np.float16)
    lo = (a * s - hi.astype(np.float32)).astype(np.float16)
    for r in range(rows):
        for c in range(cols):
            for plane, src in ((0, hi), (1, lo)):
                o = packed.byte_offset(r, c, cols, plane)
                raw[o:o + 2] = np.frombuffer(src[r, c].tobytes(), dtype=np.uint8)
    assert packed.pk_bytes(rows, cols) == 2 * 2 * packed.BLOCK_BYTES and packed.padded_rows(rows) == 256
    out = packed.unpack(torch.from_numpy(raw), rows, cols, bound)
    assert out.shape == (256, cols)
    assert float((out[:rows] - torch.from_numpy(a)).abs().max()) <= 2.0 ** -21 * float(np.abs(a).max())
    assert float(out[rows:].abs().max()) == 0.0
    # units of 8 fp16 stay contiguous and 16-byte aligned; the swizzle permutes units inside one 128-byte row only
    assert packed.byte_offset(5, 8, cols) % 16 == 0 and packed.byte_offset(5, 9, cols) == packed.byte_offset(5, 8, cols) + 2
    assert packed.byte_offset(5, 0, cols) // 128 == packed.byte_offset(5, 63, cols) // 128
    assert packed.byte_offset(0, 64, cols) == packed.BLOCK_BYTES and packed.byte_offset(128, 0, cols) == 2 * packed.BLOCK_BYTES


@pytest.mark.parametrize("ftype", [0, 1, 2])
def test_prefold_matches_per_layer_fold(ftype):
    """fcb.prefold: batched folding of all layers' filters gives the same W and the same parameter gradients as each
    layer folding its own (nn/field_conv.py:10-33 arithmetic, stacked along a leading axis)."""
    torch.manual_seed(0)
    net = torch.nn.ModuleList([fcb.FCResNetBlock(6, 6, 2, 4, ftype) for _ in range(3)] + [fcb.FieldConv(6, 4, 2, 4, ftype)])
    convs = [m for m in net.modules() if isinstance(m, fcb.FieldConv)]
    assert len(convs) == 7
    probes = [torch.randn(m.out_channels, m.in_channels, m.R, 2 * m.B + 1, dtype=torch.complex64) for m in convs]

    def loss(ws):
        return sum((w.real * p.real + w.imag * p.imag).sum() for w, p in zip(ws, probes))

    ref_w = [m.weight() for m in convs]
    loss(ref_w).backward()
    ref_g = [[p.grad.clone() for p in m.parameters()] for m in convs]
    net.zero_grad()
    fcb.prefold(net)
    assert convs[0]._prefolded is not None and convs[-1]._prefolded is None      # the odd-shaped layer folds on its own
    got_w = [m.weight() for m in convs]
    assert all(m._prefolded is None for m in convs)                               # consumed by exactly one forward
    loss(got_w).backward()
    for a, b in zip(got_w, ref_w):
        assert a.shape == b.shape and torch.equal(a, b)
    for m, gs in zip(convs, ref_g):
        for p, g in zip(m.parameters(), gs):
            assert torch.allclose(p.grad, g, rtol=1e-6, atol=1e-7)


def test_prefold_tangent_lin_embedding_matches():
    torch.manual_seed(1)
    net = torch.nn.ModuleList([fcb.TangentLin(5, 3) for _ in range(3)])
    x = torch.randn(7, 5, dtype=torch.complex64)
    from fieldconv_b200.nn import _lin_embedding
    for m in net:
        e = _lin_embedding(m.Re, m.Im)
        xr = torch.view_as_real(x).reshape(7, 10)
        y = torch.view_as_complex((xr @ e).reshape(7, 3, 2))
        y_ref = x @ torch.complex(m.Re, m.Im).t()                      # nn/tangent_lin.py:29
        assert torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6)
    fcb.prefold(net)
    for m in net:
        assert torch.equal(m._preemb[0], _lin_embedding(m.Re, m.Im))
    g = torch.autograd.grad(sum(m._preemb[0].sum() for m in net), [m.Re for m in net])
    assert all(torch.allclose(gi, torch.zeros_like(gi) + 2.0) for gi in g)   # every Re entry appears twice in E


def test_prefold_handoff_goes_stale_with_the_parameters():
    """A layer that did not run in the forward prefold() was called for must not use the pre-optimizer-step filter later
    (ADVICE r1): the hand-over is tagged with the parameters' versions."""
    torch.manual_seed(3)
    net = torch.nn.ModuleList([fcb.FieldConv(4, 4, 1, 3, 1) for _ in range(2)])
    fcb.prefold(net)
    w_old = net[1]._prefolded[0].detach().clone()
    with torch.no_grad():
        net[1].zonal.add_(1.0)                      # what an optimizer step does
    w = net[1].weight()                              # the stale hand-over is ignored and dropped
    assert net[1]._prefolded is None
    assert not torch.equal(w.detach(), w_old)
    assert torch.equal(w, fcb.nn.fold_weights(net[1].zonal, net[1].spherical, net[1].phase, 1, 1))
    assert torch.equal(net[0].weight().detach(), fcb.nn.fold_weights(net[0].zonal, net[0].spherical, net[0].phase, 1, 1).detach())


def test_shared_dense_plan_cache_is_keyed_on_the_tensor_object():
    """the dense-path plan cache must not be fooled by a recycled device address — it keys on the tensor object
    (validated through a weak reference) and one plan serves every layer."""
    from fieldconv_b200 import nn as fnn
    calls = []
    orig = fnn.build_dense_plan
    fnn.build_dense_plan = lambda e, n: calls.append((id(e), n)) or ("plan", len(calls))
    try:
        fnn._DENSE_PLANS.clear()
        e1 = torch.zeros(5, 2, dtype=torch.long)
        p1 = fnn.shared_dense_plan(e1, 4)
        assert fnn.shared_dense_plan(e1, 4) is p1 and len(calls) == 1          # shared across layers / calls
        e2 = torch.zeros(5, 2, dtype=torch.long)                                 # same shape, same contents, another object
        assert fnn.shared_dense_plan(e2, 4) is not p1 and len(calls) == 2
        e1.add_(1)                                                               # in-place edit bumps the version
        assert fnn.shared_dense_plan(e1, 4) is not p1 and len(calls) == 3
        del e1, e2
        import gc
        gc.collect()
        assert not fnn._DENSE_PLANS                                              # entries die with their tensors
    finally:
        fnn.build_dense_plan = orig


def test_operand_bound_travels_with_the_tensor_and_dies_with_an_inplace_update():
    """ops.set_bound / peek_bound (struct fcb_bounds): the bound a producer reported is an attribute of the tensor object,
    tagged with the tensor's version counter; any in-place write makes it stale, and there is no CPU computation of one."""
    from fieldconv_b200 import ops
    t = torch.randn(5, 3, dtype=torch.complex64)
    assert ops.peek_bound(t) is None
    assert ops.bound_of(t) is None                       # CPU tensor: no bound, no fallback computation
    b = torch.tensor([4.0])
    assert ops.set_bound(t, b) is t
    assert ops.peek_bound(t) is b
    t.add_(1.0)
    assert ops.peek_bound(t) is None
    u = t.clone()
    assert ops.peek_bound(u) is None                     # a copy is a new tensor: no inherited attribute
