"""GPU: the packed-operand variant of the 2xFP16 path (precision="2xf16p": fcb_fwd_pk_f32 / fcb_bwd_pk_f32 — the
aggregation kernels write the scaled fp16 (hi, lo) tile images, the contraction kernels bulk-copy them).  Same
arithmetic as "2xf16", so it is held to the fp32 path's 1e-5 normwise tolerance against the fp64 oracle and the
reference's golden outputs, and to 3e-6 against the unpacked 2xFP16 path.

First green B200 run: profiles/r01f_pytest_packed.log.  FIELDCONV_B200_TEST_PACKED=0 skips this file."""
import os

import pytest
import torch

import fieldconv_b200 as fcb
from conftest import assert_close_normwise, load_golden
from fieldconv_b200 import _lib, ops
from fieldconv_b200.synthetic import random_features, torus_mesh
from oracle import restate
from test_gpu_parity import _oracle_layer

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FIELDCONV_B200_TEST_PACKED", "1") == "0", reason="packed-path tests disabled")]
DEV = "cuda:0"
TOL = 1e-5


def _run(mesh, plan, ci, co, B, R, precision, ftype=1, seed=0, keep=None):
    torch.manual_seed(seed)
    m = fcb.FieldConv(ci, co, B, R, ftype, precision=precision).to(DEV)
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = m(x, plan)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    return m, x, gy, y


# n_side 71: BASELINE cfg 1 (5 041 vertices: 39.4 row tiles, so the tail tile is exercised); 24: cut-down cfg 2 (2K = 2880
# columns = 45 chunks: an odd chunk count, the weight gradient's last column tile has one chunk); 16: cut-down cfg 3
# (two 128-column output chunks over the same packed operand); 9: 81 vertices, less than one row tile.
@pytest.mark.parametrize("n_side,ci,co,B,R", [(71, 32, 32, 1, 6), (24, 48, 48, 2, 6), (16, 128, 128, 2, 6), (9, 32, 32, 1, 6),
                                               (30, 64, 32, 1, 4), (30, 32, 64, 1, 2)])
def test_packed_vs_fp64_oracle(n_side, ci, co, B, R):
    if not _lib.pk_supported(n_side * n_side, ci, co, B, R):
        pytest.skip("shape not taken by the packed path")
    mesh = torus_mesh(n_side, deg=40.0, seed=1, device=DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    before = _lib.launch_count()
    m, x, gy, y = _run(mesh, plan, ci, co, B, R, "2xf16p")
    assert _lib.launch_count() > before
    y_ref, gx_ref, gp_ref = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), TOL, "2xf16p y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, "2xf16p grad x")
    assert_close_normwise(m.zonal.grad, gp_ref[0].float(), TOL, "2xf16p grad zonal")
    assert_close_normwise(m.spherical.grad, gp_ref[1].float(), TOL, "2xf16p grad spherical")
    assert_close_normwise(m.phase.grad, gp_ref[2].float(), TOL, "2xf16p grad phase")


def test_packed_matches_unpacked_2xf16_and_is_deterministic():
    mesh = torus_mesh(40, deg=40.0, seed=4, device=DEV)
    ci = co = 48
    B, R = 2, 6
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    outs = {}
    for prec in ("2xf16", "2xf16p", "2xf16p "):
        m, x, gy, y = _run(mesh, plan, ci, co, B, R, prec.strip())
        outs[prec] = (y.detach().clone(), x.grad.clone(), m.zonal.grad.clone(), m.spherical.grad.clone(), m.phase.grad.clone())
    for a, b, what in zip(outs["2xf16p"], outs["2xf16"], ("y", "gx", "g_zonal", "g_spherical", "g_phase")):
        assert_close_normwise(a, b, 3e-6, "packed vs unpacked " + what)
    for a, b in zip(outs["2xf16p"], outs["2xf16p "]):
        assert torch.equal(a, b)            # fixed reduction orders: bit-identical run to run


def test_packed_gw_from_g_matches_contrib_path():
    mesh = torus_mesh(30, deg=40.0, seed=5, device=DEV)
    ci, co, B, R = 32, 32, 1, 6
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    torch.manual_seed(0)
    layer = fcb.FieldConv(ci, co, B, R, 1)
    W = layer.weight().detach().to(DEV)
    flags = _lib.GEMM_TC_2XF16 | _lib.FLAG_PACKED
    outs = []
    for keep in (True, False):
        x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
        w = W.clone().requires_grad_(True)
        gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
        y = ops.field_conv(x, w, plan, B, flags, keep_contrib=keep)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        outs.append((y.detach(), x.grad, w.grad))
    # keep=False (default): nothing of size N x K is kept, gW from G and xhat; keep=True: gW from the saved packed contrib
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert_close_normwise(outs[1][2], outs[0][2], 5e-6, "gW from G vs from contrib (packed)")


def test_packed_golden_block():
    """FCResNetBlock(…, precision="2xf16p") against the unmodified reference's outputs where the shape allows the packed
    path, else it must raise (explicit precision never silently changes kernels)."""
    g = load_golden("block_b2r6")
    blk = fcb.FCResNetBlock(g["ci"], g["co"], g["B"], g["R"], 1, precision="2xf16p")
    blk.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("p.")})
    blk = blk.to(DEV)
    plan = fcb.build_plan(g["raw_edges"].to(DEV), g["logMag"].to(DEV), g["logAng"].to(DEV), g["xp"].to(DEV),
                          g["w"].to(DEV), g["R"], g["epsilon"])
    x = g["x"].to(DEV).requires_grad_(True)
    n = x.shape[0]
    if not (_lib.pk_supported(n, g["ci"], g["co"], g["B"], g["R"]) and _lib.pk_supported(n, g["co"], g["co"], g["B"], g["R"])):
        with pytest.raises(RuntimeError, match="2xf16p"):
            blk(x, plan)
        return
    y = blk(x, plan)
    gy = g["gy"].to(DEV)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    assert_close_normwise(y, g["y"], TOL, "block y")
    assert_close_normwise(x.grad, g["gx"], TOL, "block grad x")
    for k, p in blk.named_parameters():
        assert_close_normwise(p.grad, g["g." + k], 2e-5 if "bias" in k else TOL, "grad " + k)


def test_packed_scale_bound_holds_for_extreme_inputs():
    """The a-priori operand scale max|x| * max_row sum|wxp| must cover huge and tiny feature magnitudes and a mesh whose
    vertex weights vary by 1e4 (fp16 planes overflow if the bound is wrong)."""
    mesh = torus_mesh(24, deg=40.0, seed=6, device=DEV)
    mesh.w = (mesh.w * torch.logspace(-2, 2, mesh.num_nodes, device=DEV).reshape(mesh.w.shape)).contiguous()
    ci = co = 32
    B, R = 1, 6
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    assert 0.0 < float(plan.norms[0]) <= 1.001
    for scale in (1e-12, 1.0, 1e12):
        torch.manual_seed(0)
        ref = fcb.FieldConv(ci, co, B, R, 1, precision="fp32").to(DEV)
        pk = fcb.FieldConv(ci, co, B, R, 1, precision="2xf16p").to(DEV)
        pk.load_state_dict(ref.state_dict())
        res = []
        for m in (ref, pk):
            x = (random_features(mesh.num_nodes, ci, seed=2, device=DEV) * scale).requires_grad_(True)
            gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV) / scale
            y = m(x, plan)
            (y.real * gy.real + y.imag * gy.imag).sum().backward()
            res.append((y.detach(), x.grad, m.zonal.grad))
        for a, b, what in zip(res[1], res[0], ("y", "gx", "g_zonal")):
            assert torch.isfinite(torch.view_as_real(a) if a.is_complex() else a).all(), what
            assert_close_normwise(a, b, TOL, "scale %g %s" % (scale, what))


def test_packed_full_size_config2_layer_properties():
    """BASELINE cfg 2 layer at full size (80 656 vertices, 3.24 M edges, C=48, B=2): packed vs FP32-FMA path, plus the
    adjoint identity <gy, J dx> = <J^T gy, dx> of the packed path's own forward / backward in W (linear in W)."""
    from fieldconv_b200.synthetic import merge_meshes
    mesh = merge_meshes([torus_mesh(71, deg=40.0, seed=s, device=DEV) for s in range(16)])
    ci = co = 48
    B, R = 2, 6
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    torch.manual_seed(0)
    ref = fcb.FieldConv(ci, co, B, R, 1, precision="fp32").to(DEV)
    pk = fcb.FieldConv(ci, co, B, R, 1, precision="2xf16p").to(DEV)
    pk.load_state_dict(ref.state_dict())
    res = []
    for m in (ref, pk):
        x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
        gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
        y = m(x, plan)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        res.append((y.detach(), x.grad, m.zonal.grad, m.spherical.grad, m.phase.grad))
    for a, b, what in zip(res[1], res[0], ("y", "gx", "g_zonal", "g_spherical", "g_phase")):
        assert_close_normwise(a, b, TOL, "cfg2 packed vs fp32 " + what)
    # adjoint identity in W: y is linear in W, so <gy, y(dW)> == <gW, dW>
    W = pk.weight().detach()
    flags = _lib.GEMM_TC_2XF16 | _lib.FLAG_PACKED
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    w = W.clone().requires_grad_(True)
    y = ops.field_conv(x, w, plan, B, flags)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    dW = torch.randn_like(W)
    y2 = ops.field_conv(x, dW, plan, B, flags)
    lhs = float((y2.real * gy.real + y2.imag * gy.imag).sum())
    rhs = float((w.grad.real * dW.real + w.grad.imag * dW.imag).sum())
    assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


def test_packed_rejects_unsupported_shape():
    mesh = torus_mesh(12, deg=30.0, seed=1, device=DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 3, mesh.epsilon)
    m = fcb.FieldConv(6, 6, 1, 3, 1, precision="2xf16p").to(DEV)     # 2*R*M*Ci = 108: not a multiple of 64
    x = random_features(mesh.num_nodes, 6, seed=2, device=DEV)
    with pytest.raises(RuntimeError, match="2xf16p"):
        m(x, plan)


@pytest.mark.parametrize("n_side,c,B,R", [(24, 48, 2, 6), (37, 32, 1, 6)])
def test_packed_contrib_decodes_to_fp32_contrib(n_side, c, B, R):
    """Format-level parity: the PK buffer the packing aggregation writes, decoded on the host side
    (fieldconv_b200.packed.unpack: un-swizzle, (hi + lo) / scale), equals the fp32 contrib of the unpacked path to the
    22 bits the (hi, lo) split carries; the tail rows of the last 128-row tile are zero; y agrees."""
    from fieldconv_b200 import packed
    mesh = torus_mesh(n_side, deg=40.0, seed=7, device=DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    n = mesh.num_nodes
    assert _lib.pk_supported(n, c, c, B, R)
    torch.manual_seed(0)
    W = fcb.FieldConv(c, c, B, R, 1).weight().detach().to(DEV)
    x = random_features(n, c, seed=2, device=DEV)
    args = (x, W, plan.rowptr_tgt, plan.rec_tgt, plan.rot_tgt, plan.rowptr_src, plan.rec_src, plan.rot_src, plan.norms, B, R)
    y32, c32, cmax32 = ops.fc_fwd(*args, _lib.GEMM_TC_2XF16, True)
    ypk, cpk, bound = ops.fc_fwd(*args, _lib.GEMM_TC_2XF16 | _lib.FLAG_PACKED, True)
    cols = 2 * R * (2 * B + 1) * c
    ref = torch.view_as_real(c32)[:n].reshape(n, cols)
    dec = packed.unpack(cpk, n, cols, float(bound))
    amax = float(ref.abs().max())
    assert float(cmax32) == amax                                     # the unpacked path tracks the exact maximum
    assert amax <= float(bound) <= 64.0 * amax                       # the a-priori bound holds and is not absurdly loose
    assert float((dec[:n] - ref).abs().max()) <= 2.0 ** -20 * amax
    assert float(dec[n:].abs().max()) == 0.0
    assert_close_normwise(ypk, y32, 3e-6, "y packed vs unpacked")
