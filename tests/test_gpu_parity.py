"""GPU: parity of the CUDA FieldConv path (through the C ABI) with the reference.

Tolerance (BASELINE.json north_star): fp32 path  max|a-b| <= 1e-5 max|b|  and
||a-b||_2 <= 1e-5 ||b||_2  for y, grad x and every parameter gradient; indices bit-exact (test_gpu_kernels)."""
import pytest
import torch

import fieldconv_b200 as fcb
from conftest import assert_close_normwise, golden_names, load_golden
from fieldconv_b200.synthetic import merge_meshes, random_features, torus_mesh
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _layer_from_golden(g, precision="auto"):
    m = fcb.FieldConv(g["ci"], g["co"], g["B"], g["R"], g["ftype"], precision=precision)
    m.load_state_dict({"zonal": g["zonal"], "spherical": g["spherical"], "phase": g["phase"]})
    return m.to(DEV)


def _check_against_golden(g, m, y, x):
    gy = g["gy"].to(DEV)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    assert_close_normwise(y, g["y"], TOL, "y")
    assert_close_normwise(x.grad, g["gx"], TOL, "grad x")
    assert_close_normwise(m.zonal.grad, g["g_zonal"], TOL, "grad zonal")
    assert_close_normwise(m.spherical.grad, g["g_spherical"], TOL, "grad spherical")
    if g["ftype"] == 1:
        assert_close_normwise(m.phase.grad, g["g_phase"], TOL, "grad phase")


@pytest.mark.parametrize("name", golden_names("fc_"))
def test_dropin_signature_dense_stencil(name):
    """forward(x, supp_edges, supp_sten) with the reference's own FCPrecomp outputs."""
    g = load_golden(name)
    m = _layer_from_golden(g)
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x, g["supp_edges"].to(DEV), g["supp_sten"].to(DEV))
    _check_against_golden(g, m, y, x)


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("name", golden_names("fc_"))
def test_compact_plan_path(name, precision):
    """forward(x, plan): FCPrecomp arithmetic + CSR built on the device from the raw mesh attributes.
    "fp32" = FP32-FMA contraction kernels; "auto" (module default) = 3xTF32 tensor cores where they stay
    inside the same 1e-5 budget.  Both are held to the fp32-path tolerance."""
    g = load_golden(name)
    m = _layer_from_golden(g, precision)
    plan = fcb.build_plan(g["raw_edges"].to(DEV), g["logMag"].to(DEV), g["logAng"].to(DEV), g["xp"].to(DEV),
                          g["w"].to(DEV), g["R"], g["epsilon"])
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x, plan)
    _check_against_golden(g, m, y, x)


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "2xf16", "2xf16p"])
@pytest.mark.parametrize("name", ["fc_b2r6_f0", "fc_b1r6_f1", "fc_b3r2_f1"])
def test_weight_gradient_from_g_matches_contrib_path(name, precision):
    """Default backward: nothing of size N x K is kept, gW = sum_j conj(xhat_j) G_j (csrc/api.cu backward_common).  The
    older path (contrib saved by the forward, gW = contrib^H gy) must give the same y / grad x bit for bit and the same
    gW within the fp32 budget; both are checked against the reference's golden gradient through the fold."""
    from fieldconv_b200 import _lib, nn as fnn, ops
    g = load_golden(name)
    plan = fcb.build_plan(g["raw_edges"].to(DEV), g["logMag"].to(DEV), g["logAng"].to(DEV), g["xp"].to(DEV),
                          g["w"].to(DEV), g["R"], g["epsilon"])
    W = restate.fold_weights(g["zonal"], g["spherical"], g["phase"], g["ftype"], g["B"]).to(DEV)
    flags = fnn._PRECISIONS[precision]
    if g["ci"] % 2 or g["co"] % 2:
        pytest.skip("the raw op needs even channel counts (the module pads)")
    if flags & _lib.FLAG_PACKED and not _lib.pk_supported(g["x"].shape[0], g["ci"], g["co"], g["B"], g["R"]):
        pytest.skip("packed path does not support this shape")
    outs = []
    for keep in (True, False):
        x = g["x"].to(DEV).requires_grad_(True)
        w = W.clone().requires_grad_(True)
        y = ops.field_conv(x, w, plan, g["B"], flags, keep_contrib=keep)
        gy = g["gy"].to(DEV)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        outs.append((y.detach(), x.grad, w.grad))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert_close_normwise(outs[1][2], outs[0][2], 2e-5 if precision == "3xtf32" else 5e-6, "gW from G vs from contrib")


def test_block_matches_reference():
    g = load_golden("block_b2r6")
    blk = fcb.FCResNetBlock(g["ci"], g["co"], g["B"], g["R"], 1)
    blk.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("p.")})
    blk = blk.to(DEV)
    plan = fcb.build_plan(g["raw_edges"].to(DEV), g["logMag"].to(DEV), g["logAng"].to(DEV), g["xp"].to(DEV),
                          g["w"].to(DEV), g["R"], g["epsilon"])
    for mode in ("plan", "dense"):
        blk.zero_grad()
        x = g["x"].to(DEV).requires_grad_(True)
        y = blk(x, plan) if mode == "plan" else blk(x, g["supp_edges"].to(DEV), g["supp_sten"].to(DEV))
        gy = g["gy"].to(DEV)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        assert_close_normwise(y, g["y"], TOL, mode + ": block y")
        assert_close_normwise(x.grad, g["gx"], TOL, mode + ": block grad x")
        for k, p in blk.named_parameters():
            assert_close_normwise(p.grad, g["g." + k], 2e-5 if "bias" in k else TOL, mode + ": grad " + k)


@pytest.mark.parametrize("c,B,precision", [(48, 2, "auto"), (32, 1, "auto"), (16, 1, "fp32"), (128, 1, "auto"), (32, 1, "3xtf32")])
def test_block_epilogue_fused_matches_separate_kernels(c, B, precision, monkeypatch):
    """FCResNetBlock with nonlin1 / (residual + nonlin2) applied in the contraction kernels' epilogue (fcb_fwd_act_*) against the
    same block with TangentNonLin and the add as separate kernels: outputs, grad x and every parameter gradient; and the
    fused run must not launch the stand-alone modReLU forward where the 2xFP16 kernel can carry the epilogue."""
    from fieldconv_b200 import _lib, nn as fnn
    mesh = torus_mesh(26, deg=40.0, seed=8, device=DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 6, mesh.epsilon)
    torch.manual_seed(3)
    blk = fcb.FCResNetBlock(c, c, B, 6, 1, precision=precision).to(DEV)
    with torch.no_grad():
        blk.nonlin1.bias.uniform_(-0.3, 0.3)
        blk.nonlin2.bias.uniform_(-0.3, 0.3)
    gy = random_features(mesh.num_nodes, c, seed=5, zero_frac=0, device=DEV)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setattr(fnn, "BLOCK_EPILOGUE", mode)
        blk.zero_grad()
        x = random_features(mesh.num_nodes, c, seed=4, device=DEV).requires_grad_(True)
        _lib.profile_enable(512)
        y = blk(x, plan)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        torch.cuda.synchronize()
        names = [n for n, _ in _lib.profile_collect(512)]
        res[mode] = (names, y.detach(), x.grad, {k: p.grad.clone() for k, p in blk.named_parameters()})
    assert "modrelu_fwd" in res["0"][0]
    assert "modrelu_fwd" not in res["1"][0]
    if precision == "auto":          # one un-split 2xFP16 launch carries the epilogue: no pointwise kernel at all
        assert "res_modrelu" not in res["1"][0], res["1"][0]
    else:                            # paths that cannot fuse run ONE pointwise kernel instead of add + modReLU
        assert "res_modrelu" in res["1"][0]
    assert_close_normwise(res["1"][1], res["0"][1], 2e-6, "block y")
    assert_close_normwise(res["1"][2], res["0"][2], 2e-6, "block grad x")
    for k in res["0"][3]:
        assert_close_normwise(res["1"][3][k], res["0"][3][k], 5e-6, "grad " + k)


@pytest.mark.parametrize("c,B", [(48, 2), (32, 1)])
def test_operand_bounds_reported_by_producers_and_shared_by_consumers(c, B, monkeypatch):
    """struct fcb_bounds: (1) the bounds the kernels report equal the true maxima; (2) a two-block stack run with the bounds
    travelling along the tensors matches the run where every call takes its own pass over its operands; (3) with the
    bounds the step launches fewer operand-maximum passes."""
    from fieldconv_b200 import _lib, ops
    mesh = torus_mesh(26, deg=40.0, seed=8, device=DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 6, mesh.epsilon)
    n = mesh.num_nodes
    torch.manual_seed(3)
    blocks = torch.nn.ModuleList([fcb.FCResNetBlock(c, c, B, 6, 1).to(DEV) for _ in range(2)])
    with torch.no_grad():
        for blk in blocks:
            blk.nonlin1.bias.uniform_(-0.3, 0.3)
            blk.nonlin2.bias.uniform_(-0.3, 0.3)
    # (1) producers: the activation bound of the fused epilogue, the gradient bound of the modReLU backward, fcb_bound_f32
    x = random_features(n, c, seed=4, device=DEV)
    h = blocks[0](x, plan)
    hb = ops.peek_bound(h)
    assert hb is not None, "the fused block epilogue did not attach a bound to its output"
    true = float(h.detach().abs().max())
    assert true <= float(hb) <= true * (1 + 1e-5) + 1e-30
    z = random_features(n, c, seed=6, zero_frac=0, device=DEV)
    g = random_features(n, c, seed=7, zero_frac=0, device=DEV)
    gz, _, gzb = ops.modrelu_bwd(z, blocks[0].nonlin1.bias.detach(), g)
    true = float(torch.view_as_real(gz).abs().max())
    assert float(gzb) == pytest.approx(true, rel=1e-6)
    xb = ops.bound_of(x)
    true = float(x.abs().max())
    assert true <= float(xb) <= true * (1 + 1e-5)
    assert ops.bound_of(x) is xb, "the bound is cached on the tensor"
    x2 = x.clone()
    ops.set_bound(x2, xb)
    x2.mul_(2.0)
    assert ops.peek_bound(x2) is None, "an in-place update must invalidate the attached bound"
    # (2) + (3)
    gy = random_features(n, c, seed=5, zero_frac=0, device=DEV)
    res = {}
    for mode in (True, False):
        monkeypatch.setattr(ops, "BOUNDS", mode)
        blocks.zero_grad()
        xin = random_features(n, c, seed=4, device=DEV).requires_grad_(True)
        _lib.profile_enable(2048)
        y = xin
        for blk in blocks:
            y = blk(y, plan)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        torch.cuda.synchronize()
        names = [k for k, _ in _lib.profile_collect(2048)]
        passes = sum(1 for k in names if k in ("absmax_mod", "lin_absmax_mod"))
        big = [k for k in names if k in ("absmax", "lin_absmax")]
        res[mode] = (passes, len(big), y.detach(), xin.grad, {k: p.grad.clone() for k, p in blocks.named_parameters()})
    assert res[True][0] + res[True][1] < res[False][0] + res[False][1], (res[True][:2], res[False][:2])
    assert_close_normwise(res[True][2], res[False][2], 2e-6, "y")
    assert_close_normwise(res[True][3], res[False][3], 2e-6, "grad x")
    for k in res[False][4]:
        assert_close_normwise(res[True][4][k], res[False][4][k], 5e-6, "grad " + k)


def _oracle_layer(mesh, x, m, gy, dtype=torch.complex128):
    """fp64 folded-form oracle on the CPU for a synthetic mesh."""
    B, R = m.B, m.R
    e, sten, _, _, _ = restate.fc_precomp(mesh.logMag.cpu(), mesh.logAng.cpu(), mesh.w.cpu(), mesh.supp_edges.cpu(),
                                          mesh.xp.cpu(), B, R, mesh.epsilon)
    ps = [p.detach().cpu().double().requires_grad_(True) for p in (m.zonal, m.spherical, m.phase)]
    W = restate.fold_weights(ps[0], ps[1], ps[2], m.ftype, B)
    xd = x.detach().cpu().to(dtype)
    y = restate.field_conv_lean(xd, e, sten.to(dtype), W.detach(), B)
    gx, gw = restate.field_conv_backward(xd, e, sten.to(dtype), W.detach(), B, gy.cpu().to(dtype))
    (W.real * gw.real + W.imag * gw.imag).sum().backward()
    return y, gx, [p.grad for p in ps]


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("n_side,ci,co,B,R,ftype", [(71, 32, 32, 1, 6, 1), (24, 48, 48, 2, 6, 1), (20, 18, 10, 3, 3, 2),
                                                     (16, 128, 128, 2, 6, 0)])
def test_synthetic_mesh_vs_fp64_oracle(n_side, ci, co, B, R, ftype, precision):
    """BASELINE config 1 (5k vertices, C=32, B=1, R=6) and cut-down configs 2/3 against the fp64 oracle."""
    mesh = torus_mesh(n_side, deg=40.0, seed=1, device=DEV)
    torch.manual_seed(0)
    m = fcb.FieldConv(ci, co, B, R, ftype, precision=precision).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = m(x, plan)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    y_ref, gx_ref, gp_ref = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), TOL, "y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, "grad x")
    assert_close_normwise(m.zonal.grad, gp_ref[0].float(), TOL, "grad zonal")
    assert_close_normwise(m.spherical.grad, gp_ref[1].float(), TOL, "grad spherical")
    if ftype == 1:
        assert_close_normwise(m.phase.grad, gp_ref[2].float(), TOL, "grad phase")


@pytest.mark.parametrize("n_side,ci,co,B,R,ftype", [
    (83, 128, 128, 2, 6, 1),      # BASELINE configs[2] at FULL size: 6889 vertices, C=128, band_limit 2 (tensor-core regime)
    (22, 16, 16, 1, 2, 1),        # configs[4] corners: C=16, n_rings 2
    (18, 64, 64, 3, 2, 1),        # band_limit 3, n_rings 2
    (14, 256, 256, 1, 2, 0),      # C=256
    (16, 64, 64, 3, 6, 2),        # band_limit 3, n_rings 6, complex filters
    (20, 32, 64, 2, 6, 1),        # C_in != C_out
])
def test_sweep_corner_layers_vs_fp64_oracle(n_side, ci, co, B, R, ftype):
    """Full-size cfg 3 and the corners of the cfg-5 sweep (C in {16, 64, 256}, band_limit 3, n_rings 2) as whole layers
    (forward, grad x, parameter gradients) against the fp64 oracle, default precision."""
    mesh = torus_mesh(n_side, deg=40.0, seed=3, device=DEV)
    torch.manual_seed(0)
    m = fcb.FieldConv(ci, co, B, R, ftype).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = m(x, plan)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    y_ref, gx_ref, gp_ref = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), TOL, "y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, "grad x")
    assert_close_normwise(m.zonal.grad, gp_ref[0].float(), TOL, "grad zonal")
    assert_close_normwise(m.spherical.grad, gp_ref[1].float(), TOL, "grad spherical")
    if ftype == 1:
        assert_close_normwise(m.phase.grad, gp_ref[2].float(), TOL, "grad phase")


@pytest.mark.parametrize("precision", ["fp32", "auto", "2xf16p"])
def test_degree_skewed_mesh(precision):
    """Power-law in-degrees with many isolated rows and a few very long rows (longer than the 768 plan records a CTA of the
    aggregation kernel stages at once): rows of one warp differ in length and in where their rings end."""
    from fieldconv_b200.synthetic import skewed_degree
    mesh = skewed_degree(torus_mesh(36, deg=120.0, seed=7, device=DEV), seed=11)
    deg = torch.bincount(mesh.supp_edges[:, 1], minlength=mesh.num_nodes)
    assert int((deg == 0).sum()) > 50 and int(deg.max()) > 100 and float(deg.float().std()) > 20
    ci = co = 32
    torch.manual_seed(0)
    m = fcb.FieldConv(ci, co, 1, 6, 1, precision=precision).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 6, mesh.epsilon)
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = m(x, plan)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    y_ref, gx_ref, gp_ref = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), TOL, "y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, "grad x")
    assert_close_normwise(m.zonal.grad, gp_ref[0].float(), TOL, "grad zonal")
    assert float(y[deg == 0].abs().max()) == 0.0


def test_bitwise_determinism():
    """The segmented reduction and the split reductions use fixed orders: two runs agree bit for bit
    (the reference's scatter_add uses float atomics on CUDA and does not)."""
    mesh = torus_mesh(40, deg=40.0, seed=4, device=DEV, permute=True)
    torch.manual_seed(1)
    m = fcb.FieldConv(32, 32, 1, 6, 1).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 6, mesh.epsilon)
    gy = random_features(mesh.num_nodes, 32, seed=9, zero_frac=0, device=DEV)
    res = []
    for _ in range(2):
        m.zero_grad()
        x = random_features(mesh.num_nodes, 32, seed=8, device=DEV).requires_grad_(True)
        y = m(x, plan)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        res.append([y.detach().clone(), x.grad.clone()] + [p.grad.clone() for p in m.parameters()])
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_cuda_graph_capture_replay(precision):
    """The ops enqueue on the current stream, allocate only through torch's allocator and never synchronise, so a
    whole FCResNetBlock fwd+bwd captures into one CUDA graph; replays on new inputs match eager runs bit for bit."""
    mesh = torus_mesh(30, deg=40.0, seed=6, device=DEV)
    torch.manual_seed(2)
    blk = fcb.FCResNetBlock(16, 16, 1, 6, 1, precision=precision).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 6, mesh.epsilon)
    n = mesh.num_nodes
    x_static = random_features(n, 16, seed=1, device=DEV).requires_grad_(True)
    gy = random_features(n, 16, seed=2, zero_frac=0, device=DEV)

    def fwd_bwd():
        for p in blk.parameters():
            p.grad = None
        x_static.grad = None
        y = blk(x_static, plan)
        y.backward(gy)
        return y

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fwd_bwd()                      # warm-up outside the capture (one-time cudaFuncSetAttribute calls)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y_static = fwd_bwd()
    grads_static = [x_static.grad] + [p.grad for p in blk.parameters()]
    for seed in (11, 12):
        x_new = random_features(n, 16, seed=seed, device=DEV)
        with torch.no_grad():
            x_static.copy_(x_new)
        g.replay()
        torch.cuda.synchronize()
        got = [y_static.detach().clone()] + [t.clone() for t in grads_static]
        xe = x_new.clone().requires_grad_(True)
        for p in blk.parameters():
            p.grad = None
        ye = blk(xe, plan)
        ye.backward(gy)
        want = [ye.detach(), xe.grad] + [p.grad for p in blk.parameters()]
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        for p, gs in zip(blk.parameters(), grads_static[1:]):
            p.grad = gs                     # hand the static buffers back before the next replay


def test_full_size_properties_config2_layer():
    """At BASELINE config-2 layer size (16 x 5k-vertex meshes merged, C=48, B=2, R=6) the oracle is too slow, so
    check size-independent properties: (a) block-diagonal batching == per-mesh results, (b) linearity in W,
    (c) <gy, J dx> == <J^T gy, dx> (adjoint identity between forward and backward, away from origin entries),
    (d) gauge equivariance."""
    B, R, C = 2, 6, 48
    meshes = [torus_mesh(71, deg=40.0, seed=10 + i, device=DEV) for i in range(16)]
    big = merge_meshes(meshes)
    torch.manual_seed(3)
    m = fcb.FieldConv(C, C, B, R, 1).to(DEV)
    plan = fcb.build_plan(big.supp_edges, big.logMag, big.logAng, big.xp, big.w, R, big.epsilon)
    x = random_features(big.num_nodes, C, seed=5, zero_frac=0.01, device=DEV)
    with torch.no_grad():
        y = m(x, plan)
        # (a) one mesh alone
        k = 3
        n0 = sum(mm.num_nodes for mm in meshes[:k])
        pk = fcb.build_plan(meshes[k].supp_edges, meshes[k].logMag, meshes[k].logAng, meshes[k].xp, meshes[k].w, R,
                            meshes[k].epsilon)
        yk = m(x[n0:n0 + meshes[k].num_nodes].contiguous(), pk)
        assert torch.equal(yk, y[n0:n0 + meshes[k].num_nodes])
        # (b) linearity in the filter: conv(W1 + 2 W2) = conv(W1) + 2 conv(W2)
        from fieldconv_b200 import ops
        w1 = m.weight()
        w2 = torch.randn_like(w1) * 0.05
        lhs = ops.field_conv(x, (w1 + 2 * w2).contiguous(), plan, B)
        rhs = ops.field_conv(x, w1, plan, B) + 2 * ops.field_conv(x, w2.contiguous(), plan, B)
        assert_close_normwise(lhs, rhs, 1e-5, "linearity in W")
    # (c) adjoint identity with a finite perturbation along dx (away from the origin entries)
    xg = random_features(big.num_nodes, C, seed=6, zero_frac=0.0, device=DEV).requires_grad_(True)
    gy = random_features(big.num_nodes, C, seed=7, zero_frac=0.0, device=DEV)
    yg = m(xg, plan)
    (yg.real * gy.real + yg.imag * gy.imag).sum().backward()
    dx = random_features(big.num_nodes, C, seed=8, zero_frac=0.0, device=DEV)
    h = 1e-3
    with torch.no_grad():
        yp = m((xg + h * dx).detach(), plan).to(torch.complex128)
        ym = m((xg - h * dx).detach(), plan).to(torch.complex128)
        jdx = (yp - ym) / (2 * h)
        lhs = (jdx.real * gy.real.double() + jdx.imag * gy.imag.double()).sum()
        rhs = (xg.grad.real.double() * dx.real.double() + xg.grad.imag.double() * dx.imag.double()).sum()
    assert abs(float(lhs - rhs)) <= 5e-3 * abs(float(rhs)), (float(lhs), float(rhs))
    # (d) gauge equivariance on the full batch
    g = torch.Generator(device="cpu").manual_seed(11)
    alpha = ((torch.rand(big.num_nodes, generator=g) * 2 - 1) * 3.0).to(DEV)
    src, tgt = big.supp_edges[:, 0], big.supp_edges[:, 1]
    rot = torch.polar(torch.ones_like(alpha), -alpha)
    plan2 = fcb.build_plan(big.supp_edges, big.logMag, big.logAng - alpha[src],
                           big.xp * torch.polar(torch.ones_like(alpha[src]), alpha[src] - alpha[tgt]), big.w, R, big.epsilon)
    with torch.no_grad():
        y2 = m((x * rot[:, None]).contiguous(), plan2)
    assert_close_normwise(y2, y * rot[:, None], 2e-5, "gauge equivariance")


# Tensor-core variants: tolerance stated separately from the fp32 path (BASELINE.json north_star).
#   3xTF32 : operand rounding is compensated; what remains is the truncating fp32 accumulate of the tensor core,
#            spread over several TMEM accumulators -> a few 1e-6 for K ~ 3k; bound used here 2e-5.
#   TF32   : 10-bit mantissa operands, ~1e-3; bound 5e-3.
#   2xFP16 : scaled fp16 (hi, lo) operand pairs carry 22 bits, half as many accumulating MMAs as 3xTF32; this is what
#            precision="auto" selects, so it is held to the fp32 path's 1e-5.
@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-5), ("tf32", 5e-3), ("2xf16", 1e-5)])
@pytest.mark.parametrize("name", golden_names("fc_"))
def test_tensor_core_precision_golden(name, precision, tol):
    g = load_golden(name)
    m = _layer_from_golden(g, precision)
    plan = fcb.build_plan(g["raw_edges"].to(DEV), g["logMag"].to(DEV), g["logAng"].to(DEV), g["xp"].to(DEV),
                          g["w"].to(DEV), g["R"], g["epsilon"])
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x, plan)
    gy = g["gy"].to(DEV)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    assert_close_normwise(y, g["y"], tol, precision + " y")
    assert_close_normwise(x.grad, g["gx"], tol, precision + " grad x")
    assert_close_normwise(m.zonal.grad, g["g_zonal"], tol, precision + " grad zonal")
    assert_close_normwise(m.spherical.grad, g["g_spherical"], tol, precision + " grad spherical")


@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-5), ("tf32", 5e-3), ("2xf16", 1e-5)])
@pytest.mark.parametrize("n_side,ci,co,B,R", [(71, 32, 32, 1, 6), (24, 48, 48, 2, 6), (16, 128, 128, 2, 6)])
def test_tensor_core_precision_vs_fp64_oracle(n_side, ci, co, B, R, precision, tol):
    mesh = torus_mesh(n_side, deg=40.0, seed=1, device=DEV)
    torch.manual_seed(0)
    m = fcb.FieldConv(ci, co, B, R, 1, precision=precision).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = m(x, plan)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    y_ref, gx_ref, gp_ref = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), tol, precision + " y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), tol, precision + " grad x")
    assert_close_normwise(m.zonal.grad, gp_ref[0].float(), tol, precision + " grad zonal")
    assert_close_normwise(m.spherical.grad, gp_ref[1].float(), tol, precision + " grad spherical")
    assert_close_normwise(m.phase.grad, gp_ref[2].float(), tol, precision + " grad phase")


@pytest.mark.parametrize("name", golden_names("fc_"))
def test_fcprecomp_dropin_matches_reference_outputs(name):
    """fieldconv_b200.FCPrecomp(band_limit, n_rings, epsilon)(data) returns the reference transform's four outputs
    (transforms/fc_precomp.py:53-97): kept edges bit-exact and in input order, supp_sten / ln / wxp to fp32 rounding;
    and the reference call forward(x, supp_edges, supp_sten) on those outputs rides the attached compact plan."""
    from types import SimpleNamespace
    g = load_golden(name)
    data = SimpleNamespace(supp_edges=g["raw_edges"].to(DEV), logMag=g["logMag"].to(DEV), logAng=g["logAng"].to(DEV),
                           xp=g["xp"].to(DEV), w=g["w"].to(DEV))
    pre = fcb.FCPrecomp(g["B"], g["R"], g["epsilon"])
    e, sten, ln, wxp = pre(data)
    assert e.dtype == torch.int64 and torch.equal(e.cpu(), g["supp_edges"])
    assert sten.shape == g["supp_sten"].shape and sten.dtype == torch.complex64
    assert_close_normwise(sten, g["supp_sten"], 1e-6, "supp_sten")
    assert_close_normwise(ln, g["ln"], 1e-6, "ln")
    assert_close_normwise(wxp, g["wxp"], 1e-6, "wxp")
    # zero pattern of the radial two-tap stencil is exact (ring floor f bit-identical)
    assert torch.equal(sten.cpu().abs() > 0, g["supp_sten"].abs() > 0)
    from fieldconv_b200.transforms import attached_plan
    assert attached_plan(e, sten, g["R"], g["n"]) is not None
    m = _layer_from_golden(g)
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x, e, sten)                           # compact fast path through the attached plan
    _check_against_golden(g, m, y, x)
    m2 = _layer_from_golden(g)
    x2 = g["x"].to(DEV).requires_grad_(True)
    y2 = m2(x2, e.clone(), sten.clone())        # clones carry no plan: dense-stencil path on our own FCPrecomp output
    _check_against_golden(g, m2, y2, x2)


@pytest.mark.parametrize("precision", ["fp32", "2xf16", "2xf16p"])
def test_reference_edge_cases_on_a_synthetic_mesh(precision):
    """SURVEY.md appendix B at a shape every contraction path (FP32 FMA, 2xFP16, packed 2xFP16) accepts: targets without
    incoming edges (one of them the last vertex) give y = 0, duplicate edges are summed, edges with r > epsilon are
    dropped before the weight normalisation, the edge order is arbitrary, 1 % of the features are exact zeros."""
    from fieldconv_b200.synthetic import with_edge_cases
    base = torus_mesh(20, deg=30.0, seed=9, device=DEV)
    iso = (3, 57, base.num_nodes - 1)
    mesh = with_edge_cases(base, isolated=iso, duplicate=60, shuffle_seed=1, far=45)
    ci = co = 32
    B, R = 1, 6
    torch.manual_seed(0)
    m = fcb.FieldConv(ci, co, B, R, 1, precision=precision).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    assert plan.num_edges == mesh.supp_edges.shape[0] - 45
    x = random_features(mesh.num_nodes, ci, seed=2, device=DEV).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = m(x, plan)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    assert float(y[list(iso)].abs().max()) == 0.0
    y_ref, gx_ref, gp_ref = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), TOL, precision + " y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, precision + " grad x")
    assert_close_normwise(m.zonal.grad, gp_ref[0].float(), TOL, precision + " grad zonal")
    assert_close_normwise(m.spherical.grad, gp_ref[1].float(), TOL, precision + " grad spherical")
    assert_close_normwise(m.phase.grad, gp_ref[2].float(), TOL, precision + " grad phase")
