"""GPU: the fused forward (gather -> shared-memory operand tile -> tcgen05, csrc/fused_fwd.cu; band_limit <= 1) through the
C ABI: against the fp64 oracle at the fp32 path's 1e-5, against the unfused kernels, gradients through the default
backward (gW from G and xhat), edge cases, determinism."""
import pytest
import torch

import fieldconv_b200 as fcb
from conftest import assert_close_normwise
from fieldconv_b200 import _lib, ops
from fieldconv_b200.synthetic import random_features, torus_mesh, with_edge_cases
from test_gpu_parity import _oracle_layer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5
FUSED = _lib.GEMM_TC_2XF16 | _lib.FLAG_FUSED


def _run(mesh, ci, co, B, R, flags, ftype=1, seed=0, scale=1.0):
    torch.manual_seed(seed)
    m = fcb.FieldConv(ci, co, B, R, ftype).to(DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, R, mesh.epsilon)
    x = (random_features(mesh.num_nodes, ci, seed=2, zero_frac=0.05, device=DEV) * scale).requires_grad_(True)
    gy = random_features(mesh.num_nodes, co, seed=3, zero_frac=0, device=DEV)
    y = ops.field_conv(x, m.weight(), plan, B, flags, keep_contrib=False)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    return m, x, gy, y


@pytest.mark.parametrize("n_side,ci,co,B,R,ftype", [(71, 32, 32, 1, 6, 1), (30, 64, 64, 1, 6, 0), (24, 32, 16, 1, 2, 2),
                                                     (20, 128, 128, 1, 6, 1), (26, 64, 48, 1, 3, 1), (17, 32, 2, 1, 2, 1)])
def test_fused_forward_vs_fp64_oracle(n_side, ci, co, B, R, ftype):
    assert _lib.fused_supported(ci, co, B, R)
    mesh = torus_mesh(n_side, deg=40.0, seed=1, device=DEV)
    before = _lib.launch_count()
    _lib.profile_enable(256)
    m, x, gy, y = _run(mesh, ci, co, B, R, FUSED, ftype)
    torch.cuda.synchronize()
    names = [n for n, _ in _lib.profile_collect(256)]
    assert "fused_fwd" in names and not any(n.startswith("aggregate") and not n.startswith("aggregate_T") for n in names), names
    assert _lib.launch_count() > before
    y_ref, gx_ref, gp = _oracle_layer(mesh, x, m, gy)
    assert_close_normwise(y, y_ref.to(torch.complex64), TOL, "y")
    assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, "grad x")
    assert_close_normwise(m.zonal.grad, gp[0].float(), TOL, "grad zonal")
    assert_close_normwise(m.spherical.grad, gp[1].float(), TOL, "grad spherical")


def test_fused_matches_unfused_kernels_and_is_deterministic():
    mesh = torus_mesh(40, deg=40.0, seed=3, device=DEV)
    outs = [_run(mesh, 32, 32, 1, 6, f)[3].detach() for f in (FUSED, FUSED, _lib.GEMM_TC_2XF16, _lib.GEMM_TC_2XF16 | _lib.FLAG_PACKED)]
    assert torch.equal(outs[0], outs[1])                    # fixed reduction order: bit-identical run to run
    assert_close_normwise(outs[0], outs[2], 3e-6, "fused vs fp32-operand 2xFP16")
    assert_close_normwise(outs[0], outs[3], 3e-6, "fused vs packed-operand 2xFP16")


def test_fused_edge_cases_and_scales():
    """Isolated targets (y = 0 rows), duplicate / dropped / shuffled edges, a row count that is not a multiple of the
    62-row tile, and feature scales far from 1 (the operand scale is a power of two from an a-priori bound)."""
    base = torus_mesh(23, deg=30.0, seed=4, device=DEV)
    mesh = with_edge_cases(base, isolated=(0, 7, 528), duplicate=50, shuffle_seed=9, far=40)
    for scale in (1.0, 1e-12, 1e9):
        m, x, gy, y = _run(mesh, 32, 32, 1, 6, FUSED, scale=scale)
        y_ref, gx_ref, _ = _oracle_layer(mesh, x, m, gy)
        assert_close_normwise(y, y_ref.to(torch.complex64), TOL, "y (scale %g)" % scale)
        assert_close_normwise(x.grad, gx_ref.to(torch.complex64), TOL, "grad x (scale %g)" % scale)
        assert float(y[[0, 7, 528]].abs().max()) == 0.0


def test_fused_module_policy(monkeypatch):
    """FIELDCONV_B200_FUSED=1 routes supported layers of a module through the fused kernel, others stay unfused."""
    from fieldconv_b200 import nn as fnn
    monkeypatch.setattr(fnn, "FUSED_POLICY", "1")
    mesh = torus_mesh(25, deg=40.0, seed=6, device=DEV)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, 6, mesh.epsilon)
    for (c, b, want) in ((32, 1, True), (48, 2, False)):
        torch.manual_seed(0)
        layer = fcb.FieldConv(c, c, b, 6, 1).to(DEV)
        x = random_features(mesh.num_nodes, c, seed=1, device=DEV)
        _lib.profile_enable(256)
        layer(x, plan)
        torch.cuda.synchronize()
        names = [n for n, _ in _lib.profile_collect(256)]
        assert ("fused_fwd" in names) == want, names
