"""GPU: TransField / LiftBlock (csrc/lift.cu + fieldconv_b200/lift.py) against outputs and autograd gradients of the
unmodified reference (tests/golden/lift_*.npz, nn/trans_field.py:78-113) and against the fp64 oracle restatement."""
import pytest
import torch

import fieldconv_b200 as fcb
from conftest import assert_close_normwise, golden_names, load_golden
from fieldconv_b200.synthetic import torus_mesh
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


@pytest.mark.parametrize("name", golden_names("lift_"))
def test_trans_field_matches_reference_golden(name):
    g = load_golden(name)
    m = fcb.TransField(g["ci"], g["co"], g["R"], g["ftype"])
    m.load_state_dict({"zonalAng": g["zonalAng"], "zonalMag": g["zonalMag"], "phase": g["phase"]})
    m = m.to(DEV)
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x, g["supp_edges"].to(DEV), g["lift_sten"].to(DEV))
    gy = g["gy"].to(DEV)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    assert_close_normwise(y, g["y"], TOL, "y")
    assert_close_normwise(x.grad, g["gx"], TOL, "grad x")
    assert_close_normwise(m.zonalAng.grad, g["g_zonalAng"], TOL, "grad zonalAng")
    assert_close_normwise(m.zonalMag.grad, g["g_zonalMag"], TOL, "grad zonalMag")
    if g["ftype"] == 1:
        assert_close_normwise(m.phase.grad, g["g_phase"], TOL, "grad phase")


def test_lift_block_on_a_mesh_vs_fp64_oracle_and_determinism():
    """FCPrecomp (device) -> lift stencil supp_sten[..., B:B+2] -> LiftBlock, the call sequence of the reference nets
    (segmentation.ipynb:202-206), against the fp64 restatement; two runs are bit-identical (no atomics)."""
    import types
    mesh = torus_mesh(40, deg=40.0, seed=2, device=DEV)
    B, R, ci, co = 2, 6, 3, 16
    pre = fcb.FCPrecomp(B, R, mesh.epsilon)
    edges, sten, _, _ = pre(types.SimpleNamespace(**vars(mesh)))
    lift = sten[..., B:B + 2].contiguous()
    torch.manual_seed(0)
    blk = fcb.LiftBlock(ci, co, R, 1).to(DEV)
    with torch.no_grad():
        blk.nonlin.bias.uniform_(-0.2, 0.2)
    x = torch.randn(mesh.num_nodes, ci, generator=torch.Generator().manual_seed(3)).to(DEV).requires_grad_(True)
    y = blk(x, edges, lift)
    gy = torch.randn(mesh.num_nodes, co, 2, generator=torch.Generator().manual_seed(4))
    gy = torch.view_as_complex(gy).to(DEV)
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    y2 = blk(x.detach(), edges, lift)
    assert torch.equal(y.detach(), y2)
    # fp64 oracle of the same composition
    f = blk.field
    ps = [p.detach().cpu().double().requires_grad_(True) for p in (f.zonalAng, f.zonalMag, f.phase)]
    xd = x.detach().cpu().double().requires_grad_(True)
    t = restate.trans_field_lean(xd, edges.cpu(), lift.cpu().to(torch.complex128), ps[0], ps[1], ps[2], 1)
    yr = restate.tangent_nonlin(t, blk.nonlin.bias.detach().cpu().double())
    gyd = gy.cpu().to(torch.complex128)
    (yr.real * gyd.real + yr.imag * gyd.imag).sum().backward()
    # L2 at the fp32 budget; the max-norm gets 5e-5: y sums only Ci = 3 terms m * a/|a|, and a/|a| is ill-conditioned where
    # |a| is small, so single entries of ANY fp32 evaluation (the reference's included) sit a few 1e-6..1e-5 of max|y| off fp64
    from conftest import rel_l2, rel_max
    yr32 = yr.detach().to(torch.complex64)
    assert rel_l2(y.detach().cpu(), yr32) <= TOL and rel_max(y.detach().cpu(), yr32) <= 5e-5
    # grad x is a gradient of per-edge differences x_j - x_i (out-edge and in-edge sums nearly cancel): the reference's own
    # fp32 evaluation differs from its fp64 one by 4-6e-6 normwise on this mesh (measured with oracle/restate.py), so this
    # one quantity is held to 2e-5 against fp64; against the reference's fp32 goldens above it meets 1e-5
    assert_close_normwise(x.grad, xd.grad.float(), 2e-5, "grad x")
    assert_close_normwise(f.zonalAng.grad, ps[0].grad.float(), TOL, "grad zonalAng")
    assert_close_normwise(f.zonalMag.grad, ps[1].grad.float(), TOL, "grad zonalMag")
    assert_close_normwise(f.phase.grad, ps[2].grad.float(), TOL, "grad phase")
