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
    # grad x flows through d(a/|a|)/da ~ 1/|a|: among the N*Co*Ci = 77k values of a the smallest are ~1/300 of the typical
    # modulus, so their gradient entries are amplified 300x, dominate the norm and carry 300x the fp32 rounding of a.  Any fp32
    # evaluation (the reference's own differs from its fp64 run by 4e-6..6e-6 on this mesh on the CPU, measured with
    # oracle/restate.py) sits around 1e-5 of fp64 here; the goldens above pin grad x to the reference's fp32 result at 1e-5.
    assert_close_normwise(x.grad, xd.grad.float(), 1e-4, "grad x")
    # parameter gradients against fp64 on this mesh: held to 1e-4 (measured 5e-6 .. 3e-5 run to run, the spread of the
    # 1/|a|-amplified entries above); the reference's fp32 goldens pin all three to 1e-5 in the test above
    assert_close_normwise(f.zonalAng.grad, ps[0].grad.float(), 1e-4, "grad zonalAng")
    assert_close_normwise(f.zonalMag.grad, ps[1].grad.float(), 1e-4, "grad zonalMag")
    assert_close_normwise(f.phase.grad, ps[2].grad.float(), 1e-4, "grad phase")
