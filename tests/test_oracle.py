"""CPU: the oracle restatement against golden vectors produced by the unmodified reference."""
import pytest
import torch

from conftest import assert_close_normwise, golden_names, load_golden
from oracle import restate

CASES = golden_names("fc_")


def test_golden_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_fc_precomp_matches_reference(name):
    g = load_golden(name)
    e, sten, ln, wxp, (keep, f, t) = restate.fc_precomp(g["logMag"], g["logAng"], g["w"], g["raw_edges"], g["xp"],
                                                        g["B"], g["R"], g["epsilon"])
    assert torch.equal(e, g["supp_edges"])                      # indices bit-exact
    assert torch.equal(sten, g["supp_sten"])                    # same float ops -> bit-exact on the same CPU
    assert torch.equal(ln, g["ln"]) and torch.equal(wxp, g["wxp"])
    # stencil structure: <=2 non-zero rings per edge, ring weights sum to 1 (fc_precomp.py:24-25)
    nz = (sten.abs().sum(-1) > 0).sum(1)
    assert int(nz.max()) <= 2
    assert_close_normwise(sten[:, :, g["B"]].sum(1), wxp, 1e-6, "sum_r sten[:, r, m=0] == wxp")


@pytest.mark.parametrize("name", CASES)
def test_forward_and_autograd_match_reference(name):
    g = load_golden(name)
    B, ftype = g["B"], g["ftype"]
    x = g["x"].clone().requires_grad_(True)
    ps = [g[k].clone().requires_grad_(True) for k in ("zonal", "spherical", "phase")]
    y = restate.field_conv_refstyle(x, g["supp_edges"], g["supp_sten"], ps[0], ps[1], ps[2], ftype, B)
    (y.real * g["gy"].real + y.imag * g["gy"].imag).sum().backward()
    assert_close_normwise(y, g["y"], 1e-6, "y (refstyle)")
    assert_close_normwise(x.grad, g["gx"], 1e-6, "gx (refstyle)")
    assert_close_normwise(ps[0].grad, g["g_zonal"], 1e-6, "g_zonal")
    assert_close_normwise(ps[1].grad, g["g_spherical"], 1e-6, "g_spherical")
    if ftype == 1:
        assert_close_normwise(ps[2].grad, g["g_phase"], 1e-6, "g_phase")


@pytest.mark.parametrize("name", CASES)
def test_folded_form_and_closed_form_backward(name):
    g = load_golden(name)
    B, ftype = g["B"], g["ftype"]
    ps = [g[k].clone().requires_grad_(True) for k in ("zonal", "spherical", "phase")]
    W = restate.fold_weights(ps[0], ps[1], ps[2], ftype, B)
    y = restate.field_conv_lean(g["x"], g["supp_edges"], g["supp_sten"], W.detach(), B)
    assert_close_normwise(y, g["y"], 2e-6, "y (folded)")
    gx, gw = restate.field_conv_backward(g["x"], g["supp_edges"], g["supp_sten"], W.detach(), B, g["gy"])
    assert_close_normwise(gx, g["gx"], 2e-6, "gx (closed form)")
    (W.real * gw.real + W.imag * gw.imag).sum().backward()       # chain gW to the parameters
    assert_close_normwise(ps[0].grad, g["g_zonal"], 2e-6, "g_zonal via gW")
    assert_close_normwise(ps[1].grad, g["g_spherical"], 2e-6, "g_spherical via gW")
    if ftype == 1:
        assert_close_normwise(ps[2].grad, g["g_phase"], 2e-6, "g_phase via gW")


def test_folded_form_fp64_agrees():
    g = load_golden(CASES[0])
    B = g["B"]
    W = restate.fold_weights(g["zonal"].double(), g["spherical"].double(), g["phase"].double(), g["ftype"], B)
    y = restate.field_conv_lean(g["x"].to(torch.complex128), g["supp_edges"], g["supp_sten"].to(torch.complex128), W, B)
    assert_close_normwise(y.to(torch.complex64), g["y"], 2e-6, "fp64 folded vs fp32 reference")


def test_block_matches_reference():
    g = load_golden("block_b2r6")
    p = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    x = g["x"].clone().requires_grad_(True)
    y = restate.fc_resnet_block(x, g["supp_edges"], g["supp_sten"], p, g["B"], 1)
    (y.real * g["gy"].real + y.imag * g["gy"].imag).sum().backward()
    assert_close_normwise(y, g["y"], 2e-6, "block y")
    assert_close_normwise(x.grad, g["gx"], 2e-6, "block gx")


def test_gauge_equivariance_property():
    """Rotating every vertex frame by alpha_v rotates the response by the same angle (the paper's central
    claim).  Known-answer property independent of the reference's code."""
    g = load_golden("fc_b2r6_f0")
    B, R = g["B"], g["R"]
    n = g["n"]
    gen = torch.Generator().manual_seed(5)
    alpha = (torch.rand(n, generator=gen) * 2 - 1) * 3.0
    e_raw = g["raw_edges"]
    src, tgt = e_raw[:, 0], e_raw[:, 1]
    W = restate.fold_weights(g["zonal"], g["spherical"], g["phase"], g["ftype"], B)

    def run(x, log_ang, xp):
        e, sten, _, _, _ = restate.fc_precomp(g["logMag"], log_ang, g["w"], e_raw, xp, B, R, g["epsilon"])
        return restate.field_conv_lean(x, e, sten, W, B)

    y0 = run(g["x"], g["logAng"], g["xp"])
    rot = torch.polar(torch.ones(n), -alpha)
    x1 = g["x"] * rot[:, None]
    la1 = g["logAng"] - alpha[src]          # applied to self loops too (theta multiplies ring 0 for m != 0)
    xp1 = g["xp"] * torch.polar(torch.ones_like(alpha[src]), alpha[src] - alpha[tgt])
    y1 = run(x1.to(torch.complex64), la1, xp1.to(torch.complex64))
    assert_close_normwise(y1, y0 * rot[:, None], 5e-6, "gauge equivariance")


# ---- F2 groundwork: TransField restatements against the unmodified reference's outputs
@pytest.mark.parametrize("name", golden_names("lift_"))
def test_trans_field_restatements_match_reference(name):
    g = load_golden(name)
    assert torch.equal(restate.lift_stencil(g["supp_sten"], g["B"]), g["lift_sten"])
    for fn, tol in ((restate.trans_field_refstyle, 1e-6), (restate.trans_field_lean, 2e-6)):
        x = g["x"].clone().requires_grad_(True)
        ps = [g[k].clone().requires_grad_(True) for k in ("zonalAng", "zonalMag", "phase")]
        y = fn(x, g["supp_edges"], g["lift_sten"], ps[0], ps[1], ps[2], g["ftype"])
        (y.real * g["gy"].real + y.imag * g["gy"].imag).sum().backward()
        assert_close_normwise(y, g["y"], tol, fn.__name__ + " y")
        assert_close_normwise(x.grad, g["gx"], 10 * tol, fn.__name__ + " gx")
        assert_close_normwise(ps[0].grad, g["g_zonalAng"], 10 * tol, fn.__name__ + " g_zonalAng")
        assert_close_normwise(ps[1].grad, g["g_zonalMag"], 10 * tol, fn.__name__ + " g_zonalMag")
        if g["ftype"] == 1:
            assert_close_normwise(ps[2].grad, g["g_phase"], 10 * tol, fn.__name__ + " g_phase")
    assert float(g["y"][5].abs().max()) == 0.0            # the isolated target


@pytest.mark.parametrize("name", golden_names("echo_"))
def test_echo_restatement_matches_reference(name):
    """oracle/restate.py echo_refstyle (nn/echo.py:94-148) against outputs and autograd gradients of the unmodified reference."""
    g = load_golden(name)
    x = g["x"].clone().requires_grad_(True)
    y = restate.echo_refstyle(x, g["supp_edges"], g["ln"], g["wxp"], g["n_bins"])
    (y * g["gy"]).sum().backward()
    d_map, dim = restate.echo_disk_map(g["n_bins"])
    assert dim == g["hdim"] and torch.equal(d_map, g["dMap"])
    assert float((y.detach() - g["y"]).abs().max()) <= 2e-6 * float(g["y"].abs().max())
    assert float((x.grad - g["gx"]).abs().max()) <= 2e-6 * float(g["gx"].abs().max())


def test_network_restatement_matches_reference():
    """oracle/restate.py notebook_net — LiftBlock, FCResNetBlocks (one frontloaded), TangentPerceptron residuals, ECHOBlock —
    against logits, loss and every parameter gradient of the unmodified reference (net_b2r6).  Gradients are held to
    max(1e-5, 10 x the reference's own summation-order noise for that parameter), see oracle/make_golden.py make_net."""
    g = load_golden("net_b2r6")
    p = {k[2:]: v.clone().requires_grad_(v.is_floating_point()) for k, v in g.items() if k.startswith("p.")}
    logits = restate.notebook_net(g["pos"], g["supp_edges"], g["supp_sten"], g["ln"], g["wxp"], p, g["B"], g["n_bins"], g["ftype"])
    loss = torch.nn.functional.cross_entropy(logits, g["labels"])
    loss.backward()
    assert_close_normwise(logits, g["logits"], 1e-5, "logits")
    assert abs(loss.item() - g["loss"]) <= 1e-5 * abs(g["loss"])
    for k in (k[2:] for k in g if k.startswith("g.")):
        assert_close_normwise(p[k].grad, g["g." + k], max(1e-5, 10.0 * g["noise." + k]), "grad " + k)
