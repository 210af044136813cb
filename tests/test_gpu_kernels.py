"""GPU: building blocks of the C ABI (sort, plan, aggregation, GEMM, modReLU) against the oracle."""
import pytest
import torch

import fieldconv_b200 as fcb
from conftest import assert_close_normwise, golden_names, load_golden
from fieldconv_b200 import _lib, ops
from fieldconv_b200.synthetic import random_features, torus_mesh
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sort(keys, vals, bits):
    n = keys.numel()
    k_in, v_in = keys.clone().int(), vals.clone().int()
    k_out, v_out = torch.empty_like(k_in), torch.empty_like(v_in)
    nbytes = _lib.query_bytes("fcb_sort_workspace_bytes", n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.call("fcb_sort_pairs_u32", k_in.data_ptr(), v_in.data_ptr(), k_out.data_ptr(), v_out.data_ptr(), n, bits,
              ws.data_ptr(), nbytes, _lib.stream_ptr())
    return k_out, v_out


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 8), (31, 3), (2048, 8), (2049, 9), (100003, 17), (1 << 20, 24), (300007, 31)])
def test_radix_sort_is_stable_and_exact(n, bits):
    g = torch.Generator(device="cpu").manual_seed(n + bits)
    keys = torch.randint(0, 2 ** bits, (n,), generator=g, dtype=torch.int64).to(DEV)
    if n > 10:
        keys[: n // 3] = keys[n // 2]        # many duplicates -> stability matters
    vals = torch.arange(n, device=DEV)
    k_out, v_out = _sort(keys, vals, bits)
    ref_k, ref_i = torch.sort(keys, stable=True)
    assert torch.equal(k_out.long(), ref_k)
    assert torch.equal(v_out.long(), ref_i)


def _plan_from_golden(g):
    return fcb.build_plan(g["raw_edges"].to(DEV), g["logMag"].to(DEV), g["logAng"].to(DEV), g["xp"].to(DEV),
                          g["w"].to(DEV), g["R"], g["epsilon"])


@pytest.mark.parametrize("name", golden_names("fc_") + ["block_b2r6"])
def test_plan_matches_reference_fcprecomp(name):
    g = load_golden(name)
    R, B = g["R"], g["B"]
    plan = _plan_from_golden(g)
    e_ref = g["supp_edges"]                                     # reference FCPrecomp output (kept edges, input order)
    E = e_ref.shape[0]
    assert plan.num_edges == E
    # oracle pieces (bit-exact restatement of the reference, see test_oracle.py)
    _, sten, _, wxp, (keep, f, t) = restate.fc_precomp(g["logMag"], g["logAng"], g["w"], g["raw_edges"], g["xp"], B, R,
                                                       g["epsilon"])
    for side, col in (("tgt", 1), ("src", 0)):
        perm = getattr(plan, "perm_" + side)[:E].long().cpu()
        rec = getattr(plan, "rec_" + side)[:E].cpu()
        rot = getattr(plan, "rot_" + side)[:E].cpu()
        rowptr = getattr(plan, "rowptr_" + side).long().cpu()
        # CSR order == stable sort of the kept edges by (vertex, ring floor): indices bit-exact
        key = e_ref[:, col] * (R - 1) + f
        order = torch.argsort(key, stable=True)
        assert torch.equal(perm, keep[order]), side + ": permutation"
        counts = torch.bincount(e_ref[:, col], minlength=g["n"])
        assert torch.equal(rowptr[1:] - rowptr[:-1], counts), side + ": row pointers"
        assert int(rowptr[0]) == 0 and int(rowptr[-1]) == E
        nbr = (rec[:, 0] & ((1 << 27) - 1)).long()
        ring = (rec[:, 0] >> 27) & 31
        assert torch.equal(nbr, e_ref[order, 1 - col]), side + ": neighbour ids"
        assert torch.equal(ring.long(), f[order]), side + ": ring floor"
        assert torch.equal(rec[:, 1].view(torch.float32), t[order]), side + ": ring weight t bit-exact"
        got_wxp = torch.view_as_complex(rec[:, 2:4].contiguous().view(torch.float32))
        assert_close_normwise(got_wxp, wxp[order], 1e-6, side + ": wxp")
        theta = g["logAng"][keep][order]
        assert_close_normwise(rot, torch.stack((torch.cos(theta), torch.sin(theta)), 1), 1e-6, side + ": rot")
    # expanded multiset of (j, i) equals the reference's edge list
    got = plan.edges_by_target().cpu()
    assert torch.equal(got[torch.argsort(got[:, 0] * g["n"] + got[:, 1], stable=True)],
                       e_ref[torch.argsort(e_ref[:, 0] * g["n"] + e_ref[:, 1], stable=True)])


def test_plan_empty_and_all_filtered():
    w = torch.ones(7, 1, device=DEV)
    p = fcb.build_plan(torch.zeros(0, 2, dtype=torch.long, device=DEV), torch.zeros(0, device=DEV), torch.zeros(0, device=DEV),
                       torch.zeros(0, dtype=torch.complex64, device=DEV), w, 6, 1.0)
    assert p.num_edges == 0 and int(p.rowptr_src[-1]) == 0
    e = torch.tensor([[0, 1], [2, 3], [6, 6]], device=DEV)
    p = fcb.build_plan(e, torch.full((3,), 5.0, device=DEV), torch.zeros(3, device=DEV),
                       torch.ones(3, dtype=torch.complex64, device=DEV), w, 6, 1.0)
    assert p.num_edges == 0
    m = fcb.FieldConv(4, 6).to(DEV)
    y = m(torch.randn(7, 4, dtype=torch.complex64, device=DEV), p)
    assert y.shape == (7, 6) and float(y.detach().abs().max()) == 0.0     # no in-edges -> y = 0 (field_conv.py:134 dim_size=N)


def test_plan_validate_rejects_out_of_range_endpoints():
    """The reference raises an index error on an endpoint outside [0, N) (nn/field_conv.py:130-134); build_plan(validate=True)
    does the same instead of dropping the edge."""
    w = torch.ones(7, 1, device=DEV)
    e = torch.tensor([[0, 1], [2, 7], [6, 6]], device=DEV)
    args = (torch.full((3,), 0.3, device=DEV), torch.zeros(3, device=DEV), torch.ones(3, dtype=torch.complex64, device=DEV), w, 6, 1.0)
    with pytest.raises(IndexError):
        fcb.build_plan(e, *args, validate=True)
    assert fcb.build_plan(e, *args).num_edges == 2            # default: the edge is dropped (documented)
    ok = [a[[0, 2]] if (torch.is_tensor(a) and a.ndim == 1 and a.shape[0] == 3) else a for a in args]
    assert fcb.build_plan(e[[0, 2]], *ok, validate=True).num_edges == 2


@pytest.mark.parametrize("m,n,k,trans", [(1, 4, 4, 0), (130, 96, 72, 0), (257, 20, 1000, 0), (1000, 64, 36, 0),
                                         (300, 96, 4000, 1), (2880, 96, 20000, 1), (64, 256, 515, 1), (5, 12, 7, 1)])
def test_gemm_matches_fp64(m, n, k, trans):
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n)
    k_pad = (k + 3) // 4 * 4 if not trans else k
    m_pad = (m + 3) // 4 * 4 if trans else m
    a = torch.randn((k, m_pad) if trans else (m, k_pad), generator=g).to(DEV)
    if trans:
        a[:, m:] = 0
    else:
        a[:, k:] = 0
    b = torch.randn(k_pad if not trans else k, n, generator=g).to(DEV)
    c = ops.gemm(a, b, bool(trans))
    ref = (a.double().t() if trans else a.double()) @ b.double()
    assert_close_normwise(c, ref.float(), 2e-6, "gemm")
    c2 = ops.gemm(a, b, bool(trans))
    assert torch.equal(c, c2), "split-K reduction must be deterministic"


@pytest.mark.parametrize("name", golden_names("fc_"))
def test_aggregate_matches_oracle(name):
    g = load_golden(name)
    B, R, ci = g["B"], g["R"], g["ci"]
    plan = _plan_from_golden(g)
    cpad = ci + (ci % 2)
    x = torch.zeros(g["n"], cpad, dtype=torch.complex64)
    x[:, :ci] = g["x"]
    xd = x.to(DEV)
    M = 2 * B + 1
    out = torch.empty(g["n"], R * cpad * M, dtype=torch.complex64, device=DEV)
    _lib.call("fcb_aggregate_f32", torch.view_as_real(xd).data_ptr(), plan.rowptr_tgt.data_ptr(), plan.rec_tgt.data_ptr(),
              plan.rot_tgt.data_ptr(), torch.view_as_real(out).data_ptr(), g["n"], cpad, B, R, 0, _lib.stream_ptr())
    ref = restate.aggregate(x, g["supp_edges"], g["supp_sten"], B)            # (N, C, R, M)
    got = out.cpu().reshape(g["n"], R, M, cpad).permute(0, 3, 1, 2)
    assert_close_normwise(got, ref, 3e-6, "contrib")
    # transposed gather: G[j,m,r,o] = sum_{e: src=j} conj(sten[e,r,m]) gy[tgt(e), o]
    gy = xd                                                                   # any complex field will do
    outT = torch.empty_like(out)
    _lib.call("fcb_aggregate_f32", torch.view_as_real(gy).data_ptr(), plan.rowptr_src.data_ptr(), plan.rec_src.data_ptr(),
              plan.rot_src.data_ptr(), torch.view_as_real(outT).data_ptr(), g["n"], cpad, B, R, 1, _lib.stream_ptr())
    e, sten = g["supp_edges"], g["supp_sten"]
    per_edge = x[e[:, 1]][:, None, None, :] * sten.conj().permute(0, 2, 1)[..., None]   # (E, M, R, C)
    refT = torch.zeros(g["n"], M, R, cpad, dtype=torch.complex64).index_add(0, e[:, 0], per_edge)
    assert_close_normwise(outT.cpu().reshape(g["n"], M, R, cpad), refT, 3e-6, "transposed gather")


def test_modrelu_matches_oracle():
    g = torch.Generator(device="cpu").manual_seed(3)
    for n, c in ((1000, 48), (77, 5), (300, 300)):
        x = random_features(n, c, seed=n, zero_frac=0.05)
        bias = (torch.rand(1, c, generator=g) - 0.5)
        gy = torch.complex(torch.randn(n, c, generator=g), torch.randn(n, c, generator=g))
        xr = x.clone().requires_grad_(True)
        br = bias.clone().requires_grad_(True)
        y_ref = restate.tangent_nonlin(xr, br)
        (y_ref.real * gy.real + y_ref.imag * gy.imag).sum().backward()
        xd = x.to(DEV).requires_grad_(True)
        bd = bias.to(DEV).requires_grad_(True)
        y = ops.modrelu(xd, bd)
        (y.real * gy.to(DEV).real + y.imag * gy.to(DEV).imag).sum().backward()
        assert_close_normwise(y, y_ref, 2e-6, "modrelu y")
        assert_close_normwise(xd.grad, xr.grad, 2e-6, "modrelu gx")
        assert_close_normwise(bd.grad, br.grad, 5e-6, "modrelu gb")


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-6), ("auto", 5e-6)])
def test_tangent_lin_matches_oracle(precision, tol):
    for ci, co in ((48, 48), (5, 7), (6, 3), (64, 128)):
        lin = fcb.TangentLin(ci, co, precision=precision)
        x = random_features(3000, ci, seed=ci)
        xr = x.clone().requires_grad_(True)
        y_ref = restate.tangent_lin(xr, lin.Re, lin.Im)
        gy = random_features(3000, co, seed=co + 1, zero_frac=0)
        (y_ref.real * gy.real + y_ref.imag * gy.imag).sum().backward()
        g_re, g_im = lin.Re.grad.clone(), lin.Im.grad.clone()
        lin.zero_grad()
        lin_d = lin.to(DEV)
        xd = x.to(DEV).requires_grad_(True)
        y = lin_d(xd)
        (y.real * gy.to(DEV).real + y.imag * gy.to(DEV).imag).sum().backward()
        assert_close_normwise(y, y_ref, tol, "TangentLin y")
        assert_close_normwise(xd.grad, xr.grad, tol, "TangentLin gx")
        assert_close_normwise(lin_d.Re.grad, g_re, tol, "TangentLin gRe")
        assert_close_normwise(lin_d.Im.grad, g_im, tol, "TangentLin gIm")


@pytest.mark.parametrize("mode,tol", [(1, 1e-5), (2, 3e-3), (3, 5e-6)])
@pytest.mark.parametrize("m,n,k", [(128, 96, 32), (128, 96, 2880), (1000, 96, 576), (80656, 96, 2880), (300, 64, 1152),
                                   (257, 256, 520), (130, 20, 36), (5, 12, 8), (4096, 16, 64), (640, 128, 7680),
                                   (200, 96, 100)])
def test_tensor_core_gemm(m, n, k, mode, tol):
    """tcgen05 GEMM: 3xTF32 (mode 1) and 2xFP16 (mode 3) must sit at fp32-grade accuracy, plain TF32 (mode 2) at ~1e-3."""
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    k_pad = (k + 3) // 4 * 4
    a = torch.randn(m, k_pad, generator=g).to(DEV)
    a[:, k:] = 0
    b = torch.randn(k_pad, n, generator=g).to(DEV)
    c = ops.gemm(a, b, False, mode)
    torch.cuda.synchronize()
    ref = a.double() @ b.double()
    assert_close_normwise(c, ref.float(), tol, "tensor-core gemm mode %d" % mode)
    assert torch.equal(c, ops.gemm(a, b, False, mode)), "deterministic"
    if mode == 2:   # TF32 must actually be less accurate than fp32 (i.e. the tensor path really ran)
        err = float((c.double() - ref).abs().max() / ref.abs().max())
        assert err > 1e-6 or k <= 8


@pytest.mark.parametrize("mode,tol", [(1, 1e-5), (2, 3e-3), (3, 5e-6)])
@pytest.mark.parametrize("m,n,k", [(128, 96, 32), (2880, 96, 80656), (1152, 64, 5041), (300, 256, 1000), (36, 20, 77),
                                   (7680, 256, 6889), (128, 16, 8), (256, 96, 64), (100, 72, 200)])
def test_tensor_core_gemm_transposed(m, n, k, mode, tol):
    """gW-shaped product P[m x n] = A^T B, A = (k x m), B = (k x n): MN-major UMMA operands + split reduction."""
    g = torch.Generator(device="cpu").manual_seed(m + n + k + 1)
    m_pad = (m + 3) // 4 * 4
    a = torch.randn(k, m_pad, generator=g).to(DEV)
    a[:, m:] = 0
    b = torch.randn(k, n, generator=g).to(DEV)
    c = ops.gemm(a, b, True, mode)[:m]
    torch.cuda.synchronize()
    ref = (a.double().t() @ b.double())[:m]
    assert_close_normwise(c, ref.float(), tol, "tensor-core gemm^T mode %d" % mode)
    assert torch.equal(c, ops.gemm(a, b, True, mode)[:m]), "deterministic"


@pytest.mark.parametrize("sa,sb", [(1.0, 1.0), (3e-21, 7e14), (2e18, 5e-9), (1e-15, 1e-15)])
@pytest.mark.parametrize("trans", [False, True])
def test_fp16_pair_gemm_operand_scaling(sa, sb, trans):
    """2xFP16 mode: the power-of-two operand scales (from max|A|, max|B|) make the fp16 exponent range a non-issue —
    the result is fp32-grade whatever the magnitudes, and rows 2^12 below the largest keep ~fp32 relative accuracy."""
    g = torch.Generator(device="cpu").manual_seed(5)
    m, n, k = 384, 96, 1152
    a = torch.randn(m, k, generator=g)
    a[64:128] *= 2.0 ** -12          # a block of small rows
    b = torch.randn(k, n, generator=g)
    a, b = (a * sa).to(DEV), (b * sb).to(DEV)
    if trans:
        c = ops.gemm(a.t().contiguous(), b, True, 3)
    else:
        c = ops.gemm(a, b, False, 3)
    ref = (a.double() @ b.double())
    assert torch.isfinite(c).all()
    assert_close_normwise(c, ref.float(), 5e-6, "2xFP16 gemm scaled")
    small = slice(64, 128)
    err = float((c[small].double() - ref[small]).norm() / ref[small].norm())
    assert err < 2e-4, err           # absolute error <= 2^-25 of the scaled range: still ~1e-5 relative 2^12 down


def test_fp16_pair_gemm_zero_operand():
    a = torch.zeros(256, 128, device=DEV)
    b = torch.randn(128, 64, device=DEV)
    assert torch.count_nonzero(ops.gemm(a, b, False, 3)) == 0
    assert torch.count_nonzero(ops.gemm(b, torch.zeros(128, 32, device=DEV), True, 3)) == 0
