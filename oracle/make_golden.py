"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Each file holds the inputs of one seeded case and what the reference produced for them:
FCPrecomp outputs (transforms/fc_precomp.py:53-97), FieldConv forward output and the
autograd gradients of x / zonal / spherical / phase (nn/field_conv.py:104-137), and for the
block case FCResNetBlock (nn/fc_resnet_block.py:65-88).  The loss is Re<y, gy> with a seeded
gy so that the x-gradient equals the vector-Jacobian product for upstream gradient gy.
"""
import importlib.util
import os

import numpy as np
import torch

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("_syn", os.path.join(ROOT, "fieldconv_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(syn)

CASES = [
    # name, n_side, deg, Ci, Co, B, R, ftype, eps_scale, tweaks
    ("fc_b1r6_f1", 9, 14.0, 6, 4, 1, 6, 1, 0.93, ()),
    ("fc_b2r6_f0", 8, 12.0, 4, 6, 2, 6, 0, 0.95, ()),
    ("fc_b2r4_f2", 8, 12.0, 5, 3, 2, 4, 2, 1.0, ("dups", "shuffle")),
    ("fc_b3r2_f1", 7, 10.0, 3, 5, 3, 2, 1, 0.9, ("isolated", "exact_eps")),
    ("fc_b1r3_f2", 7, 16.0, 8, 8, 1, 3, 2, 1.0, ("shuffle",)),
]


def _np(t):
    t = t.detach()
    return t.numpy()


def make_case(name, n_side, deg, ci, co, B, R, ftype, eps_scale, tweaks, seed):
    ns = ref_loader.load()
    d = syn.torus_mesh(n_side, deg=deg, seed=seed, tile=4)
    g = torch.Generator().manual_seed(seed)
    eps = float(d.epsilon * eps_scale)
    if "dups" in tweaks:                      # duplicate edges are summed (field_conv.py:134)
        pick = torch.randperm(d.supp_edges.shape[0], generator=g)[:7]
        for k in ("supp_edges", "logMag", "logAng", "xp"):
            setattr(d, k, torch.cat((getattr(d, k), getattr(d, k)[pick])))
    if "isolated" in tweaks:                  # a target with no incoming edge -> y = 0
        keep = d.supp_edges[:, 1] != 3
        for k in ("supp_edges", "logMag", "logAng", "xp"):
            setattr(d, k, getattr(d, k)[keep])
    if "exact_eps" in tweaks:                 # r/eps == 1 exactly stays in the support
        far = torch.nonzero(d.logMag > 0.5 * eps)[:2, 0]
        d.logMag[far] = torch.tensor(eps, dtype=torch.float32)
        eps = float(d.logMag[far[0]])
    if "shuffle" in tweaks:                   # any edge order is accepted
        perm = torch.randperm(d.supp_edges.shape[0], generator=g)
        for k in ("supp_edges", "logMag", "logAng", "xp"):
            setattr(d, k, getattr(d, k)[perm])
    e, sten, ln, wxp = ns.FCPrecomp(B, R, eps)(d)
    x = syn.random_features(d.num_nodes, ci, seed=seed, zero_frac=0.05)
    torch.manual_seed(seed)
    fc = ns.FieldConv(ci, co, B, R, ftype)
    xr = x.clone().requires_grad_(True)
    y = fc(xr, e, sten)
    gy = torch.complex(torch.randn(y.shape, generator=g), torch.randn(y.shape, generator=g))
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    out = dict(
        n=np.int64(d.num_nodes), ci=np.int64(ci), co=np.int64(co), B=np.int64(B), R=np.int64(R),
        ftype=np.int64(ftype), epsilon=np.float64(eps),
        raw_edges=_np(d.supp_edges), logMag=_np(d.logMag), logAng=_np(d.logAng), xp=_np(d.xp), w=_np(d.w),
        supp_edges=_np(e), supp_sten=_np(sten), ln=_np(ln), wxp=_np(wxp),
        x=_np(x), gy=_np(gy), zonal=_np(fc.zonal), spherical=_np(fc.spherical), phase=_np(fc.phase),
        y=_np(y), gx=_np(xr.grad), g_zonal=_np(fc.zonal.grad), g_spherical=_np(fc.spherical.grad),
    )
    if ftype == 1:
        out["g_phase"] = _np(fc.phase.grad)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "N", d.num_nodes, "E", e.shape[0], "|y|max", float(y.abs().max()))


def make_block(seed=11):
    ns = ref_loader.load()
    d = syn.torus_mesh(8, deg=12.0, seed=seed, tile=4)
    B, R, ci, co = 2, 6, 4, 6
    e, sten, ln, wxp = ns.FCPrecomp(B, R, float(d.epsilon))(d)
    x = syn.random_features(d.num_nodes, ci, seed=seed, zero_frac=0.05)
    torch.manual_seed(seed)
    blk = ns.FCResNetBlock(ci, co, B, R, 1)
    with torch.no_grad():
        blk.nonlin1.bias.uniform_(-0.3, 0.3)
        blk.nonlin2.bias.uniform_(-0.3, 0.3)
    xr = x.clone().requires_grad_(True)
    y = blk(xr, e, sten)
    g = torch.Generator().manual_seed(seed)
    gy = torch.complex(torch.randn(y.shape, generator=g), torch.randn(y.shape, generator=g))
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    out = dict(n=np.int64(d.num_nodes), ci=np.int64(ci), co=np.int64(co), B=np.int64(B), R=np.int64(R),
               epsilon=np.float64(d.epsilon), raw_edges=_np(d.supp_edges), logMag=_np(d.logMag),
               logAng=_np(d.logAng), xp=_np(d.xp), w=_np(d.w),
               supp_edges=_np(e), supp_sten=_np(sten), x=_np(x), gy=_np(gy), y=_np(y), gx=_np(xr.grad))
    for k, v in blk.state_dict().items():
        out["p." + k] = _np(v)
    for k, v in blk.named_parameters():
        out["g." + k] = _np(v.grad)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "block_b2r6.npz"), **out)
    print("block_b2r6", "N", d.num_nodes, "E", e.shape[0])


LIFT_CASES = [
    # name, n_side, deg, Ci, Co, B (of the FCPrecomp stencil the lift stencil is cut from), R, ftype
    ("lift_b1r6_f1", 8, 12.0, 3, 5, 1, 6, 1),
    ("lift_b2r4_f0", 7, 14.0, 4, 6, 2, 4, 0),
]


def make_lift(name, n_side, deg, ci, co, B, R, ftype, seed):
    """TransField (nn/trans_field.py:78-113) on the lift stencil supp_sten[..., B:B+2] — SURVEY.md §8(f) F2 groundwork."""
    ns = ref_loader.load()
    d = syn.torus_mesh(n_side, deg=deg, seed=seed, tile=4)
    g = torch.Generator().manual_seed(seed)
    keep = d.supp_edges[:, 1] != 5                      # an isolated target
    for k in ("supp_edges", "logMag", "logAng", "xp"):
        setattr(d, k, getattr(d, k)[keep])
    e, sten, _, _ = ns.FCPrecomp(B, R, float(d.epsilon))(d)
    lift = sten[..., B:B + 2].contiguous()
    x = torch.randn(d.num_nodes, ci, generator=g)
    x[torch.rand(x.shape, generator=g) < 0.05] = 0.0
    torch.manual_seed(seed)
    tf = ns.TransField(ci, co, n_rings=R, ftype=ftype)
    xr = x.clone().requires_grad_(True)
    y = tf(xr, e, lift)
    gy = torch.complex(torch.randn(y.shape, generator=g), torch.randn(y.shape, generator=g))
    (y.real * gy.real + y.imag * gy.imag).sum().backward()
    out = dict(n=np.int64(d.num_nodes), ci=np.int64(ci), co=np.int64(co), B=np.int64(B), R=np.int64(R), ftype=np.int64(ftype),
               epsilon=np.float64(d.epsilon), supp_edges=_np(e), supp_sten=_np(sten), lift_sten=_np(lift), x=_np(x), gy=_np(gy),
               zonalAng=_np(tf.zonalAng), zonalMag=_np(tf.zonalMag), phase=_np(tf.phase), y=_np(y), gx=_np(xr.grad),
               g_zonalAng=_np(tf.zonalAng.grad), g_zonalMag=_np(tf.zonalMag.grad))
    if ftype == 1:
        out["g_phase"] = _np(tf.phase.grad)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "N", d.num_nodes, "E", e.shape[0], "|y|max", float(y.abs().max()))


ECHO_CASES = [
    # name, n_side, deg, C, n_bins, B, R (of the FCPrecomp call that produces ln / wxp)
    ("echo_nb2", 8, 14.0, 4, 2, 1, 6),
    ("echo_nb3", 7, 12.0, 6, 3, 2, 4),
]


def make_echo(name, n_side, deg, c, n_bins, B, R, seed):
    """ECHO descriptors (nn/echo.py:94-148) on FCPrecomp's (supp_edges, ln, wxp) — SURVEY.md §8(f) F2."""
    ns = ref_loader.load()
    d = syn.torus_mesh(n_side, deg=deg, seed=seed, tile=4)
    g = torch.Generator().manual_seed(seed)
    keep = d.supp_edges[:, 1] != 3                      # an isolated target
    for k in ("supp_edges", "logMag", "logAng", "xp"):
        setattr(d, k, getattr(d, k)[keep])
    e, _, ln, wxp = ns.FCPrecomp(B, R, float(d.epsilon))(d)
    x = torch.complex(torch.randn(d.num_nodes, c, generator=g), torch.randn(d.num_nodes, c, generator=g))
    x[torch.rand(x.shape, generator=g) < 0.08] = 0.0
    echo = ns.ECHO(c, n_bins)
    xr = x.clone().requires_grad_(True)
    y = echo(xr, e, ln, wxp)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out = dict(n=np.int64(d.num_nodes), c=np.int64(c), n_bins=np.int64(n_bins), hdim=np.int64(echo.hdim), supp_edges=_np(e),
               ln=_np(ln), wxp=_np(wxp), x=_np(x), gy=_np(gy), y=_np(y), gx=_np(xr.grad), dMap=_np(echo.dMap))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "N", d.num_nodes, "E", e.shape[0], "hdim", echo.hdim, "|y|max", float(y.abs().max()))


def build_net(ns, B, R, ftype, n_classes, n_des, n_bins):
    """The reference notebooks' network shape in small: correspondence.ipynb `Net` (JSON lines of `class Net`: LiftBlock ->
    FCResNetBlocks with TangentPerceptron 'meta' residuals, one frontloaded block -> ECHOBlock), here with 4 / 8 channels.
    `ns` supplies the module classes: the unmodified reference (this script) or fieldconv_b200 (tests/test_gpu_net.py)."""
    import torch.nn as tnn

    class Net(tnn.Module):
        def __init__(self):
            super().__init__()
            self.lift = ns.LiftBlock(3, 4, n_rings=R, ftype=ftype)
            self.resnet1 = ns.FCResNetBlock(4, 8, band_limit=B, n_rings=R, ftype=ftype)
            self.resnet2 = ns.FCResNetBlock(8, 8, band_limit=B, n_rings=R, ftype=ftype)
            self.resnet3 = ns.FCResNetBlock(8, n_des, band_limit=B, n_rings=R, ftype=ftype, frontload=True)
            self.res1 = ns.TangentPerceptron(4, 8)
            self.res2 = ns.TangentPerceptron(8, n_des)
            self.echo = ns.ECHOBlock(n_des, n_classes, n_des=n_des, n_bins=n_bins, band_limit=B, n_rings=R, ftype=ftype)
            self.band_limit = B

        def forward(self, pos, supp_edges, supp_sten, ln, wxp):
            b = self.band_limit
            x1 = self.lift(pos, supp_edges, supp_sten[..., b:(b + 2)])
            x = self.resnet1(x1, supp_edges, supp_sten)
            x2 = self.resnet2(x, supp_edges, supp_sten) + self.res1(x1)
            x = self.resnet3(x2, supp_edges, supp_sten) + self.res2(x2)
            return self.echo(x, supp_edges, supp_sten, ln, wxp)

    return Net()


def make_net(name="net_b2r6", seed=21):
    """A whole network of the reference's modules, forward + cross-entropy + backward, by the unmodified reference:
    logits, loss and the gradient of EVERY parameter (LiftBlock, FCResNetBlock incl. frontload, TangentPerceptron, ECHOBlock)."""
    ns = ref_loader.load()
    B, R, ftype, n_classes, n_des, n_bins = 2, 6, 1, 5, 4, 2
    d = syn.torus_mesh(9, deg=14.0, seed=seed, tile=4)
    g = torch.Generator().manual_seed(seed)
    e, sten, ln, wxp = ns.FCPrecomp(B, R, float(d.epsilon))(d)
    pos = torch.randn(d.num_nodes, 3, generator=g)
    labels = torch.randint(0, n_classes, (d.num_nodes,), generator=g)
    torch.manual_seed(seed)
    net = build_net(ns, B, R, ftype, n_classes, n_des, n_bins)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("nonlin.bias") or ".nonlin1.bias" in k or ".nonlin2.bias" in k:
                p.uniform_(-0.2, 0.2)
    logits = net(pos, e, sten, ln, wxp)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    out = dict(n=np.int64(d.num_nodes), B=np.int64(B), R=np.int64(R), ftype=np.int64(ftype), n_classes=np.int64(n_classes),
               n_des=np.int64(n_des), n_bins=np.int64(n_bins), epsilon=np.float64(d.epsilon), raw_edges=_np(d.supp_edges),
               logMag=_np(d.logMag), logAng=_np(d.logAng), xp=_np(d.xp), w=_np(d.w), supp_edges=_np(e), supp_sten=_np(sten),
               ln=_np(ln), wxp=_np(wxp), pos=_np(pos), labels=_np(labels), logits=_np(logits), loss=np.float64(loss.item()))
    for k, v in net.state_dict().items():
        out["p." + k] = _np(v)
    for k, v in net.named_parameters():
        out["g." + k] = _np(v.grad)
    # The reference's OWN fp32 sensitivity to the summation order of scatter_add (nn/field_conv.py:134, nn/echo.py:143-147):
    # the same network on the same inputs with the edge list permuted.  Logits move by ~2e-7, but parameter gradients by up
    # to 1.5e-4 normwise (the deep chain through modReLU thresholds and ECHO's floor/ceil binning is ill-conditioned in
    # fp32), so a gradient tolerance below that would test rounding luck, not parity: "noise.<param>" = the largest
    # normwise deviation over 4 permutations, the yardstick tests/test_gpu_net.py scales its gradient tolerance by.
    noise = {k: 0.0 for k, _ in net.named_parameters()}
    noise_logits = 0.0
    state = {k: v.clone() for k, v in net.state_dict().items()}

    def dev(a, b):
        return max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)),
                   float(torch.linalg.vector_norm((a - b).reshape(-1)) / torch.linalg.vector_norm(b.reshape(-1)).clamp_min(1e-30)))

    for t in range(4):
        perm = torch.randperm(e.shape[0], generator=g)
        net2 = build_net(ns, B, R, ftype, n_classes, n_des, n_bins)
        net2.load_state_dict(state)
        lg = net2(pos, e[perm], sten[perm], ln[perm], wxp[perm])
        torch.nn.functional.cross_entropy(lg, labels).backward()
        noise_logits = max(noise_logits, dev(lg.detach(), logits.detach()))
        for (k, p2), (_, p1) in zip(net2.named_parameters(), net.named_parameters()):
            noise[k] = max(noise[k], dev(p2.grad, p1.grad))
    out["noise_logits"] = np.float64(noise_logits)
    for k, v in noise.items():
        out["noise." + k] = np.float64(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "N", d.num_nodes, "E", e.shape[0], "params", sum(p.numel() for p in net.parameters()), "loss", loss.item(),
          "order noise: logits %.1e, gradients up to %.1e" % (noise_logits, max(noise.values())))


def main():
    import sys
    if not ref_loader.available():
        raise SystemExit("reference tree not present; golden vectors can only be generated in the build container")
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else "all"      # "all" | "fc" | "lift" | "echo" | "net"
    if only in ("all", "fc"):
        for i, c in enumerate(CASES):
            make_case(*c, seed=100 + i)
        make_block()
    if only in ("all", "lift"):
        for i, c in enumerate(LIFT_CASES):
            make_lift(*c, seed=200 + i)
    if only in ("all", "echo"):
        for i, c in enumerate(ECHO_CASES):
            make_echo(*c, seed=300 + i)
    if only in ("all", "net"):
        make_net()


if __name__ == "__main__":
    main()
