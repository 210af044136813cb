#!/usr/bin/env python
"""One-process probe of an aggregation-kernel variant (FIELDCONV_B200_AGG_VARIANT is read once per process):
parity of one small layer against the fp64 oracle (the oracle is only the checker), then fwd+bwd time of one cfg-2 sized
layer.  usage: FIELDCONV_B200_AGG_VARIANT=135,135,135,135 python tools/variant_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fieldconv_b200 as fcb  # noqa: E402
from fieldconv_b200.synthetic import random_features, torus_mesh  # noqa: E402


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float(torch.linalg.vector_norm((a - b).reshape(-1)) / torch.linalg.vector_norm(b.reshape(-1)))


def main():
    dev = "cuda:0"
    out = {"variant": os.environ.get("FIELDCONV_B200_AGG_VARIANT", "default")}
    from test_gpu_parity import _oracle_layer
    prec = os.environ.get("PROBE_PRECISION", "auto")
    out["precision"] = prec
    for (side, c, b, r) in ((24, 48, 2, 6), (30, 32, 1, 6)):
        mesh = torus_mesh(side, deg=40.0, seed=1, device=dev)
        torch.manual_seed(0)
        m = fcb.FieldConv(c, c, b, r, 1, precision=prec).to(dev)
        plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, r, mesh.epsilon)
        x = random_features(mesh.num_nodes, c, seed=2, device=dev).requires_grad_(True)
        gy = random_features(mesh.num_nodes, c, seed=3, zero_frac=0, device=dev)
        y = m(x, plan)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        y_ref, gx_ref, gp = _oracle_layer(mesh, x, m, gy)
        out["err_B%d" % b] = {"y": rel(y, y_ref.to(torch.complex64)), "gx": rel(x.grad, gx_ref.to(torch.complex64)),
                              "g_zonal": rel(m.zonal.grad, gp[0].float()), "g_spherical": rel(m.spherical.grad, gp[1].float())}
    if os.environ.get("PROBE_PARITY_ONLY"):
        print(json.dumps(out), flush=True)
        return
    for (side, c, b, r, tag) in ((284, 48, 2, 6, "cfg2_layer_ms"), (1000, 32, 1, 6, "1M_c32_ms")):
        mesh = torus_mesh(side, deg=40.0, seed=0, device=dev)
        plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, r, mesh.epsilon)
        torch.manual_seed(0)
        m = fcb.FieldConv(c, c, b, r, 1).to(dev)
        x = random_features(mesh.num_nodes, c, seed=1, device=dev).requires_grad_(True)
        gy = random_features(mesh.num_nodes, c, seed=2, zero_frac=0, device=dev)

        def step():
            x.grad = None
            m(x, plan).backward(gy)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        out[tag] = e0.elapsed_time(e1) / 5
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
