#!/usr/bin/env python
"""Single-FieldConv-layer fwd+bwd timing on one GPU for any (mesh size, channels, band_limit, n_rings):
BASELINE.json configs[0] (cfg 1), configs[2] (cfg 3) and the configs[4] sweep (cfg 5).  Prints one JSON
line per configuration with edges/s, per-kernel CUDA-event times (fcb_profile_*) and the roofline figures
of SURVEY.md §8(d).  Not the driver's bench (that is bench.py); used for profiles/ and for ncu captures:

    python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6            # one layer at cfg-2 size
    python tools/layer_bench.py --sweep                                                 # cfg 5 grid on a 1M-vertex mesh
    FIELDCONV_B200_NCU=1 ncu --profile-from-start off ... python tools/layer_bench.py ...  # one profiled fwd+bwd
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def algorithmic_bytes(n, e, ci, co, b, r):
    """each array once, compact edge format (24 B/edge per CSR order)."""
    m = 2 * b + 1
    k = r * ci * m
    fwd = e * 24 + (n + 1) * 4 + n * ci * 8 + n * co * 8 + co * k * 8
    bwd = e * 24 + e * 24 + 2 * (n + 1) * 4 + n * (2 * ci + co) * 8 + 2 * co * k * 8
    return fwd, fwd + bwd


def run_one(mesh, plan, c, b, r, precision, steps, warmup, dev, ftype=1, tag="", graph=False):
    import fieldconv_b200 as fcb
    from fieldconv_b200 import _lib
    from fieldconv_b200.synthetic import random_features
    n = mesh.num_nodes
    torch.manual_seed(0)
    layer = fcb.FieldConv(c, c, b, r, ftype, precision=precision).to(dev)
    x = random_features(n, c, seed=1, device=dev).requires_grad_(True)
    gy = random_features(n, c, seed=2, zero_frac=0, device=dev)
    e = plan.num_edges

    def step():
        x.grad = None
        for p in layer.parameters():
            p.grad = None
        y = layer(x, plan)
        y.backward(gy)

    for _ in range(warmup):
        step()
    eager_step = step
    if graph:
        # the whole fwd+bwd (fold_weights, every library launch, autograd glue) as ONE CUDA graph: small meshes are
        # launch-bound in eager mode (cfg 1: 0.24 ms of kernels inside 1.0 ms of wall time)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eager_step()
        step = g.replay
        step()
    if os.environ.get("FIELDCONV_B200_NCU"):
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    _lib.profile_enable(4096)
    eager_step()
    torch.cuda.synchronize()
    recs = _lib.profile_collect(4096)
    kern = {}
    for name, t in recs:
        kern[name] = round(kern.get(name, 0.0) + t, 4)
    hbm, which = peaks()
    fwd_b, all_b = algorithmic_bytes(n, e, c, c, b, r)
    m = 2 * b + 1
    flops = 3 * 8.0 * r * c * m * c * n + 3 * 14.0 * c * m * e
    return {"tag": tag, "vertices": n, "edges": e, "channels": c, "band_limit": b, "n_rings": r, "precision": precision,
            "cuda_graph": bool(graph), "flags": layer_flags(layer), "ms_fwd_bwd": round(ms, 4), "edges_per_s": e / (ms * 1e-3),
            "algorithmic_GBps": all_b / (ms * 1e-3) / 1e9, "hbm_frac": all_b / (ms * 1e-3) / 1e9 / hbm, "hbm_peak": hbm,
            "peak_source": which, "algorithmic_TFLOPs": flops / (ms * 1e-3) / 1e12, "kernels_ms": kern,
            "library_ms": round(sum(kern.values()), 4)}


def layer_flags(layer):
    from fieldconv_b200 import nn as fnn
    ci, co = layer.in_channels, layer.out_channels
    return int(fnn._resolve_precision(layer.precision, ci + ci % 2, co + co % 2, layer.R, layer.B))   # may carry FLAG_PACKED


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=284)          # 284^2 = 80 656 vertices (cfg-2 batch size)
    ap.add_argument("--deg", type=float, default=40.0)
    ap.add_argument("--channels", type=int, default=48)
    ap.add_argument("--band", type=int, default=2)
    ap.add_argument("--rings", type=int, default=6)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays of the captured fwd+bwd")
    ap.add_argument("--permute", action="store_true", help="random vertex numbering (cache-hostile case)")
    ap.add_argument("--sweep", action="store_true", help="cfg 5: C x band_limit x n_rings grid on a 1000x1000 mesh")
    ap.add_argument("--sweep-side", type=int, default=1000)
    ap.add_argument("--tag", default=None, help="label copied into the output line (e.g. an environment setting under test)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("layer_bench: needs a CUDA device (no CPU path)")
    import fieldconv_b200 as fcb
    from fieldconv_b200.synthetic import torus_mesh
    dev = torch.device("cuda", 0)
    if not args.sweep:
        mesh = torus_mesh(args.side, deg=args.deg, seed=0, device=dev, permute=args.permute)
        plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, args.rings, mesh.epsilon)
        out = run_one(mesh, plan, args.channels, args.band, args.rings, args.precision, args.steps, args.warmup, dev,
                      tag=args.tag or ("permuted" if args.permute else "tiled"), graph=args.graph)
        if out:
            print(json.dumps(out), flush=True)
        return
    mesh = torus_mesh(args.sweep_side, deg=args.deg, seed=0, device=dev)
    for r in (2, 6):
        plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, r, mesh.epsilon)
        for b in (1, 2, 3):
            for c in (16, 32, 64, 128, 256):
                k = r * c * (2 * b + 1)
                if mesh.num_nodes * k * 8 > 70e9:             # one N x K buffer (contrib, then G) + workspace must fit the 180 GB HBM
                    print(json.dumps({"skipped": {"channels": c, "band_limit": b, "n_rings": r},
                                      "why": "one N x K operand buffer (N*K*8 B) exceeds the memory budget of one GPU"}), flush=True)
                    continue
                try:
                    out = run_one(mesh, plan, c, b, r, args.precision, max(2, args.steps // 3), 2, dev, tag="cfg5")
                    print(json.dumps(out), flush=True)
                except torch.OutOfMemoryError:
                    print(json.dumps({"skipped": {"channels": c, "band_limit": b, "n_rings": r}, "why": "OOM"}), flush=True)
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
