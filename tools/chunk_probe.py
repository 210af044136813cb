#!/usr/bin/env python
"""Feasibility probe (not a product path): does the forward of one FieldConv layer get faster when the row range is cut into
chunks whose `contrib` fits the 126 MB L2, the aggregation of chunk k+1 running on one stream while the contraction of
chunk k runs on another?  Uses only exported building blocks (fcb_aggregate_f32, fcb_gemm_f32).  One JSON line per setting.

    python tools/chunk_probe.py --side 284 --channels 48 --band 2 --rings 6 --chunks 1 16 32
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=284)
    ap.add_argument("--deg", type=float, default=40.0)
    ap.add_argument("--channels", type=int, default=48)
    ap.add_argument("--band", type=int, default=2)
    ap.add_argument("--rings", type=int, default=6)
    ap.add_argument("--chunks", type=int, nargs="+", default=[1, 8, 16, 32])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import fieldconv_b200 as fcb
    from fieldconv_b200 import _lib
    from fieldconv_b200.synthetic import torus_mesh, random_features
    dev = torch.device("cuda", 0)
    mesh = torus_mesh(args.side, deg=args.deg, seed=0, device=dev)
    plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, args.rings, mesh.epsilon)
    n, c, b, r = mesh.num_nodes, args.channels, args.band, args.rings
    m = 2 * b + 1
    k2 = 2 * r * m * c
    x = random_features(n, c, seed=1, device=dev)
    xr = torch.view_as_real(x).contiguous()
    wmat = torch.randn(k2, 2 * c, device=dev) / k2 ** 0.5
    y = torch.empty(n, 2 * c, device=dev)
    flags = 3
    s_main = torch.cuda.current_stream()
    s_agg, s_mm = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)

    def ws_for(rows):
        return torch.empty(_lib.query_bytes("fcb_gemm_workspace_bytes", rows, 2 * c, k2, 0, 1, 1, flags) + 256, dtype=torch.uint8, device=dev)

    def aggregate(a, bb, out, stream):
        _lib.call("fcb_aggregate_f32", xr.data_ptr(), plan.rowptr_tgt.data_ptr() + 4 * a, plan.rec_tgt.data_ptr(),
                  plan.rot_tgt.data_ptr(), out.data_ptr(), bb - a, c, b, r, 0, stream.cuda_stream)

    def gemm(a, bb, src, ws, stream):
        _lib.call("fcb_gemm_f32", src.data_ptr(), wmat.data_ptr(), y.data_ptr() + 4 * a * 2 * c, bb - a, 2 * c, k2, k2, 2 * c, 2 * c,
                  0, 1, 0, 0, 0, 1, None, None, ws.data_ptr(), ws.numel(), flags, stream.cuda_stream)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    ref = None
    for nch in args.chunks:
        rows = (n + nch - 1) // nch
        bounds = [(i * rows, min(n, (i + 1) * rows)) for i in range(nch) if i * rows < n]
        bufs = [torch.empty(rows, k2, device=dev) for _ in range(2 if nch > 1 else 1)]
        wss = [ws_for(rows) for _ in bufs]

        def sequential():
            for i, (a, bb) in enumerate(bounds):
                aggregate(a, bb, bufs[i % len(bufs)], s_main)
                gemm(a, bb, bufs[i % len(bufs)], wss[i % len(bufs)], s_main)

        def overlapped():
            s_agg.wait_stream(s_main)
            s_mm.wait_stream(s_main)
            done = [None, None]
            for i, (a, bb) in enumerate(bounds):
                j = i % 2
                if done[j] is not None:
                    s_agg.wait_event(done[j])              # the contraction that read this buffer two chunks ago
                aggregate(a, bb, bufs[j], s_agg)
                ev = torch.cuda.Event()
                ev.record(s_agg)
                s_mm.wait_event(ev)
                gemm(a, bb, bufs[j], wss[j], s_mm)
                done[j] = torch.cuda.Event()
                done[j].record(s_mm)
            s_main.wait_stream(s_agg)
            s_main.wait_stream(s_mm)

        out = {"chunks": len(bounds), "rows_per_chunk": rows, "contrib_MB_per_chunk": round(rows * k2 * 4 / 1e6, 1),
               "sequential_ms": round(timed(sequential), 4)}
        _lib.profile_enable(4096)
        sequential()
        torch.cuda.synchronize()
        kern = {}
        for name, t in _lib.profile_collect(4096):
            kern[name] = round(kern.get(name, 0.0) + t, 4)
        out["sequential_kernels_ms"] = kern
        if ref is None:
            ref = y.clone()
        else:
            out["max_abs_diff_vs_first"] = float((y - ref).abs().max())
        if nch > 1:
            out["overlapped_ms"] = round(timed(overlapped), 4)
            out["max_abs_diff_overlapped"] = float((y - ref).abs().max())
        print(json.dumps(out), flush=True)
        del bufs, wss
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
