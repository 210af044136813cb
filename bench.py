#!/usr/bin/env python
"""Benchmark of the FieldConv hot path (BASELINE.json metric: FieldConv fwd+bwd edges/sec).

Workload (BASELINE.json configs[1], SURVEY.md §8(d) cfg 2): SHAPESEG-style FC-ResNet segmentation
net — 5 FCResNetBlock(48,48, band_limit=2, n_rings=6, ftype=1) + |x| -> Linear(48,8) -> cross-entropy —
on a batch of 16 synthetic 5041-vertex torus meshes (~40 support edges/vertex) merged block-diagonally
per GPU; one step = forward + backward + Adam step.  edges/sec = (kept edges) x (10 FieldConv layers) /
step time.  N > 1: data parallel, one 16-mesh batch per rank (weak scaling), NCCL all-reduce of the
parameter gradients every step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the CPU port of the reference's own formulation (oracle/restate.py,
field_conv_refstyle) on a bounded sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "fieldconv_fwd_bwd_edges_per_sec"
UNIT = "edges/s"
B, R, C, N_BLOCKS, N_CLASSES = 2, 6, 48, 5, 8
MESHES_PER_RANK, N_SIDE, DEG = 16, 71, 40.0
CPU_SAMPLE_SIDE = 36            # bounded CPU sample: one FCResNetBlock on a 1296-vertex mesh


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


# ----------------------------------------------------------------------------- reference arm (CPU port)
def cpu_sample(steps, warmup, threads=None):
    """One FCResNetBlock(48,48,B=2,R=6) fwd+bwd in the reference's formulation on a small mesh."""
    from fieldconv_b200.synthetic import random_features, torus_mesh
    from fieldconv_b200 import nn as fnn
    from oracle import restate
    if threads:
        torch.set_num_threads(threads)
    mesh = torus_mesh(CPU_SAMPLE_SIDE, deg=DEG, seed=0)
    e, sten, _, _, _ = restate.fc_precomp(mesh.logMag, mesh.logAng, mesh.w, mesh.supp_edges, mesh.xp, B, R, mesh.epsilon)
    torch.manual_seed(0)
    blk = fnn.FCResNetBlock(C, C, B, R, 1)          # parameter container only (reference init); math is the oracle's
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in blk.state_dict().items()}
    x = random_features(mesh.num_nodes, C, seed=1)
    gy = random_features(mesh.num_nodes, C, seed=2, zero_frac=0)
    times = []
    for it in range(warmup + steps):
        xr = x.clone().requires_grad_(True)
        for v in params.values():
            v.grad = None
        t0 = time.perf_counter()
        y = restate.fc_resnet_block(xr, e, sten, params, B, 1, refstyle=True)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    edges = 2 * e.shape[0]                           # two FieldConv layers per block
    return edges, times, mesh.num_nodes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    edges, times, n = cpu_sample(args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = edges / (ms * 1e-3)
    sample = "1 FCResNetBlock(48,48,B=2,R=6) fwd+bwd on one %d-vertex mesh (%d edge-convs/step)" % (n, edges)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fcresnet5_c48_b2_r6_16x5k (bounded CPU sample of it)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
class Net(torch.nn.Module):
    def __init__(self, precision):
        super().__init__()
        import fieldconv_b200 as fcb
        self.blocks = torch.nn.ModuleList([fcb.FCResNetBlock(C, C, B, R, 1, precision=precision) for _ in range(N_BLOCKS)])
        self.head = torch.nn.Linear(C, N_CLASSES)

    def forward(self, x, plan):
        import fieldconv_b200 as fcb
        fcb.prefold(self)                 # the 10 layers' filters folded in one batch of torch ops (same arithmetic)
        for b in self.blocks:
            x = b(x, plan)
        return self.head(x.abs())


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.15)]
        window = "timed region"
        if len(inside) < 2:          # very short timed region: fall back to every sample taken under load (warm-up on)
            inside, window = [r for (_, r) in self.rows], "warm-up + timed region"
        for r in inside:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def algorithmic_bytes_aggregate(n, e, c, transpose=False):
    """Per launch of k_aggregate: each array once — 24 B/edge of plan records
    (16 B rec + 8 B rot), the feature rows read once (N*C*8), row pointers, and the N*R*C*M*8 B result written."""
    m = 2 * B + 1
    return e * 24 + (n + 1) * 4 + n * c * 8 + n * R * c * m * 8


def run_ours(args):
    import torch.distributed as dist
    import fieldconv_b200 as fcb
    from fieldconv_b200 import _lib
    from fieldconv_b200.synthetic import merge_meshes, random_features, torus_mesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic batch (per rank: its own 16 meshes), host-resident pinned copies for the e2e leg
    meshes = [torus_mesh(N_SIDE, deg=DEG, seed=1000 * rank + i, device=dev) for i in range(MESHES_PER_RANK)]
    batch = merge_meshes(meshes)
    del meshes
    n = batch.num_nodes
    x_dev = random_features(n, C, seed=rank, device=dev)
    labels = torch.randint(0, N_CLASSES, (n,), device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    host = {k: getattr(batch, k).cpu().pin_memory() for k in ("supp_edges", "logMag", "logAng", "xp", "w")}
    host["x"] = x_dev.cpu().pin_memory()
    host["labels"] = labels.cpu().pin_memory()

    torch.manual_seed(0)                       # identical initial weights on every rank
    net = Net(args.precision).to(dev)
    params = [p for p in net.parameters()]
    opt = torch.optim.Adam(params, lr=0.01)
    plan = fcb.build_plan(batch.supp_edges, batch.logMag, batch.logAng, batch.xp, batch.w, R, batch.epsilon)
    e_kept = plan.num_edges
    edges_per_step = e_kept * 2 * N_BLOCKS
    loss_fn = torch.nn.CrossEntropyLoss()

    def allreduce_grads():
        if world == 1:
            return
        grads = [p.grad for p in params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= world
        torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def step(x, pl, lab):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net(x, pl), lab)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(args.warmup):
        step(x_dev, plan, labels)
    if os.environ.get("FIELDCONV_B200_NCU"):     # profiler capture of exactly one step (ncu --profile-from-start off)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(x_dev, plan, labels)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    barrier()
    t_begin = time.perf_counter()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(x_dev, plan, labels)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * edges_per_step / (ms_step * 1e-3)

    # ---- end-to-end leg: pinned host buffers -> H2D -> device plan build -> fwd+bwd+step -> D2H loss.
    #      A loader stream stages step k+1 (H2D copies + plan build) while step k computes — every step's copies and
    #      plan build are still inside the timed region, the first step's included.
    loader = torch.cuda.Stream(device=dev)

    def stage():
        with torch.cuda.stream(loader):
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            pl = fcb.build_plan(d["supp_edges"], d["logMag"], d["logAng"], d["xp"], d["w"], R, batch.epsilon)
            ready = torch.cuda.Event()
            ready.record(loader)
        return d, pl, ready

    def consume(staged):
        d, pl, ready = staged
        main = torch.cuda.current_stream()
        main.wait_event(ready)
        for t in (d["x"], d["labels"]) + pl.tensors():
            t.record_stream(main)           # allocated on the loader stream, used on the compute stream
        return d["x"], pl, d["labels"]

    def e2e_run(k):
        losses = torch.empty(k, dtype=torch.float32).pin_memory()
        staged = stage()
        for i in range(k):
            xs, pl, lab = consume(staged)
            loss = step(xs, pl, lab)
            losses[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)    # D2H of the step's result
            if i + 1 < k:
                staged = stage()            # overlaps with the step just enqueued
        torch.cuda.synchronize()
        assert bool(torch.isfinite(losses).all()), "non-finite loss in the end-to-end leg"
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    e2e_run(max(2, min(args.warmup, 3)))
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    ev0.record()
    e2e_run(e2e_steps)
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0)) / e2e_steps
    e2e_wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_ms = max(e2e_ms, e2e_wall_ms)      # host-side work (copies are async, loss.item() syncs) is part of e2e
    e2e_value = world * edges_per_step / (e2e_ms * 1e-3)

    # ---- per-kernel timing of one more step with CUDA events around every library launch (recorded on rank 0;
    #      every rank runs the step because it contains the gradient all-reduce)
    if rank == 0:
        _lib.profile_enable(1 << 14)
    step(x_dev, plan, labels)
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_sus, which = peaks()
    recs = _lib.profile_collect(1 << 14)
    tot = {}
    for name, ms in recs:
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    lib_ms = sum(v[1] for v in tot.values())
    shares = {k: {"launches": v[0], "ms": round(v[1], 4), "share_of_step": round(v[1] / ms_step, 4)} for k, v in
              sorted(tot.items(), key=lambda kv: -kv[1][1])}
    # dominant kernel among the FieldConv path's own kernels (the "lin_*" records are TangentLin's small GEMMs)
    dominant = max(((k, v) for k, v in tot.items() if not k.startswith("lin_")), key=lambda kv: kv[1][1])[0]
    m = 2 * B + 1
    k_complex = R * C * m
    avg_ms = tot[dominant][1] / tot[dominant][0]
    if dominant.startswith("aggregate"):
        per_launch = algorithmic_bytes_aggregate(n, e_kept, C)
        roof = {"kernel": "k_aggregate" + ("<transpose>" if "_T" in dominant else "") + (" (packed fp16 output)" if dominant.endswith("_pk") else ""),
                "bound": "hbm", "achieved": per_launch / (avg_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "algorithmic_bytes_per_launch": per_launch,
                "algorithmic_tflops": 14.0 * C * m * e_kept / (avg_ms * 1e-3) / 1e12,
                "note": "gather + segmented reduction: E*24 B of plan records, the feature rows once, the N x R*M*C complex "
                        "result written once; ncu (profiles/) shows the kernel issue/FMA-pipe bound (65 % issue-active), "
                        "the HBM fraction is the contract's figure"}
    else:
        # contraction launches of one name share one shape family: the forward-shaped (N x 2K) @ (2K x 2Co) product;
        # algorithmic bytes = the A operand read once + B + C written once
        per_launch = n * 2 * k_complex * 4 + 2 * k_complex * 2 * C * 4 + n * 2 * C * 4
        flops = 8.0 * n * k_complex * C
        roof = {"kernel": "k_gemm (" + dominant + ")", "bound": "hbm", "achieved": per_launch / (avg_ms * 1e-3) / 1e9,
                "peak": hbm, "unit": "GB/s", "algorithmic_bytes_per_launch": per_launch,
                "fp32_tflops_achieved": flops / (avg_ms * 1e-3) / 1e12,
                "note": ("tcgen05 contraction: reads the contrib operand once (HBM floor) — fp32 split into fp16 (hi, lo) planes "
                         "by producer warps (gemm_h_*) or packed fp16 planes bulk-copied (gemm_p_*); %.0f flop/B"
                         if dominant.startswith(("gemm_h", "gemm_p", "gemm_tc")) else
                         "FP32-FMA contraction (arithmetic intensity %.0f flop/B): the binding resource is the FMA pipe, "
                         "the HBM fraction is reported because the contract asks for it") % (flops / per_launch)}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = which
    roof["avg_launch_ms"] = avg_ms
    roof["launches_averaged"] = tot[dominant][0]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same
    # layer shape (profiles/ncu_traffic.json, written from profiles/*_ncu_full_summary.md), or null for kernels without one
    roof["traffic"] = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if dominant in tr.get("cfg2_layer", {}):
            roof["traffic"] = tr["cfg2_layer"][dominant]
            roof["traffic_source"] = tr.get("source")
    except (OSError, ValueError):
        pass

    # ---- CPU baseline on the box's host cores (bounded sample)
    if args.skip_cpu_baseline:             # A/B runs of tools/*.sh only; the driver's command never passes this
        cpu = None
    else:
        edges_c, times_c, n_c = cpu_sample(steps=2, warmup=1)
        cpu_val = edges_c / (sum(times_c) / len(times_c))
        cpu = {"value": cpu_val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "1 FCResNetBlock(48,48,B=2,R=6) fwd+bwd on one %d-vertex mesh, reference formulation "
                         "(oracle/restate.py field_conv_refstyle), mean of 2 runs after 1 warm-up" % n_c}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "fcresnet5_c48_b2_r6_16x5k", "blocks": N_BLOCKS, "channels": C, "band_limit": B, "n_rings": R,
                   "meshes_per_gpu": MESHES_PER_RANK, "vertices_per_gpu": n, "edges_per_gpu": e_kept,
                   "fieldconv_layers": 2 * N_BLOCKS, "precision": args.precision, "parallelism": "dp%d" % world,
                   "l2": "working set per step (>=0.9 GB of contrib per layer) exceeds the 126 MB L2; no explicit flush",
                   "step": "forward + backward + NCCL grad all-reduce (N>1) + Adam"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "includes": "H2D of x + raw mesh attributes from pinned memory, device plan build "
                                                   "(FCPrecomp + 2 CSR sorts), fwd+bwd+Adam, async D2H of the loss into pinned memory "
                                                   "every step; step k+1 is staged on a loader stream while step k computes"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernel_shares": shares,
        "library_ms_per_step": lib_ms,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- cfg 4: one large mesh, vertex partition
def run_partitioned(args):
    """BASELINE.json configs[3]: a single side^2-vertex torus mesh (64 support edges/vertex), C=64, band_limit=1,
    n_rings=6, one FCResNetBlock (2 FieldConv layers) + |x| -> Linear -> CE; vertex-partitioned across the ranks,
    halo rows exchanged over NCCL/NVLink and overlapped with the interior rows; strong scaling."""
    import torch.distributed as dist
    import fieldconv_b200 as fcb
    from fieldconv_b200 import _lib
    from fieldconv_b200.synthetic import random_features, torus_mesh
    c4, b4, r4, deg4 = 64, 1, 6, 64.0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mesh = torus_mesh(args.side, deg=deg4, seed=0, device=dev)
    n_global = mesh.num_nodes
    part = fcb.partition_mesh(mesh, world, rank)
    del mesh
    torch.cuda.empty_cache()
    plan = part.build_plan(r4)
    e_local = plan.num_edges
    t = torch.tensor([e_local], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t)
    e_global = int(t.item())
    torch.manual_seed(0)
    blk = fcb.FCResNetBlock(c4, c4, b4, r4, 1, precision=args.precision).to(dev)
    head = torch.nn.Linear(c4, N_CLASSES).to(dev)
    params = list(blk.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=0.01)
    x_own = random_features(n_global, c4, seed=1, device=dev)[part.own_global].contiguous()
    labels = torch.randint(0, N_CLASSES, (n_global,), device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    labels = labels[part.own_global].contiguous()
    host_x, host_lab = x_own.cpu().pin_memory(), labels.cpu().pin_memory()
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")

    def step(x, lab):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(head(blk(x, part).abs()), lab) / n_global
        loss.backward()
        fcb.allreduce_gradients(params)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tt = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(args.warmup):
        step(x_own, labels)
    barrier()
    t_begin = time.perf_counter()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(x_own, labels)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    value = e_global * 2 / (ms_step * 1e-3)

    def e2e_step():
        return float(step(host_x.to(dev, non_blocking=True), host_lab.to(dev, non_blocking=True)).item())
    e2e_step()
    barrier()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_e2e
    # one more step with per-launch CUDA events (every rank runs it: the step contains collectives)
    if rank == 0:
        _lib.profile_enable(1 << 12)
    step(x_own, labels)
    barrier()
    if rank == 0:
        hbm, _, which = peaks()
        tot = {}
        for name, ms in _lib.profile_collect(1 << 12):
            a = tot.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ms
        shares = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}
        n_own = part.n_own
        agg = tot.get("aggregate", [1, 0.0])
        per_launch = e_local * 24 + (n_own + 1) * 4 + part.n_ext * c4 * 8 + n_own * r4 * c4 * (2 * b4 + 1) * 8
        avg_ms = agg[1] / max(agg[0], 1)
        roof = {"kernel": "k_aggregate", "bound": "hbm", "achieved": per_launch / max(avg_ms, 1e-9) / 1e6, "peak": hbm,
                "unit": "GB/s", "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms, "peak_source": which,
                "traffic": None}
        roof["frac"] = roof["achieved"] / roof["peak"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "partitioned_mesh_%dx%d_c64_b1_r6_deg64" % (args.side, args.side), "vertices": n_global,
                           "edges": e_global, "fieldconv_layers": 2, "precision": args.precision,
                           "parallelism": "vertex-partition x%d + NCCL halo exchange overlapped with interior rows" % world,
                           "rank0": {"owned": part.n_own, "interior": part.n_interior, "halo": part.n_halo},
                           "l2": "per-layer working set (>= 4 GB of contrib per rank) exceeds the 126 MB L2",
                           "step": "forward + backward + halo exchanges + NCCL grad all-reduce + Adam"},
                "e2e": {"value": e_global * 2 / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": host_x.numel() * 8 + host_lab.numel() * 8, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms, "includes": "H2D of this rank's features + labels (pinned), step, D2H loss; "
                                                           "the static partition/plan is built once"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_ms_rank0": shares}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("FIELDCONV_B200_PRECISION", "auto"))
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="kernel A/B runs only: omit the cpu_baseline leg")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"],
                    help="cfg2 (default, the driver's line): FC-ResNet on a batch of 16 meshes per GPU, data parallel; "
                         "cfg4: one large mesh vertex-partitioned across the GPUs with NVLink halo exchange (strong scaling)")
    ap.add_argument("--side", type=int, default=2000, help="cfg4: the mesh has side^2 vertices (2000 -> 4M)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
        if args.workload == "cfg4":
            run_partitioned(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
