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
WORKLOAD = "fcresnet5_c48_b2_r6_16x5k"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal: 148 SMs x 128 FP32 lanes x FMA at clocks.max.sm


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


def base_config(world, precision):
    """The workload description shared by both arms (the reference arm measures a bounded sample of exactly this)."""
    return {"workload": WORKLOAD, "blocks": N_BLOCKS, "channels": C, "band_limit": B, "n_rings": R,
            "meshes_per_gpu": MESHES_PER_RANK, "vertices_per_mesh": N_SIDE * N_SIDE, "edges_per_vertex": DEG,
            "fieldconv_layers": 2 * N_BLOCKS, "precision": precision, "parallelism": "dp%d" % world,
            "l2": "working set per step (>=0.9 GB of contrib per layer) exceeds the 126 MB L2; no explicit flush",
            "step": "forward + backward + NCCL grad all-reduce (N>1) + Adam"}


# ----------------------------------------------------------------------------- reference arm (the reference's CPU path)
def host_threads():
    """All host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _reference_ns():
    """The UNMODIFIED reference (nn/field_conv.py etc.) when its tree is reachable (FIELDCONV_REFERENCE or
    /root/reference — never on the GPU box), else None: the caller then times the oracle's port of the same formulation."""
    try:
        from oracle import ref_loader
        if ref_loader.available():
            return ref_loader.load()
    except Exception:
        pass
    return None


def cpu_block_sample(steps, warmup):
    """Bounded sample of the bench workload on the host: ONE FCResNetBlock(48,48,B=2,R=6) fwd+bwd on ONE of the
    5041-vertex meshes (1 of the 5 blocks x 1 of the 16 meshes of a step), in the reference's own formulation
    (nn/fc_resnet_block.py:84-88 over nn/field_conv.py:104-137: per-edge product tensor, index-sum, broadcast filter)."""
    from fieldconv_b200.synthetic import random_features, torus_mesh
    from fieldconv_b200 import nn as fnn
    from oracle import restate
    mesh = torus_mesh(N_SIDE, deg=DEG, seed=0)
    e, sten, _, _, _ = restate.fc_precomp(mesh.logMag, mesh.logAng, mesh.w, mesh.supp_edges, mesh.xp, B, R, mesh.epsilon)
    torch.manual_seed(0)
    blk = fnn.FCResNetBlock(C, C, B, R, 1)          # parameter container only (reference init)
    x = random_features(mesh.num_nodes, C, seed=1)
    gy = random_features(mesh.num_nodes, C, seed=2, zero_frac=0)
    ns = _reference_ns()
    if ns is not None:
        ref = ns.FCResNetBlock(C, C, B, R, 1)
        ref.load_state_dict(blk.state_dict())
        params = list(ref.parameters())

        def run(xr):
            return ref(xr, e, sten)
        kind = "reference"
    else:
        pd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in blk.state_dict().items()}
        params = list(pd.values())

        def run(xr):
            return restate.fc_resnet_block(xr, e, sten, pd, B, 1, refstyle=True)
        kind = "port"
    times = []
    for it in range(warmup + steps):
        xr = x.clone().requires_grad_(True)
        for v in params:
            v.grad = None
        t0 = time.perf_counter()
        y = run(xr)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    edges = 2 * e.shape[0]                           # two FieldConv layers per block
    sample = ("1 FCResNetBlock(48,48,B=2,R=6) fwd+bwd on one %d-vertex mesh of the batch (%d edge-convs/step; 1 of 5 blocks x 1 of "
              "16 meshes), %s" % (mesh.num_nodes, edges, "unmodified reference nn/fc_resnet_block.py" if kind == "reference"
                                  else "oracle/restate.py port of the reference formulation (reference tree not on this box)"))
    return edges, times, kind, sample


def cpu_cfg1_layer(threads, iters, warmup):
    """one FieldConv layer fwd+bwd at BASELINE config 1 (5041 vertices, ~40 edges/vertex, C=32, B=1,
    R=6, ftype=1) in the reference's formulation (nn/field_conv.py:104-137)."""
    from fieldconv_b200.synthetic import random_features, torus_mesh
    from fieldconv_b200 import nn as fnn
    from oracle import restate
    b1, c1 = 1, 32
    torch.set_num_threads(threads)
    mesh = torus_mesh(N_SIDE, deg=DEG, seed=0)
    e, sten, _, _, _ = restate.fc_precomp(mesh.logMag, mesh.logAng, mesh.w, mesh.supp_edges, mesh.xp, b1, R, mesh.epsilon)
    torch.manual_seed(0)
    lay = fnn.FieldConv(c1, c1, b1, R, 1)
    x = random_features(mesh.num_nodes, c1, seed=1)
    gy = random_features(mesh.num_nodes, c1, seed=2, zero_frac=0)
    ns = _reference_ns()
    if ns is not None:
        ref = ns.FieldConv(c1, c1, b1, R, 1)
        ref.load_state_dict(lay.state_dict())
        params = list(ref.parameters())

        def run(xr):
            return ref(xr, e, sten)
    else:
        params = [p.detach().clone().requires_grad_(True) for p in (lay.zonal, lay.spherical, lay.phase)]

        def run(xr):
            return restate.field_conv_refstyle(xr, e, sten, params[0], params[1], params[2], 1, b1)
    ts = []
    for it in range(warmup + iters):
        xr = x.clone().requires_grad_(True)
        for v in params:
            v.grad = None
        t0 = time.perf_counter()
        y = run(xr)
        (y.real * gy.real + y.imag * gy.imag).sum().backward()
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    ne = int(e.shape[0])
    return {"threads": torch.get_num_threads(), "edges": ne, "min_s": ts[0], "median_s": ts[len(ts) // 2],
            "edges_per_s": ne / ts[0], "kind": "reference" if ns is not None else "port"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    edges, times, kind, sample = cpu_block_sample(args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = edges / (ms * 1e-3)
    cfg1 = {"all_threads": cpu_cfg1_layer(cores, 5, 2), "one_thread": cpu_cfg1_layer(1, 2, 1),
            "what": "BASELINE.md §3 / BASELINE.json configs[0]: one FieldConv(32,32,B=1,R=6) layer fwd+bwd, 5041 vertices"}
    torch.set_num_threads(cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": base_config(args.gpus, args.precision),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "host_cpu_count": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cfg1_layer": cfg1,
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


def contract_bytes_layer(n, e, ci, co, b, r):
    """SURVEY.md §8(d) ALGORITHMIC bytes of one FieldConv layer, each array once, compact edge format, contrib / G never in
    HBM: (forward, forward + backward)."""
    m = 2 * b + 1
    k = r * ci * m
    fwd = e * 20 + (n + 1) * 4 + n * ci * 8 + n * co * 8 + co * k * 8
    return fwd, fwd + e * 20 + e * 24 + n * (2 * ci + co) * 8 + 2 * co * k * 8


def kernel_bytes(name, n, e, c, b, r):
    """Kernel-level bytes of one launch at the cfg-2 layer shape (what THIS kernel must move given that contrib / G live in
    HBM between kernels): plan records + feature rows + the N x 2K result for the aggregations, the A operand read once +
    B + C for the contractions."""
    m = 2 * b + 1
    k2 = 2 * r * c * m                                  # real columns of contrib / G
    if name.startswith("aggregate"):
        return e * 24 + (n + 1) * 4 + n * c * 8 + n * k2 * 4
    if name.startswith(("gemm_h_nn", "gemm_p_nn", "gemm_tc_nn", "gemm_nn")):
        return n * k2 * 4 + k2 * 2 * c * 4 + n * 2 * c * 4
    if name.startswith(("gemm_h_tn", "gemm_p_tn", "gemm_tc_tn", "gemm_tn")):
        return n * k2 * 4 + n * 2 * c * 4 * m + k2 * 2 * c * 4
    return None


def ncu_reference():
    """Counters of the committed `ncu --set full` capture of the same kernels at the cfg-2 layer shape (profiles/ncu_traffic.json,
    written from profiles/*_ncu_full_summary.md): DRAM bytes per launch, issue-active %, FMA-pipe %, tensor-pipe %."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        return {}


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
    # high priority: the staging work is many small kernels (plan build) next to the step's SM-filling ones — at equal priority
    # two boxes ran them only at the step's kernel boundaries (35-38 ms per step instead of 18 when the plan is rebuilt every step)
    loader = torch.cuda.Stream(device=dev, priority=-1)
    plan_cache = {}                       # mesh-batch id -> device plan (+ the device copies of its static attributes)

    def stage(batch_id, cached):
        """H2D of one step's inputs on the loader stream.  The mesh set of a training run is static (the reference's
        datasets are fixed lists of meshes): with `cached` the plan of a batch is built on its FIRST use (inside the timed
        region) and kept on the device, so later steps copy only features + labels; without, every step copies the raw
        mesh attributes and rebuilds the plan (FCPrecomp arithmetic + 2 CSR sorts) like the reference's Net.forward."""
        with torch.cuda.stream(loader):
            d = {k: host[k].to(dev, non_blocking=True) for k in ("x", "labels")}
            pl = plan_cache.get(batch_id) if cached else None
            nbytes = host["x"].numel() * host["x"].element_size() + host["labels"].numel() * host["labels"].element_size()
            if pl is None:
                m = {k: host[k].to(dev, non_blocking=True) for k in ("supp_edges", "logMag", "logAng", "xp", "w")}
                pl = fcb.build_plan(m["supp_edges"], m["logMag"], m["logAng"], m["xp"], m["w"], R, batch.epsilon)
                nbytes += sum(host[k].numel() * host[k].element_size() for k in m)
                if cached:
                    plan_cache[batch_id] = pl
            ready = torch.cuda.Event()
            ready.record(loader)
        return d, pl, ready, nbytes

    def consume(staged):
        d, pl, ready, _ = staged
        main = torch.cuda.current_stream()
        main.wait_event(ready)
        for t in (d["x"], d["labels"]) + pl.tensors():
            t.record_stream(main)           # allocated on the loader stream, used on the compute stream
        return d["x"], pl, d["labels"]

    def e2e_run(k, cached):
        losses = torch.empty(k, dtype=torch.float32).pin_memory()
        copied = []
        staged = stage(0, cached)
        for i in range(k):
            copied.append(staged[3])
            xs, pl, lab = consume(staged)
            loss = step(xs, pl, lab)
            losses[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)    # D2H of the step's result
            if i + 1 < k:
                staged = stage(0, cached)   # overlaps with the step just enqueued
        torch.cuda.synchronize()
        assert bool(torch.isfinite(losses).all()), "non-finite loss in the end-to-end leg"
        return copied

    def e2e_measure(cached):
        plan_cache.clear()
        e2e_run(max(2, min(args.warmup, 3)), cached)
        plan_cache.clear()                  # the timed region starts cold: its first step builds the plan
        barrier()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        ev0.record()
        copied = e2e_run(k, cached)
        ev1.record()
        barrier()
        ms = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0)) / k
        wall = max_over_ranks((time.perf_counter() - t0) * 1e3) / k
        return max(ms, wall), copied, k      # host-side work (copies are async) is part of e2e

    e2e_ms, copied, e2e_steps = e2e_measure(cached=True)
    e2e_value = world * edges_per_step / (e2e_ms * 1e-3)
    h2d = sum(copied) / len(copied)
    e2e_ms_rebuild, copied_rb, _ = e2e_measure(cached=False)
    e2e_rebuild = {"value": world * edges_per_step / (e2e_ms_rebuild * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_rebuild,
                   "h2d_bytes_per_step": sum(copied_rb) / len(copied_rb),
                   "what": "no plan cache: every step copies the raw mesh attributes and rebuilds the plan on the device"}
    plan_cache.clear()

    # ---- per-kernel timing of one more step with CUDA events around every library launch (recorded on rank 0;
    #      every rank runs the step because it contains the gradient all-reduce)
    if rank == 0:
        _lib.profile_enable(1 << 14)
    step(x_dev, plan, labels)
    barrier()
    recs = _lib.profile_collect(1 << 14) if rank == 0 else []
    del opt, net, params, plan, batch, x_dev, labels
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3] inside the same run: NCCL parity of the vertex-partitioned layer (N >= 2) and the 4M-vertex
    #      strong-scaling point of this N, so the driver's 1 -> 2 -> 4 -> 8 runs carry the whole curve
    parity = partition_parity(world, rank, dev) if world > 1 and not args.skip_cfg4 else None
    cfg4 = None
    if not args.skip_cfg4:
        cfg4 = measure_cfg4(args.cfg4_side, world, rank, dev, args.precision, steps=3, warmup=2, e2e=False)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_sus, which = peaks()
    tot = {}
    for name, ms in recs:
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    lib_ms = sum(v[1] for v in tot.values())
    shares = {k: {"launches": v[0], "ms": round(v[1], 4), "share_of_step": round(v[1] / ms_step, 4)} for k, v in
              sorted(tot.items(), key=lambda kv: -kv[1][1])}
    # kernel families of the FieldConv path (the "lin_*" records are TangentLin's small GEMMs)
    fam = {}
    for k, v in tot.items():
        if k.startswith("lin_"):
            continue
        f = "k_aggregate" if k.startswith("aggregate") else ("k_gemm_*_nn" if "_nn" in k and k.startswith("gemm") else
                                                               ("k_gemm_*_tn" if "_tn" in k and k.startswith("gemm") else k))
        a = fam.setdefault(f, [0, 0.0, []])
        a[0] += v[0]
        a[1] += v[1]
        a[2].append(k)
    dominant = max(fam.items(), key=lambda kv: kv[1][1])[0]
    m = 2 * B + 1
    ncu = ncu_reference()
    per_kernel = {}
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:6]:
        kb = kernel_bytes(k, n, e_kept, C, B, R)
        avg = v[1] / v[0]
        ent = {"launches": v[0], "avg_launch_ms": round(avg, 4)}
        if kb:
            ent.update({"kernel_bytes_per_launch": kb, "GBps": round(kb / (avg * 1e-3) / 1e9, 1),
                        "frac_of_hbm_peak": round(kb / (avg * 1e-3) / 1e9 / hbm, 4)})
        if isinstance(ncu.get("cfg2_layer", {}).get(k), dict):
            ent["ncu"] = ncu["cfg2_layer"][k]
        per_kernel[k] = ent
    fwd_b, all_b = contract_bytes_layer(n, e_kept, C, C, B, R)
    contract_step = all_b * 2 * N_BLOCKS
    d_launches, d_ms = fam[dominant][0], fam[dominant][1]
    avg_ms = d_ms / d_launches
    if dominant == "k_aggregate":
        per_launch = kernel_bytes("aggregate", n, e_kept, C, B, R)
        flops = 14.0 * C * m * e_kept
        roof = {"kernel": "k_aggregate (forward + transposed launches: %s)" % ", ".join(sorted(fam[dominant][2])),
                "bound": "hbm", "achieved": per_launch / (avg_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "algorithmic_bytes_per_launch": per_launch,
                "fp32_pipe": {"achieved_tflops": flops / (avg_ms * 1e-3) / 1e12, "peak_tflops_nominal": FP32_PEAK_TFLOPS,
                              "frac": flops / (avg_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                              "flops_per_launch": flops, "definition": "SURVEY.md §8(d): 14*Ci*M flops per edge"},
                "note": "gather + segmented reduction; kernel-level bytes = E*24 B of plan records + the feature rows once + the "
                        "N x 2K fp32 result written once.  The kernel is bound by FP32 instruction issue / the FMA pipe, not by "
                        "HBM (see fp32_pipe and the ncu counters): the HBM fraction is reported because the contract asks for it"}
    else:
        per_launch = kernel_bytes(fam[dominant][2][0], n, e_kept, C, B, R) or 0
        flops = 8.0 * n * R * C * m * C
        roof = {"kernel": dominant + " (" + ", ".join(sorted(fam[dominant][2])) + ")", "bound": "hbm",
                "achieved": per_launch / (avg_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "algorithmic_bytes_per_launch": per_launch, "fp32_tflops_achieved": flops / (avg_ms * 1e-3) / 1e12,
                "note": "tcgen05 contraction: reads its N x 2K operand once (HBM floor), 2xFP16 split operands"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = which
    roof["avg_launch_ms"] = avg_ms
    roof["launches_averaged"] = d_launches
    roof["share_of_step"] = round(d_ms / ms_step, 4)
    # the contract's own figure: §8(d) bytes of the whole step (contrib / G counted as on-chip) over the step time
    roof["contract"] = {"bytes_per_step": contract_step, "bytes_per_edge_fwd_bwd": all_b / e_kept,
                        "achieved_GBps": contract_step / (ms_step * 1e-3) / 1e9,
                        "frac": contract_step / (ms_step * 1e-3) / 1e9 / hbm,
                        "note": "SURVEY.md §8(d) algorithmic bytes x 10 layers / step time; the kernels actually move the N x 2K "
                                "contrib / G through HBM between aggregation and contraction (see traffic)"}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same
    # layer shape, for the dominant kernel and summed over one layer's forward + backward
    roof["traffic"] = None
    key = fam[dominant][2][0]
    ent = ncu.get("cfg2_layer", {}).get(key)
    if isinstance(ent, dict):
        roof["traffic"] = ent.get("dram_bytes")
        roof["ncu"] = {kk: ent.get(kk) for kk in ("issue_active_pct", "fma_pipe_pct", "tensor_pipe_pct", "dram_pct")}
        roof["traffic_source"] = ncu.get("source")
    elif ent is not None:
        roof["traffic"] = ent
        roof["traffic_source"] = ncu.get("source")
    if "layer_dram_bytes_fwd_bwd" in ncu:
        roof["traffic_layer_fwd_bwd"] = ncu["layer_dram_bytes_fwd_bwd"]
        roof["contract_bytes_layer_fwd_bwd"] = all_b
    roof["per_kernel"] = per_kernel

    # ---- CPU baseline on the box's host cores (bounded sample of the same workload, rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:       # --skip-cpu-baseline: A/B runs of tools/*.sh only
        cores = host_threads()
        edges_c, times_c, kind_c, sample_c = cpu_block_sample(steps=2, warmup=1)
        cpu = {"value": edges_c / (sum(times_c) / len(times_c)), "unit": UNIT, "cores": cores, "kind": kind_c,
               "sample": sample_c + "; mean of 2 runs after 1 warm-up"}

    cfg = base_config(world, args.precision)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": cfg,
        "measured": {"vertices_per_gpu": n, "edges_per_gpu": e_kept},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "h2d_bytes_first_step": copied[0], "h2d_bytes_later_steps": copied[-1],
                "includes": "every step: H2D of that step's features + labels from pinned memory, fwd+bwd+Adam, async D2H of the loss "
                            "into pinned memory; the batch's plan is built from its raw mesh attributes (H2D + FCPrecomp arithmetic + "
                            "2 CSR sorts on the device) on the FIRST step inside the timed region and cached per mesh-batch id for the "
                            "later ones (static mesh set); step k+1 is staged on a loader stream while step k computes",
                "rebuild_plan_every_step": e2e_rebuild},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernel_shares": shares,
        "library_ms_per_step": lib_ms,
        "partition_parity": parity,
        "cfg4": cfg4,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the vertex-partitioned layer disagrees with the single-GPU layer: %s" % json.dumps(parity))


# ----------------------------------------------------------------------------- cfg 4: one large mesh, vertex partition
def partition_parity(world, rank, dev):
    """NCCL parity of the vertex-partitioned path (run by every rank, N >= 2): one FCResNetBlock(32,32,B=1,R=6) on a
    ~100k-vertex mesh, partitioned across the ranks (halo exchange over NCCL, overlapped with the interior rows), against
    the same block on the whole mesh on one GPU.  Matches nn/field_conv.py:130-134 (one-hop gather) of the reference.
    Returns normwise relative errors (max-norm, L2) of y, grad x and the parameter gradients; ok = all <= 1e-5 (2e-5 for the
    modReLU bias gradients, as in tests/test_gpu_parity.py)."""
    import torch.distributed as dist
    import fieldconv_b200 as fcb
    from fieldconv_b200.synthetic import random_features, torus_mesh
    c, b, r, side = 32, 1, 6, 316
    mesh = torus_mesh(side, deg=DEG, seed=5, device=dev)
    n = mesh.num_nodes
    torch.manual_seed(1)
    blk = fcb.FCResNetBlock(c, c, b, r, 1).to(dev)
    x = random_features(n, c, seed=11, device=dev)
    gy = random_features(n, c, seed=12, zero_frac=0, device=dev)
    part = fcb.partition_mesh(mesh, world, rank)
    own = part.own_global
    xo = x[own].contiguous().requires_grad_(True)
    y = blk(xo, part)
    y.backward(gy[own].contiguous())
    fcb.allreduce_gradients(list(blk.parameters()))
    y_all = torch.zeros(n, c, dtype=torch.complex64, device=dev)
    gx_all = torch.zeros(n, c, dtype=torch.complex64, device=dev)
    y_all[own] = y.detach()
    gx_all[own] = xo.grad
    for t in (y_all, gx_all):
        dist.all_reduce(torch.view_as_real(t))
    out = None
    if rank == 0:
        gp = {k: p.grad.detach().clone() for k, p in blk.named_parameters()}
        for p in blk.parameters():
            p.grad = None
        plan = fcb.build_plan(mesh.supp_edges, mesh.logMag, mesh.logAng, mesh.xp, mesh.w, r, mesh.epsilon)
        xr = x.clone().requires_grad_(True)
        y_ref = blk(xr, plan)
        y_ref.backward(gy)

        def rel(a, bb):
            d = (a - bb).abs()
            return max(float(d.max() / bb.abs().max().clamp_min(1e-30)),
                       float(torch.linalg.vector_norm(d.reshape(-1)) / torch.linalg.vector_norm(bb.reshape(-1)).clamp_min(1e-30)))
        errs = {"y": rel(y_all, y_ref.detach()), "gx": rel(gx_all, xr.grad)}
        ok = errs["y"] <= 1e-5 and errs["gx"] <= 1e-5
        worst_p, worst_b = 0.0, 0.0
        for k, p in blk.named_parameters():
            e = rel(gp[k], p.grad)
            if "bias" in k:
                worst_b = max(worst_b, e)
            else:
                worst_p = max(worst_p, e)
        errs["param_grads"], errs["bias_grads"] = worst_p, worst_b
        ok = ok and worst_p <= 1e-5 and worst_b <= 2e-5
        out = {"ok": bool(ok), "rel_err": errs, "tolerance": 1e-5, "world": world, "vertices": n, "edges": plan.num_edges,
               "halo_rows_rank0": part.n_halo, "interior_rows_rank0": part.n_interior,
               "what": "FCResNetBlock(32,32,B=1,R=6) fwd+bwd, vertex-partitioned over NCCL vs the whole mesh on rank 0"}
    flag = torch.tensor([1 if (out is None or out["ok"]) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    del mesh, part, blk
    torch.cuda.empty_cache()
    return out


def measure_cfg4(side, world, rank, dev, precision, steps, warmup, e2e):
    """BASELINE.json configs[3]: a single side^2-vertex torus mesh (64 support edges/vertex), C=64, band_limit=1,
    n_rings=6, one FCResNetBlock (2 FieldConv layers) + |x| -> Linear -> CE; vertex-partitioned across the ranks,
    halo rows exchanged over NCCL/NVLink and overlapped with the interior rows; strong scaling.  The process group is the
    caller's.  Returns the record on rank 0 (None elsewhere)."""
    import torch.distributed as dist
    import fieldconv_b200 as fcb
    from fieldconv_b200 import _lib, partition as fpart
    from fieldconv_b200.synthetic import random_features, torus_mesh
    c4, b4, r4, deg4 = 64, 1, 6, 64.0
    t_setup = time.perf_counter()
    mesh = torus_mesh(side, deg=deg4, seed=0, device=dev)
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
    blk = fcb.FCResNetBlock(c4, c4, b4, r4, 1, precision=precision).to(dev)
    head = torch.nn.Linear(c4, N_CLASSES).to(dev)
    params = list(blk.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=0.01)
    x_own = random_features(n_global, c4, seed=1, device=dev)[part.own_global].contiguous()
    labels = torch.randint(0, N_CLASSES, (n_global,), device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    labels = labels[part.own_global].contiguous()
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")
    setup_s = time.perf_counter() - t_setup

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

    for _ in range(warmup):
        step(x_own, labels)
    barrier()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step(x_own, labels)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    value = e_global * 2 / (ms_step * 1e-3)
    rec = {"workload": "partitioned_mesh_%dx%d_c64_b1_r6_deg64" % (side, side), "vertices": n_global, "edges": e_global,
           "fieldconv_layers": 2, "n_gpus": world, "scaling": "strong", "steps": steps, "warmup": warmup,
           "ms_per_step": ms_step, "value": value, "unit": UNIT, "gpu_launches": launches, "setup_s": round(setup_s, 2),
           "parallelism": "vertex-partition x%d + NCCL halo exchange overlapped with interior rows" % world,
           "step": "forward + backward + halo exchanges + NCCL grad all-reduce + Adam",
           "backward": "no contrib kept or recomputed: gW from G and xhat (csrc/api.cu backward_common)"}
    if e2e:
        host_x, host_lab = x_own.cpu().pin_memory(), labels.cpu().pin_memory()

        def e2e_step():
            return float(step(host_x.to(dev, non_blocking=True), host_lab.to(dev, non_blocking=True)).item())
        e2e_step()
        barrier()
        n_e2e = max(2, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_e2e
        rec["e2e"] = {"value": e_global * 2 / (e2e_ms * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": host_x.numel() * 8 + host_lab.numel() * 8, "d2h_bytes_per_step": 4,
                      "ms_per_step": e2e_ms, "includes": "H2D of this rank's features + labels (pinned), step, D2H loss; "
                                                         "the static partition/plan is built once"}
    # exposed communication: events around the halo exchanges of one more step (every rank runs it: collectives inside)
    fpart.HALO_TIMING = []
    if rank == 0:
        _lib.profile_enable(1 << 12)
    step(x_own, labels)
    barrier()
    halo = fpart.collect_halo_timing()
    fpart.HALO_TIMING = None
    comm = torch.tensor([halo["exposed_ms"], halo["comm_ms"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(comm, op=dist.ReduceOp.MAX)
    rec["halo"] = {"rows_rank0": part.n_halo, "owned_rank0": part.n_own, "interior_rank0": part.n_interior,
                   "bytes_per_layer_per_direction_rank0": part.n_halo * c4 * 8,
                   "exposed_ms_per_step_max_rank": float(comm[0].item()), "comm_ms_per_step_max_rank": float(comm[1].item()),
                   "definition": "comm = device time of the halo send/recv on the side stream; exposed = time the compute stream "
                                 "waited for it (CUDA events), summed over the 2 layers x 2 directions of one step"}
    out = None
    if rank == 0:
        tot = {}
        for name, ms in _lib.profile_collect(1 << 12):
            a = tot.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ms
        rec["kernel_ms_rank0"] = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}
        out = rec
    del blk, head, opt, params, plan, part, x_own, labels
    torch.cuda.empty_cache()
    return out


def run_partitioned(args):
    """`--workload cfg4`: the cfg-4 record as the bench line itself."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    parity = partition_parity(world, rank, dev) if world > 1 else None
    t_begin = time.perf_counter()
    rec = measure_cfg4(args.side, world, rank, dev, args.precision, args.steps, args.warmup, e2e=True)
    if rank == 0:
        clocks = sampler.stop(t_begin, time.perf_counter())
        e2e = rec.pop("e2e")
        line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": rec["workload"], "vertices": rec["vertices"], "edges": rec["edges"], "fieldconv_layers": 2,
                           "precision": args.precision, "parallelism": rec["parallelism"],
                           "l2": "per-layer working set (GBs of G per rank) exceeds the 126 MB L2", "step": rec["step"]},
                "e2e": e2e, "gpu_launches": rec["gpu_launches"], "clocks": clocks, "halo": rec["halo"],
                "kernel_ms_rank0": rec["kernel_ms_rank0"], "partition_parity": parity}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: partition parity failed: %s" % json.dumps(parity))


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
    ap.add_argument("--cfg4-side", type=int, default=2000, help="side of the cfg-4 sub-record's mesh inside the default run")
    ap.add_argument("--skip-cfg4", action="store_true", help="kernel A/B runs only: omit the cfg-4 sub-record and the NCCL partition parity")
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
