"""Debug: where the e2e leg of bench.py spends its time (H2D, plan build, step)."""
import sys, time, torch
sys.path.insert(0, '.')
import bench, fieldconv_b200 as fcb
from fieldconv_b200.synthetic import merge_meshes, random_features, torus_mesh
dev = torch.device('cuda', 0)
meshes = [torus_mesh(bench.N_SIDE, deg=bench.DEG, seed=i, device=dev) for i in range(bench.MESHES_PER_RANK)]
batch = merge_meshes(meshes); n = batch.num_nodes
x = random_features(n, bench.C, seed=0, device=dev)
labels = torch.randint(0, 8, (n,), device=dev)
host = {k: getattr(batch, k).cpu().pin_memory() for k in ("supp_edges", "logMag", "logAng", "xp", "w")}
host["x"] = x.cpu().pin_memory(); host["labels"] = labels.cpu().pin_memory()
net = bench.Net("auto").to(dev); opt = torch.optim.Adam(net.parameters(), lr=0.01); lf = torch.nn.CrossEntropyLoss()
def step(x, pl, lab):
    opt.zero_grad(set_to_none=True); loss = lf(net(x, pl), lab); loss.backward(); opt.step(); return loss
plan = fcb.build_plan(batch.supp_edges, batch.logMag, batch.logAng, batch.xp, batch.w, bench.R, batch.epsilon)
for _ in range(3): step(x, plan, labels)
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter()
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}; torch.cuda.synchronize(); t1 = time.perf_counter()
    pl = fcb.build_plan(d["supp_edges"], d["logMag"], d["logAng"], d["xp"], d["w"], bench.R, batch.epsilon); torch.cuda.synchronize(); t2 = time.perf_counter()
    l = step(d["x"], pl, d["labels"]); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("h2d %.1f ms  plan %.1f ms  step %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
