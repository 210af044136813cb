import time, torch, json, os, sys
sys.path.insert(0, os.getcwd())
import fieldconv_b200 as fcb
from fieldconv_b200.synthetic import merge_meshes, torus_mesh
dev = torch.device("cuda", 0)
meshes = [torus_mesh(71, deg=40.0, seed=i, device=dev) for i in range(16)]
batch = merge_meshes(meshes)
host = {k: getattr(batch, k).cpu().pin_memory() for k in ("supp_edges", "logMag", "logAng", "xp", "w")}
nbytes = sum(v.numel() * v.element_size() for v in host.values())
s = torch.cuda.Stream()
def h2d():
    with torch.cuda.stream(s):
        return {k: v.to(dev, non_blocking=True) for k, v in host.items()}
def plan(m):
    with torch.cuda.stream(s):
        return fcb.build_plan(m["supp_edges"], m["logMag"], m["logAng"], m["xp"], m["w"], 6, batch.epsilon)
out = {}
for name, fn in (("h2d", lambda: h2d()), ("h2d+plan", lambda: plan(h2d()))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): fn()
    e1.record(s)
    t_host = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    out[name] = {"gpu_ms": e0.elapsed_time(e1) / 10, "host_ms": t_host * 1e3}
out["h2d_GBps"] = nbytes / out["h2d"]["gpu_ms"] / 1e6
out["nbytes"] = nbytes
print(json.dumps(out))
