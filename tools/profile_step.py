"""Kernel-time breakdown of the C4 bench step via torch.profiler (CUPTI), top kernels by device time."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
dev = torch.device("cuda", 0)
cfg = bench.C4
ei = bench.rmat_edge_index(cfg["nodes"], cfg["edges"], cfg["rmat_scale"], 0, dev)
gen = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((cfg["nodes"], cfg["feats"]), device=dev, generator=gen)
y = torch.randint(0, cfg["classes"], (cfg["nodes"],), device=dev, generator=gen)
layers = bench.build_model(cfg["feats"], cfg["hidden"], cfg["classes"])
params = [p for l in layers for p in l.trainable_weights]
opt = torch.optim.SGD(params, lr=1e-3)
def step():
    opt.zero_grad(set_to_none=True)
    h = x
    for l in layers: h = l([h, ei])
    loss = bench.cross_entropy(h, y); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"<.*", "", e.name)[:60]
        tot[name][0] += 1; tot[name][1] += e.device_time
s = sum(v[1] for v in tot.values())
print(f"total device time per step {s/2/1e3:.2f} ms")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{t/2/1e3:8.2f} ms/step  n={n//2:3d}  {k}")
# idle time between consecutive kernels of the two profiled steps: where the device waits for the host
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
             key=lambda e: e.time_range.start)
span = evs[-1].time_range.end - evs[0].time_range.start
gaps = []
for a, b in zip(evs, evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 3:
        gaps.append((g, re.sub(r"<.*", "", a.name)[:40], re.sub(r"<.*", "", b.name)[:40]))
print(f"span per step {span/2/1e3:.2f} ms, idle per step {sum(g for g, _, _ in gaps)/2/1e3:.2f} ms in {len(gaps)//2} gaps > 3 us")
for g, a, b in sorted(gaps, reverse=True)[:14]:
    print(f"{g:8.0f} us  after {a}  before {b}")
