#!/usr/bin/env python
"""Where does the end-to-end step (bench.py's e2e leg) spend its time?  Phase timers with a synchronize after each
phase (diagnostic only; the bench itself never synchronises inside the step)."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from redner_b200 import api
from redner_b200 import redner as rb
import scenes
dev = torch.device("cuda:0")
host = bench.make_scene(api, scenes, dev, pose=0, pinned_host=True)
ht = {"floor_v": host.shapes[0].vertices, "floor_i": host.shapes[0].indices, "blk_v": host.shapes[1].vertices.detach(), "blk_i": host.shapes[1].indices,
      "light_v": host.shapes[2].vertices, "light_i": host.shapes[2].indices, "kd0": host.materials[0].diffuse_reflectance.texels.detach(),
      "kd1": host.materials[1].diffuse_reflectance.texels}
ht = {k: v.pin_memory() for k, v in ht.items()}
st = rb.SamplerType.sobol
def sync(): torch.cuda.synchronize(); return time.perf_counter()
def step(prof):
    t = [sync()]
    d = {k: v.to(dev, non_blocking=True) for k, v in ht.items()}
    blk = d["blk_v"].requires_grad_(True); kd0 = d["kd0"].requires_grad_(True)
    m0, m1 = api.Material(diffuse_reflectance=kd0), api.Material(diffuse_reflectance=d["kd1"])
    shapes = [api.Shape(d["floor_v"], d["floor_i"], 0), api.Shape(blk, d["blk_i"], 0), api.Shape(d["light_v"], d["light_i"], 1)]
    inten = torch.tensor([1000.0, 1000.0, 1000.0], requires_grad=True)
    scn = api.Scene(host.camera, shapes, [m0, m1], [api.AreaLight(2, inten)])
    t.append(sync())
    args = api.RenderFunction.serialize_scene(scn, bench.SPP, bench.MB, sampler_type=st, device=dev)
    t.append(sync())
    img = api.RenderFunction.apply(bench.SEED, *args)
    t.append(sync())
    loss = img.pow(2).sum()
    t.append(sync())
    loss.backward()
    t.append(sync())
    outs = [loss.detach().cpu(), blk.grad.cpu(), kd0.grad.cpu(), inten.grad]
    t.append(sync())
    return [1e3 * (b - a) for a, b in zip(t, t[1:])]
for i in range(3): step(False)
names = ["h2d+scene objects", "serialize_scene", "RenderFunction.apply (build+forward)", "loss", "backward", "d2h"]
acc = [0.0] * len(names)
N = 5
for i in range(N):
    for k, v in enumerate(step(False)): acc[k] += v / N
for n, v in zip(names, acc): print("%-40s %8.2f ms" % (n, v))
print("%-40s %8.2f ms" % ("sum", sum(acc)))
pr = cProfile.Profile(); pr.enable(); step(True); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:5000])
