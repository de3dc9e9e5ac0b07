#!/usr/bin/env python
"""Exactly one profiled fwd+bwd step of the bench workload (C2) between cudaProfilerStart/Stop, after two warm steps.
Run under `ncu --profile-from-start off ...` to capture the kernels of ONE step and nothing else."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from redner_b200 import api
from redner_b200 import redner as rb
import scenes
dev = torch.device("cuda:0")
sc = bench.make_scene(api, scenes, dev, pose=0)
def step():
    for p in (sc.shapes[1].vertices, sc.materials[0].diffuse_reflectance.texels, sc.area_lights[0].intensity):
        p.grad = None
    args = api.RenderFunction.serialize_scene(sc, bench.SPP, bench.MB, sampler_type=rb.SamplerType.sobol, device=dev)
    img = api.RenderFunction.apply(bench.SEED, *args)
    img.pow(2).sum().backward()
    torch.cuda.synchronize()
for _ in range(2): step()
torch.cuda.profiler.start()
step()
torch.cuda.profiler.stop()
print("one step done")
