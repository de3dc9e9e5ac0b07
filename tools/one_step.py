#!/usr/bin/env python
"""Exactly one profiled fwd+bwd step between cudaProfilerStart/Stop, after two warm steps.
Run under `ncu --profile-from-start off ...` to capture the kernels of ONE step and nothing else.
usage: one_step.py [scene res spp max_bounces edges]     (default: the bench workload C2 = shadow_blocker 512 64 1 3)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from redner_b200 import api
from redner_b200 import redner as rb
import scenes
a = sys.argv[1:]
scene, res, spp, mb, edges = (a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4])) if len(a) >= 5 else ("shadow_blocker", 512, 64, 1, 3)
dev = torch.device("cuda:0")
sc = scenes.SCENES[scene](dev, resolution=(res, res))
def step():
    args = api.RenderFunction.serialize_scene(sc, spp, mb, sampler_type=rb.SamplerType.sobol, device=dev, use_primary_edge_sampling=bool(edges & 1),
                                              use_secondary_edge_sampling=bool(edges & 2))
    img = api.RenderFunction.apply(1, *args)
    img.pow(2).sum().backward()
    torch.cuda.synchronize()
for _ in range(2): step()
torch.cuda.profiler.start()
step()
torch.cuda.profiler.stop()
print("one step done")
