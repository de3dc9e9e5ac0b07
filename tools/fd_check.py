#!/usr/bin/env python
"""Finite-difference diagnosis on the GPU: translate C2's blocker, central differences of sum(img) at several step sizes vs the
analytic gradient split by estimator (interior / + primary edges / + secondary edges), with seed-to-seed scatter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import scenes
from redner_b200 import api
from redner_b200 import redner as rb
dev = torch.device("cuda:0")
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SCENE, SHAPE = (sys.argv[2], int(sys.argv[3])) if len(sys.argv) > 3 else ("shadow_blocker", 1)


def loss(shift, axis, spp, seed):
    sc = scenes.SCENES[SCENE](dev, resolution=(res, res), grad=False)
    v = sc.shapes[SHAPE].vertices.clone(); v[:, axis] += shift; sc.shapes[SHAPE].vertices = v
    args = api.RenderFunction.serialize_scene(sc, spp, 1, sampler_type=rb.SamplerType.sobol, device=dev, backend=rb)
    return float(api.RenderFunction.apply(seed, *args).double().sum())


def analytic(edges, spp, seed):
    sc = scenes.SCENES[SCENE](dev, resolution=(res, res))
    args = api.RenderFunction.serialize_scene(sc, spp, 1, sampler_type=rb.SamplerType.sobol, device=dev, backend=rb,
                                              use_primary_edge_sampling=bool(edges & 1), use_secondary_edge_sampling=bool(edges & 2))
    api.RenderFunction.apply(seed, *args).sum().backward()
    return sc.shapes[SHAPE].vertices.grad.double().sum(0).cpu().numpy()


seeds = list(range(1, 9))
for edges in (0, 1, 2, 3):
    a = np.stack([analytic(edges, 256, s) for s in seeds])
    print("analytic edges=%d mean %s  sem %s" % (edges, np.round(a.mean(0), 1), np.round(a.std(0, ddof=1) / np.sqrt(len(seeds)), 1)), flush=True)
for eps in (0.04, 0.02, 0.01):
    for axis in (0, 1, 2):
        f = np.array([(loss(eps, axis, 1024, s) - loss(-eps, axis, 1024, s)) / (2 * eps) for s in seeds])
        print("fd eps=%.3f axis=%d mean %.1f sem %.1f" % (eps, axis, f.mean(), f.std(ddof=1) / np.sqrt(len(seeds))), flush=True)
