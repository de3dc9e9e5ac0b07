#!/usr/bin/env python
"""GPU vs host edge trees of a scene: which words of which records differ (debug aid for tests/test_scene_build_gpu.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import scenes
from redner_b200 import api
from redner_b200 import redner as rb
dev = torch.device("cuda:0")
NAMES = ["pmin.x", "pmin.y", "pmin.z", "pmax.x", "pmax.y", "pmax.z", "dmin.x", "dmin.y", "dmin.z", "dmax.x", "dmax.y", "dmax.z", "wlen", "ref"]


def trees(scene, res):
    sc = scenes.SCENES[scene](dev, resolution=(res, res))
    args = api.RenderFunction.serialize_scene(sc, 1, 1, sampler_type=rb.SamplerType.sobol, device=dev, backend=rb, use_secondary_edge_sampling=True)
    out = None
    for rep in range(3):
        t0 = time.perf_counter()
        c = api.RenderFunction._unpack((1, 2), args)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        out = (c.scene.edge_trees(), c.scene.build_ms(), ms)
    return out


for scene, res in [(a, 32) for a in sys.argv[1:]]:
    os.environ.pop("RB_HOST_TREES", None)
    os.environ["RB_GPU_TREES"] = "1"
    (g, cs_g, ncs_g, ex_g), ms_g, wall_g = trees(scene, res)
    os.environ.pop("RB_GPU_TREES", None)
    os.environ["RB_HOST_TREES"] = "1"
    (h, cs_h, ncs_h, ex_h), ms_h, wall_h = trees(scene, res)
    print("== %s: records %d / %d, roots gpu (%d, %d) host (%d, %d), expand %.9g / %.9g; third build: gpu %s wall %.2f ms | host %s wall %.2f ms" %
          (scene, len(g), len(h), cs_g, ncs_g, cs_h, ncs_h, ex_g, ex_h, {k: round(v, 2) for k, v in ms_g.items()}, wall_g, {k: round(v, 2) for k, v in ms_h.items()}, wall_h))
    if g.shape != h.shape:
        continue
    bad = np.argwhere(g[:, :28] != h[:, :28])
    shown = 0
    for r, w in bad:
        c, k = divmod(int(w), 14)
        if k == 12:
            continue
        gv, hv = (g[r, w].view(np.int32), h[r, w].view(np.int32)) if k == 13 else (g[r, w:w + 1].view(np.float32)[0], h[r, w:w + 1].view(np.float32)[0])
        print("   record %d child %d %-6s gpu %r host %r" % (r, c, NAMES[k], gv, hv))
        shown += 1
        if shown > 24:
            break
    if len(g) <= 8:
        for r in range(len(g)):
            print("   gpu ", r, g[r, [13, 27]].view(np.int32), g[r, :6].view(np.float32), g[r, 14:20].view(np.float32))
            print("   host", r, h[r, [13, 27]].view(np.int32), h[r, :6].view(np.float32), h[r, 14:20].view(np.float32))
