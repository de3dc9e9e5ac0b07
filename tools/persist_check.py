#!/usr/bin/env python
"""Persistent-warp hierarchy pick vs the plain pick kernel (default; RB_PERSISTENT_PICK=1 selects the persistent kernel): same gradients up to the order of the atomics,
and the stage times of both on one box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import numpy as np, torch
    import parity_utils as pu
    from redner_b200 import redner as rb
    dev = torch.device("cuda:0")
    out = {}
    for name, cfg in (("teapot", dict(scene="teapot_geometry", res=64, spp=8, mb=2, sampler="sobol", edges=2)),
                      ("c2", dict(scene="shadow_blocker_all", res=64, spp=16, mb=1, sampler="sobol", edges=2)),
                      ("room", dict(scene="glossy_room", res=48, spp=8, mb=2, sampler="sobol", edges=2))):
        _, g = pu.render_case(rb, dev, cfg, 7)
        for k, v in g.items():
            out[name + "." + k] = v.numpy()
    np.savez(sys.argv[2], **out)
    sys.exit(0)
import numpy as np
res = {}
for tag, env in (("persistent", {"RB_PERSISTENT_PICK": "1"}), ("plain", {})):
    path = "/tmp/persist_%s.npz" % tag
    subprocess.run([sys.executable, __file__, "worker", path], check=True, env=dict(os.environ, **env), timeout=600)
    res[tag] = dict(np.load(path))
for k in res["plain"]:
    a, b = res["persistent"][k].astype(np.float64), res["plain"][k].astype(np.float64)
    n = np.linalg.norm(b)
    if n > 0:
        print("%-40s rel diff %.2e" % (k, np.linalg.norm(a - b) / n))
