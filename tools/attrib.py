#!/usr/bin/env python
"""Per-kernel time attribution on the GPU: C2 (or another scene) with each combination of edge samplers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import scenes  # noqa: E402
from redner_b200 import api  # noqa: E402
from redner_b200 import redner as rb  # noqa: E402

if os.environ.get("RB_LIB"):
    from redner_b200 import _lib
    _lib._lib = _lib.load(os.environ["RB_LIB"])
scene = sys.argv[1] if len(sys.argv) > 1 else "shadow_blocker"
res = int(sys.argv[2]) if len(sys.argv) > 2 else 512
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 64
mb = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
for edges in [int(e) for e in os.environ.get("RB_EDGES", "0,1,2,3").split(",")]:
    for rep in range(2):
        sc = scenes.SCENES[scene](dev, resolution=(res, res))
        args = api.RenderFunction.serialize_scene(sc, spp, mb, sampler_type=rb.SamplerType.sobol, device=dev, backend=rb,
                                                  use_primary_edge_sampling=bool(edges & 1), use_secondary_edge_sampling=bool(edges & 2))
        c = api.RenderFunction._unpack((1, 1000004), args)
        img = torch.zeros(res, res, 3, device=dev)
        rb.render(c.scene, c.options, rb.float_ptr(img.data_ptr()), rb.float_ptr(0), None, rb.float_ptr(0), rb.float_ptr(0))
        f = c.scene.last_stage_stats()[0]["k_forward"]
        ctx = type("C", (), {})()
        ctx.c, ctx.args = c, args
        api.RenderFunction.backward(ctx, (2 * img).contiguous())
        st, v, h = c.scene.last_stage_stats()
    sub = " (trace %.2f sec %.2f sweep %.2f)" % (st["k_bwd_trace"], st["k_bwd_secondary"], st["k_bwd_sweep"]) if "k_bwd_trace" in st else ""
    print("%s %dx%dx%d mb=%d edges=%d: fwd %.2f ms | bwd %.2f%s | prim %.2f | vert/sample %.3f hits/sample %.3f" %
          (scene, res, res, spp, mb, edges, f, st["k_backward"], sub, st["k_primary_edge"], v / (res * res * spp), h / (res * res * spp)),
          "| build ms", {k: round(x, 2) for k, x in c.scene.build_ms().items()})
