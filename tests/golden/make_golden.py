#!/usr/bin/env python
"""Generates the committed golden vectors from the UNMODIFIED reference (compiled by oracle/build_ref.sh from
/root/reference/src, CPU/Embree path).  Run in the build container (where /root/reference exists):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

Each case renders a seeded synthetic scene (tests/scenes.py) through the host code in redner_b200/api.py with the
reference module as backend and stores the image and every gradient of loss = sum(img^2).
Cases with edge sampling whose samples cannot be reproduced sample-by-sample (secondary edges: the reference indexes
that stream by the rank of the pixel in its compacted active list, src/pathtracer.cpp:504-505) store the MEAN over
several seeds together with the standard error of that mean, for a statistical comparison.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import ref_loader  # noqa: E402
from parity_utils import CASES, GBUFFER_CASES, REFSTREAM_CASES, SCREEN_CASES, STAT_CASES, render_case, render_gbuffer, render_screen_gradient, render_stat_case  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = ref_loader.load()
    dev = torch.device("cpu")
    only = set(sys.argv[1:])  # optional: regenerate only the named cases
    for name, cfg in list(CASES.items()) + list(REFSTREAM_CASES.items()):
        if only and name not in only:
            continue
        img, grads = render_case(ref, dev, cfg, cfg["seed"])
        arrs = {"image": img.numpy()}
        for k, v in grads.items():
            arrs["grad." + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(name, "image mean %.6f" % img.mean().item(), {k: float(v.norm()) for k, v in grads.items()})
    for name, cfg in GBUFFER_CASES.items():
        if only and name not in only:
            continue
        img = render_gbuffer(ref, dev, cfg)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img.numpy())
        print(name, tuple(img.shape), "mean %.6f" % img.mean().item())
    for name, cfg in SCREEN_CASES.items():
        if only and name not in only:
            continue
        img = render_screen_gradient(ref, dev, cfg)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img.numpy())
        print(name, tuple(img.shape), "norm %.6f" % img.norm().item())
    if not only or "c2_full_size_forward_blocks" in only:
        # the headline configuration itself (C2, 512 x 512 x 64 spp, forward): 8 x 8 block means of the reference's image
        cfg = dict(scene="shadow_blocker", res=512, spp=64, mb=1, sampler="sobol", edges=0)
        img, _ = render_case(ref, dev, cfg, 1, backward=False)
        np.savez_compressed(os.path.join(OUT, "c2_full_size_forward_blocks.npz"), blocks=img.numpy().reshape(64, 8, 64, 8, 3).mean((1, 3)))
        print("c2_full_size_forward_blocks", "mean %.6f" % img.mean().item())
    for name, cfg in STAT_CASES.items():
        if only and name not in only:
            continue
        acc = render_stat_case(ref, dev, name)
        arrs = {}
        for k, lst in acc.items():
            a = np.stack(lst)
            arrs["mean." + k] = a.mean(0)
            arrs["sem." + k] = a.std(0, ddof=1) / np.sqrt(len(lst))
            if cfg.get("test") == "ranks":  # per-seed values for the two-sample rank test (heavy-tailed estimators)
                arrs["samples." + k] = a
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(name, {k: float(np.linalg.norm(v)) for k, v in arrs.items()})


if __name__ == "__main__":
    main()
