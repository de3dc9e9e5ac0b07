#!/usr/bin/env python
"""Development aid (GPU box): render the same seeded scene with the compiled reference (oracle/_ref, CPU/Embree) and
with the CUDA path, and print relative-L2 differences of the image and of every gradient.

    python tools/compare.py [scene] [--res N] [--spp N] [--mb N] [--edges 0|1] [--sampler sobol|independent]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import ref_loader  # noqa: E402
import scenes  # noqa: E402
from redner_b200 import api  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d


def grads_of(scene):
    import parity_utils
    return {k: v.clone() for k, v in parity_utils.collect_grads(scene).items()}


def run(backend, device, name, res, spp, mb, edges, sampler, seed, do_backward=True, **kw):
    sc = scenes.SCENES[name](device, resolution=(res, res), **kw)
    st = backend.SamplerType.sobol if sampler == "sobol" else backend.SamplerType.independent
    args = api.RenderFunction.serialize_scene(sc, spp, mb, sampler_type=st, device=device, backend=backend,
                                              use_primary_edge_sampling=bool(edges & 1), use_secondary_edge_sampling=bool(edges & 2))
    t0 = time.time()
    img = api.RenderFunction.apply(seed, *args)
    if device.type == "cuda":
        torch.cuda.synchronize()
    t1 = time.time()
    g = {}
    if do_backward and img.requires_grad:
        img.pow(2).sum().backward()
        if device.type == "cuda":
            torch.cuda.synchronize()
        g = grads_of(sc)
    t2 = time.time()
    return img.detach(), g, (t1 - t0, t2 - t1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", nargs="?", default="single_triangle")
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--mb", type=int, default=1)
    ap.add_argument("--edges", type=int, default=0, help="bit 0: primary, bit 1: secondary")
    ap.add_argument("--sampler", default="sobol")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--lib", default=None, help="alternative libredner_b200 (e.g. the f64 validation build)")
    ap.add_argument("--emu", action="store_true", help="debug only: run the host-compiled emulator (tools/cpu_emu) instead of the GPU")
    ap.add_argument("--edges-ref", type=int, default=None, help="edge-sampling flags for the reference run (default: same)")
    a = ap.parse_args()
    from redner_b200 import _lib
    dev = torch.device("cuda:0")
    if a.emu:
        _lib._lib = _lib._bind(__import__("ctypes").CDLL(os.path.join(ROOT, "tools", "cpu_emu", "libredner_b200_emu.so")))
        dev = torch.device("cpu")
    elif a.lib:
        _lib._lib = _lib.load(a.lib)
    from redner_b200 import redner as rb
    ref = ref_loader.load()
    img_r, g_r, t_r = run(ref, torch.device("cpu"), a.scene, a.res, a.spp, a.mb, a.edges if a.edges_ref is None else a.edges_ref, a.sampler, a.seed)
    img_c, g_c, t_c = run(rb, dev, a.scene, a.res, a.spp, a.mb, a.edges, a.sampler, a.seed)
    print("scene=%s res=%d spp=%d mb=%d edges=%d sampler=%s" % (a.scene, a.res, a.spp, a.mb, a.edges, a.sampler))
    print("time ref fwd %.3fs bwd %.3fs | cuda fwd %.3fs bwd %.3fs (incl. scene build, first call)" % (t_r + t_c))
    print("image  relL2 = %.3e   (mean ref %.5f, mean cuda %.5f)" % (rel(img_c, img_r), img_r.mean().item(), img_c.cpu().mean().item()))
    diff = (img_c.cpu() - img_r).abs().sum(-1)
    nbad = int((diff > 1e-3 * max(1e-6, img_r.abs().max().item())).sum())
    print("pixels differing by > 1e-3 of max: %d of %d" % (nbad, diff.numel()))
    for k in sorted(g_r):
        if k in g_c:
            print("grad %-28s relL2 = %.3e   |ref| = %.4e" % (k, rel(g_c[k], g_r[k]), g_r[k].norm().item()))
        else:
            print("grad %-28s MISSING on cuda" % k)


if __name__ == "__main__":
    main()
