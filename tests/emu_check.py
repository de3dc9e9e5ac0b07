"""Helper of tests/test_device_code_cpu.py (run as a subprocess, never imported by the product).

Binds the Python shim to the host-compiled build of the per-sample device headers (tools/cpu_emu: the same
rb_*.cuh sources the sm_100a kernels are made of, compiled with g++ and driven by plain loops) and checks the named
golden cases with the tolerances of the GPU suite.  This verifies the MATH of the device code on a machine without a GPU;
it says nothing about the kernels' launch structure, compaction, sorting or atomics, which only `-m gpu` covers.

usage: python tests/emu_check.py <emulator.so> <case> [<case> ...]
"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def check_batch_of_views(rb, dev):
    """api.render_batch (one native scene re-targeted per view with rb_scene_set_camera) == one full Scene per view: images bit for bit,
    gradients to 1e-5 -- the host logic of SURVEY.md section 8f rank 2, on the host build of the device headers."""
    import torch
    import parity_utils as pu
    import scenes
    from redner_b200 import api

    def make():
        views = [scenes.glossy_room(dev, resolution=(24, 24)) for _ in range(3)]
        cams = [([0.3, 1.4, -4.5], [0.0, 0.6, 0.0]), ([1.2, 1.1, -4.0], [0.1, 0.5, 0.1]), ([-0.8, 1.8, -4.2], [0.0, 0.7, 0.2])]
        for v, (p, l) in zip(views, cams):
            v.camera = api.Camera(position=torch.tensor(p, requires_grad=True), look_at=torch.tensor(l, requires_grad=True),
                                  up=torch.tensor([0.0, 1.0, 0.0], requires_grad=True), fov=torch.tensor([40.0]), clip_near=1e-2, resolution=(24, 24))
            v.shapes, v.materials, v.area_lights = views[0].shapes, views[0].materials, views[0].area_lights
        return views
    kw = dict(sampler_type=rb.SamplerType.sobol, device=dev, backend=rb)
    views = make()
    imgs = api.render_batch(views, 4, 2, [11, 12, 13], **kw)
    imgs.pow(2).sum().backward()
    g_batch = pu.collect_grads(views[0])
    cam_batch = [v.camera.position.grad.clone() for v in views]
    views = make()
    singles = [api.RenderFunction.apply(11 + k, *api.RenderFunction.serialize_scene(v, 4, 2, **kw)) for k, v in enumerate(views)]
    sum(s.pow(2).sum() for s in singles).backward()
    g_single = pu.collect_grads(views[0])
    for k in range(3):
        assert torch.equal(imgs[k], singles[k]), k
        assert pu.rel_l2(cam_batch[k].numpy(), views[k].camera.position.grad.numpy()) < 1e-5
    for key in g_single:
        if not key.startswith("cam."):
            assert pu.rel_l2(g_batch[key].numpy(), g_single[key].numpy()) < 1e-5, key


def main():
    so, names = sys.argv[1], sys.argv[2:]
    import torch
    from redner_b200 import _lib
    _lib._lib = _lib._bind(ctypes.CDLL(so))  # this process only: the emulator exports the same C ABI with host pointers
    from redner_b200 import redner as rb
    import parity_utils as pu
    dev = torch.device("cpu")
    for name in names:
        if name == "batch_of_views":
            check_batch_of_views(rb, dev)
            print("ok", name, flush=True)
            continue
        if name in pu.STAT_CASES:
            pu.assert_stat_matches_golden(name, pu.render_stat_case(rb, dev, name))
        elif name in pu.SCREEN_CASES:
            pu.assert_screen_gradient_matches_golden(name, pu.render_screen_gradient(rb, dev, pu.SCREEN_CASES[name]).numpy())
        elif name in pu.GBUFFER_CASES:
            pu.assert_gbuffer_matches_golden(name, pu.render_gbuffer(rb, dev, pu.GBUFFER_CASES[name]).numpy())
        else:
            cfg = pu.CASES[name] if name in pu.CASES else pu.REFSTREAM_CASES[name]
            img, grads = pu.render_case(rb, dev, cfg, cfg["seed"])
            pu.assert_matches_golden(name, img.numpy(), grads)
        print("ok", name, flush=True)


if __name__ == "__main__":
    main()
