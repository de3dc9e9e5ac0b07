"""Helper of the drop-in tests (run as a subprocess): the UNMODIFIED pyredner package renders and differentiates one scene through
its own serialize / unpack / forward / backward on a chosen native `redner` module and saves every number.

usage: python tests/dropin_script.py <native> <out.npz> <dir containing pyredner/> <cpu|cuda>
  native = "reference"  the reference's own pybind module (oracle/_ref, CPU / Embree)
           "cuda"       redner_b200/dropin/redner.py on libredner_b200.so (the product, needs a GPU)
           <path.so>    redner_b200/dropin/redner.py bound to the host build of the device headers (tools/cpu_emu, CPU suite only)
"""
import ctypes
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
native, out, pyredner_dir, devname = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
sys.path.insert(0, pyredner_dir)
sys.path.insert(0, ROOT)
for name in ("skimage", "skimage.io", "skimage.transform", "imageio"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["skimage"].io = sys.modules["skimage.io"]
sys.modules["skimage"].transform = sys.modules["skimage.transform"]
if native == "reference":      # the reference's own pybind module
    d = os.path.join(ROOT, "oracle", "_ref")
    for lib in ("libtbbmalloc.so.2", "libtbb.so.2", "libembree3.so.3"):
        ctypes.CDLL(os.path.join(d, lib), mode=ctypes.RTLD_GLOBAL)
    sys.path.insert(0, d)
else:                          # our drop-in module: the product library, or the host build of the device headers (this process only)
    if native != "cuda":
        from redner_b200 import _lib
        _lib._lib = _lib._bind(ctypes.CDLL(native))
    sys.path.insert(0, os.path.join(ROOT, "redner_b200", "dropin"))
import numpy as np, torch, redner, pyredner
pyredner.set_use_gpu(devname == "cuda")
pyredner.set_print_timing(False)
dev = pyredner.get_device()
_tensor = torch.tensor


def T(data, requires_grad=False, dtype=None):  # scene tensors live on the render device, camera tensors on the host
    return _tensor(data, dtype=dtype, device=dev, requires_grad=requires_grad)


g = torch.Generator().manual_seed(7)
cam = pyredner.Camera(position=torch.tensor([0.1, 1.2, -4.0], requires_grad=True), look_at=torch.tensor([0.0, 0.5, 0.0], requires_grad=True),
                      up=torch.tensor([0.0, 1.0, 0.0], requires_grad=True), fov=torch.tensor([45.0]), clip_near=1e-2, resolution=(24, 28))
tex = (0.2 + 0.6 * torch.rand(8, 8, 3, generator=g)).to(dev).requires_grad_(True)
m_floor = pyredner.Material(diffuse_reflectance=pyredner.Texture(tex, uv_scale=T([2.0, 2.0])))
m_tri = pyredner.Material(diffuse_reflectance=T([0.4, 0.5, 0.3], requires_grad=True), specular_reflectance=T([0.3, 0.3, 0.3], requires_grad=True),
                          roughness=T([0.2], requires_grad=True))
m_l = pyredner.Material(diffuse_reflectance=T([0.0, 0.0, 0.0]))
floor = pyredner.Shape(vertices=T([[-2.0, 0.0, -2.0], [-2.0, 0.0, 2.0], [2.0, 0.0, -2.0], [2.0, 0.0, 2.0]]), indices=T([[0, 1, 2], [1, 3, 2]], dtype=torch.int32),
                       uvs=T([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]]), normals=None, material_id=0)
tri_v = T([[-0.8, 0.3, 0.2], [0.7, 0.4, -0.1], [0.0, 1.6, 0.3]], requires_grad=True)
tri = pyredner.Shape(vertices=tri_v, indices=T([[0, 1, 2]], dtype=torch.int32), uvs=None, normals=None, material_id=1)
lamp = pyredner.Shape(vertices=T([[-0.5, 2.8, -0.5], [-0.5, 2.8, 0.5], [0.5, 2.8, -0.5], [0.5, 2.8, 0.5]]), indices=T([[0, 2, 1], [1, 2, 3]], dtype=torch.int32),
                      uvs=None, normals=None, material_id=2)
inten = torch.tensor([20.0, 19.0, 18.0], requires_grad=True)
scene = pyredner.Scene(cam, [floor, tri, lamp], [m_floor, m_tri, m_l], [pyredner.AreaLight(shape_id=2, intensity=inten)])
res = {}
# pass 1: textured floor, two channels, interior derivatives only (every number is sample-exact)
args = pyredner.RenderFunction.serialize_scene(scene=scene, num_samples=4, max_bounces=2, sampler_type=redner.SamplerType.sobol,
                                               channels=[redner.channels.radiance, redner.channels.depth],
                                               use_primary_edge_sampling=False, use_secondary_edge_sampling=False)
img = pyredner.RenderFunction.apply(3, *args)
(img * T([1.0, 1.0, 1.0, 0.2])).pow(2).sum().backward()
res.update(image=img.detach().cpu().numpy(), position=cam.position.grad.cpu().numpy().copy(), look_at=cam.look_at.grad.cpu().numpy().copy(), up=cam.up.grad.cpu().numpy().copy(),
           tex=tex.grad.cpu().numpy().copy(), kd=m_tri.diffuse_reflectance.texels.grad.cpu().numpy().copy(), ks=m_tri.specular_reflectance.texels.grad.cpu().numpy().copy(),
           ro=m_tri.roughness.texels.grad.cpu().numpy().copy(), tri=tri_v.grad.cpu().numpy().copy(), inten=inten.grad.cpu().numpy().copy())
# pass 2: primary edge sampling (silhouette derivatives) on a one-colour floor -- what an edge ray sees must not depend on the
# filter footprint for the comparison to be sample-exact
for t in (cam.position, cam.look_at, cam.up, tri_v):
    t.grad = None
scene.materials[0] = pyredner.Material(diffuse_reflectance=T([0.5, 0.45, 0.4]))
args = pyredner.RenderFunction.serialize_scene(scene=scene, num_samples=4, max_bounces=1, sampler_type=redner.SamplerType.sobol,
                                               use_primary_edge_sampling=True, use_secondary_edge_sampling=False)
img2 = pyredner.RenderFunction.apply(5, *args)
img2.pow(2).sum().backward()
res.update(edge_image=img2.detach().cpu().numpy(), edge_position=cam.position.grad.cpu().numpy(), edge_look_at=cam.look_at.grad.cpu().numpy(), edge_up=cam.up.grad.cpu().numpy(),
           edge_tri=tri_v.grad.cpu().numpy())
# pass 3: the callers either side of the path, untouched: deferred shading of a G-buffer and a
# BATCH of two scenes through pyredner.render_g_buffer (pyredner/render_utils.py:104-313, :431-503)
for t in (cam.position, cam.look_at, cam.up, tri_v):
    t.grad = None
dl = [pyredner.PointLight(position=T([0.5, 2.5, -1.0]), intensity=T([8.0, 8.0, 8.0])),
      pyredner.AmbientLight(intensity=T([0.1, 0.1, 0.1]))]
img3 = pyredner.render_deferred(scene, lights=dl, aa_samples=2, seed=11, device=dev)
img3.pow(2).sum().backward()
res.update(deferred_image=img3.detach().cpu().numpy(), deferred_position=cam.position.grad.cpu().numpy().copy(), deferred_tri=tri_v.grad.cpu().numpy().copy())
cam2 = pyredner.Camera(position=torch.tensor([-0.6, 1.0, -3.5]), look_at=torch.tensor([0.0, 0.5, 0.0]), up=torch.tensor([0.0, 1.0, 0.0]), fov=torch.tensor([50.0]),
                       clip_near=1e-2, resolution=(24, 28))
scene2 = pyredner.Scene(cam2, scene.shapes, scene.materials, scene.area_lights)
with torch.no_grad():
    gb = pyredner.render_g_buffer([scene, scene2], channels=[pyredner.channels.position, pyredner.channels.shading_normal, pyredner.channels.diffuse_reflectance],
                                  num_samples=(2, 2), seed=[13, 14], device=dev)
res.update(batch_gbuffer_image=gb.cpu().numpy())
# pass 4: a short inverse-rendering loop in the style of tests/test_single_triangle.py: move the triangle towards a target
# image with Adam, primary-edge (silhouette) gradients driving it; the loss curves must coincide
with torch.no_grad():
    target_v = tri_v + T([[0.15, -0.1, 0.0], [-0.1, 0.1, 0.05], [0.05, -0.15, 0.0]])
scene.shapes[1].vertices = target_v
args = pyredner.RenderFunction.serialize_scene(scene=scene, num_samples=4, max_bounces=1, sampler_type=redner.SamplerType.sobol, use_secondary_edge_sampling=False)
target = pyredner.RenderFunction.apply(1, *args).detach()
v = tri_v.detach().clone().requires_grad_(True)
scene.shapes[1].vertices = v
opt = torch.optim.Adam([v], lr=2e-2)
losses = []
for it in range(8):
    opt.zero_grad()
    args = pyredner.RenderFunction.serialize_scene(scene=scene, num_samples=4, max_bounces=1, sampler_type=redner.SamplerType.sobol, use_secondary_edge_sampling=False)
    loss = (pyredner.RenderFunction.apply(it + 2, *args) - target).pow(2).sum()
    loss.backward()
    opt.step()
    losses.append(float(loss))
res.update(opt_losses=np.array(losses), opt_vertices=v.detach().cpu().numpy())
np.savez(out, **res)
print("DONE")
