"""Seeded synthetic scenes shared by the parity tests, the golden-vector generator and bench.py.

Geometry follows the reference's own test scripts:
  C1  tests/test_single_triangle.py:17-90      one grey triangle + quad light, 256x256, max_bounces 1
  C2  tests/test_shadow_blocker.py:11-59       floor + blocker + small light (geometry), 512x512, max_bounces 1
plus small procedural scenes (`glossy_room`, `sphere_box`) that exercise the specular lobe, shading normals, uv /
textures and multi-bounce paths without needing mesh files.
"""
import math
import os

import numpy as np
import torch

from redner_b200 import api


def _t(x, device, dtype=torch.float32, grad=False):
    t = torch.tensor(x, dtype=dtype, device=device)
    if grad:
        t.requires_grad_(True)
    return t


def single_triangle(device, resolution=(256, 256), grad=True):
    cam = api.Camera(position=torch.tensor([0.0, 0.0, -5.0]), look_at=torch.tensor([0.0, 0.0, 0.0]), up=torch.tensor([0.0, 1.0, 0.0]),
                     fov=torch.tensor([45.0]), clip_near=1e-2, resolution=resolution)
    if grad:
        cam.position.requires_grad_(True)
    mat_grey = api.Material(diffuse_reflectance=_t([0.5, 0.5, 0.5], device, grad=grad))
    tri = api.Shape(_t([[-2.0, 1.5, 0.3], [0.9, 1.2, -0.3], [-0.4, -1.4, 0.2]], device, grad=grad), _t([[0, 1, 2]], device, torch.int32), 0)
    light = api.Shape(_t([[-1.0, -1.0, -7.0], [1.0, -1.0, -7.0], [-1.0, 1.0, -7.0], [1.0, 1.0, -7.0]], device),
                      _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 0)
    al = api.AreaLight(1, torch.tensor([20.0, 20.0, 20.0], requires_grad=grad))
    return api.Scene(cam, [tri, light], [mat_grey], [al])


def shadow_blocker(device, resolution=(512, 512), grad=True, grad_all=False):
    """C2 geometry (tests/test_shadow_blocker.py).  `grad_all`: floor and lamp vertices differentiable too (the lamp's
    boundary is where the two boundary-sampling strategies of the reference meet)."""
    cam = api.Camera(position=torch.tensor([0.0, 2.0, -5.0]), look_at=torch.tensor([0.0, 0.0, 0.0]), up=torch.tensor([0.0, 1.0, 0.0]),
                     fov=torch.tensor([45.0]), clip_near=1e-2, resolution=resolution)
    mat_grey = api.Material(diffuse_reflectance=_t([0.5, 0.5, 0.5], device, grad=grad))
    mat_black = api.Material(diffuse_reflectance=_t([0.0, 0.0, 0.0], device))
    floor = api.Shape(_t([[-2.0, 0.0, -2.0], [-2.0, 0.0, 2.0], [2.0, 0.0, -2.0], [2.0, 0.0, 2.0]], device, grad=grad and grad_all),
                      _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 0)
    blocker = api.Shape(_t([[-0.2, 3.5, -0.8], [-0.8, 3.0, 0.3], [0.4, 2.8, -0.8], [0.3, 3.2, 1.0]], device, grad=grad),
                        _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 0)
    light = api.Shape(_t([[-0.1, 5, -0.1], [-0.1, 5, 0.1], [0.1, 5, -0.1], [0.1, 5, 0.1]], device, grad=grad and grad_all),
                      _t([[0, 2, 1], [1, 2, 3]], device, torch.int32), 1)
    al = api.AreaLight(2, torch.tensor([1000.0, 1000.0, 1000.0], requires_grad=grad))
    return api.Scene(cam, [floor, blocker, light], [mat_grey, mat_black], [al])


def uv_sphere(device, radius, center, n_theta=12, n_phi=24, grad=False):
    """Closed sphere with per-vertex normals and uvs (exercises shading normals + uv derivatives)."""
    verts, uvs, normals, idx = [], [], [], []
    for i in range(n_theta + 1):
        th = math.pi * i / n_theta
        for j in range(n_phi + 1):
            ph = 2 * math.pi * j / n_phi
            n = (math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph))
            verts.append([center[0] + radius * n[0], center[1] + radius * n[1], center[2] + radius * n[2]])
            normals.append(list(n))
            uvs.append([j / n_phi, i / n_theta])
    for i in range(n_theta):
        for j in range(n_phi):
            a = i * (n_phi + 1) + j
            b = a + n_phi + 1
            if i != 0:
                idx.append([a, a + 1, b])
            if i != n_theta - 1:
                idx.append([a + 1, b + 1, b])
    return (_t(verts, device, grad=grad), _t(idx, device, torch.int32), _t(uvs, device), _t(normals, device))


def glossy_room(device, resolution=(128, 128), grad=True, textured=True, nmap=False, sphere_res=(12, 24), camera_type=0, distortion=False):
    """Open box (floor, back wall, side wall) with a glossy textured floor, a Phong-shaded sphere and two area lights."""
    g = torch.Generator().manual_seed(7)
    cam = api.Camera(position=torch.tensor([0.3, 1.4, -4.5]), look_at=torch.tensor([0.0, 0.6, 0.0]), up=torch.tensor([0.0, 1.0, 0.0]),
                     fov=torch.tensor([40.0]), clip_near=1e-2, resolution=resolution)
    if distortion:  # Brown-Conrady lens model with differentiable parameters and pose
        cam = api.Camera(position=torch.tensor([0.3, 1.4, -4.5], requires_grad=grad), look_at=torch.tensor([0.0, 0.6, 0.0], requires_grad=grad),
                         up=torch.tensor([0.0, 1.0, 0.0], requires_grad=grad), fov=torch.tensor([40.0]), clip_near=1e-2, resolution=resolution,
                         distortion_params=torch.tensor([0.12, -0.05, 0.01, 0.02, 0.01, -0.004, 0.01, -0.015], requires_grad=grad))
    if camera_type == 1:  # orthographic: same pose, the intrinsic matrix scales the film to the room
        cam = api.Camera(position=torch.tensor([0.3, 1.4, -4.5], requires_grad=grad), look_at=torch.tensor([0.0, 0.6, 0.0], requires_grad=grad),
                         up=torch.tensor([0.0, 1.0, 0.0], requires_grad=grad), clip_near=1e-2, resolution=resolution,
                         intrinsic_mat=torch.tensor([[0.4, 0.0, 0.0], [0.0, 0.4, 0.0], [0.0, 0.0, 1.0]]), camera_type=1)
    elif camera_type != 0:  # fisheye / panorama: from inside the room, differentiable pose
        cam = api.Camera(position=torch.tensor([0.4, 1.2, -1.6], requires_grad=grad), look_at=torch.tensor([0.1, 0.7, 0.2], requires_grad=grad),
                         up=torch.tensor([0.0, 1.0, 0.0], requires_grad=grad), fov=torch.tensor([40.0]), clip_near=1e-2, resolution=resolution,
                         camera_type=camera_type)
    if textured:
        tex = (0.2 + 0.6 * torch.rand(16, 16, 3, generator=g)).to(device)
        rough = (0.05 + 0.3 * torch.rand(16, 16, 1, generator=g)).to(device)
    else:
        tex = torch.tensor([0.45, 0.4, 0.35], device=device)
        rough = torch.tensor([0.15], device=device)
    if grad:
        tex.requires_grad_(True)
        rough.requires_grad_(True)
    m_floor = api.Material(diffuse_reflectance=api.Texture(tex, torch.tensor([2.0, 2.0], device=device)),
                           specular_reflectance=_t([0.3, 0.3, 0.3], device, grad=grad), roughness=api.Texture(rough, torch.tensor([2.0, 2.0], device=device)))
    m_wall = api.Material(diffuse_reflectance=_t([0.6, 0.3, 0.25], device, grad=grad), two_sided=True)
    if nmap:
        # bumpy ball: mip-mapped normal map + mip-mapped specular texture with a differentiable uv_scale
        nm = torch.tensor([0.5, 0.5, 1.0]) + 0.25 * (torch.rand(8, 8, 3, generator=g) - 0.5)
        nm = nm.to(device).requires_grad_(grad)
        spec = (0.2 + 0.5 * torch.rand(8, 8, 3, generator=g)).to(device).requires_grad_(grad)
        m_ball = api.Material(diffuse_reflectance=_t([0.2, 0.35, 0.6], device, grad=grad),
                              specular_reflectance=api.Texture(spec, torch.tensor([3.0, 2.0], device=device, requires_grad=grad)),
                              roughness=_t([0.2], device, grad=grad), normal_map=api.Texture(nm, torch.tensor([2.0, 2.0], device=device)))
    else:
        m_ball = api.Material(diffuse_reflectance=_t([0.2, 0.35, 0.6], device, grad=grad), specular_reflectance=_t([0.5, 0.5, 0.5], device, grad=grad),
                              roughness=_t([0.2], device, grad=grad))
    m_light = api.Material(diffuse_reflectance=_t([0.0, 0.0, 0.0], device))
    floor = api.Shape(_t([[-2.5, 0.0, -2.5], [-2.5, 0.0, 2.5], [2.5, 0.0, -2.5], [2.5, 0.0, 2.5]], device),
                      _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 0, uvs=_t([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]], device))
    back = api.Shape(_t([[-2.5, 0.0, 2.5], [-2.5, 3.0, 2.5], [2.5, 0.0, 2.5], [2.5, 3.0, 2.5]], device),
                     _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 1)
    side = api.Shape(_t([[-2.5, 0.0, -2.5], [-2.5, 3.0, -2.5], [-2.5, 0.0, 2.5], [-2.5, 3.0, 2.5]], device),
                     _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 1)
    v, i, uv, n = uv_sphere(device, 0.7, (0.2, 0.7, 0.3), n_theta=sphere_res[0], n_phi=sphere_res[1], grad=grad)
    ball = api.Shape(v, i, 2, uvs=uv, normals=n)
    l1 = api.Shape(_t([[-0.6, 2.9, -0.6], [-0.6, 2.9, 0.6], [0.6, 2.9, -0.6], [0.6, 2.9, 0.6]], device),
                   _t([[0, 2, 1], [1, 2, 3]], device, torch.int32), 3)
    l2 = api.Shape(_t([[1.8, 0.8, -1.5], [1.8, 1.6, -1.5], [2.2, 0.8, -0.9]], device), _t([[0, 1, 2]], device, torch.int32), 3)
    lights = [api.AreaLight(4, torch.tensor([25.0, 24.0, 22.0], requires_grad=grad)),
              api.AreaLight(5, torch.tensor([8.0, 10.0, 14.0], requires_grad=grad), two_sided=True)]
    return api.Scene(cam, [floor, back, side, ball, l1, l2], [m_floor, m_wall, m_ball, m_light], lights)


def random_soup(device, num_tris=2000, resolution=(256, 256), seed=3, grad=False):
    """Random triangle soup under one light: BVH / traversal stress (many shapes' worth of edges, deep tree)."""
    g = torch.Generator().manual_seed(seed)
    c = (torch.rand(num_tris, 1, 3, generator=g) - 0.5) * torch.tensor([4.0, 2.5, 3.0]) + torch.tensor([0.0, 1.3, 0.5])
    v = (c + 0.25 * (torch.rand(num_tris, 3, 3, generator=g) - 0.5)).reshape(-1, 3).contiguous()
    i = torch.arange(3 * num_tris, dtype=torch.int32).reshape(-1, 3).contiguous()
    cam = api.Camera(position=torch.tensor([0.0, 1.5, -6.0]), look_at=torch.tensor([0.0, 1.0, 0.0]), up=torch.tensor([0.0, 1.0, 0.0]),
                     fov=torch.tensor([45.0]), clip_near=1e-2, resolution=resolution)
    m = api.Material(diffuse_reflectance=_t([0.6, 0.6, 0.6], device), two_sided=True)
    m_l = api.Material(diffuse_reflectance=_t([0.0, 0.0, 0.0], device))
    soup = api.Shape(v.to(device), i.to(device), 0)
    floor = api.Shape(_t([[-4.0, 0.0, -4.0], [-4.0, 0.0, 4.0], [4.0, 0.0, -4.0], [4.0, 0.0, 4.0]], device), _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 0)
    light = api.Shape(_t([[-1.0, 4.5, -1.0], [-1.0, 4.5, 1.0], [1.0, 4.5, -1.0], [1.0, 4.5, 1.0]], device), _t([[0, 2, 1], [1, 2, 3]], device, torch.int32), 1)
    return api.Scene(cam, [soup, floor, light], [m, m_l], [api.AreaLight(2, torch.tensor([30.0, 30.0, 30.0]))])


def env_ball(device, resolution=(64, 64), grad=True, constant_sky=False, camera_type=0, cam_grad=False):
    """A glossy ball and a textured floor under an environment map (plus one small area light, so that light selection
    mixes both kinds); the camera sees the sky directly.  `cam_grad`: differentiable pose (the sky's gradient w.r.t. the
    camera flows through the primary rays that leave the scene)."""
    g = torch.Generator().manual_seed(11)
    pose = [[0.2, 1.1, -4.0], [0.0, 0.6, 0.0]] if camera_type == 0 else [[0.1, 0.9, -1.6], [0.0, 0.7, 0.0]]
    cam = api.Camera(position=torch.tensor(pose[0], requires_grad=cam_grad), look_at=torch.tensor(pose[1], requires_grad=cam_grad),
                     up=torch.tensor([0.0, 1.0, 0.0], requires_grad=cam_grad), fov=torch.tensor([45.0]), clip_near=1e-2, resolution=resolution,
                     camera_type=camera_type)
    sky = (0.2 + 1.5 * torch.rand(16, 32, 3, generator=g))
    if constant_sky:  # same image size, one colour: no dependence on the mip level
        sky = torch.ones(16, 32, 3) * torch.tensor([0.6, 0.7, 0.9])
    sky = sky.to(device).requires_grad_(grad)
    a = 0.4
    e2w = torch.tensor([[math.cos(a), 0.0, math.sin(a), 0.0], [0.0, 1.0, 0.0, 0.0], [-math.sin(a), 0.0, math.cos(a), 0.0], [0.0, 0.0, 0.0, 1.0]],
                       requires_grad=grad)
    env = api.EnvironmentMap(sky, e2w)
    tex = (0.2 + 0.6 * torch.rand(8, 8, 3, generator=g)).to(device).requires_grad_(grad)
    m_floor = api.Material(diffuse_reflectance=api.Texture(tex, torch.tensor([2.0, 2.0], device=device)))
    m_ball = api.Material(diffuse_reflectance=_t([0.3, 0.25, 0.5], device, grad=grad), specular_reflectance=_t([0.4, 0.4, 0.4], device, grad=grad),
                          roughness=_t([0.25], device, grad=grad))
    m_light = api.Material(diffuse_reflectance=_t([0.0, 0.0, 0.0], device))
    floor = api.Shape(_t([[-2.0, 0.0, -2.0], [-2.0, 0.0, 2.0], [2.0, 0.0, -2.0], [2.0, 0.0, 2.0]], device),
                      _t([[0, 1, 2], [1, 3, 2]], device, torch.int32), 0, uvs=_t([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]], device))
    v, i, uv, n = uv_sphere(device, 0.6, (0.1, 0.6, 0.2), grad=grad)
    ball = api.Shape(v, i, 1, uvs=uv, normals=n)
    lamp = api.Shape(_t([[-0.4, 2.4, -0.4], [-0.4, 2.4, 0.4], [0.4, 2.4, -0.4], [0.4, 2.4, 0.4]], device), _t([[0, 2, 1], [1, 2, 3]], device, torch.int32), 2)
    lights = [api.AreaLight(2, torch.tensor([6.0, 6.0, 5.0], requires_grad=grad))]
    return api.Scene(cam, [floor, ball, lamp], [m_floor, m_ball, m_light], lights, envmap=env)


def corner_ball(device, resolution=(32, 36), variant="plain", grad=True):
    """Small scene that exercises the less travelled inputs of the boundary: separate uv / normal index buffers, vertex
    colours (`vcolor`), a viewport crop (`viewport`), a generic texture (`generic`), a light the camera cannot see
    (`invisible`), two-sided materials and lights, a differentiable camera pose; non-square image."""
    g = torch.Generator().manual_seed(5)

    def T(x, gr=False, dt=torch.float32):
        t = torch.tensor(x, dtype=dt)
        return t.to(device).requires_grad_(True) if gr and grad else t.to(device)

    def C(x):  # camera parameters live on the host
        return torch.tensor(x, dtype=torch.float32, requires_grad=grad)
    vp = (4, 6, 28, 30) if variant == "viewport" else None
    cam = api.Camera(position=C([0.2, 1.3, -4.0]), look_at=C([0.0, 0.6, 0.0]), up=C([0.0, 1.0, 0.0]), fov=torch.tensor([45.0]), clip_near=1e-2,
                     resolution=resolution, viewport=vp)
    v, i, uv, n = uv_sphere(device, 0.7, (0.1, 0.7, 0.2), grad=grad)
    cols = torch.rand(v.shape[0], 3, generator=g).to(device).requires_grad_(grad)
    perm = torch.randperm(uv.shape[0], generator=g)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(uv.shape[0])
    uv2, n2 = uv.detach().cpu()[perm].to(device).requires_grad_(grad), n.detach().cpu()[perm].to(device).requires_grad_(grad)
    idx2 = inv[i.cpu().long()].int().to(device)
    gen = torch.rand(8, 8, 5, generator=g).to(device).requires_grad_(grad)
    m_ball = api.Material(diffuse_reflectance=T([0.4, 0.4, 0.4]), use_vertex_color=(variant == "vcolor"),
                          generic_texture=api.Texture(gen) if variant == "generic" else None, specular_reflectance=T([0.3, 0.3, 0.3], True), roughness=T([0.3], True))
    m_floor = api.Material(diffuse_reflectance=T([0.5, 0.45, 0.4], True), two_sided=True)
    m_l = api.Material(diffuse_reflectance=T([0.0, 0.0, 0.0]))
    ball = api.Shape(v, i, 0, uvs=uv2, normals=n2, uv_indices=idx2, normal_indices=idx2.clone(), colors=cols)
    floor = api.Shape(T([[-2.5, 0.0, -2.5], [-2.5, 0.0, 2.5], [2.5, 0.0, -2.5], [2.5, 0.0, 2.5]], True), T([[0, 1, 2], [1, 3, 2]], dt=torch.int32), 1)
    l1 = api.Shape(T([[-0.6, 2.9, -0.6], [-0.6, 2.9, 0.6], [0.6, 2.9, -0.6], [0.6, 2.9, 0.6]]), T([[0, 2, 1], [1, 2, 3]], dt=torch.int32), 2)
    l2 = api.Shape(T([[1.8, 0.8, -1.5], [1.8, 1.6, -1.5], [2.2, 0.8, -0.9]]), T([[0, 1, 2]], dt=torch.int32), 2)
    lights = [api.AreaLight(2, torch.tensor([20.0, 19.0, 18.0], requires_grad=grad)),
              api.AreaLight(3, torch.tensor([8.0, 10.0, 14.0], requires_grad=grad), two_sided=True, directly_visible=(variant != "invisible"))]
    return api.Scene(cam, [ball, floor, l1, l2], [m_ball, m_floor, m_l], lights)


def hires_room(device, **kw):
    """glossy_room with a 32 k-triangle ball (48 k edges): scene-build and traversal cost at the BASELINE C3 / C4 scale."""
    return glossy_room(device, sphere_res=(90, 180), **kw)


def ortho_room(device, **kw):
    return glossy_room(device, camera_type=1, **kw)


def distort_room(device, **kw):
    return glossy_room(device, distortion=True, **kw)


def fisheye_room(device, **kw):
    return glossy_room(device, camera_type=2, **kw)


def panorama_room(device, **kw):
    return glossy_room(device, camera_type=3, **kw)


def env_ball_flat_sky(device, **kw):
    return env_ball(device, constant_sky=True, **kw)


def shadow_blocker_all(device, **kw):
    return shadow_blocker(device, grad_all=True, **kw)


def env_ball_fisheye(device, **kw):
    return env_ball(device, camera_type=2, cam_grad=True, **kw)


def nmap_room(device, **kw):
    """glossy_room with a normal-mapped, specular-textured ball (normal-map and uv_scale adjoints)."""
    return glossy_room(device, nmap=True, **kw)


# ---------------------------------------------------------------------------------------------------- reference meshes
# BASELINE configs C3 - C5 use the reference's own test scenes.  /root/reference is not on the GPU box, so their arrays
# were exported once with the reference's unmodified loaders (tests/golden/export_ref_scenes.py) into .npz fixtures.
_FIXTURE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_fixture_cache = {}


def _fixture(name):
    if name not in _fixture_cache:
        _fixture_cache[name] = dict(np.load(os.path.join(_FIXTURE_DIR, name + ".npz")))
    return _fixture_cache[name]


def _mitsuba_fixture_scene(name, device, resolution, grad_shapes=(), grad_materials=(), cam_translation=None, cam_grad=False, materials_override=None):
    f = _fixture(name)
    res = tuple(int(x) for x in f["cam.resolution"]) if resolution is None else tuple(resolution)
    pos, look = torch.from_numpy(f["cam.position"]).clone(), torch.from_numpy(f["cam.look_at"]).clone()
    if cam_translation is not None:
        pos, look = pos + cam_translation, look + cam_translation
    up = torch.from_numpy(f["cam.up"]).clone()
    if cam_grad:
        pos.requires_grad_(True), look.requires_grad_(True), up.requires_grad_(True)
    cam = api.Camera(position=pos, look_at=look, up=up, fov=torch.from_numpy(f["cam.fov"]).clone(), clip_near=float(f["cam.clip_near"]), resolution=res)
    shapes = []
    for i in range(int(f["num_shapes"])):
        p = "shape%d." % i
        opt = {k: torch.from_numpy(f[p + k]).to(device) for k in ("uvs", "normals", "uv_indices", "normal_indices") if p + k in f}
        v = torch.from_numpy(f[p + "vertices"]).to(device)
        if i in grad_shapes or (i - int(f["num_shapes"])) in grad_shapes:
            v.requires_grad_(True)
        shapes.append(api.Shape(v, torch.from_numpy(f[p + "indices"]).to(device), int(f[p + "material_id"]), **opt))
    mats = []
    for i in range(int(f["num_materials"])):
        p = "mat%d." % i
        g = i in grad_materials or (i - int(f["num_materials"])) in grad_materials
        tex = {}
        for k in ("diffuse", "specular", "roughness"):
            t = torch.from_numpy(f[p + k + ".texels"]).to(device)
            if materials_override and (i, k) in materials_override:
                t = torch.tensor(materials_override[(i, k)], dtype=torch.float32, device=device)
            if g:
                t.requires_grad_(True)
            tex[k] = api.Texture(t, torch.from_numpy(f[p + k + ".uv_scale"]).to(device))
        mats.append(api.Material(diffuse_reflectance=tex["diffuse"], specular_reflectance=tex["specular"], roughness=tex["roughness"],
                                 two_sided=bool(f[p + "two_sided"])))
    lights = [api.AreaLight(int(f["light%d.shape_id" % i]), torch.from_numpy(f["light%d.intensity" % i]).clone(), two_sided=bool(f["light%d.two_sided" % i]))
              for i in range(int(f["num_lights"]))]
    return api.Scene(cam, shapes, mats, lights)


def teapot(device, resolution=(512, 512), grad=True):
    """C3, tests/test_teapot_reflectance.py:12-60 (tests/scenes/teapot.xml, 15 712 triangles, 3 lamps, textured floor): the
    optimisation's initial guess -- teapot material diffuse 0.3 / specular 0.5 / roughness 0.2 and the camera translated by
    (-0.2, 0.2, -0.2), all differentiable (SVBRDF + camera-pose gradients)."""
    n = int(_fixture("scene_teapot")["num_materials"])
    over = {(n - 1, "diffuse"): [0.3, 0.3, 0.3], (n - 1, "specular"): [0.5, 0.5, 0.5], (n - 1, "roughness"): [0.2]}
    return _mitsuba_fixture_scene("scene_teapot", device, resolution, grad_materials=(-1,) if grad else (), cam_translation=torch.tensor([-0.2, 0.2, -0.2]),
                                  cam_grad=grad, materials_override=over)


def teapot_geometry(device, resolution=(512, 512), grad=True):
    """C3's scene with the teapot's vertices (lid = shape 4, body = shape 5) differentiable: boundary terms on a 15 k-triangle mesh."""
    return _mitsuba_fixture_scene("scene_teapot", device, resolution, grad_shapes=(4, 5) if grad else (), cam_translation=torch.tensor([-0.2, 0.2, -0.2]))


def bunny_box(device, resolution=(1024, 1024), grad=True, cam_translation=None):
    """C4, tests/test_bunny_box.py:8-36 (tests/scenes/bunny_box.xml, 14 416 triangles, Cornell-style closed box, deep
    paths): the bunny's vertices are differentiable (the test's translation / rotation parameters are linear functions of them)."""
    return _mitsuba_fixture_scene("scene_bunny_box", device, resolution, grad_shapes=(-1,) if grad else (), cam_translation=cam_translation)


def bunny_box_shifted(device, **kw):
    """C4's meshes seen from a camera moved off the box's axis by (0.013, 0.007, 0).  The file's camera sits exactly on the
    axis of the box, so the floor / wall corners project onto exact pixel diagonals, and the (0, 0), (.5, .5), (.25, .75) ...
    points of a Sobol pattern put primary rays EXACTLY on an edge shared by two shapes: which of the two a ray then hits is decided
    by the last bit of two hit distances (in Embree as well as here) and the rest of the path follows.  Sample-exact
    comparisons with the Sobol sampler use this camera; with the independent sampler the file's camera is used."""
    return bunny_box(device, cam_translation=torch.tensor([0.013, 0.007, 0.0]), **kw)


def teapot_pose(device, k, num_poses=64, resolution=(512, 512), grad=True):
    """C5 (BASELINE.json configs[4]): pose k of `num_poses` cameras on a circle around the teapot of C3, at the distance
    and height of the file's camera, looking at the teapot's centre; the pose is differentiable."""
    sc = teapot(device, resolution=resolution, grad=grad)
    f = _fixture("scene_teapot")
    centre = torch.tensor([0.3, 2.6, 0.0])  # centre of the teapot's bounding box (shapes 4 and 5)
    p0 = torch.from_numpy(f["cam.position"])
    r = float(torch.linalg.norm((p0 - centre)[[0, 2]]))
    a = 2.0 * math.pi * k / num_poses
    pos = torch.tensor([float(centre[0]) + r * math.sin(a), float(p0[1]), float(centre[2]) + r * math.cos(a)])
    sc.camera = api.Camera(position=pos.requires_grad_(grad), look_at=centre.clone().requires_grad_(grad), up=torch.tensor([0.0, 1.0, 0.0], requires_grad=grad),
                           fov=torch.from_numpy(f["cam.fov"]).clone(), clip_near=float(f["cam.clip_near"]), resolution=tuple(resolution))
    return sc


SCENES = {"teapot": teapot, "bunny_box": bunny_box, "bunny_box_shifted": bunny_box_shifted, "teapot_geometry": teapot_geometry, "env_ball_fisheye": env_ball_fisheye, "shadow_blocker_all": shadow_blocker_all, "single_triangle": single_triangle, "shadow_blocker": shadow_blocker, "glossy_room": glossy_room, "random_soup": random_soup,
          "nmap_room": nmap_room, "corner_ball": corner_ball, "hires_room": hires_room, "ortho_room": ortho_room, "distort_room": distort_room, "fisheye_room": fisheye_room, "panorama_room": panorama_room, "env_ball": env_ball, "env_ball_flat_sky": env_ball_flat_sky}
