"""Drop-in replacement for the reference's pybind11 module `redner` (src/redner.cpp:20-272), restricted to the names
`pyredner/render_pytorch.py` uses on the RenderFunction.forward/backward path.  Same class / enum / function names,
same constructor argument order and meaning; objects are thin POD holders that are marshalled into the C ABI of
libredner_b200.so (include/redner_b200.h) when a Scene is constructed or `render` is called.

To run an unmodified pyredner on top of the B200 kernels, put `redner_b200/dropin` in front of `sys.path`
(see INTEGRATION.md); `import redner` then resolves to this module.

Error behaviour: where the reference assert()s / exit(1)s, this module raises RuntimeError carrying rb_last_error().
"""
import ctypes as C
import os
import enum

from . import _lib as L


class float_ptr:  # src/ptr.h:9-23, src/redner.cpp:23-24
    def __init__(self, addr):
        self.addr = int(addr)


class int_ptr:  # src/redner.cpp:25-26
    def __init__(self, addr):
        self.addr = int(addr)


def _addr(p):
    if p is None:
        return 0
    return int(p.addr)


def _read_floats(p, n):
    """Host-read small parameters at construction time (the reference dereferences these pointers on the host:
    src/camera.h:44-62, src/area_light.h:18-20)."""
    a = _addr(p)
    if a == 0:
        return None
    return list((C.c_float * n).from_address(a))


class CameraType(enum.IntEnum):  # src/redner.cpp:28-32
    perspective = 0
    orthographic = 1
    fisheye = 2
    panorama = 3


class SamplerType(enum.IntEnum):  # src/redner.cpp:203-205
    independent = 0
    sobol = 1


class channels(enum.IntEnum):  # src/redner.cpp:183-199
    radiance = 0
    alpha = 1
    depth = 2
    position = 3
    geometry_normal = 4
    shading_normal = 5
    uv = 6
    barycentric_coordinates = 7
    diffuse_reflectance = 8
    specular_reflectance = 9
    roughness = 10
    generic_texture = 11
    vertex_color = 12
    shape_id = 13
    triangle_id = 14
    material_id = 15


class Vector2i:  # src/redner.cpp:218-221
    def __init__(self, x, y):
        self.x = int(x)
        self.y = int(y)


class Camera:  # src/redner.cpp:34-50, src/camera.h:22-66
    def __init__(self, width, height, position, look, up, cam_to_world, world_to_cam, intrinsic_mat_inv, intrinsic_mat,
                 distortion_params, clip_near, camera_type, viewport_beg, viewport_end):
        c = L.rb_camera()
        c.width, c.height = int(width), int(height)
        c2w = _read_floats(cam_to_world, 16)
        if c2w is not None:
            w2c = _read_floats(world_to_cam, 16)
            c.cam_to_world[:] = c2w
            c.world_to_cam[:] = w2c
            c.use_look_at = 0
        else:
            c.position[:] = _read_floats(position, 3)
            c.look[:] = _read_floats(look, 3)
            c.up[:] = _read_floats(up, 3)
            c.use_look_at = 1
        c.intrinsic_mat_inv[:] = _read_floats(intrinsic_mat_inv, 9)
        c.intrinsic_mat[:] = _read_floats(intrinsic_mat, 9)
        d = _read_floats(distortion_params, 8)
        c.has_distortion = 0 if d is None else 1
        if d is not None:
            c.distortion[:] = d
        c.clip_near = float(clip_near)
        c.camera_type = int(camera_type)
        c.viewport_beg[:] = [viewport_beg.x, viewport_beg.y]
        c.viewport_end[:] = [viewport_end.x, viewport_end.y]
        self._c = c
        self.use_look_at = bool(c.use_look_at)

    def has_distortion_params(self):
        return bool(self._c.has_distortion)


class DCamera:  # src/redner.cpp:52-60
    def __init__(self, position, look, up, cam_to_world, world_to_cam, intrinsic_mat_inv, intrinsic_mat, distortion_params):
        d = L.rb_dcamera()
        d.position, d.look, d.up = _addr(position) or None, _addr(look) or None, _addr(up) or None
        d.cam_to_world, d.world_to_cam = _addr(cam_to_world) or None, _addr(world_to_cam) or None
        d.intrinsic_mat_inv, d.intrinsic_mat = _addr(intrinsic_mat_inv) or None, _addr(intrinsic_mat) or None
        d.distortion = _addr(distortion_params) or None
        self._c = d


class Shape:  # src/redner.cpp:84-104, src/shape.h:9-63
    def __init__(self, vertices, indices, uvs, normals, uv_indices, normal_indices, colors, num_vertices, num_uv_vertices,
                 num_normal_vertices, num_triangles, material_id, light_id):
        s = L.rb_shape()
        s.vertices, s.indices = _addr(vertices) or None, _addr(indices) or None
        s.uvs, s.normals = _addr(uvs) or None, _addr(normals) or None
        s.uv_indices, s.normal_indices = _addr(uv_indices) or None, _addr(normal_indices) or None
        s.colors = _addr(colors) or None
        s.num_vertices, s.num_uv_vertices = int(num_vertices), int(num_uv_vertices)
        s.num_normal_vertices, s.num_triangles = int(num_normal_vertices), int(num_triangles)
        s.material_id, s.light_id = int(material_id), int(light_id)
        self._c = s
        self.num_vertices = s.num_vertices
        self.num_uv_vertices = s.num_uv_vertices
        self.num_normal_vertices = s.num_normal_vertices

    def has_uvs(self):
        return bool(self._c.uvs)

    def has_normals(self):
        return bool(self._c.normals)

    def has_colors(self):
        return bool(self._c.colors)


class DShape:  # src/redner.cpp:106-110
    def __init__(self, vertices, uvs, normals, colors):
        d = L.rb_dshape()
        d.vertices, d.uvs = _addr(vertices) or None, _addr(uvs) or None
        d.normals, d.colors = _addr(normals) or None, _addr(colors) or None
        self._c = d


class _Texture:  # src/redner.cpp:112-131, src/texture.h:14-46
    _channels = None

    def __init__(self, texels, width, height, channels, uv_scale):
        assert len(texels) == len(width) == len(height)
        t = L.rb_texture()
        n = min(len(texels), L.RB_MAX_MIP_LEVELS)
        for i in range(n):
            t.texels[i] = _addr(texels[i]) or None
            t.width[i] = int(width[i])
            t.height[i] = int(height[i])
        t.num_levels = n
        t.channels = int(channels) if self._channels is None else self._channels
        t.uv_scale = _addr(uv_scale) or None
        self._c = t


class Texture1(_Texture):
    _channels = 1


class Texture3(_Texture):
    _channels = 3


class TextureN(_Texture):
    _channels = None


class Material:  # src/redner.cpp:133-151, src/material.h:12-91
    def __init__(self, diffuse_reflectance, specular_reflectance, roughness, generic_texture, normal_map, compute_specular_lighting,
                 two_sided, use_vertex_color):
        m = L.rb_material()
        m.diffuse_reflectance = diffuse_reflectance._c
        m.specular_reflectance = specular_reflectance._c
        m.roughness = roughness._c
        m.generic_texture = generic_texture._c
        m.normal_map = normal_map._c
        m.compute_specular_lighting = int(bool(compute_specular_lighting))
        m.two_sided = int(bool(two_sided))
        m.use_vertex_color = int(bool(use_vertex_color))
        self._c = m

    def _levels(self, t):
        return int(t.num_levels)

    def _size(self, t, i):
        return (int(t.width[i]), int(t.height[i]))

    def get_diffuse_levels(self):
        return self._levels(self._c.diffuse_reflectance)

    def get_diffuse_size(self, i):
        return self._size(self._c.diffuse_reflectance, i)

    def get_specular_levels(self):
        return self._levels(self._c.specular_reflectance)

    def get_specular_size(self, i):
        return self._size(self._c.specular_reflectance, i)

    def get_roughness_levels(self):
        return self._levels(self._c.roughness)

    def get_roughness_size(self, i):
        return self._size(self._c.roughness, i)

    def get_generic_levels(self):
        return self._levels(self._c.generic_texture)

    def get_generic_size(self, i):
        t = self._c.generic_texture
        return (int(t.channels), int(t.width[i]), int(t.height[i]))

    def get_normal_map_levels(self):
        return self._levels(self._c.normal_map)

    def get_normal_map_size(self, i):
        return self._size(self._c.normal_map, i)


class DMaterial:  # src/redner.cpp:153-158
    def __init__(self, diffuse_reflectance, specular_reflectance, roughness, generic_texture, normal_map):
        m = L.rb_material()
        m.diffuse_reflectance = diffuse_reflectance._c
        m.specular_reflectance = specular_reflectance._c
        m.roughness = roughness._c
        m.generic_texture = generic_texture._c
        m.normal_map = normal_map._c
        self._c = m


class AreaLight:  # src/redner.cpp:160-164, src/area_light.h:8-36
    def __init__(self, shape_id, intensity, two_sided, directly_visible):
        a = L.rb_area_light()
        a.shape_id = int(shape_id)
        a.intensity[:] = _read_floats(intensity, 3)
        a.two_sided = int(bool(two_sided))
        a.directly_visible = int(bool(directly_visible))
        self._c = a


class DAreaLight:  # src/redner.cpp:166-167
    def __init__(self, intensity):
        self.addr = _addr(intensity)


class EnvironmentMap:  # src/redner.cpp:169-178
    def __init__(self, values, env_to_world, world_to_env, sample_cdf_ys, sample_cdf_xs, pdf_norm, directly_visible):
        e = L.rb_envmap()
        e.values = values._c
        e.env_to_world[:] = _read_floats(env_to_world, 16)
        e.world_to_env[:] = _read_floats(world_to_env, 16)
        e.sample_cdf_ys, e.sample_cdf_xs = _addr(sample_cdf_ys) or None, _addr(sample_cdf_xs) or None
        e.pdf_norm = float(pdf_norm)
        e.directly_visible = int(bool(directly_visible))
        self._c = e

    def get_levels(self):
        return int(self._c.values.num_levels)

    def get_size(self, i):
        return (int(self._c.values.width[i]), int(self._c.values.height[i]))


class DEnvironmentMap:  # src/redner.cpp:179-181
    def __init__(self, values, world_to_env):
        e = L.rb_denvmap()
        e.values = values._c
        e.world_to_env = _addr(world_to_env) or None
        self._c = e


class RenderOptions:  # src/redner.cpp:207-216, src/pathtracer.h:16-23
    def __init__(self, seed, num_samples, max_bounces, channels, sampler_type, sample_pixel_center):
        self.seed = int(seed)
        self.num_samples = int(num_samples)
        self.max_bounces = int(max_bounces)
        self.channels = [int(c) for c in channels]
        self.sampler_type = int(sampler_type)
        self.sample_pixel_center = bool(sample_pixel_center)


def compute_num_channels(chs, max_generic_texture_dimension):  # src/redner.cpp:201
    lib = L.load()
    arr = (C.c_int * max(1, len(chs)))(*[int(c) for c in chs])
    n = lib.rb_compute_num_channels(arr, len(chs), int(max_generic_texture_dimension))
    if n < 0:
        raise RuntimeError("compute_num_channels: unknown channel")
    return n


class Scene:  # src/redner.cpp:62-73, src/scene.cpp:63-307
    def __init__(self, camera, shapes, materials, area_lights, envmap, use_gpu, gpu_index, use_primary_edge_sampling,
                 use_secondary_edge_sampling):
        lib = L.load()
        self._lib = lib
        self._handle = None
        d = L.rb_scene_desc()
        d.camera = camera._c
        self._shapes = (L.rb_shape * max(1, len(shapes)))(*[s._c for s in shapes])
        self._materials = (L.rb_material * max(1, len(materials)))(*[m._c for m in materials])
        self._lights = (L.rb_area_light * max(1, len(area_lights)))(*[a._c for a in area_lights])
        d.num_shapes, d.shapes = len(shapes), self._shapes
        d.num_materials, d.materials = len(materials), self._materials
        d.num_lights, d.lights = len(area_lights), self._lights
        self._env = envmap._c if envmap is not None else None
        d.envmap = C.pointer(self._env) if self._env is not None else None
        d.use_gpu = int(bool(use_gpu))
        d.gpu_index = int(gpu_index)
        d.use_primary_edge_sampling = int(bool(use_primary_edge_sampling))
        if envmap is not None and use_secondary_edge_sampling:
            # The C ABI rejects this combination (the reference differentiates sky-side edge rays at stale hit points, DESIGN.md
            # section 7).  pyredner switches both edge samplers on by default, so an environment-lit scene arrives here with the flag
            # set: fail loudly unless the caller accepts the difference (RB_ENVMAP_WITHOUT_SECONDARY_EDGES=1, or pass
            # use_secondary_edge_sampling=False), in which case the scene is rendered WITHOUT shadow / interreflection boundary terms.
            if os.environ.get("RB_ENVMAP_WITHOUT_SECONDARY_EDGES", "") in ("", "0"):
                raise RuntimeError("redner_b200: secondary edge sampling together with an environment map is not available (DESIGN.md section 7). "
                                   "Pass use_secondary_edge_sampling=False, or set RB_ENVMAP_WITHOUT_SECONDARY_EDGES=1 to have it switched off "
                                   "automatically: such scenes then lose their secondary (shadow / interreflection) boundary gradients compared "
                                   "with the reference.")
            import warnings
            warnings.warn("redner_b200: secondary edge sampling is switched off for this scene (environment map, "
                          "RB_ENVMAP_WITHOUT_SECONDARY_EDGES=1); interior terms and primary edges are rendered and differentiated")
            use_secondary_edge_sampling = False
        d.use_secondary_edge_sampling = int(bool(use_secondary_edge_sampling))
        h = C.c_void_p()
        # build on the current PyTorch stream of the scene's device: the geometry tensors were produced there
        stream = 0
        try:
            import torch
            if use_gpu and torch.cuda.is_available():
                stream = torch.cuda.current_stream(gpu_index if gpu_index >= 0 else None).cuda_stream
        except ImportError:
            pass
        if hasattr(lib, "rb_scene_create_on_stream"):
            rc = lib.rb_scene_create_on_stream(C.byref(d), C.byref(h), C.c_void_p(stream or 0))
        else:
            rc = lib.rb_scene_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise RuntimeError("redner.Scene: " + L.last_error(lib))
        self._handle = h
        self.max_generic_texture_dimension = lib.rb_scene_max_generic_texture_dimension(h)
        self.use_gpu = bool(use_gpu)
        self.gpu_index = int(gpu_index)

    # --- redner_b200 extensions (no reference counterpart) ---
    def set_partition(self, part, num_parts, rows_per_stripe=16):
        if self._lib.rb_scene_set_partition(self._handle, int(part), int(num_parts), int(rows_per_stripe)) != 0:
            raise RuntimeError("redner.Scene.set_partition: " + L.last_error(self._lib))

    def set_camera(self, camera):
        """Re-target this scene at another camera; only the camera-dependent tables are rebuilt, on the device (rb_scene_set_camera)."""
        if self._lib.rb_scene_set_camera(self._handle, C.byref(camera._c)) != 0:
            raise RuntimeError("redner.Scene.set_camera: " + L.last_error(self._lib))

    def last_stats(self):
        n = C.c_int(0)
        ms = C.c_float(0)
        self._lib.rb_scene_last_stats(self._handle, C.byref(n), C.byref(ms))
        return n.value, ms.value

    def last_stage_stats(self):
        """({kernel name: ms}, path_vertices, primary_hits) of the last render on this scene."""
        ms = (C.c_float * 4)()
        v, h = C.c_double(0), C.c_double(0)
        self._lib.rb_scene_last_stage_stats(self._handle, ms, C.byref(v), C.byref(h))
        out = dict(zip(("k_forward", "k_backward", "k_primary_edge", "k_finish_camera"), list(ms)))
        if hasattr(self._lib, "rb_scene_last_backward_stats"):
            # "k_backward" is the sum over the backward bands; its three stages follow
            b = (C.c_float * 3)()
            self._lib.rb_scene_last_backward_stats(self._handle, b)
            out.update(dict(zip(("k_bwd_trace", "k_bwd_secondary", "k_bwd_sweep"), list(b))))
        return out, v.value, h.value

    def edge_trees(self):
        """(records [n, 32] as uint32 words, root of the camera-silhouette tree, root of the other tree, billboard size): test hook."""
        import numpy as np
        info = (C.c_int * 3)()
        ex = C.c_float(0)
        self._lib.rb_scene_edge_trees(self._handle, info, C.byref(ex), None, 0)
        rec = np.zeros((max(info[0], 0), 32), dtype=np.uint32)
        if info[0] > 0:
            self._lib.rb_scene_edge_trees(self._handle, info, C.byref(ex), rec.ctypes.data_as(C.c_void_p), rec.nbytes)
        return rec, info[1], info[2], ex.value

    def edge_list(self):
        """[num_edges, 5] int32 rows (shape, v0, v1, f0, f1): test hook."""
        import numpy as np
        n = C.c_int(0)
        self._lib.rb_scene_edge_list(self._handle, C.byref(n), None, 0)
        out = np.zeros((max(n.value, 0), 5), dtype=np.int32)
        if n.value > 0:
            self._lib.rb_scene_edge_list(self._handle, C.byref(n), out.ctypes.data_as(C.POINTER(C.c_int)), out.nbytes)
        return out

    def build_ms(self):
        ms = (C.c_float * 3)()
        self._lib.rb_scene_build_ms(self._handle, ms)
        return dict(zip(("bvh", "lights", "edges"), list(ms)))

    def __del__(self):
        try:
            if self._handle is not None and self._handle.value:
                self._lib.rb_scene_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


class DScene:  # src/redner.cpp:75-82
    def __init__(self, camera, shapes, materials, area_lights, envmap, use_gpu, gpu_index):
        d = L.rb_dscene_desc()
        d.camera = camera._c
        self._shapes = (L.rb_dshape * max(1, len(shapes)))(*[s._c for s in shapes])
        self._materials = (L.rb_material * max(1, len(materials)))(*[m._c for m in materials])
        self._lights = (C.c_void_p * max(1, len(area_lights)))(*[a.addr or None for a in area_lights])
        d.num_shapes, d.shapes = len(shapes), self._shapes
        d.num_materials, d.materials = len(materials), self._materials
        d.num_lights, d.light_intensity = len(area_lights), self._lights
        self.envmap = envmap
        d.envmap = C.pointer(envmap._c) if envmap is not None else None
        self._c = d


def render(scene, options, rendered_image, d_rendered_image, d_scene, screen_gradient_image, debug_image, stream=None):
    """src/redner.cpp:257 / src/pathtracer.cpp:177-183.  `debug_image` is accepted and ignored (the reference never
    reads it).  `stream` (a cudaStream_t value) is a redner_b200 extension; by default the current PyTorch stream is used
    when torch is importable, else the legacy default stream."""
    lib = scene._lib
    o = L.rb_options()
    o.seed = options.seed
    o.num_samples = options.num_samples
    o.max_bounces = options.max_bounces
    chs = (C.c_int * max(1, len(options.channels)))(*options.channels)
    o.num_channels = len(options.channels)
    o.channels = chs
    o.sampler_type = options.sampler_type
    o.sample_pixel_center = int(options.sample_pixel_center)
    if stream is None:
        try:
            import torch
            # (the stream of the SCENE's device: a stream handle of another device would be invalid there)
            stream = torch.cuda.current_stream(scene.gpu_index if scene.gpu_index >= 0 else None).cuda_stream if torch.cuda.is_available() else 0
        except Exception:
            stream = 0
    rc = lib.rb_render(scene._handle, C.byref(o), _addr(rendered_image) or None, _addr(d_rendered_image) or None,
                       C.byref(d_scene._c) if d_scene is not None else None, _addr(screen_gradient_image) or None, C.c_void_p(stream or 0))
    if rc != 0:
        raise RuntimeError("redner.render: " + L.last_error(lib))
