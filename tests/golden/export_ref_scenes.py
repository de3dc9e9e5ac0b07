#!/usr/bin/env python
"""Exports the meshes / materials / lights / camera of the reference's own test scenes (BASELINE configs C3, C4, C5)
to small .npz fixtures, loaded with the reference's UNMODIFIED Python loaders in the build container:

    tests/scenes/teapot.xml     -> tests/golden/scene_teapot.npz      (test_teapot_reflectance.py:12, 15 712 triangles)
    tests/scenes/bunny_box.xml  -> tests/golden/scene_bunny_box.npz   (test_bunny_box.py:12)
    tests/scenes/teapot.obj     -> tests/golden/scene_teapot_obj.npz  (test_batch.py / C5 mesh)

/root/reference does not exist on the GPU box, so the parity tests and bench.py read these fixtures (tests/scenes.py).
skimage / imageio are not installed here; the two functions pyredner.imread needs are provided by PIL.
"""
import os
import re
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _stub_image_modules():
    from PIL import Image
    sk, skio, skt, iio = (types.ModuleType(n) for n in ("skimage", "skimage.io", "skimage.transform", "imageio"))
    skio.imread = lambda fn: np.asarray(Image.open(fn))
    sk.img_as_float = lambda a: a.astype(np.float64) / 255.0 if a.dtype == np.uint8 else a.astype(np.float64)
    sk.io, sk.transform = skio, skt
    for m in (sk, skio, skt, iio):
        sys.modules[m.__name__] = m


def _tolerant_fromstring():
    """numpy >= 2.3 raises where the numpy pyredner was written for stopped at the first unparsable token with a
    DeprecationWarning; load_mitsuba.parse_vector relies on the latter ('0.64 0.64 0.64' with sep=',' -> [0.64] -> retry)."""
    orig = np.fromstring

    def fromstring(s, dtype=float, count=-1, sep=""):
        try:
            return orig(s, dtype=dtype, count=count, sep=sep)
        except ValueError:
            vals = []
            for tok in s.split(sep):
                try:
                    vals.append(float(tok))
                except ValueError:
                    m = re.match(r"\s*[-+0-9.eE]+", tok)
                    if m:
                        vals.append(float(m.group(0)))
                    break
            return np.asarray(vals, dtype=dtype)
    np.fromstring = fromstring


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _tex(prefix, tex, out):
    # pyredner.Texture: `texels` is [h, w, c] or [c]; the mip pyramid is rebuilt by api.Texture like pyredner/texture.py
    out[prefix + ".texels"] = _np(tex.texels)
    out[prefix + ".uv_scale"] = _np(tex.uv_scale)


def export(scene, path):
    out = {}
    cam = scene.camera
    out["cam.position"], out["cam.look_at"], out["cam.up"] = _np(cam.position), _np(cam.look_at), _np(cam.up)
    out["cam.fov"] = _np(cam.fov)
    out["cam.clip_near"] = np.float32(cam.clip_near)
    out["cam.resolution"] = np.asarray(cam.resolution, dtype=np.int32)
    out["num_shapes"], out["num_materials"], out["num_lights"] = (np.int32(len(x)) for x in (scene.shapes, scene.materials, scene.area_lights))
    for i, s in enumerate(scene.shapes):
        p = "shape%d." % i
        out[p + "vertices"], out[p + "indices"] = _np(s.vertices), _np(s.indices)
        for k in ("uvs", "normals", "uv_indices", "normal_indices"):
            v = getattr(s, k, None)
            if v is not None:
                out[p + k] = _np(v)
        out[p + "material_id"] = np.int32(s.material_id)
    for i, m in enumerate(scene.materials):
        p = "mat%d." % i
        _tex(p + "diffuse", m.diffuse_reflectance, out)
        _tex(p + "specular", m.specular_reflectance, out)
        _tex(p + "roughness", m.roughness, out)
        out[p + "two_sided"] = np.int32(bool(m.two_sided))
    for i, l in enumerate(scene.area_lights):
        p = "light%d." % i
        out[p + "shape_id"] = np.int32(l.shape_id)
        out[p + "intensity"] = _np(l.intensity)
        out[p + "two_sided"] = np.int32(bool(l.two_sided))
    np.savez_compressed(path, **out)
    nt = sum(int(s.indices.shape[0]) for s in scene.shapes)
    print(os.path.basename(path), "shapes", len(scene.shapes), "triangles", nt, "bytes", os.path.getsize(path))


def main():
    _stub_image_modules()
    _tolerant_fromstring()
    import ref_loader
    sys.modules["redner"] = ref_loader.load()
    sys.path.insert(0, REF)
    import torch
    import pyredner
    pyredner.set_use_gpu(False)
    os.chdir(os.path.join(REF, "tests"))
    export(pyredner.load_mitsuba("scenes/teapot.xml"), os.path.join(OUT, "scene_teapot.npz"))
    export(pyredner.load_mitsuba("scenes/bunny_box.xml"), os.path.join(OUT, "scene_bunny_box.npz"))
    # C5 / tests/test_batch.py style: the OBJ teapot (one object per material), a grey floor and a lamp are added by tests/scenes.py
    mats, meshes, _ = pyredner.load_obj("scenes/teapot.obj")
    out = {"num_meshes": np.int32(len(meshes))}
    for i, (mtl_name, mesh) in enumerate(meshes):
        p = "mesh%d." % i
        out[p + "vertices"], out[p + "indices"] = _np(mesh.vertices), _np(mesh.indices)
        for k in ("uvs", "normals", "uv_indices", "normal_indices"):
            v = getattr(mesh, k, None)
            if v is not None:
                out[p + k] = _np(v)
    np.savez_compressed(os.path.join(OUT, "scene_teapot_obj.npz"), **out)
    print("scene_teapot_obj.npz meshes", len(meshes), "triangles", sum(int(m.indices.shape[0]) for _, m in meshes),
          "bytes", os.path.getsize(os.path.join(OUT, "scene_teapot_obj.npz")))


if __name__ == "__main__":
    main()
