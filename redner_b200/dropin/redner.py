"""Drop-in `redner` module: put this directory in front of sys.path and an UNMODIFIED pyredner
(pyredner/render_pytorch.py:3 `import redner`) runs on the sm_100a kernels of libredner_b200.so.
See INTEGRATION.md."""
import os
import sys

_pkg_parent = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _pkg_parent not in sys.path:
    sys.path.insert(0, _pkg_parent)

from redner_b200.redner import *  # noqa: F401,F403,E402
from redner_b200.redner import (AreaLight, Camera, CameraType, DAreaLight, DCamera, DEnvironmentMap, DMaterial, DScene, DShape,  # noqa: F401,E402
                                EnvironmentMap, Material, RenderOptions, SamplerType, Scene, Shape, Texture1, Texture3, TextureN, Vector2i,
                                channels, compute_num_channels, float_ptr, int_ptr, render)


def _not_on_the_hot_path(name):
    def f(*a, **k):
        raise NotImplementedError("redner.%s is asset IO / preprocessing outside the render hot path (SURVEY.md section 2 rows 19-21); "
                                  "use the reference build for it" % name)
    return f


load_serialized = _not_on_the_hot_path("load_serialized")
automatic_uv_map = _not_on_the_hot_path("automatic_uv_map")
copy_texture_atlas = _not_on_the_hot_path("copy_texture_atlas")


class UVTriMesh:  # referenced at import time by pyredner/shape.py only inside functions
    def __init__(self, *a, **k):
        _not_on_the_hot_path("UVTriMesh")()


class TextureAtlas:
    def __init__(self, *a, **k):
        _not_on_the_hot_path("TextureAtlas")()
