"""CPU suite, part 2: the C-ABI library exists, loads without a GPU, exports every symbol include/redner_b200.h declares,
agrees with the ctypes mirror on struct layout, and fails LOUDLY (no fallback) when asked to render without CUDA."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest
import torch

from redner_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "redner_b200.h")


def test_library_is_built_and_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m redner_b200.build` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", open(HEADER).read()))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libredner_b200.so does not export %s" % name
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))


def test_struct_layout_matches_header():
    structs = ["rb_camera", "rb_shape", "rb_texture", "rb_material", "rb_area_light", "rb_envmap", "rb_scene_desc", "rb_options", "rb_dshape",
               "rb_dcamera", "rb_dscene_desc"]
    src = '#include <stdio.h>\n#include "redner_b200.h"\nint main(){' + "".join('printf("%s %%zu\\n", sizeof(%s));' % (s, s) for s in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for s in structs:
        assert int(sizes[s]) == ctypes.sizeof(getattr(_lib, s)), s


def test_compute_num_channels_without_gpu():
    from redner_b200 import redner as rb
    ch = rb.channels
    assert rb.compute_num_channels([ch.radiance], 0) == 3
    assert rb.compute_num_channels([ch.radiance, ch.alpha, ch.depth, ch.uv, ch.shape_id], 0) == 3 + 1 + 1 + 2 + 1
    assert rb.compute_num_channels([ch.generic_texture, ch.position], 5) == 8
    assert _lib.load().rb_version().decode().startswith("redner_b200")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_silent_cpu_fallback():
    """Without a CUDA device the product must raise, not fall back to the oracle or to PyTorch."""
    import scenes
    from redner_b200 import api
    from redner_b200 import redner as rb
    sc = scenes.single_triangle(torch.device("cpu"), resolution=(8, 8))
    args = api.RenderFunction.serialize_scene(sc, 1, 1, device=torch.device("cpu"), backend=rb)
    with pytest.raises(RuntimeError) as e:
        api.RenderFunction.apply(0, *args)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(RuntimeError):
        api.get_device()


def test_product_never_imports_the_oracle():
    """The package under redner_b200/ must not reference oracle/ or the emulator (they are test infrastructure)."""
    pkg = os.path.join(ROOT, "redner_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ref_loader" not in txt and "oracle/" not in txt.replace("the oracle", "") or f in ("api.py",), f
                assert "cpu_emu" not in txt or "tools/cpu_emu" in txt, f
