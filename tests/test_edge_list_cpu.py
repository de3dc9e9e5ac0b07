"""CPU suite: the data-parallel edge-list steps of redner_b200/csrc/rb_edge_list.cuh (what rb_edge_list.cu runs in kernels between CUB
sorts and scans) give the same list as host_build_edges, the step-by-step restatement of src/edge.cpp:233-296 -- on every fixture scene
(the teapot's UV seams and the bunny included) and on random scenes built to hit the corner cases: seam twins, groups of more than 32
equal segments (the reference's non-strict comparator reverses them), edges shared by more than two faces, degenerate and coplanar
triangles, empty shapes.  tests/edge_list_check.cpp holds the serial driver; the device driver is checked by tests/test_scene_build_gpu.py."""
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("edge_list") / "edge_list_check")
    cmd = ["g++", "-O2", "-std=c++17", "-w", "-include", os.path.join(ROOT, "tools", "cpu_emu", "emu_shim.h"), "-I/usr/local/cuda/include",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "edge_list_check.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, timeout=900)
    return exe


def test_fixture_scenes(checker, tmp_path):
    files = []
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scene_*.npz"))):
        d = np.load(path)
        out = str(tmp_path / (os.path.basename(path)[:-4] + ".bin"))
        with open(out, "wb") as f:
            stem = "shape" if "num_shapes" in d.files else "mesh"
            S = int(d["num_%ss" % ("shape" if stem == "shape" else "meshe")])
            np.array([S], np.int32).tofile(f)
            for s in range(S):
                v, i = d["%s%d.vertices" % (stem, s)].astype(np.float32), d["%s%d.indices" % (stem, s)].astype(np.int32)
                np.array([v.shape[0], i.shape[0]], np.int32).tofile(f)
                v.tofile(f)
                i.tofile(f)
        files.append(out)
    assert len(files) >= 2
    r = subprocess.run([checker] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    ok = [l for l in r.stdout.splitlines() if l.startswith("ok ")]
    assert len(ok) == len(files), r.stdout
    assert max(int(l.split()[5]) for l in ok) > 15000  # (the teapot scene: 18 084 edges)


def test_random_scenes(checker):
    r = subprocess.run([checker, "--random", "4000"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.strip().splitlines()[-1] == "random scenes 4000 mismatching 0"
