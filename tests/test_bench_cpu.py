"""CPU suite: the parts of bench.py that do not need a GPU -- the workloads' scenes load from the committed fixtures, the roofline
accounting is the SURVEY.md section 8(d) formula, and the reference arm (`--impl reference`: the unmodified reference timed on the host
cores) prints a contract-shaped JSON line within its time budget."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_scenes_load_from_fixtures():
    dev = torch.device("cpu")
    tris = {}
    for name, wl in bench.WORKLOADS.items():
        small = dict(wl, res=32)
        sc = bench.make_scene(small, dev, pose=3 if wl.get("poses") else None)
        tris[name] = sum(int(s.indices.shape[0]) for s in sc.shapes)
        assert len(bench.leaf_params(sc)) > 0, name
    assert tris == {"c2": 6, "c3": 15712, "c4": 14416, "c5": 15712}


def test_algorithmic_bytes_follow_the_survey_formula():
    a = bench.bytes_per_sample(1.0, 0.5, True, True)
    assert a["k_forward"] == 750 + 1630 and a["k_bwd_trace"] == a["k_forward"]
    assert a["k_bwd_secondary"] == 2650 + 3260 * 0.5 and a["k_bwd_sweep"] == 448 + 1280 + 305 and a["k_primary_edge"] == 1700 + 3260
    assert bench.bytes_per_sample(1.0, 0.5, False, False)["k_primary_edge"] == 0.0
    assert bench.pick_sample(bench.WORKLOADS["c2"], 1.0, 1e9) == (512, 64) and bench.pick_sample(bench.WORKLOADS["c2"], 1.0, 1e-9) == (64, 4)


@pytest.mark.skipif(not bench.reference_available(), reason="oracle/_ref not built")
def test_reference_arm_prints_a_contract_line():
    env = dict(os.environ, RB_REF_TOTAL_S="4", RB_REF_FULL_STEP_MAX_S="0")
    r = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"] == "fwd+bwd megasamples/s at 512x512x64spp" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "same_config" in d["config"] and "sample" in d["config"]
