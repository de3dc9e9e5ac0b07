"""Diagnostic for the statistical secondary-edge cases: per-seed gradients of one library build (CUDA or the host
emulator) saved to an .npz, plus the z statistics against the golden.

usage: python tools/diag_stat.py <case> <out.npz> [--lib path.so] [--emu path.so] [--seeds N]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch


def main():
    name, out = sys.argv[1], sys.argv[2]
    args = sys.argv[3:]
    from redner_b200 import _lib
    dev = torch.device("cuda:0")
    if "--emu" in args:
        _lib._lib = _lib._bind(ctypes.CDLL(args[args.index("--emu") + 1]))
        dev = torch.device("cpu")
    elif "--lib" in args:
        _lib._lib = _lib.load(args[args.index("--lib") + 1])
    from redner_b200 import redner as rb
    import parity_utils as pu
    cfg = dict(pu.STAT_CASES[name])
    if "--seeds" in args:
        cfg["seeds"] = cfg["seeds"][:int(args[args.index("--seeds") + 1])]
    acc = {k: [] for k in cfg["keys"]}
    for seed in cfg["seeds"]:
        _, grads = pu.render_case(rb, dev, cfg, seed)
        for k in cfg["keys"]:
            acc[k].append(grads[k].numpy())
    np.savez(out, **{k: np.stack(v) for k, v in acc.items()})
    g = pu.load_golden(name)
    for k in cfg["keys"]:
        a = np.stack(acc[k]).astype(np.float64)
        mean, sem = a.mean(0), a.std(0, ddof=1) / np.sqrt(a.shape[0])
        ref_mean, ref_sem = g["mean." + k], g["sem." + k]
        err = np.linalg.norm(mean - ref_mean)
        noise = np.sqrt(np.linalg.norm(sem) ** 2 + np.linalg.norm(ref_sem) ** 2)
        floor = 1e-3 * np.abs(ref_mean).max()
        z = (mean - ref_mean) / np.maximum(np.sqrt(sem ** 2 + ref_sem ** 2), floor)
        print("%-40s %-16s err/noise %.2f  rel %.3f  z_rms %.2f  max|z| %.2f" % (os.path.basename(out), k, err / noise, err / np.linalg.norm(ref_mean),
                                                                              np.sqrt((z ** 2).mean()), np.abs(z).max()), flush=True)


if __name__ == "__main__":
    main()
