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


def main():
    so, names = sys.argv[1], sys.argv[2:]
    import torch
    from redner_b200 import _lib
    _lib._lib = _lib._bind(ctypes.CDLL(so))  # this process only: the emulator exports the same C ABI with host pointers
    from redner_b200 import redner as rb
    import parity_utils as pu
    dev = torch.device("cpu")
    for name in names:
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
