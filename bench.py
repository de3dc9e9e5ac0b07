#!/usr/bin/env python
"""Benchmark of the hot path: pyredner.RenderFunction forward + backward == two redner.render() calls.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode poses|tiles]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): the shadow-blocker scene of the reference's
tests/test_shadow_blocker.py at 512 x 512 pixels x 64 spp, max_bounces = 1, Sobol sampler, primary + secondary edge
sampling, loss = sum(img^2) (dense d_image = 2 img), forward seed s / backward seed s + 1000003.
One "step" = one forward render + one backward render of that scene == W*H*spp pixel samples through the whole
differentiable path tracer.  metric = fwd+bwd megasamples/s = W*H*spp / (t_forward_call + t_backward_call) / 1e6.

  value  whole-job throughput with the scene tensors resident in HBM; timed per step with CUDA events on the render
         stream around the two rb_render calls (scene construction -- BVH / light tables / edge tree -- is timed
         separately and reported in config.scene_build_ms, as BASELINE.md section 2 prescribes);
  e2e    the same metric through the public API (redner_b200.api.RenderFunction) starting from HOST tensors in pinned
         memory: host->device copies of every scene tensor, scene construction, forward, loss, backward and the
         device->host read of the loss and of all gradients are inside the timed region;
  roofline      dominant kernel / stage of the step, algorithmic bytes of SURVEY.md section 8(d) over its CUDA-event duration;
  cpu_baseline  the unmodified reference (oracle/_ref, CPU/Embree, all host cores) on a bounded sample of the same
                workload (rank 0, N = 1 only).

N > 1 (one process per GPU): `--mode poses` (default) renders one C2 image per rank with a different camera pose and
all-reduces the parameter gradients over NCCL (weak scaling, BASELINE config 5 pattern); `--mode tiles` splits ONE image
into row stripes across the ranks and all-reduces framebuffer and gradients (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

RES, SPP, MB, SEED = 512, 64, 1, 1


def bytes_per_sample(d_bar, hit_frac, use_primary, use_secondary):
    """Algorithmic bytes per pixel sample after SURVEY.md section 8(d) (fp32 state of every stage functor of the reference,
    counted once per write and once per consuming read), split by the kernel that does that work here:
      k_forward        750 + 1630 d
      k_bwd_trace      750 + 1630 d                     (primal replay)
      k_bwd_secondary  2650 d + 3260 (d - h)            (boundary terms: edge sample + two sub-paths)
      k_bwd_sweep      448 + 1280 d + 610 h             (reverse sweep, first-hit and camera adjoints)
      k_primary_edge   1700 + 3260 d
    d = mean executed bounces per sample (measured by the backward pass), h = measured primary-hit fraction."""
    a = {"k_forward": 750 + 1630 * d_bar, "k_bwd_trace": 750 + 1630 * d_bar, "k_bwd_sweep": 448 + 1280 * d_bar + 610 * hit_frac}
    a["k_bwd_secondary"] = (2650 * d_bar + 3260 * max(0.0, d_bar - hit_frac)) if use_secondary else 0.0
    a["k_primary_edge"] = (1700 + 3260 * d_bar) if use_primary else 0.0
    return a


def measured_traffic():
    """DRAM bytes per step of each kernel from the committed ncu captures (profiles/r01_dram_traffic.json), or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))["dram_bytes_per_step"]
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        # NVML in-process (microseconds per sample); forking nvidia-smi every 200 ms perturbed the host-timed e2e leg
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            while not self.stop_flag:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))] +
                                 ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.05)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def make_scene(api, scenes, device, pose=0, pinned_host=False):
    sc = scenes.shadow_blocker(device if not pinned_host else torch.device("cpu"), resolution=(RES, RES), grad=True)
    if pose:
        import math
        a = 0.05 * pose
        sc.camera.position = torch.tensor([5.0 * math.sin(a), 2.0, -5.0 * math.cos(a)])
    return sc


def run_ours(args, rank, world, local_rank):
    from redner_b200 import api, dist as rdist
    from redner_b200 import redner as rb
    import scenes
    dev = torch.device("cuda:%d" % local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    st = rb.SamplerType.sobol
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    # ---------------- device-resident throughput (value) ----------------
    sc = make_scene(api, scenes, dev, pose=rank if args.mode == "poses" else 0)
    params = [sc.shapes[1].vertices, sc.materials[0].diffuse_reflectance.texels]

    def step_resident(timed):
        for p in params:
            p.grad = None
        sc.area_lights[0].intensity.grad = None
        if world > 1 and args.mode == "tiles":
            fargs = api.RenderFunction.serialize_scene(sc, SPP, MB, sampler_type=st, device=dev)
            t_build0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            img = rdist.TileRenderFunction.apply(SEED, None, 16, *fargs)
            img.pow(2).sum().backward()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1), None, (time.perf_counter() - t_build0) * 1e3
        fargs = api.RenderFunction.serialize_scene(sc, SPP, MB, sampler_type=st, device=dev)
        # time exactly the two render() calls with CUDA events on the render stream
        ctx = type("Ctx", (), {})()
        e0, e1, e1b, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        t0 = time.perf_counter()
        c = api.RenderFunction._unpack((SEED, SEED + 1000003), fargs)  # scene construction (BVH, lights, edge tree)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        nch = rb.compute_num_channels(c.channels, c.scene.max_generic_texture_dimension)
        img = torch.zeros(RES, RES, nch, device=dev)
        e0.record()
        rb.render(c.scene, c.options, rb.float_ptr(img.data_ptr()), rb.float_ptr(0), None, rb.float_ptr(0), rb.float_ptr(0))
        e1.record()
        fwd_stats = c.scene.last_stage_stats()
        d_img = (2 * img).contiguous()
        ctx.c, ctx.args = c, fargs
        e1b.record()
        grads = api.RenderFunction.backward(ctx, d_img)
        e2.record()
        torch.cuda.synchronize()
        bwd_stats = c.scene.last_stage_stats()
        if world > 1:  # poses mode: data-parallel gradient exchange
            tens = [g.to(dev) for g in grads if isinstance(g, torch.Tensor)]
            e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e3.record()
            rdist.all_reduce_packed(tens)
            e4.record()
            torch.cuda.synchronize()
            comm_ms = e3.elapsed_time(e4)
        else:
            comm_ms = 0.0
        t_ms = e0.elapsed_time(e1) + e1b.elapsed_time(e2) + comm_ms
        info = dict(fwd_ms=e0.elapsed_time(e1), bwd_ms=e1b.elapsed_time(e2), comm_ms=comm_ms, build=c.scene.build_ms(), fwd_k=fwd_stats[0],
                    bwd_k=bwd_stats[0], vertices=bwd_stats[1], hits=bwd_stats[2], launches=c.scene.last_stats()[0])
        return t_ms, info, build_ms

    for _ in range(args.warmup):
        step_resident(False)
        flush.zero_()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    times, infos, builds = [], [], []
    for _ in range(args.steps):
        flush.zero_()  # evict the previous step's working set from L2
        torch.cuda.synchronize()
        t, info, b = step_resident(True)
        times.append(t)
        infos.append(info)
        builds.append(b)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    total_ms = sum(times)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        total_ms = tt.item()
    ms_per_step = total_ms / args.steps
    samples_per_step = RES * RES * SPP * (world if args.mode == "poses" else 1)
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---------------- end to end from pinned host memory (e2e) ----------------
    host = make_scene(api, scenes, dev, pose=rank if args.mode == "poses" else 0, pinned_host=True)
    host_tensors = {"floor_v": host.shapes[0].vertices, "floor_i": host.shapes[0].indices, "blk_v": host.shapes[1].vertices.detach(),
                    "blk_i": host.shapes[1].indices, "light_v": host.shapes[2].vertices, "light_i": host.shapes[2].indices,
                    "kd0": host.materials[0].diffuse_reflectance.texels.detach(), "kd1": host.materials[1].diffuse_reflectance.texels}
    host_tensors = {k: v.pin_memory() for k, v in host_tensors.items()}
    h2d = sum(v.numel() * v.element_size() for v in host_tensors.values())

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host_tensors.items()}
        blk = d["blk_v"].requires_grad_(True)
        kd0 = d["kd0"].requires_grad_(True)
        m0, m1 = api.Material(diffuse_reflectance=kd0), api.Material(diffuse_reflectance=d["kd1"])
        shapes = [api.Shape(d["floor_v"], d["floor_i"], 0), api.Shape(blk, d["blk_i"], 0), api.Shape(d["light_v"], d["light_i"], 1)]
        inten = torch.tensor([1000.0, 1000.0, 1000.0], requires_grad=True)
        scn = api.Scene(host.camera, shapes, [m0, m1], [api.AreaLight(2, inten)])
        if world > 1 and args.mode == "tiles":
            img = rdist.render_tiles(scn, SPP, MB, SEED, sampler_type=st, device=dev)
        else:
            img = api.RenderFunction.apply(SEED, *api.RenderFunction.serialize_scene(scn, SPP, MB, sampler_type=st, device=dev))
        loss = img.pow(2).sum()
        loss.backward()
        outs = [loss.detach().cpu(), blk.grad.cpu(), kd0.grad.cpu(), inten.grad]
        return sum(o.numel() * o.element_size() for o in outs)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        d2h = step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = tt.item()
    e2e_value = samples_per_step / (e2e_ms * 1e-3) / 1e6
    if rank == 0:
        clocks.stop_flag = True
        clocks.join(timeout=2)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    roofline = None
    info = next((i for i in reversed(infos) if i), None)
    cfg = {"workload": "C2 shadow_blocker %dx%dx%dspp max_bounces=%d sobol, primary+secondary edge sampling, loss=sum(img^2)" % (RES, RES, SPP, MB),
           "parallelism": ("single GPU" if world == 1 else ("%d ranks, %s, NCCL all-reduce" % (world, args.mode))), "l2": "256 MB flush between timed steps",
           "scene_build_ms": sum(builds) / len(builds) if builds and builds[0] is not None else None}
    if info:
        n_samples = RES * RES * SPP
        d_bar = info["vertices"] / n_samples
        hit_frac = info["hits"] / n_samples
        alg = bytes_per_sample(d_bar, hit_frac, True, True)
        kms = {"k_forward": info["fwd_k"]["k_forward"], **{k: v for k, v in info["bwd_k"].items() if k in alg and k != "k_forward"}}
        traffic = measured_traffic()
        # the boundary-term stage is timed as a whole: compaction + k_bwd_sec_pick + radix sort + k_bwd_sec_shade
        stage_kernels = {"k_bwd_secondary": ("k_bwd_sec_pick", "k_bwd_sec_shade", "k_bwd_compact"), "k_primary_edge": ("k_primary_edge", "k_prim_keys")}
        for stage, names in stage_kernels.items():
            if any(n in traffic for n in names):
                traffic[stage] = sum(traffic.get(n, 0.0) for n in names)
        per_kernel = {k: {"ms": kms[k], "algorithmic_GB": alg[k] * n_samples / 1e9, "achieved_GBps": alg[k] * n_samples / (kms[k] * 1e-3) / 1e9,
                          "dram_GB_measured": (traffic[k] / 1e9 if k in traffic else None)} for k in kms if kms[k] > 0}
        dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
        achieved = per_kernel[dom]["achieved_GBps"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                    "traffic": traffic.get(dom), "peak_source": peak_src, "algorithmic_bytes_per_sample": alg[dom], "kernel_ms": kms[dom],
                    "mean_bounces_per_sample": d_bar, "per_kernel": per_kernel,
                    "note": "achieved = SURVEY.md 8(d) algorithmic bytes of the dominant kernel per step / its CUDA-event time summed over the "
                            "step's band launches; traffic = dram read+write bytes of those launches (ncu, profiles/).  The kernels keep the "
                            "reference's per-stage state in registers, so DRAM traffic is far below the algorithmic bytes: they are "
                            "issue/latency bound (DESIGN.md section 3), the HBM fraction is the contract's metric, not the limiter"}
        cfg["kernel_ms"] = {"k_forward": info["fwd_k"]["k_forward"], **{k: v for k, v in info["bwd_k"].items() if k != "k_forward"}}
        cfg["scene_build_detail_ms"] = info["build"]
        cfg["fwd_ms"], cfg["bwd_ms"], cfg["comm_ms"] = info["fwd_ms"], info["bwd_ms"], info["comm_ms"]
    out = {"metric": "fwd+bwd megasamples/s at 512x512x64spp", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.mode == "poses" else "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": cfg, "clocks": clocks.summary(),
           "e2e": {"value": e2e_value, "unit": "Msamples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": (info["launches"] + 1) * args.steps if info else 0, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(budget_s=20.0)
    print(json.dumps(out))


def reference_step(ref, res, spp):
    """One fwd+bwd of the C2 workload with the unmodified reference (CPU/Embree); returns seconds of the two render calls."""
    from redner_b200 import api
    import scenes
    dev = torch.device("cpu")
    sc = scenes.shadow_blocker(dev, resolution=(res, res), grad=True)
    fargs = api.RenderFunction.serialize_scene(sc, spp, MB, sampler_type=ref.SamplerType.sobol, device=dev, backend=ref)
    c = api.RenderFunction._unpack((SEED, SEED + 1000003), fargs)
    nch = ref.compute_num_channels(c.channels, c.scene.max_generic_texture_dimension)
    img = torch.zeros(res, res, nch)
    t0 = time.perf_counter()
    ref.render(c.scene, c.options, ref.float_ptr(img.data_ptr()), ref.float_ptr(0), None, ref.float_ptr(0), ref.float_ptr(0))
    t1 = time.perf_counter()
    ctx = type("Ctx", (), {})()
    ctx.c, ctx.args = c, fargs
    d_img = (2 * img).contiguous()
    api.RenderFunction.backward(ctx, d_img)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    return ref_loader.load()


def pick_sample(ref, budget_s):
    """Choose a bounded sample (res x res x spp) of the C2 workload that costs about budget_s seconds per step."""
    f, b = reference_step(ref, 64, 4)
    per_sample = (f + b) / (64 * 64 * 4)
    for res, spp in ((512, 64), (512, 16), (256, 16), (256, 8), (128, 16), (128, 8), (128, 4), (64, 4)):
        if per_sample * res * res * spp <= budget_s:
            return res, spp
    return 64, 4


def cpu_baseline(budget_s):
    try:
        ref = load_reference()
    except Exception as e:  # oracle/_ref not built on this machine
        return {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: %s" % e}
    res, spp = pick_sample(ref, budget_s)
    f, b = reference_step(ref, res, spp)
    return {"value": res * res * spp / (f + b) / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": "C2 shadow_blocker %dx%dx%dspp (same scene, sampler, edge sampling and loss; reduced size), fwd %.2fs bwd %.2fs" % (res, res, spp, f, b),
            "fwd_s": f, "bwd_s": b}


def run_reference(args, rank, world):
    if rank != 0:
        return
    try:
        ref = load_reference()
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built: %s" % str(e).splitlines()[0]}))
        return
    budget = max(2.0, 150.0 / max(1, args.steps + args.warmup))
    res, spp = pick_sample(ref, budget)
    for _ in range(args.warmup):
        reference_step(ref, res, spp)
    ts = [reference_step(ref, res, spp) for _ in range(args.steps)]
    f = sum(t[0] for t in ts) / len(ts)
    b = sum(t[1] for t in ts) / len(ts)
    v = res * res * spp / (f + b) / 1e6
    sample = "C2 shadow_blocker %dx%dx%dspp per step (bounded sample of the 512x512x64 workload), fwd %.2fs bwd %.2fs" % (res, res, spp, f, b)
    print(json.dumps({"impl": "reference", "metric": "fwd+bwd megasamples/s at 512x512x64spp", "value": v, "unit": "Msamples/s", "n_gpus": world,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": (f + b) * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "C2 shadow_blocker sobol max_bounces=1 primary+secondary edge sampling, loss=sum(img^2)", "sample": sample,
                                 "parallelism": "reference CPU/Embree path, %d host threads" % (os.cpu_count() or 1)},
                      "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": sample},
                      "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="poses", choices=["poses", "tiles"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the redner_b200 render path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    run_ours(args, rank, world, local_rank)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
