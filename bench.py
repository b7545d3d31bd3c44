#!/usr/bin/env python
"""bench.py -- LiDAR range-view frames/sec (forward + backward) of the rasterizer hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json `metric`, configs[2] = SURVEY.md §8d "cfg 3"): 2 M Gaussians, 64 x 2048 range
image, one "step" = one full operator train step = visible_filter on P/6 anchors + forward + backward
with a fixed random upstream gradient; seeded synthetic data.  With N > 1 every rank renders its own
frame (own sensor pose) of the same replicated Gaussians and the 13 P fp32 parameter gradients are
summed with ONE NCCL all-reduce per step (weak scaling: N frames per step).

value      : frames/s, inputs resident in HBM, through the C ABI (include/lgs_rasterizer.h)
e2e        : frames/s through the reference-facing Python operator (diff_lidargs_rasterization.
             GaussianRasterizer, autograd) with HOST inputs: per step H2D of all Gaussian attributes from
             pinned memory and D2H of images + gradients
roofline   : dominant kernel, algorithmic bytes / CUDA-event time measured inside the timed region
cpu_baseline / --impl reference : the CPU restatement of the reference algorithm (oracle/, C + OpenMP,
             all host threads) -- the reference itself ships no CPU rasterizer.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lidar-gs_b200"))

METRIC = "LiDAR range-view frames/sec (fwd+bwd) @2M Gaussians, 64x2048"
CFG = 3
WORKLOAD = "cfg3: 2M Gaussians, 64x2048, visible_filter(P/6 anchors)+forward+backward, seeded synthetic"


def rank_pose(sc, rank):
    """Sensor pose of rank k: the shared world Gaussians seen from a pose shifted / yawed by k."""
    if rank == 0:
        return sc["viewmatrix"]
    yaw = np.deg2rad(5.0 * rank)
    c, s = np.cos(yaw), np.sin(yaw)
    W2L = np.eye(4)
    W2L[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    W2L[:3, 3] = [0.5 * rank, -0.25 * rank, 0.0]
    return np.ascontiguousarray((W2L @ sc["viewmatrix"].T.astype(np.float64)).T, dtype=np.float32)


def make_anchors(sc, seed=99):
    """A = P/6 anchors with scales[A, 6] (prefilter_voxel passes the slice [:, :3], gaussian_renderer:252)."""
    rng = np.random.default_rng(seed)
    A = sc["P"] // 6
    return dict(means=np.ascontiguousarray(sc["means3D"][:A]),
                scales6=np.exp(rng.uniform(np.log(0.05), np.log(0.5), (A, 6))).astype(np.float32),
                rots=np.ascontiguousarray(sc["rotations"][:A]))


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        self.marks = []

    def mark(self):
        self.marks.append(time.time())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


def reference_cuda_timing(sc, anc, iters=20):
    """Context for the reference arm: the UNMODIFIED reference CUDA rasterizer (oracle/_ref, built from
    /root/reference by oracle/build_ref.py) on the same step, inputs resident -- reported beside the CPU figure,
    never as `value`.  None when the build or a GPU is absent."""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        import build_ref
        import make_goldens as MG
        ref = build_ref.load()
        if ref is None:
            return None
        dev = torch.device("cuda:0")
        d = MG.to_dev(sc, dev)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        am, a6, ar = t(anc["means"]), t(anc["scales6"]), t(anc["rots"])
        empty = torch.Tensor([])

        def step():
            ref.rasterize_aussians_filter(am, a6[:, :3], ar, 1.0, empty, d["viewmatrix"], d["projmatrix"], d["campos"], 1.0, 1.0,
                                          int(sc["H"]), int(sc["W"]), d["beams"], False, int(sc["far"]), int(sc["near"]), False)
            MG.run_ref(ref, sc, dev, d=d)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"ms_per_step": ms, "frames_per_s": 1e3 / ms, "steps": iters,
                "note": "reference CUDA source compiled for sm_100a, same cfg3 step (filter+fwd+bwd), inputs resident"}
    except Exception as e:  # context only: never fail the reference arm on it
        return {"unavailable": repr(e)[:200]}


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement (oracle/) with all host threads; rank 0 only."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread it can get
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lgs_oracle as O
    from lgs_b200 import synth
    sc = synth.make_config(CFG)
    anc = make_anchors(sc)
    cores = O.num_threads()

    def step():
        O.visible_filter(sc, means3D=anc["means"], scales=np.ascontiguousarray(anc["scales6"][:, :3]), rotations=anc["rots"])
        f = O.Forward(sc)
        f.backward(sc["g_color"], sc["g_depth"], sc["g_occ"])
        f.close()

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / max(args.steps, 1)
    v = 1.0 / dt
    extra = {"reference_cuda_sm100a": reference_cuda_timing(sc, anc)}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": native_config(sc["P"], sc["H"], sc["W"], anc["means"].shape[0], 8, args.gpus, None),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "every step = one full cfg3 frame (filter+fwd+bwd) on the CPU restatement "
                                       "oracle/lgs_oracle.c (C + OpenMP); the reference ships no CPU rasterizer"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": extra}
    print(json.dumps(line), flush=True)


def native_config(P, H, W, A, rows_per_bin, world, exchange):
    """`config` of the JSON line -- the same keys for the native and the reference arm (the driver compares them)."""
    return {"workload": WORKLOAD, "P": int(P), "H": int(H), "W": int(W), "anchors": int(A), "rows_per_bin": int(rows_per_bin),
            "parallelism": (f"frames sharded over {world} GPU(s), one exchange of the 13P fp32 parameter grads per step"
                            + (f": {exchange}" if exchange else "")) if world > 1 else "1 GPU",
            "l2": "no explicit flush: per-step working set (inputs 104 MB + records 128 MB + lists + 296 MB of "
                  "gradient buffers) exceeds the 126 MB L2"}


SURFEL_METRIC = "LiDAR range-view frames/sec (fwd+bwd), surfel path @5M surfels, 128x2048"
SURFEL_WORKLOAD = "cfg5: 5M surfels (diff_lidargs_surfel_rasterization), 128x2048, forward+backward, seeded synthetic"


def surfel_reference_cuda_timing(sc, iters=5):
    """The reference surfel CUDA rasterizer (oracle/_ref/lidargs_surfel_ref_C.so) on the same step, inputs resident."""
    try:
        import torch
        import build_ref
        import make_goldens_surfel as MG
        ref = build_ref.load_surfel()
        if ref is None or not torch.cuda.is_available():
            return None
        dev = torch.device("cuda:0")
        d = MG.to_dev(sc, dev)
        for _ in range(2):
            MG.run_ref(ref, sc, dev, d=d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            MG.run_ref(ref, sc, dev, d=d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"ms_per_step": ms, "frames_per_s": 1e3 / ms, "steps": iters,
                "note": "reference surfel CUDA source compiled for sm_100a (printf silenced), same cfg5 step (fwd+bwd), inputs resident"}
    except Exception as e:
        return {"unavailable": repr(e)[:200]}


def run_surfel(args, rank, world, local):
    """--workload surfel: BASELINE config 5 (5M surfels, 128x2048, forward + backward) on the surfel path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from lgs_b200 import synth
    sc = synth.make_surfel_config(5)
    P, H, W = sc["P"], sc["H"], sc["W"]
    if args.impl == "reference":
        if rank != 0:
            return
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1
        import lgs_oracle_surfel as S
        steps, warm = (2 if args.steps is None else args.steps), (1 if args.warmup is None else args.warmup)

        def cstep():
            f = S.Forward(sc)
            f.backward(sc["g_color"], sc["g_others"])
            f.close()
        for _ in range(warm):
            cstep()
        t0 = time.time()
        for _ in range(steps):
            cstep()
        dt = (time.time() - t0) / max(steps, 1)
        v = 1.0 / dt
        print(json.dumps({"impl": "reference", "metric": SURFEL_METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": SURFEL_WORKLOAD},
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": S.num_threads(), "kind": "port",
                                           "sample": "every step = one full cfg5 frame (fwd+bwd) on oracle/lgs_oracle_surfel.c (C + OpenMP)"},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "extra": {"reference_cuda_sm100a": surfel_reference_cuda_timing(sc)}}), flush=True)
        return
    steps = 100 if args.steps is None else args.steps
    warm = 5 if args.warmup is None else max(args.warmup, 3)
    import torch
    import torch.distributed as dist
    from lgs_b200 import capi
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = capi.load()
    L.lgs_set_rows_per_bin(args.rows_per_bin)
    if args.forward_mode is not None:
        L.lgs_set_forward_split(args.forward_mode)
    if args.no_order_history:
        L.lgs_set_order_history(0)
    sc["viewmatrix"] = rank_pose(sc, rank if args.pose_rank is None else args.pose_rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = {k: t(v) for k, v in sc.items() if isinstance(v, np.ndarray)}
    out = dict(color=torch.empty((2, H, W), device=dev), others=torch.empty((7, H, W), device=dev),
               radii=torch.empty((P,), dtype=torch.int32, device=dev))
    bucket = torch.empty(12 * P, device=dev)  # means3D 3, scales 2, rot 4, opacity 1, colours 2: the all-reduce message
    views, o = {}, 0
    for name, c in (("means3D", 3), ("scales", 2), ("rotations", 4), ("opacities", 1), ("colors", 2)):
        views[name] = bucket[o:o + c * P].view(P, c)
        o += c * P
    grads = dict(views, means2D=torch.empty((P, 4), device=dev), transMat=None, depth=None,
                 scratch=torch.empty(L.lgs_surfel_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))
    fr = capi.SurfelFrame(dev)

    def step():
        fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], d["viewmatrix"], d["beams"],
                   H, W, sc["far"], sc["near"], out=out)
        fr.backward(d["g_color"], d["g_others"], grads=grads)
        if world > 1:
            dist.all_reduce(bucket)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(warm):
        step()
    barrier()
    capi.timing_enable(True)
    n0 = L.lgs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    launches = L.lgs_launch_count() - n0
    stages = capi.timing_collect()
    capi.timing_enable(False)
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / steps
    value = world * 1e3 / ms_per_step
    R, V, Ninst = fr.num_rendered, int((out["radii"] > 0).sum().item()), fr.num_instances

    e2e = None
    if not args.no_e2e:
        import diff_lidargs_surfel_rasterization as dlr
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_in = {k: pin(sc[k]) for k in ("means3D", "scales", "rotations", "opacities", "colors")}
        h_out = dict(color=torch.empty((2, H, W)).pin_memory(), others=torch.empty((7, H, W)).pin_memory(),
                     bucket=torch.empty(12 * P).pin_memory(), means2D=torch.empty((P, 4)).pin_memory())
        settings = dlr.GaussianRasterizationSettings(
            image_height=H, image_width=W, bg=d["bg"], scale_modifier=1.0, depth_threshold=0.37, viewmatrix=d["viewmatrix"],
            projmatrix=d["projmatrix"], sh_degree=1, campos=d["campos"], prefiltered=False, beam_inclinations=d["beams"],
            lidar_far=sc["far"], lidar_near=sc["near"], debug=False)
        rast = dlr.GaussianRasterizer(settings)
        h2d = sum(v.numel() * v.element_size() for v in h_in.values())
        d2h = sum(v.numel() * v.element_size() for v in h_out.values())

        def estep():
            g = {k: v.to(dev, non_blocking=True).requires_grad_(True) for k, v in h_in.items()}
            m2d = torch.zeros((P, 4), device=dev, requires_grad=True)
            color, radii, others, _pix = rast(means3D=g["means3D"], means2D=m2d, shs=None, colors_precomp=g["colors"],
                                              opacities=g["opacities"], scales=g["scales"], rotations=g["rotations"],
                                              cov3D_precomp=None)
            torch.autograd.backward([color, others], [d["g_color"], d["g_others"]])
            flat = torch.cat([g[k].grad.reshape(-1) for k in ("means3D", "scales", "rotations", "opacities", "colors")])
            if world > 1:
                dist.all_reduce(flat)
            h_out["color"].copy_(color.detach(), non_blocking=True)
            h_out["others"].copy_(others.detach(), non_blocking=True)
            h_out["bucket"].copy_(flat, non_blocking=True)
            h_out["means2D"].copy_(m2d.grad, non_blocking=True)

        ke = max(3, min(steps, 20))
        for _ in range(2):
            estep()
        barrier()
        e0.record()
        for _ in range(ke):
            estep()
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms_e], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt.item())
        e2e = {"value": world * 1e3 / (ms_e / ke), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": ke, "ms_per_step": ms_e / ke,
               "api": "diff_lidargs_surfel_rasterization.GaussianRasterizer (autograd), pinned host buffers, one stream"}
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    HW = H * W
    nbins = ((W + 15) // 16) * ((H + 7) // 8)
    alg = {  # algorithmic bytes per launch (surfel kernels)
        "clear": 80.0 * P + 8.0 * nbins * 64,
        "project": 48.0 * P + 80.0 * V + 16.0 * P + 4.0 * P,
        "scan": 8.0 * nbins * 64,
        "scatter": 16.0 * P + 16.0 * Ninst,
        "render_fwd": (32.0 + 80.0) * Ninst + 76.0 * HW,
        "render_bwd": (16.0 + 80.0) * Ninst + 76.0 * HW + 80.0 * V,
        "finalize_bwd": (80.0 + 40.0 + 4.0) * P + 48.0 * P + 16.0 * P,
    }
    per = {}
    for s_, (tot_ms, n) in stages.items():
        if n and s_ in alg:
            per[s_] = {"ms_per_step": tot_ms / steps, "launches_per_step": n / steps, "gbs": alg[s_] / (tot_ms / steps * 1e-3) / 1e9}
    dom = max((s_ for s_ in per if s_ != "clear"), key=lambda s_: per[s_]["ms_per_step"])
    roof = {"bound": "hbm", "kernel": dom, "achieved": per[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": per[dom]["gbs"] / peak,
            "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "alg_bytes_per_launch": alg[dom], "avg_launch_ms": per[dom]["ms_per_step"] / max(per[dom]["launches_per_step"], 1),
            "note": "render byte counts are an upper bound (every instance); the lazy walk reads only the consumed prefixes"}
    cpu = None
    if world == 1 and not args.no_cpu:
        import lgs_oracle_surfel as S
        ts = []
        for _ in range(3):
            t0 = time.time()
            f = S.Forward(sc)
            f.backward(sc["g_color"], sc["g_others"])
            f.close()
            ts.append(time.time() - t0)
        cpu = {"value": 1.0 / float(np.median(ts)), "unit": "frames/s", "cores": S.num_threads(), "kind": "port",
               "sample": f"3 full cfg5 frames (fwd+bwd), median of {['%.2f' % x for x in ts]} s, oracle/lgs_oracle_surfel.c (C + OpenMP)"}
    line = {"metric": SURFEL_METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": SURFEL_WORKLOAD, "P": P, "H": H, "W": W, "parallelism": f"{world} GPU(s)",
                       "l2": "no explicit flush: per-step working set (inputs 220 MB + records 400 MB + 400 MB accumulator + gradients) exceeds the 126 MB L2"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "extra": {"num_rendered": R, "num_visible": V, "num_instances": Ninst, "stages": per}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_workloads(dev, L, steps=30):
    """extra.workloads: how the frame time moves with the workload -- BASELINE config 2, config 3 with opacities that never
    saturate a ray, config 3 from a shifted pose (what rank 7 renders), config 5 (surfels) -- each next to the UNMODIFIED
    reference CUDA source (oracle/_ref, compiled for sm_100a) on the same inputs, forward + backward, inputs resident."""
    import torch
    from lgs_b200 import capi, synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    out = {}

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in evs])
        return {"ms_median": float(np.median(t)), "ms_p10": float(np.percentile(t, 10)), "ms_p90": float(np.percentile(t, 90)), "steps": n}

    def ours_3d(sc):
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
        fr = capi.Frame(dev)
        P, H, W = sc["P"], sc["H"], sc["W"]
        o = dict(color=torch.empty((2, H, W), device=dev), depth=torch.empty((1, H, W), device=dev),
                 occ=torch.empty((1, H, W), device=dev), radii=torch.empty((P,), dtype=torch.int32, device=dev))
        f = lambda *s_: torch.empty(s_, dtype=torch.float32, device=dev)
        g = dict(means2D=f(P, 4), opacities=f(P, 1), colors=f(P, 2), means3D=f(P, 3), cov3D=None, scales=f(P, 3), rotations=f(P, 4),
                 scratch=torch.empty(L.lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))

        def fn():
            fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], d["viewmatrix"], d["beams"],
                       H, W, sc["far"], sc["near"], out=o)
            fr.backward(d["g_color"], d["g_depth"], d["g_occ"], grads=g)
        r = timed(fn, steps)
        r.update(num_rendered=int(fr.num_rendered), num_instances=int(fr.num_instances))
        return r

    def ref_3d(sc, n=5):
        try:
            import build_ref
            import make_goldens as MG
            ref = build_ref.load()
            if ref is None:
                return None
            d = MG.to_dev(sc, dev)
            return timed(lambda: MG.run_ref(ref, sc, dev, d=d), n)
        except Exception as e:
            return {"unavailable": repr(e)[:160]}

    def entry(name, ours, ref):
        e = {"this_repo": ours, "reference_cuda_sm100a": ref}
        if ref and "ms_median" in ref:
            e["speedup_vs_reference_cuda"] = ref["ms_median"] / ours["ms_median"]
        out[name] = e
        torch.cuda.empty_cache()

    sc = synth.make_config(2)
    entry("cfg2: 500k Gaussians, 64x1024, fwd+bwd", ours_3d(sc), ref_3d(sc))
    sc = synth.make_config(CFG)
    lo = dict(sc)
    lo["opacities"] = np.random.default_rng(7).uniform(0.01, 0.1, sc["opacities"].shape).astype(np.float32)
    entry("cfg3 with opacities U(0.01, 0.1) (rays do not saturate), fwd+bwd", ours_3d(lo), ref_3d(lo))
    p7 = dict(sc)
    p7["viewmatrix"] = rank_pose(sc, 7)
    entry("cfg3 from the pose rank 7 renders, fwd+bwd", ours_3d(p7), ref_3d(p7))
    wall = dict(sc)  # a surface at one range: every bin's list sits in one or two depth buckets of thousands of entries
    m = sc["means3D"].astype(np.float64)
    m = m / np.linalg.norm(m, axis=1, keepdims=True) * (30.0 + np.random.default_rng(8).uniform(-0.5, 0.5, (m.shape[0], 1)))
    wall["means3D"] = np.ascontiguousarray(m, np.float32)
    entry("cfg3 Gaussians moved onto a 1 m thick shell at 30 m (oversized depth buckets), fwd+bwd", ours_3d(wall), ref_3d(wall))
    del sc, lo, p7, wall, m
    try:
        ss = synth.make_surfel_config(5)
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in ss.items() if isinstance(v, np.ndarray)}
        fr = capi.SurfelFrame(dev)

        def sfn():
            fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], d["viewmatrix"], d["beams"],
                       ss["H"], ss["W"], ss["far"], ss["near"])
            fr.backward(d["g_color"], d["g_others"])
        ours = timed(sfn, 10)
        ours.update(num_rendered=int(fr.num_rendered), num_instances=int(fr.num_instances))
        del fr, d
        torch.cuda.empty_cache()
        ref = surfel_reference_cuda_timing(ss, iters=3)
        if ref and "ms_per_step" in ref:
            ref = {"ms_median": ref["ms_per_step"], "steps": ref["steps"]}
        entry("cfg5: 5M surfels, 128x2048, fwd+bwd (surfel rasterizer)", ours, ref)
    except Exception as e:
        out["cfg5"] = {"unavailable": repr(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "surfel"],
                    help="cfg3 (default, BASELINE.json's metric) or surfel (BASELINE config 5, the second rasterizer)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows-per-bin", type=int, default=0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "sparse", "dense"],
                    help="N > 1: peer = fused pack + pull over NVLink peer memory (default); sparse = NCCL all-gather of the "
                         "packed rows; dense = NCCL all-reduce of the 13P-float bucket")
    ap.add_argument("--dense-allreduce", action="store_true", help="same as --exchange dense")
    ap.add_argument("--serial-exchange", action="store_true",
                    help="N > 1: run the exchange on the frame's own stream instead of a side stream (by default the exchange of "
                         "step k overlaps filter + forward of step k + 1 and is waited for before that step's backward rewrites the bucket)")
    ap.add_argument("--no-workloads", action="store_true", help="skip extra.workloads (other configs / poses, each vs the reference CUDA)")
    ap.add_argument("--forward-mode", type=int, default=None, help="lgs_set_forward_split(mode): 0 one pipelined kernel, 1 split, 2 evaluate/blend warps")
    ap.add_argument("--no-order-history", action="store_true", help="lgs_set_order_history(0): launch order from list lengths only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-dense-grads", action="store_true",
                    help="end-to-end leg copies the dense gradient arrays to the host instead of the non-zero rows")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pose-rank", type=int, default=None, help="debug: render the frame rank K would render")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "surfel":
        return run_surfel(args, rank, world, local)
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args, rank, world)
    args.steps = 200 if args.steps is None else args.steps
    args.warmup = 10 if args.warmup is None else max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from lgs_b200 import capi, dp, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU baseline)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = capi.load()
    L.lgs_set_rows_per_bin(args.rows_per_bin)
    if args.forward_mode is not None:
        L.lgs_set_forward_split(args.forward_mode)
    if args.no_order_history:
        L.lgs_set_order_history(0)

    sc = synth.make_config(CFG)
    sc["viewmatrix"] = rank_pose(sc, rank if args.pose_rank is None else args.pose_rank)
    anc = make_anchors(sc)
    P, H, W = sc["P"], sc["H"], sc["W"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = {k: t(v) for k, v in sc.items() if isinstance(v, np.ndarray)}
    a_means, a_scales6, a_rots = t(anc["means"]), t(anc["scales6"]), t(anc["rots"])
    a_scales3 = a_scales6[:, :3].contiguous()

    # outputs + one flat gradient bucket (the all-reduce message: means3D 3, scales 3, rot 4, opacity 1, colours 2)
    out = dict(color=torch.empty((2, H, W), device=dev), depth=torch.empty((1, H, W), device=dev),
               occ=torch.empty((1, H, W), device=dev), radii=torch.empty((P,), dtype=torch.int32, device=dev))
    bucket = torch.empty(13 * P, device=dev)
    views, o = {}, 0
    for name, c in (("means3D", 3), ("scales", 3), ("rotations", 4), ("opacities", 1), ("colors", 2)):
        views[name] = bucket[o:o + c * P].view(P, c)
        o += c * P
    grads = dict(views, means2D=torch.empty((P, 4), device=dev), cov3D=None,
                 scratch=torch.empty(L.lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))
    fr = capi.Frame(dev)
    xchg, xmode = None, "dense"
    if args.dense_allreduce:
        args.exchange = "dense"
    if world > 1 and args.exchange != "dense":
        xmode = args.exchange
        if xmode == "peer":
            try:
                xchg = dp.PeerExchange(P, dev)
            except Exception as e:  # no peer access between the GPUs of this box: the NCCL path still works
                print(f"[bench] rank {rank}: PeerExchange unavailable ({e!r}); using the NCCL all-gather", file=sys.stderr)
                xmode = "sparse"
            agree = torch.tensor([1 if xmode == "peer" else 0], device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)  # all ranks or none
            if int(agree.item()) == 0 and xmode == "peer":
                xchg.close()
                xmode = "sparse"
        if xmode == "sparse":
            xchg = dp.SparseExchange(P, dev)

    overlap = world > 1 and not args.serial_exchange
    cur = torch.cuda.current_stream(dev)
    xs = torch.cuda.Stream(dev) if overlap else None
    ev_bwd, ev_x = torch.cuda.Event(), torch.cuda.Event()
    ev_x.record(cur)

    def render():
        capi.visible_filter(a_means, a_scales3, a_rots, d["viewmatrix"], d["beams"], H, W, sc["far"], sc["near"])
        fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], d["viewmatrix"],
                   d["beams"], H, W, sc["far"], sc["near"], out=out)
        if overlap:
            cur.wait_event(ev_x)  # the previous step's exchange is done with the bucket this backward rewrites
        fr.backward(d["g_color"], d["g_depth"], d["g_occ"], grads=grads)

    def exchange():
        if xchg is not None:
            xchg.exchange(grads["scratch"], bucket, views)  # the other ranks' rows added locally == all-reduce
        elif world > 1:
            dist.all_reduce(bucket)

    ev = []  # per timed step: (start, after the local frame, after the exchange)

    def step(timed=False):
        if timed:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
        render()
        if timed:
            e[1].record()
        if overlap:
            ev_bwd.record(cur)
            with torch.cuda.stream(xs):
                xs.wait_event(ev_bwd)
                if timed:
                    e[2].record()
                exchange()
                if timed:
                    e[3].record()
                ev_x.record(xs)
        else:
            if timed:
                e[2].record()
            exchange()
            if timed:
                e[3].record()
        if timed:
            ev.append(e)

    def drain():
        if overlap:
            cur.wait_event(ev_x)  # the last exchange belongs to the timed region

    if xchg is not None:  # once, before anything is timed: the sparse exchange must equal the dense all-reduce
        render()
        dense = bucket.clone()
        dist.all_reduce(dense)
        xchg.exchange(grads["scratch"], bucket, views)
        err = float((bucket - dense).abs().max().item()) / max(float(dense.abs().max().item()), 1e-30)
        assert err < 1e-5, f"{xmode} gradient exchange differs from the dense all-reduce: {err}"
        del dense

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    # ---- the timed region: K uninstrumented steps between two events ----
    n0 = L.lgs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    barrier()
    wall = time.time() - t0
    launches = L.lgs_launch_count() - n0
    ms = e0.elapsed_time(e1)
    # ---- the same steps again with the library's per-stage events and per-step events (roofline, percentiles, per-rank split):
    # kept out of the region above so that the instrumentation does not cost the headline anything ----
    nstage = max(10, min(args.steps, 100))
    capi.timing_enable(True)
    for _ in range(nstage):
        step(timed=True)
    barrier()
    stages = capi.timing_collect()
    capi.timing_enable(False)
    if xmode == "peer" and xchg is not None:
        xchg.status()  # raises if any step was skipped on the device (a rank had more rows than the buffer holds)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step
    R, V, Ninst = fr.num_rendered, int((out["radii"] > 0).sum().item()), fr.num_instances
    # per rank: the local frame (filter + forward + backward) and the exchange (which includes waiting for slower ranks)
    f_ms = np.array([e[0].elapsed_time(e[1]) for e in ev]); x_ms = np.array([e[2].elapsed_time(e[3]) for e in ev])
    s_ms = np.array([a_[0].elapsed_time(b_[0]) for a_, b_ in zip(ev[:-1], ev[1:])])  # start of a step to the start of the next
    mine = [float(f_ms.mean()), float(x_ms.mean()), float(np.percentile(s_ms, 10)), float(np.median(s_ms)), float(np.percentile(s_ms, 90)),
            int(L.lgs_last_forward_mode())]
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    per_rank = [dict(rank=r, frame_ms=a[0], exchange_ms=a[1], step_ms_p10=a[2], step_ms_median=a[3], step_ms_p90=a[4], forward_mode=a[5])
                for r, a in enumerate(allr)]
    # Gaussians the backward pass of THIS rank visited (the library's own list), read before anything overwrites the scratch
    import ctypes as _C
    _ids, _cnt = _C.c_void_p(), _C.c_void_p()
    L.lgs_backward_touched(_C.c_void_p(grads["scratch"].data_ptr()), P, _C.byref(_ids), _C.byref(_cnt))
    touched_local = int(dp._device_u32(_cnt.value, dev).item())

    # ---- end to end through the reference-facing operator, host buffers ----
    e2e = None
    if not args.no_e2e:
        import diff_lidargs_rasterization as dlr
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_in = {k: pin(sc[k]) for k in ("means3D", "scales", "rotations", "opacities", "colors")}
        h_anc = dict(means=pin(anc["means"]), scales6=pin(anc["scales6"]), rots=pin(anc["rots"]))
        h_out = dict(color=torch.empty((2, H, W)).pin_memory(), depth=torch.empty((1, H, W)).pin_memory(),
                     occ=torch.empty((1, H, W)).pin_memory(),
                     anchor_radii=torch.empty(anc["means"].shape[0], dtype=torch.int32).pin_memory())
        GRAD_KEYS = ("means3D", "scales", "rotations", "opacities", "colors")
        GRAD_COLS = dict(means3D=3, scales=3, rotations=4, opacities=1, colors=2)
        sparse_rows = not args.e2e_dense_grads
        row_cap = 0
        settings = dlr.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=d["bg"], scale_modifier=1.0,
            viewmatrix=d["viewmatrix"], projmatrix=d["projmatrix"], sh_degree=1, campos=d["campos"], prefiltered=False,
            beam_inclinations=d["beams"], lidar_far=sc["far"], lidar_near=sc["near"], debug=False)
        rast = dlr.GaussianRasterizer(settings)

        def frame_grads(b):
            """forward + backward of one frame through the operator -> (images, anchor radii, dict of dense gradients [P, c])"""
            g = {k: b[k].detach().requires_grad_(True) for k in h_in}
            a_radii = rast.visible_filter(b["a_means"], b["a_scales6"][:, :3], b["a_rots"])
            m2d = torch.zeros((P, 4), device=dev, requires_grad=True)
            color, depth, occ, radii = rast(means3D=g["means3D"], means2D=m2d, shs=None, colors_precomp=g["colors"],
                                            opacities=g["opacities"], scales=g["scales"], rotations=g["rotations"],
                                            cov3D_precomp=None)
            torch.autograd.backward([color, depth, occ], [d["g_color"], d["g_depth"], d["g_occ"]])
            grads = {k: g[k].grad for k in GRAD_KEYS}
            flat = None
            if world > 1 or not sparse_rows:
                flat = torch.cat([grads[k].reshape(-1) for k in GRAD_KEYS])
                if world > 1:
                    dist.all_reduce(flat)
                    o = 0
                    for k in GRAD_KEYS:
                        grads[k] = flat[o:o + GRAD_COLS[k] * P].view(P, GRAD_COLS[k])
                        o += GRAD_COLS[k] * P
            return (color.detach(), depth.detach(), occ.detach()), a_radii, grads, m2d.grad, flat

        if sparse_rows:
            # The host gets every gradient row that is not zero (lossless: dp.unpack_rows() rebuilds the dense arrays) instead of
            # the 17 P floats of the autograd surface.  Size the row buffer from one untimed frame and check the round trip.
            b0 = {k: v.to(dev) for k, v in dict(h_in, a_means=h_anc["means"], a_scales6=h_anc["scales6"], a_rots=h_anc["rots"]).items()}
            _, _, gr0, m2g0, _ = frame_grads(b0)
            probe = dp.pack_nonzero_rows(gr0, m2g0, 0)
            found0 = int(probe[0, :1].view(torch.int32).item())
            row_cap = (int(found0 * 1.25) + 4096 + 4095) // 4096 * 4096
            back, found = dp.unpack_rows(dp.pack_nonzero_rows(gr0, m2g0, row_cap).cpu().numpy(), P)
            assert found == found0 <= row_cap
            for k in GRAD_KEYS:
                assert np.array_equal(back[k], gr0[k].cpu().numpy()), f"sparse read-back differs from the dense gradient: {k}"
            assert np.array_equal(back["means2D"], m2g0.cpu().numpy())
            h_out["grad_rows"] = torch.empty((row_cap + 1, dp.ROW_FLOATS)).pin_memory()
            del b0, gr0, m2g0, probe, back
        else:
            h_out["bucket"] = torch.empty(13 * P).pin_memory()
            h_out["means2D"] = torch.empty((P, 4)).pin_memory()
        h2d = sum(v.numel() * v.element_size() for v in h_in.values()) + sum(v.numel() * v.element_size() for v in h_anc.values())
        d2h = sum(v.numel() * v.element_size() for v in h_out.values())

        # Three-stream pipeline: while step k computes, the inputs of step k + 1 come up from pinned host memory and
        # the results of step k - 1 (images + every gradient) go back; every step's copies are inside the timed region.
        s_in, s_out, cur = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        NBUF = 2
        h_all = dict(h_in, a_means=h_anc["means"], a_scales6=h_anc["scales6"], a_rots=h_anc["rots"])
        d_bufs = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in h_all.items()} for _ in range(NBUF)]
        h_outs = [h_out] + [{k: torch.empty_like(v).pin_memory() for k, v in h_out.items()} for _ in range(NBUF - 1)]
        ev_in = [torch.cuda.Event() for _ in range(NBUF)]
        ev_free = [torch.cuda.Event() for _ in range(NBUF)]
        for e in ev_free:
            e.record(cur)

        dbg = [] if os.environ.get("LGS_E2E_DEBUG") else None

        def upload(i):
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[i])  # the step that last read buffer set i has finished
                if dbg is not None:
                    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ea.record(s_in)
                for k, v in h_all.items():
                    d_bufs[i][k].copy_(v, non_blocking=True)
                if dbg is not None:
                    eb.record(s_in)
                    dbg.append(("h2d", ea, eb))
                ev_in[i].record(s_in)

        def compute_download(i):
            cur.wait_event(ev_in[i])
            if dbg is not None:
                ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ca.record(cur)
            b = d_bufs[i]
            (color, depth, occ), a_radii, grads, m2g, flat = frame_grads(b)
            if sparse_rows:
                rows = dp.pack_nonzero_rows(grads, m2g, row_cap)
            ev_free[i].record(cur)
            done = torch.cuda.Event()
            done.record(cur)
            if dbg is not None:
                cb.record(cur)
                dbg.append(("compute", ca, cb))
                oa, ob = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            outs = dict(anchor_radii=a_radii, color=color, depth=depth, occ=occ)
            if sparse_rows:
                outs["grad_rows"] = rows
            else:
                outs.update(bucket=flat, means2D=m2g)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                if dbg is not None:
                    oa.record(s_out)
                for k, t_ in outs.items():
                    t_.record_stream(s_out)
                    h_outs[i][k].copy_(t_, non_blocking=True)
                if dbg is not None:
                    ob.record(s_out)
                    dbg.append(("d2h", oa, ob))

        def run_pipeline(n):
            upload(0)
            for k in range(n):
                if k + 1 < n:
                    upload((k + 1) % NBUF)
                compute_download(k % NBUF)
            cur.wait_stream(s_out)

        ke = max(4, min(args.steps, 50))
        run_pipeline(4)
        barrier()
        e0.record()
        run_pipeline(ke)
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1)
        if sparse_rows:  # no timed step may have found more rows than the buffer holds
            worst = max(int(h["grad_rows"][0, :1].view(torch.int32).item()) for h in h_outs)
            assert 0 < worst <= row_cap, f"sparse gradient read-back overflowed: {worst} rows > {row_cap}"
        if dbg:
            first = dbg[-3 * ke][1]
            for name, a_, b_ in dbg[-30:]:
                print(f"[e2e] {name:8s} start {first.elapsed_time(a_):8.3f} ms  dur {a_.elapsed_time(b_):6.3f} ms", file=sys.stderr)
        if world > 1:
            tt = torch.tensor([ms_e], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt.item())
        e2e = {"value": world * 1e3 / (ms_e / ke), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": ke, "ms_per_step": ms_e / ke,
               "api": "diff_lidargs_rasterization.GaussianRasterizer (autograd) + visible_filter; H2D / compute / D2H on three "
                      "streams, double buffered",
               "results_read_back": ("images + anchor radii + every non-zero gradient row (80 B rows of lgs_grad_pack_nonzero, "
                                     f"capacity {row_cap}; dp.unpack_rows() == the dense gradients, asserted before timing)")
               if sparse_rows else "images + anchor radii + all dense gradients (17 P floats)"}
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md §"Kernels") ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    A = anc["means"].shape[0]
    HW = H * W
    from lgs_b200.inspect import consumed_entries
    cons = consumed_entries(fr, H, W, args.rows_per_bin)
    cons["touched"] = touched_local  # rank-local for every world size (the exchanged bucket holds the other ranks' rows too)
    alg = {  # bytes per launch
        "clear": 68.0 * P + 16.0 * cons["replayed"] + 80.0 * cons["touched"] + P / 8.0 + 8.0 * cons["nbins"] * 64,
        "project": 52.0 * P + 64.0 * V + 16.0 * P + 4.0 * P,
        "scan": 8.0 * cons["nbins"] * 64,
        "scatter": 16.0 * P + 16.0 * Ninst,
        "render_fwd": 32.0 * cons["sorted"] + 64.0 * cons["sorted"] + 24.0 * HW,
        "render_bwd": 16.0 * cons["replayed"] + 64.0 * cons["replayed"] + 24.0 * HW + 80.0 * cons["touched"],
        "finalize_bwd": (80.0 + 44.0 + 68.0 + 4.0) * cons["touched"],
        "filter": 44.0 * A,
    }
    per = {}
    for s, (tot_ms, n) in stages.items():
        if n:
            per[s] = {"ms_per_step": tot_ms / nstage, "launches_per_step": n / nstage,
                      "gbs": alg[s] / (tot_ms / nstage * 1e-3) / 1e9}
    dom = max((s for s in per if s != "clear"), key=lambda s: per[s]["ms_per_step"])
    traffic, traffic_src, issue = None, None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic, traffic_src = tj.get(dom), tj.get("_source")
        wi = tj.get("_warp_instructions", {})
        if world == 1 and args.pose_rank is None:  # the counts belong to this workload only
            # what actually bounds these kernels: warp instructions issued per second against 148 SMs x 4 schedulers x 1 per clock
            mhz = float(clocks["sm_mhz"]) if clocks and clocks.get("sm_mhz") else 1965.0
            issue = {k: {"warp_instructions": wi[k], "issue_frac": wi[k] / (per[k]["ms_per_step"] / max(per[k]["launches_per_step"], 1) * 1e-3)
                         / (148 * 4 * mhz * 1e6)} for k in wi if k in per}
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": per[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": per[dom]["gbs"] / peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "alg_bytes_per_launch": alg[dom], "avg_launch_ms": per[dom]["ms_per_step"] / max(per[dom]["launches_per_step"], 1),
            "traffic_source": traffic_src,
            "note": "the compositing kernels are bound by instruction issue, not by DRAM: their algorithmic bytes are the "
                    "entries + records of the list prefixes actually consumed (extra.consumed), a few percent of the lists"}
    frame_bytes = 124.0 * P + 276.0 * V + 164.0 * R + 48.0 * HW  # SURVEY.md §8d full-sort byte model
    extra = {"num_rendered": R, "num_visible": V, "num_instances": Ninst, "consumed": cons,
             "stages": per, "issue": issue, "frame_model_bytes": frame_bytes,
             "frame_model_note": "SURVEY 8d byte model of the reference's full-sort algorithm -- NOT traffic this implementation moves",
             "longest_walk_chunks": int(L.lgs_last_longest_walk()),
             "forward_mode": int(L.lgs_last_forward_mode()),  # 0: worker warp per 2 pixel rows, 3: per row (automatic)
             "per_rank": per_rank, "exchange": xmode if world > 1 else None,
             "exchange_overlap": ("side stream: the exchange of step k runs beside filter + forward of step k + 1 and is waited for before "
                                  "that step's backward rewrites the gradient bucket; the last exchange is inside the timed region")
             if overlap else ("same stream" if world > 1 else None),
             "step_ms": {"p10": per_rank[0]["step_ms_p10"], "median": per_rank[0]["step_ms_median"], "p90": per_rank[0]["step_ms_p90"]},
             "instrumented_steps": nstage,  # stages / step_ms / per_rank come from these (per-stage events cost ~0.05 ms a step)
             "wall_s": wall}
    if world == 1 and not args.no_workloads:
        extra["workloads"] = other_workloads(dev, L)

    cpu = None
    if world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lgs_oracle as O
        ts = []
        for _ in range(4):
            t0 = time.time()
            O.visible_filter(sc, means3D=anc["means"], scales=np.ascontiguousarray(anc["scales6"][:, :3]), rotations=anc["rots"])
            f = O.Forward(sc)
            f.backward(sc["g_color"], sc["g_depth"], sc["g_occ"])
            f.close()
            ts.append(time.time() - t0)
        cpu = {"value": 1.0 / float(np.median(ts)), "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
               "sample": f"4 full cfg3 frames (filter+fwd+bwd), median of {['%.2f' % x for x in ts]} s, "
                         "oracle/lgs_oracle.c (C + OpenMP restatement; the reference ships no CPU rasterizer)"}

    xdesc = None
    if world > 1:
        rows = xchg.last["rows"] if xchg is not None else 0
        xdesc = {"peer": f"fused pack + pull of the touched rows over NVLink peer memory (<= {rows} rows x 64 B per rank), no host "
                         "sync, verified equal to the dense all-reduce",
                 "sparse": f"NCCL all-gather of the touched rows ({rows} x 64 B per rank) + local add, verified equal to the dense all-reduce",
                 "dense": "dense NCCL all-reduce"}[xmode]
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": native_config(P, H, W, A, cons["rows_per_bin"], world, xdesc),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "extra": extra}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
