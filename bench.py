#!/usr/bin/env python
"""bench.py -- benchmark of the render hot path (BASELINE.json: Mpaths/s & Mrays/s at 1/2/4/8 B200; headline 1080p @ 64 spp).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C1..C5]      # this repo's CUDA path (default config C2 = the headline)
  python bench.py --impl reference ...                                        # the reference algorithm on the host CPU cores

Workloads = BASELINE.json's five configs as concrete synthetic scenes (computational_ray_tracer_b200/scenes.py CONFIGS).
One "step" = one complete render of the frame at the config's sample count.  The scene is built once (octree on the GPU, same tree as
the reference's builder), flattened and uploaded once.  With N GPUs (one process per GPU) the sample indices are split into N contiguous
ranges -- or, with --partition tiles, the image into interleaved 32x32 tiles -- and the per-GPU films are summed onto rank 0 by ONE
ncclReduce inside the library (crt_film_reduce) within the timed region ("scaling": "strong": the frame is fixed).

The JSON line carries: value (scene resident, CUDA-event timed), e2e (through the C ABI with host buffers), roofline of the dominant
kernel (octree traversal) with its binding resource named, cpu_baseline (the oracle on the host cores, bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
NODE_BYTES, TRI_BYTES, RAY_IN_BYTES, HIT_OUT_BYTES = 32, 48, 32, 16     # SURVEY.md 8(d) "algorithmic bytes per unit of work"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crt", choices=["crt", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"], help="BASELINE.json config (default C2, the headline)")
    ap.add_argument("--spp", type=int, default=None, help="override the config's samples per pixel (profiling runs)")
    ap.add_argument("--quads", type=int, default=None, help="override the height-field resolution of C2 / C4 / C5 (reduced scenes for CPU-only checks)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--mode", type=int, default=1, help="1 = path integrator with NEE (the configs), 0 = the reference's own Li (primary ray + one-bounce shading)")
    ap.add_argument("--trace-mode", type=int, default=3, help="0 exact BFS kernel only, 3 ordered traversal + exact re-trace of order-sensitive rays")
    ap.add_argument("--shade-mode", type=int, default=0, help="0 automatic (staged per material type when the scene has analytic shapes), 1 fused, 2 staged")
    ap.add_argument("--partition", default="spp", choices=["spp", "tiles"])
    ap.add_argument("--host-build", action="store_true", help="build the octree with the host incremental builder instead of the GPU builder")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def config_of(a):
    """The workload: scenes.CONFIGS[a.config] (pure Python / numpy: importing it does not load the CUDA library)."""
    from computational_ray_tracer_b200 import scenes
    c = dict(scenes.CONFIGS[a.config])
    if a.spp is not None:
        c["spp"] = a.spp
        c["label"] = c["label"].rsplit("@", 1)[0] + f"@ {a.spp} spp (of the config's {scenes.CONFIGS[a.config]['spp']})"
    if a.quads is not None and a.config in ("C2", "C4", "C5"):
        q = a.quads
        c["meshes"] = {"C2": lambda: scenes.heightfield(q), "C4": lambda: scenes.many_light_scene(q), "C5": lambda: scenes.heightfield(q, seed=5)}[a.config]
        c["label"] += f" [height field reduced to {q}x{q} quads]"
    if a.width is not None or a.height is not None:
        c["width"], c["height"] = a.width or c["width"], a.height or c["height"]
        c["label"] += f" [film {c['width']}x{c['height']}]"
    c["camera"] = scenes.CAMERA
    return c


def metric_name(c):
    return f"Mpaths/s ({c['width']}x{c['height']}, {c['spp']} spp)"


def config_block(c, mode):
    return {"workload": c["label"], "integrator": c["integrator"] if mode == 1 else "reference Li (primary ray + one-bounce shading)",
            "sampler": f"StratifiedSampler({c['xs']},{c['ys']},jitter)", "filter": "BoxFilter(0.5)"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ host CPU arm (the oracle; tests/oracle_lib.py)
def oracle_camera(c):
    """CameraBase / PerspectiveCamera matrices from the oracle (bit-identical to the compiled reference's and to the product's)."""
    import oracle_lib as O
    k = c["camera"]
    return O.camera_matrices(k["kind"], k["near"], k["far"], 0.0, 0.0, k["fov"], k["pos"], k["look"], k["right"], k["up"], float(c["width"]), float(c["height"]))


def oracle_scene(c, mode):
    import oracle_lib as O
    sc = O.OracleScene()
    sc.set_model(c["meshes"]())
    sc.build_octree()
    if mode == 1:
        sc.set_mesh_materials(c["materials"](sc))
    return sc


def cpu_sample_rate(c, sc, mode, seconds, nthreads, spp, faithful=0):
    """Time the oracle on every `stride`-th pixel x `spp` sample indices, sized for about `seconds` of work."""
    import oracle_lib as O
    r2c, c2w = oracle_camera(c)
    w, h = c["width"], c["height"]
    kw = dict(xs=c["xs"], ys=c["ys"], jitter=1, mode=mode, max_depth=c["max_depth"], rr_depth=c["rr_depth"], nthreads=nthreads, faithful=faithful)
    npix = w * h
    stride = 1021 if npix > 1021 * 64 else 61                  # calibrate on a thin sample
    p = O.make_params(w, h, r2c, c2w, spp_begin=0, spp_end=1, pixel_stride=stride, **kw)
    r = sc.render(p, counters=True)
    rate = max(r["counters"]["paths"], 1) / max(r["seconds"], 1e-6)
    stride = int(max(1, min(4093, round(npix * spp / max(rate * seconds, 1)))))
    while stride > 1 and (w % stride == 0 or stride % 2 == 0):      # avoid sampling whole columns only
        stride += 1
    p = O.make_params(w, h, r2c, c2w, spp_begin=0, spp_end=spp, pixel_stride=stride, **kw)
    r = sc.render(p, counters=True)
    k = r["counters"]
    return dict(paths=k["paths"], rays=k["closest_rays"] + k["shadow_rays"], seconds=r["seconds"], stride=stride, spp=spp,
                nodes_per_ray=k["nodes"] / max(k["rays"], 1), tris_per_ray=k["tris"] / max(k["rays"], 1))


def compiled_reference_tier_a(c, seconds, nthreads):
    """The REFERENCE'S OWN renderer (evaluate_pixel + Li, RayTracerTestApp.h:218-345, compiled unmodified into oracle/_ref/libcrt_ref.so) on
    this workload's mesh: its own octree build, primary ray + one-bounce shading, threaded over static pixel ranges like the reference.
    None where the compiled reference did not travel."""
    import ref_lib as R
    if not R.available():
        return None
    sc = R.RefScene()
    sc.set_model(c["meshes"]())
    t0 = time.time(); nodes = sc.build_octree(); t_build = time.time() - t0
    w, h = c["width"], c["height"]

    def timed(stride, spp):
        p = R.make_params(w, h, sampler_kind=1, xs=c["xs"], ys=c["ys"], jitter=1, spp_begin=0, spp_end=spp, nthreads=nthreads, pixel_stride=stride)
        t = time.perf_counter(); film = sc.render_tier_a(p); dt = time.perf_counter() - t
        return int(film[:, 3].sum()), dt
    n0, dt0 = timed(1021, 1)
    rate = n0 / max(dt0, 1e-6)
    spp = min(c["spp"], 4)
    stride = int(max(1, min(4093, round(w * h * spp / max(rate * seconds, 1)))))
    while stride > 1 and (w % stride == 0 or stride % 2 == 0):
        stride += 1
    n, dt = timed(stride, spp)
    sc.close()
    return {"value": n / dt / 1e6, "unit": "Mpaths/s", "cores": nthreads, "kind": "reference",
            "what": "reference Li (primary ray + one-bounce shading), the reference's own compiled code (oracle/_ref)",
            "sample": f"every {stride}th pixel x {spp} sample indices = {n} paths in {dt:.1f} s",
            "octree_build_s": round(t_build, 2), "octree_nodes": nodes}


def run_reference(a):
    """The reference arm: the reference algorithm on the host cores.  It imports nothing that loads libcrt_b200.so: scene generators
    are numpy code, camera matrices and the renderer come from the oracle (tests/oracle_lib.py -> oracle/_build/liboracle.so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_lib as O
    O.build()
    c = config_of(a)
    mode = a.mode
    nthreads = os.cpu_count() or 1
    sc = oracle_scene(c, mode)
    per_step_seconds = 4.0
    spp = min(c["spp"], 16)
    samples = []
    info = None
    for i in range(a.warmup + a.steps):
        info = cpu_sample_rate(c, sc, mode, per_step_seconds, nthreads, spp)
        if i >= a.warmup:
            samples.append(info)
    tot_paths = sum(s["paths"] for s in samples); tot_rays = sum(s["rays"] for s in samples); tot_s = sum(s["seconds"] for s in samples)
    value = tot_paths / tot_s / 1e6
    sample = f"every {info['stride']}th pixel x {info['spp']} sample indices per step ({info['paths']} paths), oracle port, {nthreads} threads"
    line = {
        "impl": "reference", "metric": metric_name(c), "value": value, "unit": "Mpaths/s", "mrays_per_s": tot_rays / tot_s / 1e6,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot_s / max(a.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(c, mode),
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": nthreads, "kind": "port", "sample": sample,
                         "why_port": "the configs' integrator (multi-bounce path + NEE) does not exist in the reference (Integrator.h is "
                                     "comment-only); the reference's own renderer is timed under reference_tier_a"},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if a.config in ("C2", "C4"):
        try:
            line["cpu_baseline"]["reference_tier_a"] = compiled_reference_tier_a(c, 8.0, nthreads)
        except Exception as e:                                  # the arm's own number must not depend on the optional .so
            line["cpu_baseline"]["reference_tier_a"] = {"unavailable": repr(e)}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm
def ncu_facts(kernel="k_trace_wide"):
    """Counters of the dominant kernel from the committed ncu capture (profiles/r02_traffic.json).  NOT measured by this run: every
    value taken from there is labelled with its source."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(kernel)


def run_crt(a):
    import torch
    import torch.distributed as dist
    from computational_ray_tracer_b200 import api
    from computational_ray_tracer_b200 import build as crt_build

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback; use --impl reference for the host baseline)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # A collective that cannot complete (a rank died, or ranks disagree about how many steps to run) must not sit on eight GPUs for
        # NCCL's default ten minutes: nothing here legitimately waits longer than the one-off build on rank 0.
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        crt_build.build()
    if world > 1:
        dist.barrier()
    c = config_of(a)
    mode, trace_mode = a.mode, a.trace_mode
    w, h, spp = c["width"], c["height"], c["spp"]
    npix = w * h

    # One explicit (non-default) stream carries everything: the library's kernels, its ncclReduce, torch's fills / copies; the CUDA events
    # that time the run are recorded on it.  (Handle 0 would mean "the context's own stream".)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = api.Context(local)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:                                   # the library's own communicator: torch.distributed only carries the 128-byte unique id
        uid = torch.from_numpy(api.Context.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        ctx.nccl_init(world, rank, uid.cpu().numpy())

    # ---- scene (built once, uploaded once)
    meshes = c["meshes"]()
    ms = api.MeshSet(meshes)
    t0 = time.time()
    oct_ = api.Octtree_Model(ms) if a.host_build else api.Octtree_Model(ms, algorithm=api.BUILD_GPU, ctx=ctx)
    t_build = time.time() - t0
    scene = api.Scene(ctx)
    mats = c["materials"](scene) if mode == 1 else None
    scene.set_model(oct_, mesh_materials=mats)
    scene.commit()
    film_t = torch.zeros(npix * 4, dtype=torch.float32, device=dev)
    film = api.Film(ctx, w, h)
    film.attach(film_t.data_ptr())
    host_film = torch.empty(npix * 4, dtype=torch.float32).pin_memory()
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
    k = c["camera"]
    r2c, c2w = api.camera_matrices(k["kind"], k["near"], k["far"], k["fov"], k["pos"], k["look"], k["up"], w, h, right=k["right"])
    part = 1 if a.partition == "spp" else 0
    base = dict(xs=c["xs"], ys=c["ys"], jitter=1, mode=mode, max_depth=c["max_depth"], rr_depth=c["rr_depth"], spp_begin=0, spp_end=spp,
                rank=rank, world=world, partition=part, trace_mode=trace_mode, shade_mode=a.shade_mode)
    cfg = api.make_config(w, h, r2c, c2w, time_kernels=1, **base)
    cfg_stats = api.make_config(w, h, r2c, c2w, collect_stats=1, **dict(base, spp_end=min(spp, 16 * world)))

    def step():
        flush.fill_(1)                              # L2 flush between iterations (252 MiB write)
        film_t.zero_()
        st = scene.render(film, cfg)
        film.reduce(0)                              # one ncclReduce(sum) inside the library; no-op on one GPU
        return st

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(a.warmup):
        step()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = dict(paths=0, closest_rays=0, shadow_rays=0, kernel_launches=0, trace_launches=0, trace_ms=0.0, exact_retraced_rays=0)
    ev0.record(stream)
    for _ in range(a.steps):
        st = step()
        for key in acc:
            acc[key] += st[key]
    ev1.record(stream)
    sync_all()
    ms_total = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    film_sum = float(film_t.double().sum().item()) if rank == 0 else 0.0
    # Every rank must take the same decision and run the same number of steps below (a step contains a collective): both come from the
    # maximum of the ranks' timed regions, not from this rank's own clock.
    ms_ref = ms_total
    if world > 1:
        t_ref = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ref, op=dist.ReduceOp.MAX)
        ms_ref = float(t_ref.item())
    if ms_ref < 1000.0:
        # A timed region shorter than a few nvidia-smi periods (C1: milliseconds) yields no usable clock samples: repeat the same step,
        # untimed, for 1.5 s right after it and sample the clocks over that (every rank does the work, rank 0 samples).
        probe = ClockSampler(local)
        if rank == 0:
            probe.start()
        n_probe = max(3, int(1500.0 / max(ms_ref / a.steps, 1e-3)))
        for _ in range(n_probe):
            step()
        sync_all()
        if rank == 0:
            clk2 = probe.stop()
            if clk2.get("sm_mhz") and (not clk.get("sm_mhz") or clk.get("samples", 0) < 3):
                clk = dict(clk2, note=f"the timed region lasted {ms_total:.1f} ms, shorter than nvidia-smi's sampling period; sampled over {n_probe} untimed "
                                      "repetitions of the same step immediately after it")
    nccl_ok = ctx.nccl_async_error() == 0 if world > 1 else True

    # ---- end to end through the C ABI with host buffers, two ways:
    #      "upload": every step re-sends the whole flattened scene from page-locked host memory (crt_scene_commit) -- the strict reading of
    #                "host->device copy of the step's inputs";
    #      "resident": the scene stays on the device (uploaded once, as the north star words it) and a step's input is its render config.
    #      Both end with the film (rgbsum, weightsum) read back to pinned host memory on rank 0.
    def e2e_step(upload):
        if upload:
            scene.commit()
        film_t.zero_()
        scene.render(film, cfg)
        film.reduce(0)
        if rank == 0:
            host_film.copy_(film_t, non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_ms = {}
    for upload in (True, False):
        e2e_step(upload)
        sync_all()
        t_e0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(upload)
        sync_all()
        e2e_ms[upload] = (time.perf_counter() - t_e0) * 1e3
    scene_bytes = scene.device_bytes()

    # ---- traversal statistics (instrumented kernels, untimed, at most 16 sample indices per GPU) for the algorithmic-bytes model
    film_t.zero_()
    sst = scene.render(film, cfg_stats)
    torch.cuda.synchronize(dev)

    # ---- the reference's OWN integrator (Li, mode 0) on the same scene, device-timed, so that cpu_baseline.reference_tier_a (the compiled
    #      reference on the host) has a like-for-like GPU figure beside it
    tier_a_gpu = None
    if world == 1 and not a.no_cpu_baseline and a.config in ("C2", "C4"):
        cfg_a = api.make_config(w, h, r2c, c2w, **dict(base, mode=0))
        for _ in range(2):
            flush.fill_(1); film_t.zero_(); scene.render(film, cfg_a)
        ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        ea0.record(stream)
        na = 0
        for _ in range(3):
            flush.fill_(1); film_t.zero_()
            na += scene.render(film, cfg_a)["paths"]
        ea1.record(stream)
        torch.cuda.synchronize(dev)
        tier_a_gpu = na / (ea0.elapsed_time(ea1) / 1e3) / 1e6

    # ---- reduce over ranks
    vals = torch.tensor([ms_total, e2e_ms[True], e2e_ms[False], acc["trace_ms"]], dtype=torch.float64, device=dev)
    sums = torch.tensor([acc["paths"], acc["closest_rays"] + acc["shadow_rays"], acc["kernel_launches"], acc["trace_launches"],
                         sst["nodes_visited"], sst["tris_tested"], sst["closest_rays"] + sst["shadow_rays"], acc["exact_retraced_rays"]],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total, e2e_up_ms, e2e_res_ms, trace_ms = [float(x) for x in vals.tolist()]
    paths, rays, launches, trace_launches, nodes, tris, stat_rays, retraced = [float(x) for x in sums.tolist()]

    if rank == 0:
        secs = ms_total / 1e3
        nodes_per_ray = nodes / max(stat_rays, 1); tris_per_ray = tris / max(stat_rays, 1)
        bytes_per_ray = RAY_IN_BYTES + HIT_OUT_BYTES + NODE_BYTES * nodes_per_ray + TRI_BYTES * tris_per_ray
        # dominant kernel = octree traversal; trace_ms is the max over ranks of the per-rank sum of event-bracketed launches
        rays_per_rank = rays / world
        achieved = (rays_per_rank * bytes_per_ray) / (trace_ms / 1e3) / 1e9 if trace_ms > 0 else None
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
        # A one-leaf octree (C1, C3) is traversed inside the shading kernels: there is no traversal launch, and the kernels timed as
        # "traversal" are the ones that contain it (crt_render brackets them when time_kernels is set).
        ost = oct_.stats()
        root_leaf = mode == 1 and trace_mode == 3 and ost["nodes"] == 1 and 0 < ost["max_leaf"] <= 16
        staged = mode == 1 and (a.shade_mode == 2 or (a.shade_mode == 0 and scene.n_shapes > 0 and w * h >= 1 << 18))
        if root_leaf:
            kernel_key = "C3/k_path_hit" if staged else "C1/k_path_shade"
            kernel_name = (("k_path_hit" if staged else "k_path_shade") + " + k_shadow_resolve (the one-leaf octree and the shape hierarchy are traversed "
                           "inside the shading kernels: no traversal launch exists; timed = these kernels, which also form the surface record" +
                           ("" if staged else " and shade the bounce") + ")")
        else:
            kernel_key = "k_trace_wide"
            kernel_name = {0: "k_trace", 3: "k_trace_wide"}[trace_mode] + " (octree closest/any hit)"
        facts = ncu_facts(kernel_key) if trace_mode == 3 else None
        binding = None
        traffic = None
        if facts:
            traffic = facts.get("dram_bytes_per_launch")
            issue_bound = facts["issue_slots_busy_pct"] >= 60.0
            binding = {"resource": "instruction issue" if issue_bound else "latency (warps waiting on dependent loads; issue slots mostly idle)",
                       "frac": facts["issue_slots_busy_pct"] / 100.0 * facts["active_threads_per_warp"] / 32.0,
                       "issue_slots_busy_pct": facts["issue_slots_busy_pct"], "active_threads_per_warp": facts["active_threads_per_warp"],
                       "l1tex_throughput_pct": facts.get("l1tex_throughput_pct"), "l2_throughput_pct": facts.get("l2_throughput_pct"),
                       "dram_throughput_pct": facts.get("dram_throughput_pct"), "achieved_occupancy_pct": facts.get("achieved_occupancy_pct"),
                       "from": facts.get("from"), "measured_by_this_run": False}
        line = {
            "metric": metric_name(c), "value": paths / secs / 1e6, "unit": "Mpaths/s", "mrays_per_s": rays / secs / 1e6,
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_block(c, mode), partition=f"{a.partition} x{world}",
                           traversal={0: "exact BFS (warp per ray)",
                                      3: "ordered, 1 ray/lane descent + pooled super-packet / sub-packet / triangle batches + exact BFS re-trace of order-sensitive rays"}[trace_mode],
                           shading=("staged: surface-record kernel, then one kernel per material type over that type's queue"
                                    if mode == 1 and (a.shade_mode == 2 or (a.shade_mode == 0 and scene.n_shapes > 0 and w * h >= 1 << 18)) else
                                    "fused: surface record + bounce in one kernel" if mode == 1 else "k_shade_li (the reference's Li)"),
                           l2="252 MiB write between steps (L2 flush); the wave state alone exceeds the 126 MB L2",
                           octree=oct_.stats(), octree_build_s=round(t_build, 3),
                           octree_builder="host incremental (reference order)" if a.host_build else "GPU level-synchronous (identical layout)",
                           film_reduce="crt_film_reduce: one in-library ncclReduce(sum) per step" if world > 1 else "none (one GPU)"),
            "e2e": {"value": paths / (e2e_up_ms / 1e3) / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": int(scene_bytes) * world,
                    "d2h_bytes_per_step": npix * 16, "ms_per_step": e2e_up_ms / a.steps,
                    "what": "crt_scene_commit (whole scene host->device from page-locked memory, on every GPU) + crt_render + crt_film_reduce + film download to pinned host",
                    "resident": {"value": paths / (e2e_res_ms / 1e3) / 1e6, "ms_per_step": e2e_res_ms / a.steps, "h2d_bytes_per_step": 312 * world,
                                 "what": "scene uploaded once (north star); per step: render config by value + crt_render + crt_film_reduce + film download"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": ("issue" if binding["resource"] == "instruction issue" else "latency") if binding else "hbm", "kernel": kernel_name,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_from": facts.get("from") if facts else None, "peak_source": peak_src,
                         "what": "achieved = ALGORITHMIC bytes (ray 32 + hit 16 + 32 per box test + 48 per triangle test, counted by the instrumented kernel of "
                                 "this run) / the traversal launches' CUDA-event time of this run; these bytes are served by L1/L2 (the scene is cache "
                                 "resident), so this is not the binding roofline -- `binding` names the resource ncu shows limiting the kernel",
                         "binding": binding,
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "launches": int(trace_launches), "avg_launch_ms": trace_ms / max(trace_launches / world, 1),
                         "kernel_share_of_step": trace_ms / ms_total},
            "exact_retraced_rays": int(retraced),
            "film_checksum": film_sum,
            "nccl": {"version": api.Context.nccl_version()[0], "async_error_free": nccl_ok} if world > 1 else None,
            "clocks": clk,
        }
        if not a.no_cpu_baseline and world == 1:
            import oracle_lib as O
            O.build()
            nthreads = os.cpu_count() or 1
            osc = oracle_scene(c, mode)
            s = cpu_sample_rate(c, osc, mode, a.cpu_seconds, nthreads, min(spp, 16))
            sf = cpu_sample_rate(c, osc, mode, max(a.cpu_seconds / 3, 3.0), nthreads, min(spp, 4), faithful=1)
            ref_bytes = RAY_IN_BYTES + HIT_OUT_BYTES + NODE_BYTES * s["nodes_per_ray"] + TRI_BYTES * s["tris_per_ray"]
            # SURVEY 8(d) fixes the per-ray figure with the REFERENCE algorithm's visit counts (BFS over every pierced leaf); the ordered
            # traversal does not do that work, so this is reported as a comparison only
            line["roofline"]["reference_bfs_bytes_per_ray"] = ref_bytes
            line["roofline"]["reference_bfs_nodes_per_ray"] = s["nodes_per_ray"]
            line["roofline"]["reference_bfs_tris_per_ray"] = s["tris_per_ray"]
            line["cpu_baseline"] = {"value": s["paths"] / s["seconds"] / 1e6, "unit": "Mpaths/s", "cores": nthreads, "kind": "port",
                                    "mrays_per_s": s["rays"] / s["seconds"] / 1e6,
                                    "sample": f"every {s['stride']}th pixel x {s['spp']} sample indices = {s['paths']} paths in {s['seconds']:.1f} s",
                                    "faithful_value": sf["paths"] / sf["seconds"] / 1e6,
                                    "faithful_note": "same algorithm with the reference's per-triangle map lookups / chrono / vertex transforms kept"}
            osc.close()
            if a.config in ("C2", "C4"):
                try:
                    rta = compiled_reference_tier_a(c, min(a.cpu_seconds, 10.0), nthreads)
                except Exception as e:
                    rta = {"unavailable": repr(e)}
                if rta is not None:
                    rta["gpu_value_same_integrator"] = tier_a_gpu
                line["cpu_baseline"]["reference_tier_a"] = rta
        print(json.dumps(line))
    film.close(); scene.close(); oct_.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print (NCCL's version banner, torchrun notices) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import builtins
    _print = builtins.print

    def emit(*args, **kw):
        if kw.get("file") is None and len(args) == 1 and isinstance(args[0], str) and args[0].startswith("{"):
            os.write(real_stdout, (args[0] + "\n").encode())
        else:
            _print(*args, **kw)
    builtins.print = emit
    try:
        if a.impl == "reference":
            return run_reference(a)
        return run_crt(a)
    finally:
        builtins.print = _print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)


if __name__ == "__main__":
    sys.exit(main())
