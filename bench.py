#!/usr/bin/env python
"""bench.py — headline benchmark of the light-transport hot path (BASELINE.json: Mrays/s + PPM iterations/s at
1080p, 1M photons/iteration, 1/2/4/8 B200).

A "step" is one PPM iteration (camera pass, photon pass, photon-grid build, gather, direct + output) of the
synthetic ~1M-triangle textured scene at 1920x1080 with a 1024x1024 photon launch, through the C ABI
(opr_render_next_iteration).  With N > 1 ranks (torchrun, one process per GPU) rank g renders the global
iterations B+g, B+g+N, ... with the closed-form radius schedule and the per-GPU running sums are combined with one
NCCL reduce — issued by the C++ library itself, opr_reduce_output — inside the timed region;
value = rays of all ranks / max-over-ranks device time.

Window.  The PPM radius shrinks as r_i ~ i^(-1/6), and the gather costs ~ r^2: a rank of an 8-GPU job walks the
schedule eight times faster than a single GPU, so windows that start at iteration 0 give every N different work
per step (round 1's curve showed 1.04-1.09 "efficiency" for that reason).  All arms and all N therefore time the
steady state of a progressive render: the window starts at global iteration B = --first-iteration (default 1024),
where r^2 changes by < 1.5 % over the 8 x (W + 5 K) iterations of the largest job.  The first iterations of a
render (large radius, gather-heavy) are measured separately at N = 1 and reported under "early_iterations".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sponza|conference|cornell]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "ppm_1080p_1Mphotons_rays_per_second"   # BASELINE.json: "Mrays/s + PPM iterations/s at 1080p (1M photons/iter)"; it/s is in `iterations_per_s`
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sponza", choices=["sponza", "conference", "cornell", "cornellsmall"])
    ap.add_argument("--method", default="ppm", choices=["ppm", "bdpt"], help="PROGRESSIVE_PHOTON_MAPPING (headline) or VCM_BIDIRECTIONAL_PATH_TRACING")
    ap.add_argument("--max-path-length", type=int, default=8, help="BDPT only (BASELINE config 2 names 8; the reference default is 10)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--photons", type=int, default=1024, help="photon launch is photons x photons")
    ap.add_argument("--detail", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--relmse-target", type=float, default=0.0,
                    help="extra leg (BASELINE config 5): device time until the relMSE of the running mean against a longer reference render drops below this")
    ap.add_argument("--relmse-reference-iterations", type=int, default=256)
    ap.add_argument("--first-iteration", type=int, default=1024, help="global iteration the timed window starts at (see the module docstring)")
    ap.add_argument("--repeats", type=int, default=5, help="the K-step window is timed this many times back to back; ms_per_step is the median window")
    ap.add_argument("--cpu-sample", default="full", help="--impl reference: WxH:photonLaunch of one step, or 'full' = the configuration itself")
    ap.add_argument("--cpu-baseline-sample", default="full", help="cpu_baseline leg of the default arm (3 timed iterations + 1 warm-up)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.samples, self.maxmhz = [], 0      # (monotonic time, sm MHz, reasons)
        self.proc = None
        self.index = index
        self.windows = []

    def mark(self):
        return time.monotonic()

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                rs = [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")]
                self.samples.append((time.monotonic(), float(f[0]), rs))
                self.maxmhz = max(self.maxmhz, float(f[1]))
            except Exception:
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        inside = [x for x in self.samples if any(a <= x[0] <= b for a, b in self.windows)]
        s = sorted(x[1] for x in inside)
        reasons = sorted({r for x in inside for r in x[2]})
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.maxmhz or None, "reasons": reasons, "samples": len(s),
                "window": "device-timed region + end-to-end region (both under load); the sampler runs from before the warm-up"}


def build_scene(args):
    import oppositerenderer_b200 as opr
    if args.workload == "cornell":
        scene = opr.Cornell()
        cam = scene.getDefaultCamera()
    elif args.workload == "cornellsmall":
        scene = opr.CornellSmall()
        cam = scene.getDefaultCamera()
    else:
        from oppositerenderer_b200 import synthetic
        scene = synthetic.make_scene(args.workload, seed=1, detail=args.detail)
        cam = scene.getDefaultCamera()
    cam.setAspectRatio(args.width / args.height)
    return scene, cam


def cpu_reference_run(args, scene, cam, sample, steps, warmup, which):
    """The reference's CPU implementation of the path (oracle/_ref when built, else the oracle port) on a bounded
    sample of the same workload, all host threads.  Returns (Mrays/s, it/s, cores, description)."""
    import oracle_lib as ol
    from oppositerenderer_b200 import capi
    if sample == "full":
        w, h, pl = args.width, args.height, args.photons
    else:
        wh, pl = sample.split(":")
        w, h = (int(x) for x in wh.split("x"))
        pl = int(pl)
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    o = ol.OracleRenderer(which, (pl, pl), args.max_path_length if args.method == "bdpt" else 10, 1645301512, threads=len(os.sched_getaffinity(0)))
    o.init_scene(scene.getSceneDescription())
    det = ol.make_details(cam.pod, capi.VCM_BIDIRECTIONAL_PATH_TRACING if args.method == "bdpt" else capi.PROGRESSIVE_PHOTON_MAPPING, w, h)
    r0 = scene.getSceneInitialPPMRadiusEstimate()
    B = args.first_iteration
    lib = capi.load_library()
    radius = lambda i: float(lib.opr_ppm_radius_for_iteration(float(r0), int(B + i), 2.0 / 3.0))   # same window as the GPU arm
    for i in range(warmup):
        o.render_next_iteration(B + i, i, radius(i), det)
    c0 = o.ray_counts()
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        o.render_next_iteration(B + i, i, radius(i), det)
    dt = time.perf_counter() - t0
    c1 = o.ray_counts()
    rays = (c1[0] - c0[0]) + (c1[1] - c0[1])
    cores = len(os.sched_getaffinity(0))
    desc = "%d %s iteration(s) (global iterations %d..%d) of the same scene at %dx%d%s (%s; OpenMP, %d threads)" % (
        steps, args.method.upper(), B + warmup, B + warmup + steps - 1, w, h,
        " with a %dx%d photon launch" % (pl, pl) if args.method == "ppm" else ", max path length %d" % args.max_path_length,
        "the full configuration" if (w, h, pl) == (args.width, args.height, args.photons) else "a REDUCED sample of the configuration", cores)
    return rays / dt / 1e6, steps / dt, cores, desc


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    how = ("PPM %dx%d, %dx%d photon launch" % (args.width, args.height, args.photons, args.photons) if args.method == "ppm"
           else "BDPT (vertex connection, recursive MIS, max path length %d) %dx%d" % (args.max_path_length, args.width, args.height))
    workload_name = {"sponza": "synthetic ~1M-tri Sponza-scale textured scene, ", "conference": "synthetic ~300K-tri Conference-scale scene with glass/mirror caustics, ",
                     "cornell": "built-in Cornell box, ", "cornellsmall": "built-in CornellSmall (LightArea|Blocks), "}[args.workload] + how
    ws_mb = (args.photons * args.photons * 4 * 36 if args.method == "ppm" else args.width * args.height * 96 * 2) / 1e6 + args.width * args.height * 64 / 1e6
    config = {"workload": workload_name, "method": args.method, "width": args.width, "height": args.height,
              "photons_per_iteration": args.photons * args.photons if args.method == "ppm" else None, "detail": args.detail,
              "l2": ("per-iteration working set (path state + photon/light-vertex records, %.0f MB, plus ray queues) exceeds the 126 MB L2" % ws_mb) if ws_mb > 126
                    else "working set %.0f MB fits the L2; a 256 MB scratch buffer is overwritten between timed iterations" % ws_mb,
              "first_iteration": args.first_iteration,
              "window": "steady state of a progressive render: global iterations from %d on (radius schedule flat: the same work per step for every N)" % args.first_iteration,
              "parallelism": "iterations partitioned over %d GPU(s), scene replicated, one NCCL reduce of the sum buffer (opr_reduce_output, C++)" % world}
    flush_l2 = ws_mb <= 126

    if args.impl == "reference":
        if rank != 0:
            return
        import oracle_lib as ol
        scene, cam = build_scene(args)
        which = "ref" if ol.have_ref() else "oracle"
        steps = max(1, args.steps)      # K steps and W warm-up steps as asked; each step is one iteration of the bounded sample
        v, its, cores, desc = cpu_reference_run(args, scene, cam, args.cpu_sample, steps, max(0, args.warmup), which)
        if args.cpu_sample != "full":     # the line must describe what ran
            wh, pls = args.cpu_sample.split(":")
            config["width"], config["height"] = (int(x) for x in wh.split("x"))
            config["photons_per_iteration"] = int(pls) ** 2 if args.method == "ppm" else None
            config["workload"] += " — REDUCED to %s for this CPU run" % args.cpu_sample
        config["parallelism"] = "host CPU, OpenMP over launch indices, %d threads; loads liboppositerenderer_b200.so only for scene generation and host helpers" % cores
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(0, args.warmup),
                "ms_per_step": 1000.0 / its, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "iterations_per_s": its,
                "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "reference" if which == "ref" else "port", "sample": desc},
                "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import oppositerenderer_b200 as opr
    from oppositerenderer_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t0 = time.perf_counter()
    scene, cam = build_scene(args)
    t_scene = time.perf_counter() - t0
    W, H = args.width, args.height
    # distinct RNG streams per rank (the reference's servers got this from wall-clock seeds, SpatialHash.cu:322)
    from oppositerenderer_b200 import distributed as D
    r = opr.OptixRenderer(photon_launch=(args.photons, args.photons), vcm_max_path_length=args.max_path_length if args.method == "bdpt" else 0,
                          seed=D.rank_seed(1645301512, rank))
    r.initialize(opr.ComputeDevice(local_rank))
    t0 = time.perf_counter()
    r.initScene(scene)
    r.synchronize()
    t_init = time.perf_counter() - t0
    method = opr.RenderMethod.VCM_BIDIRECTIONAL_PATH_TRACING if args.method == "bdpt" else opr.RenderMethod.PROGRESSIVE_PHOTON_MAPPING
    details = opr.RenderServerRenderRequestDetails(cam, scene.getSceneName(), method, W, H, 2.0 / 3.0)
    pod = details.pod()
    r0 = scene.getSceneInitialPPMRadiusEstimate()
    lib = capi.load_library()
    D.check_state_count(max(args.photons, W) * max(args.photons, H))     # per-rank seed ranges stay disjoint over the whole state buffer
    B = args.first_iteration
    radius = lambda it: D.radius_for_iteration(r0, it)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(lib.opr_get_stream(r._h), device=dev)
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_l2 else None
    if world > 1:
        # the collective of the data path is the product's own (C++ -> NCCL, opr_reduce_output); torch.distributed only carries the 128-byte id
        ids = [opr.renderer.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        r.comm_init_rank(ids[0], rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cursor = [0]     # next local iteration of this rank; global iteration = B + local * world + rank (SURVEY.md §8e partition, offset by B)

    def run(steps, readback=None, first_iteration=None):
        base = B if first_iteration is None else first_iteration
        for _ in range(steps):
            local = cursor[0]
            cursor[0] += 1
            it = base + local * world + rank
            if scratch is not None:
                with torch.cuda.stream(stream):
                    scratch.fill_(local & 0xff)        # evict the previous iteration's buffers from the 126 MB L2
            r.renderNextIteration(it, local, radius(it), True, pod)
            if readback == "blocking":
                r.getOutputBuffer(host_out[0])
            elif readback == "async":
                r.getOutputBufferAsync(host_out[local & 1])   # returns at once; lands while the next iteration renders
        if readback == "async":
            r.waitOutput()
        r.join()        # the gather / output of the last iteration run on the renderer's second stream: fold them into `stream`
        if world > 1:
            r.reduce_output(0)      # ncclReduce(sum) on the renderer's stream, into a scratch buffer on rank 0

    def timed(steps, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record()
        run(steps, **kw)
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        return e0.elapsed_time(e1)

    host_out = [torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]

    # ---- warm-up, then the device-timed region (inputs resident in HBM): the K-step window, `repeats` times back to back
    sampler = ClockSampler(local_rank)
    sampler.start()
    run(args.warmup)
    barrier()
    st0 = r.stats()          # counters are cumulative: everything reported below is the difference over the timed region
    launches0 = st0["launches"]
    w0 = sampler.mark()
    windows_ms = [timed(args.steps) for _ in range(max(1, args.repeats))]
    sampler.window(w0, sampler.mark())
    st = r.stats()
    nwin = len(windows_ms)
    ms = sorted(windows_ms)[nwin // 2]
    rays = ((st["closest_rays"] + st["any_rays"]) - (st0["closest_rays"] + st0["any_rays"])) / nwin     # per window; the windows differ by < 0.1 %
    launches_per_window = (st["launches"] - launches0) / nwin

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline of the dominant kernel
    roof = None
    kernel_ms = {}
    ray_queue = None
    nprof = 3
    if True:   # every rank renders the same iterations (keeps the partition aligned); rank 0 reports
        r.set_timing(True)
        acc = {}
        for k in range(nprof):
            local = cursor[0]
            cursor[0] += 1
            it = B + local * world + rank
            r.renderNextIteration(it, local, radius(it), True, pod)
            for name, v in r.timings():
                acc[name] = acc.get(name, 0.0) + v / nprof
        r.set_timing(False)
        kernel_ms = acc
        stp = r.stats()
        px = W * H
        peak, how = peaks()
        alg = {}
        if args.method == "ppm":
            g = r.debug_buffer("grid", np.uint32, 8)
            V = int(g[7])
            ncells = int(g[1]) * int(g[2]) * int(g[3])
            N = args.photons * args.photons * 4
            # SURVEY.md §8d per-unit figures x the units one launch processes
            alg = {
                "ppm.gather": 76.0 * px + 36.0 * V + 8.0 * ncells,
                "trace.shadow": 36.0 * (stp["any_rays"] - st["any_rays"]) / nprof,
                "trace.camera0": 52.0 * px,
                "trace.photon0": 52.0 * args.photons * args.photons,
                "grid.build": 36.0 * N + 132.0 * V + 8.0 * ncells,
            }
            acc["grid.build"] = sum(acc.get(k, 0.0) for k in ("grid.bbox", "grid.keys", "grid.sort", "grid.offsets", "grid.reorder"))
            qc = r.debug_buffer("queue_counts", np.uint32, 34)
            ray_queue = {"capacity": int(qc[0]), "camera_rays_per_bounce": [int(x) for x in qc[2:18] if x], "photon_rays_per_bounce": [int(x) for x in qc[18:34] if x],
                         "shadow_rays": int(qc[1]), "shadow_capacity": 4 * px,
                         "note": "wavefront bounces, then the last entry is what the per-path tail kernel starts from; occupancy = rays / capacity"}
        else:
            rays_it = ((stp["closest_rays"] + stp["any_rays"]) - (st["closest_rays"] + st["any_rays"])) / nprof
            alg = {"vcm.light": 48.0 * rays_it / 2, "vcm.camera": 48.0 * rays_it / 2}
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")   # dram bytes per launch from the committed `ncu --set full` capture
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        cands = [(acc.get(k, 0.0), k) for k in alg if k != "grid.build" and acc.get(k, 0.0) > 0]
        if cands:
            dur, name = max(cands)
            achieved = alg[name] / (dur * 1e-3) / 1e9
            tr = traffic.get(name, {}) if args.workload == "sponza" and args.method == "ppm" and (W, H) == (1920, 1080) else {}
            roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": tr.get("dram_bytes"), "traffic_source": tr.get("source"),
                    "actual_limiter": {k: tr.get(k) for k in ("issue_active_pct", "l1_data_pipe_wavefronts_pct", "l2_throughput_pct", "dram_throughput_pct",
                                                              "lanes_per_instruction")} if tr else None,
                    "issue_frac": (tr["issue_active_pct"] / 100.0 if tr and tr.get("issue_active_pct") else None),   # share of the SM issue slots in use: the roofline that does bound this kernel
                    "peak_source": how, "kernel_ms": dur, "algorithmic_bytes": alg[name],
                    "note": "gather and traversal are issue- / L1- / latency-bound, not HBM-bound (DESIGN.md §4, SURVEY.md §8d); profiles/ holds issue, L1 and L2 utilisation of the same launches",
                    "others": {k: {"ms": acc[k], "achieved_gbs": alg[k] / (acc[k] * 1e-3) / 1e9, "frac": alg[k] / (acc[k] * 1e-3) / 1e9 / peak}
                               for k in alg if k != name and acc.get(k, 0.0) > 0}}

    # ---- the overlapped schedule as the device ran it: events on every pass's own stream, no synchronisation (opr_set_timing 2);
    # the third of four pipelined iterations is reported, times relative to the start of its camera pass
    timeline = None
    if args.method == "ppm":
        barrier()
        r.set_timing(2)
        run(4)
        tl = r.timeline()
        r.set_timing(0)
        starts = [i for i, e in enumerate(tl) if e[0] == "camera.gen"]
        if len(starts) >= 4:
            t0_, t1_ = tl[starts[2]][2], tl[starts[3]][2]
            ev = tl[starts[2]:starts[3]]
            timeline = {"iteration_ms": t1_ - t0_,
                        "passes": [[n, st_, round(a_ - t0_, 3), round(b_ - t0_, 3)] for n, st_, a_, b_ in ev if n not in ("ppm.camera", "ppm.photon", "ppm.direct")],
                        "note": "[pass, stream, start, end] in ms from the start of this iteration's camera pass; stream A = RNG-ordered path passes, "
                                "B = photon grid + gather + output of the same iteration (overlaps the next iteration's A), C = tails"}

    # ---- end to end through the C ABI with host buffers: request struct in (120 B), W*H*3 floats out to pinned host memory EVERY step.
    # (i) the pipelined driver loop: opr_get_output_buffer_async after every iteration (the copy of image i overlaps iteration i+1),
    # opr_wait_output at the end; (ii) the reference's blocking loop (StandaloneRenderManager.cpp:114-128), reported beside it.
    def wall(mode):
        barrier()
        s0 = r.stats()
        t0 = time.perf_counter()
        run(args.steps, readback=mode)
        barrier()
        dt = time.perf_counter() - t0
        s1 = r.stats()
        return dt, (s1["closest_rays"] + s1["any_rays"]) - (s0["closest_rays"] + s0["any_rays"])
    run(2, readback="async")      # first use allocates the snapshots
    w1 = sampler.mark()
    e2e_s, rays_e2e = wall("async")
    e2e_blocking_s, rays_e2e_blocking = wall("blocking")
    sampler.window(w1, sampler.mark())
    clocks = sampler.stop()

    # ---- the first iterations of a render (large radius: the gather dominates), N = 1 only: round 1 quoted these
    early = None
    if world == 1 and args.method == "ppm":
        cursor[0] = 0
        run(3, first_iteration=0)
        st_e0 = r.stats()
        ms_e = timed(args.steps, first_iteration=0)
        st_e1 = r.stats()
        rays_e = (st_e1["closest_rays"] + st_e1["any_rays"]) - (st_e0["closest_rays"] + st_e0["any_rays"])
        early = {"iterations": "3..%d" % (2 + args.steps), "ms_per_step": ms_e / args.steps, "value": rays_e / (ms_e * 1e-3) / 1e6, "unit": "Mrays/s",
                 "note": "same workload at the START of a progressive render (radius 1.0-0.6 r0, gather 3-6x the steady-state cost)"}

    t = torch.tensor([ms, float(rays), e2e_s, float(rays_e2e), e2e_blocking_s, float(rays_e2e_blocking)] + windows_ms, dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        windows_ms = [float(x) for x in tmax[6:]]            # every window: max over ranks
        ms = sorted(windows_ms)[nwin // 2]
        e2e_s, e2e_blocking_s = float(tmax[2]), float(tmax[4])
        rays, rays_e2e, rays_e2e_blocking = float(tsum[1]), float(tsum[3]), float(tsum[5])
    total_steps = args.steps * world

    # ---- optional: time to a target relMSE (BASELINE config 5), iteration-partitioned like everything else
    ttt = None
    if args.relmse_target > 0:
        nref = args.relmse_reference_iterations
        ref_r = opr.OptixRenderer(photon_launch=(args.photons, args.photons), vcm_max_path_length=args.max_path_length if args.method == "bdpt" else 0,
                                  seed=D.rank_seed(987654321, rank))
        ref_r.initialize(opr.ComputeDevice(local_rank))
        ref_r.initScene(scene)
        for k in range(nref):                                   # every rank renders the same reference (independent seed set)
            ref_r.renderNextIteration(k, k, radius(k), True, pod)
        ref = D.output_as_tensor(ref_r, dev).clone() / nref
        ref2 = (ref ** 2).sum(-1)
        eps = 1e-3 * float(ref2.mean())
        ref_r.close()
        barrier()
        check_every, max_local = 4, 4096
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            t0e.record()
        reached, done_local, err = None, 0, None
        while done_local < max_local and reached is None:
            for local in range(done_local, done_local + check_every):
                it = local * world + rank
                r.renderNextIteration(it, local, radius(it), True, pod)
            done_local += check_every
            r.join()
            with torch.cuda.stream(stream):
                total = D.output_as_tensor(r, dev).clone()
                if world > 1:
                    dist.all_reduce(total)
                mean = total / (done_local * world)
                err = float((((mean - ref) ** 2).sum(-1) / (ref2 + eps)).mean())
            if err <= args.relmse_target:
                reached = done_local * world
        with torch.cuda.stream(stream):
            t1e.record()
        barrier()
        tsec = torch.tensor([t0e.elapsed_time(t1e) * 1e-3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        ttt = {"target_relmse": args.relmse_target, "reached": reached is not None, "iterations": reached, "seconds": float(tsec[0]), "final_relmse": err,
               "reference_iterations": nref, "note": "relMSE of the running mean of all ranks against an independent %d-iteration render; includes a check (clone + all-reduce) every %d local iterations" % (nref, check_every)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            import oracle_lib as ol
            which = "ref" if ol.have_ref() else "oracle"
            v, its, cores, desc = cpu_reference_run(args, scene, cam, args.cpu_baseline_sample, 3, 1, which)
            cpu = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "reference" if which == "ref" else "port", "sample": desc, "iterations_per_s": its}
        except Exception as e:  # the oracle is a reported baseline, never a dependency of the product numbers
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}

    if rank == 0:
        line = {"metric": METRIC, "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "iterations_per_s": total_steps / (ms * 1e-3), "rays_per_iteration": rays / total_steps,
                "photons_stored_per_iteration": (st["photons_stored"] - st0["photons_stored"]) / max(1, args.steps * nwin),
                "windows_ms": windows_ms, "repeats": nwin,
                "e2e": {"value": rays_e2e / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(capi.opr_request_details),
                        "d2h_bytes_per_step": W * H * 12, "iterations_per_s": total_steps / e2e_s,
                        "how": "opr_render_next_iteration + opr_get_output_buffer_async into pinned host memory EVERY step (image i is copied out while "
                               "iteration i+1 renders), opr_wait_output after the last step; wall clock, max over ranks",
                        "blocking": {"value": rays_e2e_blocking / e2e_blocking_s / 1e6, "iterations_per_s": total_steps / e2e_blocking_s,
                                     "how": "the reference's driver loop: blocking opr_get_output_buffer after every step"}},
                "early_iterations": early, "timeline": timeline,
                "gpu_launches": int(round(launches_per_window)), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "ray_queue": ray_queue, "time_to_relmse": ttt,
                "kernel_ms": kernel_ms, "scene": {"triangles": scene.getNumTriangles(), "meshes": scene.getNumMeshes(), "generate_s": t_scene, "init_scene_s": t_init}}
        print(json.dumps(line))
    if world > 1:
        r.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
