#!/usr/bin/env python
"""bench.py -- chain mutations/s of the drmlt hot path on N B200s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --steps K --warmup W    (CPU arm: the reference's own DRMLT integrator, oracle/_ref, on all host
                                                            cores; the oracle port only if oracle/_ref is absent)

Workload (config.workload): C5 -- procedural occluded-light "door" scene, ~1M triangles, 1280x720,
drmlt type=orbital technique=mmlt, maxDepth 8, directSamples=-1, synthetic data generated here.
A step = every resident chain of every rank advances `--mutations` iterations of the MLT loop
(dr_job_run); per-GPU work is fixed as N grows (weak scaling; chains are independent, so there is
no data-path collective -- b all-reduce and film reduce happen once per job and are part of e2e).
`value` = mutations of all ranks / max-over-ranks device time (CUDA events on the launching stream,
scene + chain state resident in HBM).  `e2e` = the same metric through the public whole-job call
(scene re-upload H2D + bootstrap + b all-reduce + chains + film reduce + develop + image D2H).
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

import numpy as np  # noqa: E402

METRIC = "chain_mutations_per_sec"
UNIT = "mutations/s"
# BASELINE.json configs[0..4] (C1..C5): scene generator, film, integrator parameters.  C5 is the headline workload.
CONFIGS = {
    "C1": ("cornell", dict(film=(256, 256)), dict(integrator="pssmlt", technique="path", sigma=1.0 / 64, maxDepth=8, directSamples=-1)),
    "C2": ("cornell", dict(film=(256, 256)), dict(integrator="drmlt", type="mira", technique="path", scaleSecond=0.1, acceptanceMap=True,
                                                  rfilter="box", maxDepth=8, directSamples=-1)),
    "C3": ("glossy", dict(film=(512, 512)), dict(integrator="drmlt", type="green", technique="bdpt", directSampling=False, maxDepth=8, directSamples=-1)),
    "C4": ("caustic", dict(film=(512, 512)), dict(integrator="drmlt", type="orbital", technique="mmlt", fixEmitterPath=True, maxDepth=8, directSamples=-1)),
    "C5": ("door", dict(), dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=8, directSamples=-1)),
}
PARAMS = CONFIGS["C5"][2]


def config_params(args):
    return dict(CONFIGS[args.config][2])


def workload_name(config, data):
    scene_name, _, params = CONFIGS[config]
    extra = " ".join("%s=%s" % (k, v) for k, v in params.items() if k not in ("integrator", "type", "technique", "maxDepth", "directSamples"))
    return "%s %s scene %d tris %dx%d %s %s %s maxDepth=%d directSamples=%d%s" % (
        config, scene_name, data.n_triangles, data.film[0], data.film[1], params["integrator"], params.get("type", ""), params["technique"],
        params["maxDepth"], params["directSamples"], (" " + extra) if extra else "")


def build_scene(args):
    from drmlt_mitsuba_b200 import scenes
    scene_name, kw, _ = CONFIGS[args.config]
    return scenes.SCENES[scene_name](**kw)


def ray_bytes(n_tris):
    """B_ray(T) = 64*ceil(log2(T/4)) + 4*48 + 64."""
    return 64 * int(np.ceil(np.log2(max(n_tris, 8) / 4.0))) + 4 * 48 + 64


def mutation_bytes(n_tris, paths_per_mut, rays_per_path, dims, footprint):
    """B_mut = p*d*B_ray + B_state + B_splat; dims = D_s + D_e + D_d of the chain."""
    b_state = 2 * dims * 4 * (1 + paths_per_mut)
    b_splat = 3 * 16 * footprint * 3
    return paths_per_mut * rays_per_path * ray_bytes(n_tris) + b_state + b_splat


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------- CPU arm
def oracle_scene(data):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.build()
    return oracle_lib, oracle_lib.OracleScene(data)


def cpu_config(seed=99):
    """The oracle in the reference's default precision (double, Epsilon 1e-7) on the same parameters."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from drmlt_mitsuba_b200 import abi
    return oracle_lib.default_config(integrator=abi.DR_INTEGRATOR_DRMLT, technique=abi.DR_TECH_MMLT, type=abi.DR_TYPE_ORBITAL,
                                     max_depth=8, direct_samples=-1, direct_sampling=0, kelemen_style_weights=0, seed=seed)


def cpu_sample(orc, cfg, threads, target_s, n_boot=40000):
    """A bounded sample of the workload on the host cores: bootstrap, then `threads*4` chains."""
    lum, dep = orc.bootstrap(cfg, 0, n_boot)
    seeds = np.nonzero(lum > 0)[0]
    b = float(lum.mean() * 8)
    n_chains = threads * 4
    rng = np.random.RandomState(1)
    p = lum[seeds] / lum[seeds].sum()
    pick = seeds[rng.choice(len(seeds), n_chains, p=p)]
    ids = np.arange(n_chains, dtype=np.uint64)

    def run(steps):
        t0 = time.perf_counter()
        _, _, st = orc.chain_steps(cfg, b, pick.astype(np.uint64), dep[pick], ids, steps, want_film=True, threads=threads, want_records=False)
        return st.mutations / (time.perf_counter() - t0), st
    rate, _ = run(200)                                            # calibration
    steps = int(max(200, min(2_000_000, rate * target_s / n_chains)))
    rate, st = run(steps)
    return rate, n_chains, steps, st


REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_path.so")


def reference_sample(config, spp, timeout_s=900):
    """One bounded run of the REFERENCE'S OWN DRMLT integrator (oracle/_ref, compiled from the reference's sources) on all
    host cores, in a subprocess (tools/ref_sample.py).  None if oracle/_ref is absent or the run fails."""
    if not os.path.exists(REF_LIB):
        return None
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_sample.py"), "--config", config, "--spp", str(spp)],
                             capture_output=True, text=True, timeout=timeout_s).stdout
        for ln in out.splitlines():                     # the reference's progress bar shares the line
            if "REF_SAMPLE " in ln:
                return json.loads(ln[ln.index("REF_SAMPLE ") + len("REF_SAMPLE "):])
    except Exception:
        pass
    return None


def reference_baseline(r):
    return {"value": r["mutations_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference", "render_s": r["render_s"],
            "scene_build_s": r["scene_build_s"], "cold_value": r["mutations"] / (r["render_s"] + r["scene_build_s"]),
            "sample": "the reference's own integrator (DRMLT::render / PSSMLT::render of oracle/_ref: the reference's sources, -O3 -march=nocona, double precision, SAH kd-tree) on the same scene and parameters: %d mutations/pixel = "
                      "%d mutations in %.1f s render time incl. its bootstrap, %d worker threads; kd-tree build %.1f s not counted"
                      % (r["spp"], r["mutations"], r["render_s"], r["threads"], r["scene_build_s"]),
            "acceptance_percent": r["stats_percent"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if os.path.exists(REF_LIB):
        # the real reference: warm-up steps at a small sample, timed steps sized to <= ~2 minutes in total
        runs = []
        for i in range(args.warmup + args.steps):
            timed = i >= args.warmup
            # the e2e job's own 64 mutations/pixel when the K timed steps then still end within minutes (K <= 5), a bounded sample otherwise
            ref_spp = args.ref_spp if args.ref_spp > 0 else max(8, min(64, 320 // max(1, args.steps)))
            spp = ref_spp if timed else 2
            r = reference_sample(args.config, spp)
            if r is None:
                runs = None
                break
            if timed:
                runs.append(r)
        if runs:
            data = build_scene(args)
            value = float(sum(r["mutations"] for r in runs) / sum(r["render_s"] for r in runs))
            base = reference_baseline(runs[-1])
            base["value"] = value
            line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": 1e3 * float(np.mean([r["render_s"] for r in runs])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(args.config, data), "mutations_per_pixel_per_step": runs[-1]["spp"]},
                    "cpu_baseline": base,
                    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
            print(json.dumps(line))
            return 0
    data = build_scene(args)
    _, orc = oracle_scene(data)
    cfg = cpu_config()
    threads = os.cpu_count() or 1
    per_step_s = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, info = [], None
    for i in range(args.warmup + args.steps):
        rate, n_chains, steps, st = cpu_sample(orc, cfg, threads, per_step_s)
        if i >= args.warmup:
            rates.append(rate)
        info = (n_chains, steps)
    value = float(np.mean(rates))
    sample = "%d chains x %d mutations per step, oracle port (double precision) of DRMLTRenderer::process, %d threads" % (info[0], info[1], threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * info[0] * info[1] / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(args.config, data)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def scaled_traffic(traffic, stage, units_per_launch):
    """DRAM bytes (ncu dram__bytes_read.sum + dram__bytes_write.sum) of one launch of `stage`, scaled from the captured
    launch of profiles/traffic.json to the units (rays or paths) one launch of THIS run processes."""
    if stage not in traffic:
        return None
    cap = traffic.get("paths_in_captured_launch" if stage == "k_chain" else "rays_in_captured_launch")
    return int(traffic[stage] * units_per_launch / cap) if cap else traffic[stage]


# ---------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    # libraries (NCCL's version banner, ...) may write to stdout; the contract is ONE JSON line there
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import __graft_entry__
    from drmlt_mitsuba_b200 import abi, distributed
    from drmlt_mitsuba_b200.integrator import Job, Scene, make_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    __graft_entry__.build()

    data = build_scene(args)
    t0 = time.perf_counter()
    scene = Scene(data, device=local)
    scene_create_s = time.perf_counter() - t0
    scene_bytes = scene.reupload()

    params = dict(config_params(args), seed=args.seed, sampleCount=args.spp * world)   # weak scaling: W*H*spp mutations per GPU
    if args.chains:
        params["chains"] = args.chains
    if args.lanes:
        params["lanes"] = args.lanes
    cfg = make_config(rank=rank, worldSize=world, **params)
    job = Job(scene, cfg)
    s, c = job.bootstrap()
    b = distributed.all_reduce_normalization(s, c, True, cfg.max_depth, dist if world > 1 else None, dev)
    job.seed_chains(b)
    n_chains = job.num_chains
    M = args.mutations
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        job.run(M)
    barrier()
    st0 = job.stats()
    wall0 = time.perf_counter()
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            job.run(M)                                              # device-timed inside (events on the launching stream)
        barrier()
    wall = time.perf_counter() - wall0
    st1 = job.stats()
    dev_ms = st1.chains_ms - st0.chains_ms
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    muts_rank = st1.mutations - st0.mutations
    muts_all = muts_rank * world
    value = muts_all / (max_ms * 1e-3)
    launches = int(st1.kernel_launches - st0.kernel_launches)

    # ---- per-stage device time: one extra step with CUDA events around every stage of every round (the timed
    #      steps above run without them), on the same resident job
    job.profile(True)
    flush.fill_(1)
    torch.cuda.synchronize()
    p0 = job.stats()
    job.run(M)
    p1 = job.stats()
    job.profile(False)
    stage_ms = {"k_trace": p1.trace_ms - p0.trace_ms, "k_walk+k_connect": p1.walk_ms - p0.walk_ms, "k_chain": p1.chain_ms - p0.chain_ms}
    stage_launches = {"k_trace": p1.trace_launches - p0.trace_launches, "k_walk+k_connect": p1.walk_launches - p0.walk_launches,
                      "k_chain": p1.chain_launches - p0.chain_launches}
    prof_rounds = max(1, p1.rounds - p0.rounds)
    prof_rays, prof_paths, prof_muts = p1.rays - p0.rays, p1.paths - p0.paths, p1.mutations - p0.mutations
    tot_stage = sum(stage_ms.values()) or 1.0

    # ---- roofline.  Algorithmic bytes per unit of each stage (DESIGN.md "Algorithmic bytes"):
    #   k_trace : B_ray(T) per ray = 64*ceil(log2(T/4)) + 4*48 + 64
    #   k_walk  : lane records of one vertex step, read + written = 1 056 B per ray that hits
    #   k_chain : lane records + coordinate buffers of one finished path, read + written = 1 664 B per path
    paths_per_mut = (st1.paths - st0.paths) / max(1, muts_rank)
    rays_per_path = (st1.rays - st0.rays) / max(1, st1.paths - st0.paths)
    mean_depth = 4.5                                                # depths 1..8 equally likely
    dims = 2 * (3 * (mean_depth + 2)) + 1
    b_mut = mutation_bytes(data.n_triangles, paths_per_mut, rays_per_path, dims, 16)
    peak, peak_src = 6650.0, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); peak_src = "measured"
    except Exception:
        pass
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    unit_bytes = {"k_trace": (ray_bytes(data.n_triangles), prof_rays), "k_walk+k_connect": (1056, prof_rays), "k_chain": (1664, prof_paths)}
    stages = {}
    for k, ms in stage_ms.items():
        bpu, units = unit_bytes[k]
        stages[k] = {"ms": ms, "share": ms / tot_stage, "launches": int(stage_launches[k]),
                     "achieved_gbs": bpu * units / (ms * 1e-3) / 1e9 if ms > 0 else None, "bytes_per_unit": bpu, "units": int(units)}
    dom = max(stage_ms, key=stage_ms.get)
    dom_launch_ms = stage_ms[dom] / max(1, stage_launches[dom])
    achieved = stages[dom]["achieved_gbs"] or 0.0
    step_gbs = b_mut * muts_rank / (dev_ms * 1e-3) / 1e9
    # What ncu measured for the kernels (profiles/traffic.json, one --set full capture per kernel): the traversal kernel serves
    # its BVH from L1 / L2, so the ALGORITHMIC bytes of SURVEY 8d (the `achieved` / `frac` of the contract) are not its DRAM bytes --
    # its real DRAM utilisation is a few percent and its bound is L1 wavefronts + the latency of dependent node fetches under
    # divergence.  Reported side by side.
    ncu = traffic.get("ncu", {})
    roofline = {"bound": ncu.get(dom, {}).get("bound", "hbm"), "bound_of_byte_model": "hbm", "kernel": dom, "achieved": achieved, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": scaled_traffic(traffic, dom, stages[dom]["units"] / max(1, stage_launches[dom])),
                "traffic_source": traffic.get("source"), "ms_per_launch": dom_launch_ms,
                "units_per_launch": stages[dom]["units"] / max(1, stage_launches[dom]), "bytes_per_unit": stages[dom]["bytes_per_unit"],
                "ncu": ncu.get(dom),
                "stages": stages, "whole_step": {"bytes_per_mutation": b_mut, "achieved": step_gbs, "frac": step_gbs / peak},
                "paths_per_mutation": paths_per_mut, "rays_per_path": rays_per_path,
                "mrays_per_s": (st1.rays - st0.rays) * world / (max_ms * 1e-3) / 1e6,
                "rounds_per_step": prof_rounds, "profiled_step_mutations": int(prof_muts)}
    # the stage group that moves the most DRAM bytes (ncu): its measured traffic per launch over its live launch time
    dram_stage = max((k for k in stage_ms if k in traffic), key=lambda k: traffic[k] / max(1, traffic.get("paths_in_captured_launch" if k == "k_chain" else "rays_in_captured_launch", 1)) * stages[k]["units"], default=None)
    if dram_stage:
        t_launch = scaled_traffic(traffic, dram_stage, stages[dram_stage]["units"] / max(1, stage_launches[dram_stage]))
        ms_launch = stage_ms[dram_stage] / max(1, stage_launches[dram_stage])
        roofline["largest_dram_stage"] = {"kernel": dram_stage, "bound": "hbm", "dram_bytes_per_launch": t_launch, "ms_per_launch": ms_launch,
                                          "achieved": t_launch / (ms_launch * 1e-3) / 1e9 if ms_launch > 0 else None,
                                          "frac": t_launch / (ms_launch * 1e-3) / 1e9 / peak if ms_launch > 0 else None, "ncu": ncu.get(dram_stage)}
    job.close()

    # ---- e2e: whole job through the public API with host buffers
    e2e_params = dict(config_params(args), seed=args.seed + 1, sampleCount=args.e2e_spp * world)
    if args.e2e_chains:
        e2e_params["chains"] = args.e2e_chains
    if args.e2e_lanes:
        e2e_params["lanes"] = args.e2e_lanes
    barrier()
    e0 = time.perf_counter()
    h2d = scene.reupload()
    reupload_ms = (time.perf_counter() - e0) * 1e3
    img, est, eb = distributed.render(scene, e2e_params, dist if world > 1 else None, rank, world)
    barrier()
    e_wall = time.perf_counter() - e0
    te = torch.tensor([float(est.mutations)], dtype=torch.float64, device=dev)
    tw = torch.tensor([e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    W, H = data.film
    e2e = {"value": float(te.item()) / float(tw.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(W * H * 3 * 4), "seconds": float(tw.item()), "mutations": int(te.item()),
           "rank0_phases_ms": dict(getattr(distributed.render, "last_timing", {}), scene_reupload_ms=reupload_ms,
                                   device_bootstrap_and_seeding_ms=est.bootstrap_ms, device_chains_ms=est.chains_ms),
           "what": "dr_scene_reupload + bootstrap + b all-reduce + chains + film reduce + develop + image D2H; sampleCount=%d" % (args.e2e_spp * world)}

    # ---- strong scaling: ONE job of W*H*strong_spp mutations in total, shared by the N GPUs (bootstrap, b all-reduce, chains,
    #      film reduce, develop, image D2H; wall clock, max over ranks).  The driver's per-N runs give the curve.
    strong = None
    if args.strong_spp > 0:
        sp = dict(config_params(args), seed=args.seed + 2, sampleCount=args.strong_spp)
        barrier()
        s0 = time.perf_counter()
        _, sst, _ = distributed.render(scene, sp, dist if world > 1 else None, rank, world)
        barrier()
        s_wall = time.perf_counter() - s0
        ts = torch.tensor([float(sst.mutations), s_wall], dtype=torch.float64, device=dev)
        tm = ts.clone()
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "mutations_per_pixel_total": args.strong_spp, "mutations": int(ts[0].item()), "seconds": float(tm[1].item()),
                  "value": float(ts[0].item()) / float(tm[1].item()), "unit": UNIT,
                  "rank0_phases_ms": dict(getattr(distributed.render, "last_timing", {}))}

    # cold start: the same job including dr_scene_create (flattening, host SAH BVH build, 4-wide collapse, upload) -- what
    # `mitsuba scene.xml` pays once per scene; beside it (cpu_baseline.sample) the reference's kd-tree build + render
    e2e["cold"] = {"scene_create_s": scene_create_s, "seconds": scene_create_s + e2e["seconds"],
                   "value": e2e["mutations"] / (scene_create_s + e2e["seconds"]), "unit": UNIT, "bvh": scene.bvh_info()}
    # ... and with the tree built on the device (dr_scene_create_ex, DR_SCENE_BVH_GPU: what the plugin shim uses for its one-shot
    # renders): scene creation + the same job on that scene, one GPU
    if world == 1:
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        cold_scene = Scene(data, device=local, gpu_bvh=True)
        c_create = time.perf_counter() - c0
        _, cst, _ = distributed.render(cold_scene, e2e_params, None, rank, world)
        torch.cuda.synchronize()
        c_wall = time.perf_counter() - c0
        e2e["cold_gpu_bvh"] = {"scene_create_s": c_create, "seconds": c_wall, "value": float(cst.mutations) / c_wall, "unit": UNIT,
                               "bvh": cold_scene.bvh_info(), "device_chains_ms": cst.chains_ms}
        cold_scene.close()

    def pct(a, b):
        return round(100.0 * a / max(1, b), 2)
    # the reference's statistics counters (drmlt_proc.cpp:34-49) of rank 0's job, comparable with cpu_baseline.acceptance_percent
    e2e["acceptance_percent"] = {"Accepted 1st-stage mutations": pct(est.first_accept, est.first_base),
                                 "Accepted 2nd-stage mutations": pct(est.second_accept, est.second_base),
                                 "Accepted bold mutation in the 1st-stage mutations": pct(est.bold_accept, est.bold_base),
                                 "Accepted large mutations in the 1st-stage mutations": pct(est.large_accept, est.large_base),
                                 "Overall acceptance rate": pct(est.accept, est.accept_base)}
    e2e["b"] = est.luminance

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu:
            try:
                ref = reference_sample(args.config, args.cpu_spp)
                if ref is not None:
                    raise StopIteration
                _, orc = oracle_scene(data)
                threads = os.cpu_count() or 1
                rate, nch, steps, _ = cpu_sample(orc, cpu_config(), threads, args.cpu_seconds)
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": "%d chains x %d mutations of the same workload, oracle port (double) of DRMLTRenderer::process" % (nch, steps)}
            except StopIteration:
                cpu = reference_baseline(ref)
            except Exception as ex:                                   # the baseline is reported, never required
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.config, data), "chains_per_gpu": int(n_chains), "mutations_per_chain_per_step": M,
                           "l2": "flushed between timed steps (512 MiB device write)", "b": b, "scene_create_s": scene_create_s,
                           "scene_bytes": int(scene_bytes), "wall_s_timed_region": wall, "e2e_mutations_per_pixel": args.e2e_spp,
                           "cpu_baseline_mutations_per_pixel": args.cpu_spp, "reference_arm_mutations_per_pixel_per_step": args.ref_spp if args.ref_spp > 0 else max(8, min(64, 320 // max(1, args.steps)))},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "strong_scaling": strong, "gpu_launches": launches, "clocks": clk.summary()}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C5", choices=sorted(CONFIGS), help="BASELINE.json configuration (C5 = the headline workload)")
    ap.add_argument("--mutations", type=int, default=64, help="mutations per chain per step")
    ap.add_argument("--chains", type=int, default=-1, help="chains resident per GPU in the timed steps (0 = auto; default: 4194304 for C5, auto otherwise)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes of the wavefront machine in the timed steps (0 = one per chain; fewer: work-unit queue)")
    ap.add_argument("--e2e-chains", type=int, default=0, dest="e2e_chains")
    ap.add_argument("--e2e-lanes", type=int, default=0, dest="e2e_lanes")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--e2e-spp", type=int, default=64, dest="e2e_spp", help="sampleCount of the whole-job e2e render (64 = the C5 workload)")
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, dest="cpu_seconds")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--ref-spp", type=int, default=0, dest="ref_spp",
                    help="--impl reference: mutations per pixel of one timed step; 0 = 64 (the e2e job's own) for K <= 5 steps, else 320 // K "
                         "(a bounded sample: K steps must end within minutes)")
    ap.add_argument("--cpu-spp", type=int, default=64, dest="cpu_spp",
                    help="cpu_baseline leg of the GPU arm: mutations per pixel of the ONE run of the reference's integrator (64 = the e2e job's own)")
    ap.add_argument("--strong-spp", type=int, default=512, dest="strong_spp",
                    help="strong-scaling leg: a job of W*H*strong_spp mutations IN TOTAL, shared by the N GPUs (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.chains < 0:
        args.chains = 4194304 if args.config == "C5" else 0
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
