"""Equal-time relMSE (BASELINE.json: "relMSE at equal time"): the GPU path, the reference's own DRMLT integrator (oracle/_ref) and
the CPU oracle, the latter two on all host cores, render the same scene for the same time budget; all are compared with a long
converged render.

    python tools/equal_time.py [--scene door] [--seconds 10] [--film 320x180] [--out profiles/…json]

relMSE = mean((I - R)^2 / (R^2 + 1e-2))  (SURVEY.md section 8d).  The reference image R is a GPU render with --ref-factor times the
mutations of the timed GPU run and another seed; because the timed GPU run and R share an implementation, the oracle
also renders a long image R_cpu (different code, double precision) and relMSE(R, R_cpu) is reported as the noise floor
of the comparison.  Wall-clock budgets include everything a user pays: bootstrap, seeding, chains, develop."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def relmse(img, ref, eps=1e-2):
    img, ref = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="door")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--film", default="320x180")
    ap.add_argument("--ref-factor", type=int, default=8, dest="ref_factor", help="the reference image gets this many times the mutations of the timed GPU run")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "equal_time.json"))
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    import oracle_lib
    from drmlt_mitsuba_b200 import abi, scenes
    from drmlt_mitsuba_b200.integrator import Scene, make_config

    W, H = [int(x) for x in args.film.split("x")]
    if args.scene == "door":
        data = scenes.door_scene(film=(W, H))
    else:
        data = scenes.SCENES[args.scene](film=(W, H))
    params = dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=8, directSamples=-1)
    gpu = Scene(data)

    def gpu_render(spp, seed):
        cfg = make_config(seed=seed, sampleCount=spp, **params)
        t0 = time.perf_counter()
        img, st = gpu.render(cfg)
        return img, time.perf_counter() - t0, st

    # ---- calibrate and run the GPU for the budget
    _, t1, st1 = gpu_render(256, 1)
    _, t2, st2 = gpu_render(1024, 2)
    rate = (st2.mutations - st1.mutations) / max(1e-6, t2 - t1)            # marginal mutations / s
    fixed = max(0.0, t1 - st1.mutations / rate)
    spp = max(1, int((args.seconds - fixed) * rate / (W * H)))
    img_gpu, t_gpu, st_gpu = gpu_render(spp, 3)
    # ---- long reference renders
    ref_spp = spp * args.ref_factor
    ref, t_ref, st_ref = gpu_render(ref_spp, 1234567)

    # ---- CPU oracle on all host cores for the same budget
    orc = oracle_lib.OracleScene(data)
    threads = os.cpu_count() or 1
    ocfg = oracle_lib.default_config(integrator=abi.DR_INTEGRATOR_DRMLT, technique=abi.DR_TECH_MMLT, type=abi.DR_TYPE_ORBITAL, max_depth=8,
                                     direct_samples=-1, direct_sampling=0, kelemen_style_weights=0, seed=5, ray_epsilon=1e-4, shadow_epsilon=1e-3)
    n_boot, n_chains = 100000 * 8 // 4, threads * 8                         # reference sizing: luminanceSamples x maxDepth (a quarter, it is timed)

    def cpu_render(steps, seed):
        ocfg.seed = seed
        t0 = time.perf_counter()
        r, img, st, sec = orc.render(ocfg, n_boot, n_chains, steps, threads)
        assert r == 0
        return img, time.perf_counter() - t0, st

    _, c1, s1 = cpu_render(200, 7)
    _, c2, s2 = cpu_render(2000, 8)
    crate = (s2.mutations - s1.mutations) / max(1e-6, c2 - c1)
    cfixed = max(0.0, c1 - s1.mutations / crate)
    steps = max(1, int((args.seconds - cfixed) * crate / n_chains))
    img_cpu, t_cpu, st_cpu = cpu_render(steps, 9)
    # long CPU render as an independent check of the reference image (bounded: ~6x the budget)
    img_cpu_long, t_cpu_long, st_cpu_long = cpu_render(steps * 6, 10)

    # ---- the reference's OWN integrator (oracle/_ref, tools/ref_sample.py) on all host cores for the same budget of render time
    # (its bootstrap included, its kd-tree build excluded -- as the BVH build is for the GPU)
    import subprocess
    import tempfile

    def ref_render(spp_, tag):
        path = os.path.join(tempfile.gettempdir(), "equal_time_ref_%s.npy" % tag)
        outp = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_sample.py"), "--config", {"door": "C5", "cornell": "C1", "glossy": "C3", "caustic": "C4"}.get(args.scene, "C5"), "--film", "%dx%d" % (W, H),
                               "--spp", str(spp_), "--out-image", path], capture_output=True, text=True, timeout=1200).stdout
        for ln in outp.splitlines():
            if "REF_SAMPLE " in ln:
                return np.load(path), json.loads(ln[ln.index("REF_SAMPLE ") + len("REF_SAMPLE "):])
        raise RuntimeError("tools/ref_sample.py produced no result")

    real = None
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_path.so")):
        try:
            lo = max(2, -(-(threads + 1) * 100000 // (W * H)))
            _, r1 = ref_render(lo, "a")
            _, r2 = ref_render(4 * r1["spp"], "b")
            rrate = (r2["mutations"] - r1["mutations"]) / max(1e-6, r2["render_s"] - r1["render_s"])
            rfixed = max(0.0, r1["render_s"] - r1["mutations"] / rrate)
            rspp = max(r1["spp"], int((args.seconds - rfixed) * rrate / (W * H)))
            img_real, r3 = ref_render(rspp, "c")
            img_real_long, r4 = ref_render(6 * rspp, "d")
            real = {"seconds": r3["render_s"], "mutations": r3["mutations"], "spp": r3["spp"], "threads": r3["threads"], "relMSE": relmse(img_real, ref),
                    "b": r3["mean_luminance"], "acceptance_percent": r3["stats_percent"], "kind": "the reference's own DRMLT::render (oracle/_ref)",
                    "long_run": {"seconds": r4["render_s"], "mutations": r4["mutations"], "relMSE_vs_reference_image": relmse(img_real_long, ref),
                                 "mean": float(img_real_long.mean())}}
        except Exception as ex:
            real = {"failed": repr(ex)}

    out = {"scene": args.scene, "triangles": int(data.n_triangles), "film": [W, H], "params": params, "budget_s": args.seconds,
           "gpu": {"seconds": t_gpu, "mutations": int(st_gpu.mutations), "spp": spp, "relMSE": relmse(img_gpu, ref), "b": st_gpu.luminance},
           "cpu_oracle": {"seconds": t_cpu, "mutations": int(st_cpu.mutations), "threads": threads, "relMSE": relmse(img_cpu, ref), "b": st_cpu.luminance,
                          "kind": "oracle port (double), not the reference binary"},
           "reference_image": {"spp": ref_spp, "mutations": int(st_ref.mutations), "seconds": t_ref, "mean": float(ref.mean())},
           "cross_check": {"cpu_long_seconds": t_cpu_long, "cpu_long_mutations": int(st_cpu_long.mutations),
                           "relMSE_cpu_long_vs_reference": relmse(img_cpu_long, ref), "mean_cpu_long": float(img_cpu_long.mean()),
                           "mean_gpu": float(img_gpu.mean()), "mean_cpu": float(img_cpu.mean())},
           "eps": 1e-2}
    out["cpu_reference"] = real
    if real and "relMSE" in real:
        out["relMSE_ratio_reference_over_gpu"] = real["relMSE"] / max(out["gpu"]["relMSE"], 1e-300)
    out["relMSE_ratio_cpu_over_gpu"] = out["cpu_oracle"]["relMSE"] / max(out["gpu"]["relMSE"], 1e-300)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
