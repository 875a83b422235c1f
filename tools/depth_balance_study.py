"""Depth-balanced chain lengths (dr_config.depth_balance) against equal lengths at EQUAL mutation count (GPU box):
relMSE of single renders and of the mean of the seeds (bias check: the mean's error must keep falling ~ 1 / seeds) against a
long equal-length reference, plus the wall time of each.   python tools/depth_balance_study.py [--spp 64] [--scene door]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__
__graft_entry__.build()
from drmlt_mitsuba_b200 import scenes
from drmlt_mitsuba_b200.integrator import Scene, make_config

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--film", default="320x180")
ap.add_argument("--scene", default="door")
ap.add_argument("--seeds", type=int, default=8)
ap.add_argument("--per", type=int, default=0, help="mutations per chain (0: the library's automatic chain count)")
ap.add_argument("--ref-spp", type=int, default=8192, dest="ref_spp")
args = ap.parse_args()
W, H = [int(x) for x in args.film.split("x")]
data = scenes.door_scene(film=(W, H)) if args.scene == "door" else scenes.SCENES[args.scene](film=(W, H))
gpu = Scene(data)
params = dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=8, directSamples=-1)


def relmse(img, ref, eps=1e-2):
    img, ref = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))


def render(spp, seed, **extra):
    if args.per and "chains" not in extra:
        extra["chains"] = W * H * spp // args.per
    cfg = make_config(seed=seed, sampleCount=spp, **dict(params, **extra))
    t0 = time.perf_counter()
    img, st = gpu.render(cfg)
    return img, time.perf_counter() - t0, st


ref, tr, sr = render(args.ref_spp, 99, depthBalance=False, chains=131072)
print("reference: %d mutations, %.1f s, b %.6f" % (sr.mutations, tr, sr.luminance))
out = {"film": [W, H], "scene": args.scene, "spp": args.spp, "ref_spp": args.ref_spp, "rows": []}
for balance in (False, True):
    errs, secs, imgs, st = [], [], [], None
    render(args.spp, 1000, depthBalance=balance)          # warm
    for seed in range(1, args.seeds + 1):
        img, t, st = render(args.spp, seed, depthBalance=balance)
        errs.append(relmse(img, ref)); secs.append(t); imgs.append(np.asarray(img, np.float64))
    mean_img = np.mean(imgs, axis=0)
    row = {"depthBalance": balance, "per": args.per, "relMSE": float(np.mean(errs)), "relMSE_runs": errs, "relMSE_of_mean": relmse(mean_img, ref),
           "mean_luminance": float(mean_img.mean()), "ref_mean": float(np.asarray(ref).mean()),
           "seconds": float(np.mean(secs)), "mutations": int(st.mutations), "rounds": int(st.rounds), "rays": int(st.rays), "chains_ms": st.chains_ms}
    out["rows"].append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "depth_balance_%s_per%d.json" % (args.scene, args.per)), "w"), indent=1)
