"""Statistical efficiency vs number of resident chains at EQUAL mutation count (GPU box).
relMSE against a long render made of few, long chains.  python tools/chain_length_study.py [--spp 1024]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__
__graft_entry__.build()
from drmlt_mitsuba_b200 import scenes
from drmlt_mitsuba_b200.integrator import Scene, make_config

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--film", default="320x180")
ap.add_argument("--ref-spp", type=int, default=16384, dest="ref_spp")
args = ap.parse_args()
W, H = [int(x) for x in args.film.split("x")]
data = scenes.door_scene(film=(W, H))
gpu = Scene(data)
params = dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=8, directSamples=-1)

def relmse(img, ref, eps=1e-2):
    img, ref = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))

def render(spp, chains, seed, **extra):
    cfg = make_config(seed=seed, sampleCount=spp, chains=chains, **dict(params, **extra))
    t0 = time.perf_counter()
    img, st = gpu.render(cfg)
    return img, time.perf_counter() - t0, st

ref, tr, sr = render(args.ref_spp, 131072, 99)
print("reference: %d mutations, %.1f s" % (sr.mutations, tr))
rows = []
for chains in (16384, 65536, 262144, 1048576, 4194304):
    for lum in (100000, 2000000):
        errs, secs = [], []
        for seed in (1, 2, 3):
            img, t, st = render(args.spp, chains, seed, luminanceSamples=lum)
            errs.append(relmse(img, ref)); secs.append(t)
        row = {"chains": chains, "luminanceSamples": lum, "mutations_per_chain": st.mutations // chains, "relMSE": float(np.mean(errs)), "relMSE_runs": errs,
               "seconds": float(np.mean(secs)), "mutations": int(st.mutations), "relMSE_x_seconds": float(np.mean(errs) * np.mean(secs))}
        rows.append(row)
        print(json.dumps(row))
json.dump({"film": [W, H], "spp": args.spp, "ref_spp": args.ref_spp, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "chain_length_study.json"), "w"), indent=1)
