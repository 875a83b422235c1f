#!/usr/bin/env python
"""bvh_gpu_check.py -- the device-built LBVH (dr_scene_create_ex, DR_SCENE_BVH_GPU) against the host's binned-SAH tree on the GPU
box: build times, node counts, identical closest hits on random rays, traversal time of the same rays.  A tuning aid."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__
__graft_entry__.build()
import torch
from drmlt_mitsuba_b200 import scenes, abi
from drmlt_mitsuba_b200.integrator import Scene, make_config
from test_gpu_parity import _random_rays

dt = [("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<i4")]
for name, data in (("door", scenes.door_scene()), ("glossy", scenes.glossy_scene(film=(128, 128), subdiv=5)), ("caustic", scenes.caustic_scene(film=(128, 128), grid=96))):
    row = {"scene": name, "tris": int(data.n_triangles)}
    hits = {}
    rays = _random_rays(data, 400000, 3)
    for which in (False, True, True):
        t0 = time.perf_counter()
        s = Scene(data, gpu_bvh=which)
        wall = time.perf_counter() - t0
        info = s.bvh_info()
        s.trace(rays)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        h = np.frombuffer(s.trace(rays), dtype=dt).copy()
        trace_s = time.perf_counter() - t1
        key = "gpu" if which else "host"
        row[key] = dict(info, scene_create_s=wall, trace_call_s=trace_s)
        hits[key] = h
        if data.n_triangles < 200000:
            img, st = s.render(make_config(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6, directSamples=-1, sampleCount=16, seed=3))
            row[key]["render"] = dict(mutations=int(st.mutations), rays=int(st.rays), accept=int(st.accept), lum=st.luminance, mean=float(img.mean()))
            hits[key + "_img"] = img
        s.close()
    a, b = hits["host"], hits["gpu"]
    same_hit = (a["prim"] >= 0) == (b["prim"] >= 0)
    both = (a["prim"] >= 0) & (b["prim"] >= 0)
    row["hit_existence_equal"] = float(same_hit.mean())
    row["same_prim"] = float((a["prim"][both] == b["prim"][both]).mean())
    row["t_equal"] = float((a["t"][both] == b["t"][both]).mean())
    row["hit_rate"] = float((a["prim"] >= 0).mean())
    if "host_img" in hits:
        row["image_max_abs_diff"] = float(np.abs(hits["host_img"] - hits["gpu_img"]).max())
    print(json.dumps(row), flush=True)
