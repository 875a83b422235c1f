#!/usr/bin/env python
"""sweep_e2e.py -- the whole C5 job (bench.py's e2e leg) under a list of launch-geometry settings, one process.

  python tools/sweep_e2e.py "DRMLT_GROUPS=3" "DRMLT_GROUPS=6,DRMLT_TRACE_CTAS=3" ...

Each argument is a comma-separated list of KEY=VALUE environment settings (the library reads its DRMLT_* knobs when a
job is created); "chains=N" / "lanes=N" / "spp=N" / "balance=0|1" are passed to the configuration instead.  Prints one JSON line per setting: device time of
the chain phase, whole-job wall time, mutations/s.  A tuning aid -- its numbers are not bench values.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__
    from drmlt_mitsuba_b200 import distributed, scenes
    from drmlt_mitsuba_b200.integrator import Scene
    __graft_entry__.build()
    torch.cuda.set_device(0)
    datas = {}

    def scene_data(name):
        if name not in datas:
            if name == "door":
                datas[name] = scenes.door_scene()
            elif name == "door_small":               # the same room at ~100 k triangles (cache-footprint sensitivity)
                datas[name] = scenes.door_scene(floor_grid=180, n_spheres=16, sphere_subdiv=3)
            else:
                datas[name] = scenes.SCENES[name]()
        return datas[name]
    reps = int(os.environ.get("SWEEP_REPS", "3"))
    base = dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=8, directSamples=-1)
    settings = sys.argv[1:] or [""]
    saved = {k: v for k, v in os.environ.items() if k.startswith("DRMLT_")}
    for setting in settings:
        for k in [k for k in os.environ if k.startswith("DRMLT_")]:
            del os.environ[k]
        os.environ.update(saved)
        params = dict(base, sampleCount=64)
        scene_name = "door"
        for kv in [x for x in setting.split(",") if x]:
            k, v = kv.split("=")
            if k == "scene":
                scene_name = v
            elif k == "spp":
                params["sampleCount"] = int(v)
            elif k in ("chains", "lanes"):
                params[k] = int(v)
            elif k == "balance":
                params["depthBalance"] = bool(int(v))
            else:
                os.environ[k] = v
        data = scene_data(scene_name)
        scene = Scene(data, device=0)                 # per setting: the BVH knobs are read at scene creation
        best = None
        for r in range(reps + 1):                     # first run warms the memory pool
            params["seed"] = 100 + r
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            scene.reupload()
            img, st, b = distributed.render(scene, params)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            row = {"setting": setting, "wall_s": wall, "chains_ms": st.chains_ms, "bootstrap_ms": st.bootstrap_ms, "mutations": int(st.mutations),
                   "e2e_Mmut_s": st.mutations / wall / 1e6, "chain_Mmut_s": st.mutations / st.chains_ms / 1e3, "rounds": int(st.rounds),
                   "rays": int(st.rays), "tris": int(data.n_triangles), "b": b, "mean": float(img.mean()),
                   "trace_ms": st.trace_ms, "walk_ms": st.walk_ms, "chain_ms": st.chain_ms}
            if r > 0 and (best is None or row["wall_s"] < best["wall_s"]):
                best = row
        print(json.dumps(best), flush=True)
        scene.close()


if __name__ == "__main__":
    main()
