import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from drmlt_mitsuba_b200 import scenes
from drmlt_mitsuba_b200.integrator import Scene, make_config
out = {}
for tech, extra in (("mmlt", dict(type="orbital")), ("path", dict(type="mira")), ("bdpt", dict(type="green", directSampling=False))):
    for label in ("textured", "averages"):
        data = scenes.cornell_box_textured(film=(512, 512), tess=8)
        if label == "averages":
            for m in data.materials:
                m.flags &= 0xff
        sc = Scene(data)
        cfg = make_config(integrator="drmlt", technique=tech, maxDepth=8, directSamples=-1, sampleCount=64, seed=3, **extra)
        sc.render(cfg)
        best = 1e9
        for _ in range(3):
            t = time.perf_counter(); img, st = sc.render(cfg); best = min(best, time.perf_counter() - t)
        out["%s_%s" % (tech, label)] = dict(seconds=best, mutations=int(st.mutations), Mmut_s=st.mutations / best / 1e6, rays=int(st.rays))
        sc.close()
print(json.dumps(out))
