"""plugin_render.py CASE OUT.npy -- one render job through the drop-in Mitsuba plugin (test infrastructure).

oracle/_ref/plugins/{drmlt,pssmlt}.so are drmlt-mitsuba_b200/shim/mts_plugin.cpp compiled against the reference's headers and
libraries (oracle/ref/Makefile).  With REF_PLUGIN_DIR set, the reference's plugin manager (oracle/ref/ref_path.cpp) dlopen()s
them the way the reference loads plugins/<name>.so, and the job runs through the reference's own RenderJob -> Scene::render ->
Integrator::render.  Prints one JSON line: the statistics counters the plugin published (by the reference's names), render time.
Run in a process of its own (tests/ do, with a timeout): the reference's scheduler threads belong to this process."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_path_cases as RP  # noqa: E402

name, out = sys.argv[1], sys.argv[2]
spp = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else None
if "--analytic" in sys.argv:            # a room of ANALYTIC shapes (rectangles, a sphere; ref_path.cpp scene_create): the plugin tessellates them
    os.environ["REF_ANALYTIC_SCENE"] = "1"
data = None
if "--textured" in sys.argv:            # the Cornell box with bitmap textures: the plugin enumerates and flattens the BSDFs' textures
    from drmlt_mitsuba_b200 import scenes
    data = scenes.cornell_box_textured(film=(64, 64), tess=4)
for a in sys.argv:
    if a.startswith("--scene="):        # any scene of tests/ref_path_cases.py SCENES (at its test size)
        data = RP.SCENES[a[len("--scene="):]]()
if "--reference" not in sys.argv:       # (--reference: the reference's own integrator instead of the plugin -- writes the fixture)
    os.environ["REF_PLUGIN_DIR"] = os.path.join(ROOT, "oracle", "_ref", "plugins")
params, case_spp = RP.RENDER_CASES[name]
try:
    img, sec, _, stats = RP.run_render_ref(C.CDLL(RP.REF_PATH), params, spp or case_spp, threads=2 if "--reference" not in sys.argv else (os.cpu_count() or 2), data=data)
except AssertionError as e:
    print("PLUGIN_RENDER " + json.dumps({"ok": False, "error": str(e)}), flush=True)
    os._exit(3)
np.save(out, img)
print("PLUGIN_RENDER " + json.dumps({"ok": True, "stats": stats, "seconds": sec}), flush=True)
os._exit(0)
