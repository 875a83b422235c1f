"""Writes tests/golden/ref_texture.npz: the REFERENCE'S OWN bitmap-texture code (oracle/_ref/libref_path.so, compiled by oracle/ref/Makefile
from /root/reference) on the seeded cases of tests/ref_path_cases.py --
  tex_*     Texture2D::eval -> TMIPMap::evalBilinear / evalBox / evalTexel (texture.cpp:112-121, mipmap.h:503-596) at 4 000 uv pairs for
            every wrap mode, both filters, scaled / offset coordinates
  path_textured_*   PathSampler::sampleSplats (MMLT / BDPT / PT) on the textured Cornell box: the BSDF plugins with <texture> children, TriMesh
            texture coordinates and UV tangents (skdtree.h:343-426, trimesh.cpp:708-760).
Run in the build container only (needs /root/reference compiled into oracle/_ref); the fixture is what travels."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_path_cases as RP  # noqa: E402

lib = C.CDLL(RP.REF_PATH)
out = RP.run_texture(lib, "ref_")
for case in RP.TEXTURE_CASES:
    r = RP.run_paths_ref(lib, case)
    k = RP.case_key(case)
    out[k + "_lum"] = r["lum"]
    out[k + "_st"] = np.stack([r["s"], r["t"], r["n_splats"]], 1).astype(np.int8)
    out[k + "_pos0"] = r["pos0"]
    out[k + "_value0"] = r["value0"]
    print(k, "contributing", int((r["lum"] > 0).sum()), "of", len(r["lum"]))
# ---- the reference's own DRMLT integrator end to end on the textured Cornell box (three runs: it seeds from /dev/urandom), each in a
# process of its own (tools/plugin_render.py --reference --textured): what the drop-in plugin's job is held against
import json  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402
runs = []
for i in range(3):
    with tempfile.TemporaryDirectory() as td:
        npy = os.path.join(td, "img.npy")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "plugin_render.py"), "drmlt_orbital_mmlt", npy, "--reference", "--textured"],
                           capture_output=True, text=True, timeout=1200)
        line = [ln for ln in p.stdout.splitlines() if "PLUGIN_RENDER " in ln][-1]
        info = json.loads(line[line.index("PLUGIN_RENDER ") + len("PLUGIN_RENDER "):])
        assert info["ok"], info
        runs.append((np.load(npy), info["stats"]))
names = sorted(runs[0][1])
out["textured_image"] = runs[0][0]
out["textured_stats_names"] = np.array(names)
out["textured_stats"] = np.array([[r[1][k] for k in names] for r in runs])
out["textured_b"] = np.array([RP.luminance(r[0]).mean() for r in runs])
print("reference render of the textured box: b =", out["textured_b"], dict(zip(names, out["textured_stats"].T.round(2).tolist())))
np.savez_compressed(RP.GOLDEN_TEXTURE, **out)
print("wrote", RP.GOLDEN_TEXTURE, os.path.getsize(RP.GOLDEN_TEXTURE), "bytes")
os._exit(0)
