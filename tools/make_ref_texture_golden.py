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
np.savez_compressed(RP.GOLDEN_TEXTURE, **out)
print("wrote", RP.GOLDEN_TEXTURE, os.path.getsize(RP.GOLDEN_TEXTURE), "bytes")
os._exit(0)
