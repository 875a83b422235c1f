"""Writes tests/golden/ref_rough_tables.npz -- the reference's own RoughTransmittance (src/bsdfs/rtrans.h, data/microfacet/*.dat) reduced to
the (distribution, eta, alpha) of the roughplastic test materials, plus eval() of the reduced table at probe angles -- and ADDS the
roughplastic entries (BSDF plugin sample / eval / pdf, PathSampler::sampleSplats on the roughplastic Cornell box) to
tests/golden/ref_path.npz, leaving its other entries untouched.  Run where /root/reference is; the fixtures travel."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_path_cases as RP  # noqa: E402

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref")])
lib = C.CDLL(RP.REF_PATH)
PD = C.POINTER(C.c_double)
lib.ref_rough_table.argtypes = [C.c_int, C.c_double, C.c_double, PD, PD, C.c_int, PD]
probe = np.concatenate([np.linspace(0.0, 1.0, 257), [-0.25, 1e-9, 0.999999999]])
out = {"probe": probe}
for key, (ggx, eta, alpha) in RP.ROUGH_TABLES.items():
    a32 = float(np.float32(alpha))
    a = (a32 + a32 + a32) * float(np.float32(1.0) / np.float32(3.0))        # Spectrum::average of the constant alpha texture (spectrum.h:481-486)
    table, got = np.zeros(104), np.zeros(len(probe))
    assert lib.ref_rough_table(ggx, float(np.float32(eta)), a, table.ctypes.data_as(PD), probe.ctypes.data_as(PD), len(probe), got.ctypes.data_as(PD)) == 0
    out[key] = table
    out[key + "_probe"] = got
np.savez_compressed(RP.GOLDEN_ROUGH, **out)
print("wrote", RP.GOLDEN_ROUGH, os.path.getsize(RP.GOLDEN_ROUGH), "bytes")

gold = dict(np.load(RP.GOLDEN))
for k, v in RP.run_bsdf(lib, "ref_").items():
    if "roughplastic" in k:
        gold[k] = v
for case in RP.PATH_CASES:
    if case[0] != "roughplastic":
        continue
    r = RP.run_paths_ref(lib, case)
    k = RP.case_key(case)
    gold[k + "_lum"] = r["lum"]
    gold[k + "_st"] = np.stack([r["s"], r["t"], r["n_splats"]], 1).astype(np.int8)
    gold[k + "_pos0"] = r["pos0"]
    gold[k + "_value0"] = r["value0"]
    print(k, "contributing", int((r["lum"] > 0).sum()), "of", len(r["lum"]))
np.savez_compressed(RP.GOLDEN, **gold)
print("wrote", RP.GOLDEN, os.path.getsize(RP.GOLDEN), "bytes")
os._exit(0)
