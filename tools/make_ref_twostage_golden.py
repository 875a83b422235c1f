"""Writes tests/golden/ref_twostage.npz: the REFERENCE'S OWN two-stage MLT (twoStage=true: BidirectionalUtils::mltLuminancePass,
src/libbidir/util.cpp:96-199 -- nested low-resolution job, luminance map, Bitmap::resample -- then the main job with
SplatList::normalize(importanceMap) and develop x importance) end to end in oracle/_ref, three runs (it seeds from /dev/urandom).
Two local workers: the nested job has 6 work units, and the reference renders nothing when workUnits < cores.
Run in the container that has /root/reference."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_path_cases as RP  # noqa: E402

lib = C.CDLL(RP.REF_PATH)
runs = [RP.run_render_ref(lib, RP.TWOSTAGE_PARAMS, RP.TWOSTAGE_SPP, threads=2) for _ in range(3)]
names = sorted(runs[0][3])
out = {"twostage_image": runs[0][0], "twostage_b": np.array([RP.luminance(r[0]).mean() for r in runs]), "twostage_stats_names": np.array(names),
       "twostage_stats": np.array([[r[3][k] for k in names] for r in runs])}
print(out["twostage_b"], dict(zip(names, out["twostage_stats"].T.round(2).tolist())))
np.savez_compressed(RP.GOLDEN_TWOSTAGE, **out)
print("wrote", RP.GOLDEN_TWOSTAGE, os.path.getsize(RP.GOLDEN_TWOSTAGE), "bytes")
os._exit(0)
