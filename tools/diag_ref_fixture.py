"""Diagnostic: the CUDA path against the reference's own PathSampler outputs (tests/golden/ref_path.npz), with the
float-build epsilons the product defaults to and with the double-build epsilons the reference was compiled with."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_path_cases as RP  # noqa: E402
from drmlt_mitsuba_b200.integrator import Scene  # noqa: E402

gold = np.load(RP.GOLDEN)
scenes = {}
for case in RP.PATH_CASES:
    k = RP.case_key(case)
    if case[0] not in scenes:
        scenes[case[0]] = Scene(RP.SCENES[case[0]](), device=0)
    us, ue, ud, depth = RP.case_inputs(case)
    want = gold[k + "_lum"]
    for eps in ((0.0, 0.0), (1e-7, 1e-5)):
        cfg = RP.case_config(case)
        cfg.ray_epsilon, cfg.shadow_epsilon = eps
        out = scenes[case[0]].eval_paths(cfg, us, ue, ud, depth)
        n = len(want)
        import ctypes as C
        from drmlt_mitsuba_b200 import abi
        r = np.frombuffer(out, dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
        lum = r[:, 0:4].copy().view("<f4")[:, 0].astype(np.float64)
        top = want.max()
        a = np.where(lum < 1e-18 * top, 0, lum)
        b = np.where(want < 1e-18 * top, 0, want)
        sup = (a > 0) == (b > 0)
        both = (a > 0) & (b > 0)
        rel = np.abs(a[both] - b[both]) / b[both]
        ok = sup.copy()
        ok[np.nonzero(both)[0][rel >= 1e-4]] = False
        print("%-34s eps %-14s within 1e-4: %.5f  support mismatches %4d  median %.2e  99.9%% %.2e" % (k, eps, ok.mean(), (~sup).sum(), np.median(rel), np.percentile(rel, 99.9)), flush=True)
