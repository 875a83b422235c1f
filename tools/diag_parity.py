"""Diagnostic (GPU box): where do f(u) mismatches between the CUDA path and the oracle come from?"""
import ctypes as C
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__
__graft_entry__.build()
import oracle_lib
from drmlt_mitsuba_b200 import abi, scenes
from drmlt_mitsuba_b200.integrator import Scene, make_config
import test_gpu_parity as T

for case in T.CASES:
    name, params = case
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    gpu, orc, data = T.pair(name)
    cfg = make_config(seed=3, **params)
    n = 40000
    rng = np.random.RandomState(5)
    md = cfg.max_depth
    depth = rng.randint(1, md + 1, n).astype(np.int32)
    ds, de, dd = (50, 2, 2) if cfg.technique == abi.DR_TECH_PATH else (3 * (md + 2), 3 * (md + 2), 1)
    us, ue, ud = [rng.rand(n, k).astype(np.float32) for k in (ds, de, dd)]
    og = gpu.eval_paths(cfg, us, ue, ud, depth)
    oc, lum64 = orc.eval_paths(T.ocfg(cfg), us, ue, ud, depth)
    g = np.frombuffer(og, dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
    c = np.frombuffer(oc, dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
    lg = g[:, 0:4].copy().view("<f4")[:, 0].astype(np.float64)
    s = c[:, 8:12].copy().view("<i4")[:, 0]; t = c[:, 12:16].copy().view("<i4")[:, 0]
    mg, mc = g[:, 16:20].copy().view("<f4")[:, 0], c[:, 16:20].copy().view("<f4")[:, 0]
    rg, rc = g[:, -4:].copy().view("<i4")[:, 0], c[:, -4:].copy().view("<i4")[:, 0]
    both = (lg > 0) & (lum64 > 0)
    supp = (lg > 0) != (lum64 > 0)
    rel = np.zeros(n); rel[both] = np.abs(lg[both] - lum64[both]) / lum64[both]
    bad = both & (rel >= 1e-4)
    print("== %s: n=%d contributing=%d support-mismatch=%d bad=%d frac_all=%.5f frac_contrib=%.5f raydiff=%d" % (
        T._case_id(case), n, both.sum(), supp.sum(), bad.sum(), 1 - (bad.sum() + supp.sum()) / n, 1 - bad.sum() / max(1, both.sum()), (rg != rc).sum()))
    if bad.any():
        q = np.percentile(rel[bad], [50, 90, 99])
        print("   rel err of bad paths: median %.2e p90 %.2e p99 %.2e; bad with same raycount %d; mis-weight rel err median %.2e" % (
            q[0], q[1], q[2], (bad & (rg == rc)).sum(), np.median(np.abs(mg[bad] - mc[bad]) / np.maximum(mc[bad], 1e-30))))
        if cfg.technique == abi.DR_TECH_MMLT:
            for tt in range(0, 10):
                m = both & (t == tt)
                if m.sum():
                    print("   t=%d: contributing %d bad %d (%.4f)" % (tt, m.sum(), (bad & m).sum(), (bad & m).sum() / m.sum()))
        idx = np.nonzero(bad)[0][:6]
        for i in idx:
            print("   ex %d: depth %d s %d t %d gpu %.7g cpu %.7g rel %.2e mis %.6g/%.6g rays %d/%d" % (i, depth[i], s[i], t[i], lg[i], lum64[i], rel[i], mg[i], mc[i], rg[i], rc[i]))
    if supp.any():
        print('   supp: gpu zero %d, cpu zero %d; by depth' % ((supp & (lg == 0)).sum(), (supp & (lum64 == 0)).sum()), np.bincount(depth[supp], minlength=9), 's', np.bincount(s[supp].clip(0), minlength=9), 'gpu lum percentiles', np.percentile(lg[supp & (lg > 0)], [0, 50, 100]) if (supp & (lg > 0)).any() else None, 'cpu', np.percentile(lum64[supp & (lum64 > 0)], [0, 50, 100]) if (supp & (lum64 > 0)).any() else None)
        print('   all contributing lum percentiles', np.percentile(lum64[both], [0, 1, 50, 99, 100]))
        idx = np.nonzero(supp)[0][:12]
        for i in idx:
            print("   supp %d: depth %d s %d t %d gpu %.7g cpu %.7g rays %d/%d" % (i, depth[i], s[i], t[i], lg[i], lum64[i], rg[i], rc[i]))
