"""Diagnostic (GPU box): BDPT f(u) parity under parameter variants, to localise mismatches."""
import ctypes as C
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__
__graft_entry__.build()
from drmlt_mitsuba_b200 import abi
from drmlt_mitsuba_b200.integrator import make_config
import test_gpu_parity as T

scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
gpu, orc, data = T.pair(scene)
variants = [dict(maxDepth=6), dict(maxDepth=8, rrDepth=2)]
for var in variants:
    params = dict(integrator="pssmlt", technique="bdpt", directSamples=-1, directSampling=False)
    params.update(var)
    cfg = make_config(seed=3, **params)
    n = 20000
    rng = np.random.RandomState(5)
    md = cfg.max_depth
    depth = np.ones(n, np.int32)
    ds = de = 3 * (md + 2)
    us, ue, ud = [rng.rand(n, k).astype(np.float32) for k in (ds, de, 1)]
    og = gpu.eval_paths(cfg, us, ue, ud, depth)
    oc, lum64 = orc.eval_paths(T.ocfg(cfg), us, ue, ud, depth)
    g = np.frombuffer(og, dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
    c = np.frombuffer(oc, dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
    lg = g[:, 0:4].copy().view("<f4")[:, 0].astype(np.float64)
    ng_, nc_ = g[:, 4:8].copy().view("<i4")[:, 0], c[:, 4:8].copy().view("<i4")[:, 0]
    rg, rc = g[:, -4:].copy().view("<i4")[:, 0], c[:, -4:].copy().view("<i4")[:, 0]
    both = (lg > 0) & (lum64 > 0)
    rel = np.zeros(n); rel[both] = np.abs(lg[both] - lum64[both]) / lum64[both]
    bad = both & (rel > 1e-4)
    print("%s: contributing %d, bad %d (%.4f), support mismatch %d, nsplat mismatch %d, raydiff %d, gpu-minus-cpu lum sign: +%d -%d, rays gpu>cpu %d gpu<cpu %d" % (
        var, both.sum(), bad.sum(), bad.mean(), ((lg > 0) != (lum64 > 0)).sum(), (ng_ != nc_).sum(), (rg != rc).sum(),
        (bad & (lg > lum64)).sum(), (bad & (lg < lum64)).sum(), (rg > rc).sum(), (rg < rc).sum()))
    sg, tg = g[:, 8:12].copy().view("<i4")[:, 0], g[:, 12:16].copy().view("<i4")[:, 0]
    sc_, tc = c[:, 8:12].copy().view("<i4")[:, 0], c[:, 12:16].copy().view("<i4")[:, 0]
    print("    ns differs %d (gpu>cpu %d), nt differs %d (gpu>cpu %d); ns hist gpu %s cpu %s" % ((sg != sc_).sum(), (sg > sc_).sum(), (tg != tc).sum(), (tg > tc).sum(),
          np.bincount(sg.clip(0), minlength=10), np.bincount(sc_.clip(0), minlength=10)))
    for i in np.nonzero(bad)[0][:3]:
        print("    ex %d: gpu %.7g cpu %.7g rel %.2e nsplat %d/%d rays %d/%d" % (i, lg[i], lum64[i], rel[i], ng_[i], nc_[i], rg[i], rc[i]))
