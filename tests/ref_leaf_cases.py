"""Seeded leaf cases run through a library that exports the numerical leaves of the hot path with a prefix:
`ref_` = oracle/_ref/libref_leaf.so (the reference's OWN code, compiled from /root/reference by oracle/ref/Makefile),
`orc_` = oracle/_build/liboracle.so (the restatement).  Both sides get the same inputs; the outputs of the `ref_`
side are committed as tests/golden/ref_leaf.npz by tools/make_ref_golden.py.  Test infrastructure only.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LEAF = os.path.join(ROOT, "oracle", "_ref", "libref_leaf.so")
ORACLE = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_leaf.npz")

D = C.c_double
PD = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(PD)


def _unit(rng, n, upper=False):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if upper:
        v[:, 2] = np.abs(v[:, 2])
    return np.ascontiguousarray(v)


def run_cases(lib, prefix, n=4000, seed=20261018):
    """-> dict name -> ndarray of outputs; inputs depend on (n, seed) only."""
    rng = np.random.default_rng(seed)
    out = {}

    def fn(name, restype, argtypes):
        f = getattr(lib, prefix + name)
        f.restype, f.argtypes = restype, argtypes
        return f

    uv = rng.random((n, 2))
    uv[:8] = [[0, 0], [1, 1], [0.5, 0.5], [0, 1], [1, 0], [0.25, 0.75], [0.5, 0], [1e-9, 1 - 1e-9]]
    for name, k in (("squareToCosineHemisphere", 3), ("squareToUniformDiskConcentric", 2), ("squareToUniformTriangle", 2)):
        f = fn(name, None, [D, D, PD])
        r = np.zeros((n, k))
        for i in range(n):
            f(uv[i, 0], uv[i, 1], _p(r[i]))
        out[name] = r

    # Fresnel terms
    cos_i = rng.uniform(-1, 1, n)
    cos_i[:4] = [1.0, -1.0, 0.0, 1e-6]
    etas = rng.choice([1.0, 1.5, 1 / 1.5, 1.33, 2.4, 1.0001], n)
    f = fn("fresnelDielectricExt", D, [D, D, PD])
    r = np.zeros((n, 2))
    for i in range(n):
        ct = D(0)
        r[i, 0] = f(cos_i[i], etas[i], C.byref(ct))
        r[i, 1] = ct.value
    out["fresnelDielectricExt"] = r
    f = fn("fresnelConductorExact", None, [D, PD, PD, PD])
    ce, ck = rng.uniform(0.1, 3.0, (n, 3)), rng.uniform(0.0, 7.0, (n, 3))
    r = np.zeros((n, 3))
    for i in range(n):
        f(abs(cos_i[i]), _p(ce[i]), _p(ck[i]), _p(r[i]))
    out["fresnelConductorExact"] = r
    if prefix == "ref_":
        f = fn("fresnelDiffuseReflectance", D, [D, C.c_int])
        out["fresnelDiffuseReflectance"] = np.array([f(e, 0) for e in (1.5, 1 / 1.5, 1.33, 1 / 1.33, 1.9, 1 / 1.9)])
    else:
        f = fn("fresnelDiffuseReflectance", D, [D])
        out["fresnelDiffuseReflectance"] = np.array([f(e) for e in (1.5, 1 / 1.5, 1.33, 1 / 1.33, 1.9, 1 / 1.9)])

    # frames and luminance
    a = _unit(rng, n)
    a[:3] = [[0, 0, 1], [1, 0, 0], [0, 1, 0]]
    f = fn("coordinateSystem", None, [PD, PD, PD])
    r = np.zeros((n, 6))
    for i in range(n):
        f(_p(a[i]), _p(r[i, :3]), _p(r[i, 3:]))
    out["coordinateSystem"] = r
    rgb = rng.uniform(0, 20, (n, 3))
    f = fn("luminance", D, [PD])
    out["luminance"] = np.array([f(_p(rgb[i])) for i in range(n)])

    # TriAccel: rays aimed at (and near the edges of) random triangles
    tri = rng.uniform(-1, 1, (n, 3, 3))
    bary = rng.dirichlet([1, 1, 1], n)
    bary[::5] += rng.normal(scale=0.3, size=(len(bary[::5]), 3))          # a fifth of the rays miss
    target = np.einsum("nk,nkd->nd", bary, tri)
    org = rng.uniform(-3, 3, (n, 3))
    d = target - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org, d = np.ascontiguousarray(org), np.ascontiguousarray(d)
    f = fn("triAccel", C.c_int, [PD, PD, PD, PD, PD, D, D, PD])
    r = np.zeros((n, 4))
    for i in range(n):
        t = np.ascontiguousarray(tri[i])
        hit = f(_p(t[0]), _p(t[1]), _p(t[2]), _p(org[i]), _p(d[i]), 1e-4, 1e30, _p(r[i, 1:]))
        r[i, 0] = hit
        if hit != 1:
            r[i, 1:] = 0
    out["triAccel"] = r

    # DiscreteDistribution (with zero-probability entries, pmf.h:124-135)
    w = rng.uniform(0, 1, 257)
    w[rng.random(257) < 0.2] = 0.0
    w[0] = 0.0
    w[-1] = 0.0
    xi = rng.random(n)
    xi[:3] = [0.0, 1.0 - 2 ** -53, 0.5]
    f = fn("pmf", D, [PD, C.c_int, PD, C.c_int, C.POINTER(C.c_int32), PD, PD])
    idx, reused, pmf = np.zeros(n, np.int32), np.zeros(n), np.zeros(257)
    total = f(_p(w), 257, _p(xi), n, idx.ctypes.data_as(C.POINTER(C.c_int32)), _p(reused), _p(pmf))
    out["pmf_sum"] = np.array([total])
    out["pmf_index"] = idx.astype(np.float64)
    out["pmf_reused"] = reused
    out["pmf_pmf"] = pmf

    # microfacet distribution: Beckmann / GGX x sampleAll / sampleVisible
    f = fn("microfacet", None, [C.c_int, D, C.c_int, PD, PD, D, D, PD])
    wi, mm = _unit(rng, n, upper=True), _unit(rng, n, upper=True)
    wi[:, 2] = np.maximum(wi[:, 2], 1e-3)
    wi /= np.linalg.norm(wi, axis=1, keepdims=True)
    wi = np.ascontiguousarray(wi)
    alphas = rng.choice([0.05, 0.1, 0.3, 0.6], n)
    uv2 = rng.random((n, 2))
    for t in (0, 1):
        for vis in (0, 1):
            r = np.zeros((n, 8))
            for i in range(n):
                f(t, alphas[i], vis, _p(wi[i]), _p(mm[i]), uv2[i, 0], uv2[i, 1], _p(r[i]))
            out["microfacet_%s_%s" % ("ggx" if t else "beckmann", "visible" if vis else "all")] = r

    # transition kernels (drmlt/tools/transition.h)
    xi1, xi2 = rng.random(n), rng.random(n)
    s2 = 1.0 / 64
    s1 = s2 / 16
    f = fn("kelemen_sample", D, [D, D, D])
    ks = np.array([f(s1, s2, x) for x in xi1])
    out["kelemen_sample"] = ks
    f = fn("kelemen_pdf", D, [D, D, D])
    du = np.concatenate([ks[: n // 2], rng.uniform(-2 * s2, 2 * s2, n - n // 2)])
    out["kelemen_pdf"] = np.array([f(s1, s2, x) for x in du])
    f = fn("kelemen_logpdf", D, [D, D, D])
    out["kelemen_logpdf"] = np.array([f(s1, s2, x) for x in ks])
    f = fn("gaussian_sample", D, [D, D, D])
    out["gaussian_sample"] = np.array([f(s2 * 0.1, a_, b_) for a_, b_ in zip(xi1, xi2)])
    f = fn("cauchy_sample", D, [D, D])
    out["cauchy_sample"] = np.array([f(float(np.exp(-0.25)), x) for x in xi1])
    return out


def load(path):
    return C.CDLL(path)
