"""CPU tests of the oracle (oracle/): known answers and the reference's own run-time invariants.

The reference ships no golden vectors for this path, so the oracle is pinned by:
published known answers (Philox4x32-10), closed-form values of the transition kernels
(transition.h), the invariants the reference asserts at run time (seed replay, 0 <= a <= 1,
MMLT depth consistency), analytic checks (PT == BDPT == MMLT normalisation, energy conservation)
and BSDF sample/eval/pdf consistency modelled on src/tests/test_chisquare.cpp.
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib
from drmlt_mitsuba_b200 import abi, scenes


# ------------------------------------------------------------------ RNG
def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    kats = [
        ((0, 0, 0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), 0x299f31d0a4093822, (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    out = (C.c_uint32 * 4)()
    for ctr, key, want in kats:
        oracle.orc_philox(key, *ctr, out)
        assert tuple(out) == want


def test_keyed_uniform_range_and_addressing(oracle):
    vals = np.array([oracle.orc_uniform(42, 1, i, 0, j) for i in range(64) for j in range(8)])
    assert vals.min() >= 0.0 and vals.max() < 1.0
    assert abs(vals.mean() - 0.5) < 0.06
    # the uniform is (word >> 8) * 2^-24 of the addressed Philox word
    out = (C.c_uint32 * 4)()
    oracle.orc_philox(42, 5, 0, 2, (3 << 24) | (6 >> 2), out)
    assert oracle.orc_uniform(42, 3, 5, 2, 6) == np.float32((out[6 & 3] >> 8) * 2.0 ** -24)


# ------------------------------------------------------------------ transition kernels (transition.h:23-190)
def test_kelemen_kernel_closed_form(oracle):
    s1, s2 = 1 / 1024, 1 / 64
    assert oracle.orc_kelemen_sample(s1, s2, 0.0) == pytest.approx(s1, rel=1e-12)            # xi -> 0: +s1
    assert oracle.orc_kelemen_sample(s1, s2, 0.5) == pytest.approx(-s1, rel=1e-12)           # second half: negative
    assert oracle.orc_kelemen_sample(s1, s2, 0.5 - 1e-12) == pytest.approx(s2, rel=1e-9)
    assert oracle.orc_kelemen_sample(s1, s2, 0.25) == pytest.approx(math.sqrt(s1 * s2), rel=1e-12)
    # pdf = 1 / (2 |d| ln(s2/s1)) on [s1, s2], zero outside; integrates to one
    assert oracle.orc_kelemen_pdf(s1, s2, s1 / 2) == 0.0 and oracle.orc_kelemen_pdf(s1, s2, 2 * s2) == 0.0
    d = 0.004
    assert oracle.orc_kelemen_pdf(s1, s2, -d) == pytest.approx(1 / (2 * d * math.log(s2 / s1)), rel=1e-12)
    xs = np.exp(np.linspace(math.log(s1), math.log(s2), 20001))
    p = np.array([oracle.orc_kelemen_pdf(s1, s2, x) for x in xs])
    assert 2 * np.trapezoid(p, xs) == pytest.approx(1.0, rel=1e-4)   # end points may round outside [s1, s2]


def test_gaussian_and_cauchy_kernels(oracle):
    sigma = 0.1 / 64
    assert oracle.orc_gaussian_sample(sigma, 0.0, 0.3) == 0.0
    u1, u2 = 0.7, 0.2
    want = math.sqrt(-2 * math.log(1 - u1)) * math.cos(2 * math.pi * u2) * sigma
    assert oracle.orc_gaussian_sample(sigma, u1, u2) == pytest.approx(want, rel=1e-12)
    rho = math.exp(-0.25)
    # wrapped Cauchy by CDF inversion: symmetric, in [-pi, pi], median of |theta| = 2 atan((1-rho)/(1+rho))
    xs = (np.arange(20000) + 0.5) / 20000
    th = np.array([oracle.orc_cauchy_sample(rho, x) for x in xs])
    assert np.all(np.abs(th) <= math.pi + 1e-12)
    assert abs(np.mean(th)) < 1e-9
    assert np.median(np.abs(th)) == pytest.approx(2 * math.atan((1 - rho) / (1 + rho)), rel=2e-3)


def test_wrap_reflect(oracle):
    for y, want in [(0.25, 0.25), (1.25, 0.75), (-0.25, 0.25), (0.0, 0.0), (1.0, 1.0), (2.0, 0.0)]:
        assert oracle.orc_wrap(y) == pytest.approx(want)


# ------------------------------------------------------------------ findMaxDimensions (pssmlt_utils.h:27-77)
def test_max_dimensions(oracle):
    def dims(**kw):
        depth = kw.pop("depth", -1)
        cfg = oracle_lib.default_config(**kw)
        s, e, d = C.c_int(), C.c_int(), C.c_int()
        oracle.orc_max_dimensions(C.byref(cfg), depth, C.byref(s), C.byref(e), C.byref(d))
        return s.value, e.value, d.value
    assert dims(technique=abi.DR_TECH_PATH, max_depth=8) == (50, 0, 0)
    assert dims(technique=abi.DR_TECH_BDPT, max_depth=8, direct_sampling=0) == (30, 30, 0)    # C3
    assert dims(technique=abi.DR_TECH_BDPT, max_depth=8, direct_sampling=1) == (30, 30, 8)
    for d in range(1, 9):
        m = 3 * (d + 2)
        m += m & 1
        assert dims(technique=abi.DR_TECH_MMLT, max_depth=8, depth=d) == (m, m, 1)
    assert dims(technique=abi.DR_TECH_PATH, max_depth=4, rr_depth=5) == (24, 0, 0)            # no RR dimension


# ------------------------------------------------------------------ BSDFs (modelled on test_chisquare.cpp)
def _mat(type_, flags=0, alpha=0.2):
    m = abi.dr_material()
    m.type, m.flags = type_, flags
    m.reflectance[:] = (0.8, 0.7, 0.6)
    m.transmittance[:] = (1, 1, 1)
    m.eta[:] = (0.2, 0.92, 1.1) if type_ in (abi.DR_BSDF_CONDUCTOR, abi.DR_BSDF_ROUGHCONDUCTOR) else (1.5, 0, 0)
    m.k[:] = (3.9, 2.45, 2.14)
    m.alpha = alpha
    return m


def _sample(oracle, m, wi, mode, u1, u2):
    wo, w, pdf, ty = (C.c_double * 3)(), (C.c_double * 3)(), C.c_double(), C.c_int()
    oracle.orc_bsdf_sample(C.byref(m), (C.c_double * 3)(*wi), mode, u1, u2, wo, w, C.byref(pdf), C.byref(ty))
    return np.array(wo), np.array(w), pdf.value, ty.value


def _eval(oracle, m, wi, wo, mode, measure):
    v, pdf = (C.c_double * 3)(), C.c_double()
    oracle.orc_bsdf_eval(C.byref(m), (C.c_double * 3)(*wi), (C.c_double * 3)(*wo), mode, measure, v, C.byref(pdf))
    return np.array(v), pdf.value


@pytest.mark.parametrize("name,mat", [
    ("diffuse", _mat(abi.DR_BSDF_DIFFUSE)),
    ("twosided-diffuse", _mat(abi.DR_BSDF_DIFFUSE, abi.DR_MAT_TWOSIDED)),
    ("rough-ggx-vis", _mat(abi.DR_BSDF_ROUGHCONDUCTOR, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE)),
    ("rough-ggx-all", _mat(abi.DR_BSDF_ROUGHCONDUCTOR, abi.DR_MAT_GGX)),
    ("rough-beckmann-vis", _mat(abi.DR_BSDF_ROUGHCONDUCTOR, abi.DR_MAT_SAMPLE_VISIBLE, alpha=0.3)),
    ("rough-beckmann-all", _mat(abi.DR_BSDF_ROUGHCONDUCTOR, 0, alpha=0.3)),
])
def test_smooth_bsdf_sample_matches_eval_over_pdf(oracle, name, mat):
    rng = np.random.RandomState(3)
    wi = np.array([0.3, -0.2, 0.0]); wi[2] = math.sqrt(1 - wi[0] ** 2 - wi[1] ** 2)
    n_ok = 0
    for _ in range(400):
        u1, u2 = rng.rand(2)
        wo, w, pdf, ty = _sample(oracle, mat, wi, 0, u1, u2)
        if not w.any():
            continue
        f, p = _eval(oracle, mat, wi, wo, 0, 1)
        assert p == pytest.approx(pdf, rel=1e-6), name
        tol = 2e-2 if "beckmann-vis" in name else 1e-6   # Newton inversion is approximate (microfacet.h:573-644)
        assert np.allclose(f / p, w, rtol=tol, atol=1e-9), name
        n_ok += 1
    assert n_ok > 300


@pytest.mark.parametrize("flags", [abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE, abi.DR_MAT_GGX, 0])
def test_rough_conductor_pdf_integrates_to_one(oracle, flags):
    mat = _mat(abi.DR_BSDF_ROUGHCONDUCTOR, flags, alpha=0.4)
    wi = np.array([0.5, 0.1, 0.0]); wi[2] = math.sqrt(1 - 0.26)
    n = 200
    th = (np.arange(n) + 0.5) / n * (math.pi / 2)
    ph = (np.arange(2 * n) + 0.5) / (2 * n) * 2 * math.pi
    total = 0.0
    for t in th:
        for p in ph[::4]:
            wo = (math.sin(t) * math.cos(p), math.sin(t) * math.sin(p), math.cos(t))
            total += _eval(oracle, mat, wi, wo, 0, 1)[1] * math.sin(t)
    total *= (math.pi / 2 / n) * (2 * math.pi / (2 * n) * 4)
    # sampleAll loses the mass of microfacet normals facing away from wi / below the horizon
    assert 0.8 < total <= 1.01


def test_dielectric_is_delta_and_energy_conserving(oracle):
    mat = _mat(abi.DR_BSDF_DIELECTRIC)
    wi = np.array([0.6, 0.0, 0.8])
    wo_r, w_r, pdf_r, ty_r = _sample(oracle, mat, wi, 1, 0.0, 0.5)       # u <= F: reflection
    wo_t, w_t, pdf_t, ty_t = _sample(oracle, mat, wi, 1, 0.999, 0.5)     # refraction
    assert ty_r == 4 and ty_t == 8
    assert pdf_r + pdf_t == pytest.approx(1.0)
    assert np.allclose(wo_r, [-0.6, 0.0, 0.8])
    # Snell: sin_t = sin_i / eta
    assert math.hypot(wo_t[0], wo_t[1]) == pytest.approx(0.6 / 1.5, rel=1e-12) and wo_t[2] < 0
    # radiance transport scales by 1/eta^2 entering the denser medium (dielectric.cpp:318-320)
    _, w_rad, _, _ = _sample(oracle, mat, wi, 0, 0.999, 0.5)
    assert np.allclose(w_rad, 1 / 1.5 ** 2)
    f, p = _eval(oracle, mat, wi, wo_t, 1, 4)
    assert p == pytest.approx(pdf_t) and np.allclose(f / p, w_t)


# ------------------------------------------------------------------ film (imageblock.h:149-196, rfilter.cpp:37-55)
def test_film_splat_box_and_gaussian(oracle):
    W = H = 8
    film = np.zeros((H, W, 3), np.float32)
    pos = np.array([[3.5, 2.5]], np.float32)
    rgb = np.array([[1.0, 2.0, 3.0]], np.float32)
    oracle.orc_splat(W, H, abi.DR_FILTER_BOX, oracle_lib.fptr(pos), oracle_lib.fptr(rgb), 1, oracle_lib.fptr(film))
    assert np.count_nonzero(film[..., 0]) == 1
    w = 1.0 / (2 * (0.5 + 1e-5)) ** 2
    assert film[2, 3, 1] == pytest.approx(2.0 * w, rel=1e-5)
    film[:] = 0
    pos = np.array([[3.7, 2.2]], np.float32)
    oracle.orc_splat(W, H, abi.DR_FILTER_GAUSSIAN, oracle_lib.fptr(pos), oracle_lib.fptr(rgb), 1, oracle_lib.fptr(film))
    assert np.count_nonzero(film[..., 0]) == 16                       # radius 2 => 4x4 footprint
    assert film[..., 0].sum() == pytest.approx(1.0, rel=0.08)         # tabulated, normalised filter
    assert np.allclose(film[..., 2], 3 * film[..., 0])
    # negative or non-finite values are rejected (imageblock.h:151-160)
    bad = np.array([[1.0, -1.0, 0.0]], np.float32)
    before = film.copy()
    oracle.orc_splat(W, H, abi.DR_FILTER_GAUSSIAN, oracle_lib.fptr(pos), oracle_lib.fptr(bad), 1, oracle_lib.fptr(film))
    assert before.any() and np.array_equal(film, np.zeros_like(film))


# ------------------------------------------------------------------ path sampling invariants
@pytest.fixture(scope="module")
def cornell(oracle):
    return oracle_lib.OracleScene(scenes.cornell_box(film=(64, 64), tess=2))


def _cfg(**kw):
    kw.setdefault("max_depth", 6)
    kw.setdefault("direct_samples", -1)
    kw.setdefault("direct_sampling", 0)
    return oracle_lib.default_config(**kw)


def test_normalisation_agrees_across_techniques(cornell):
    """b from PT, BDPT and MMLT are unbiased estimators of the same integral once they cover the
    same path space.  With separateDirect (directSamples >= 0) all three drop depth <= 2
    (pathsampler.cpp:279-284, :559-561) and must agree.  With directSamples = -1 PT and MMLT still
    agree (both skip the directly visible emitter: path.cpp:152 / pathsampler.cpp:131-135) while
    BDPT additionally sees the emitter directly (s=0, t=2), a depth-independent surplus."""
    n = 60000
    for ds, md in ((0, 6), (-1, 4)):
        b = {}
        for name, tech in (("path", abi.DR_TECH_PATH), ("bdpt", abi.DR_TECH_BDPT), ("mmlt", abi.DR_TECH_MMLT)):
            cfg = _cfg(technique=tech, seed=11, direct_samples=ds, max_depth=md)
            k = md if tech == abi.DR_TECH_MMLT else 1
            lum, _ = cornell.bootstrap(cfg, 0, n * k)
            b[name] = lum.mean() * k
        assert b["path"] == pytest.approx(b["mmlt"], rel=0.03), (ds, b)
        if ds == 0:
            assert b["bdpt"] == pytest.approx(b["mmlt"], rel=0.03), (ds, b)
        else:
            assert b["bdpt"] > b["mmlt"] * 1.5, (ds, b)


def test_seed_replay_invariant_and_acceptance_bounds(cornell):
    """drmlt_proc.cpp:509-512: the replayed seed must reproduce the bootstrap luminance;
    flipCoin asserts 0 <= a <= 1 (drmlt_proc.cpp:419-422)."""
    for integ, tech, typ in ((abi.DR_INTEGRATOR_DRMLT, abi.DR_TECH_MMLT, abi.DR_TYPE_ORBITAL),
                             (abi.DR_INTEGRATOR_DRMLT, abi.DR_TECH_PATH, abi.DR_TYPE_MIRA),
                             (abi.DR_INTEGRATOR_DRMLT, abi.DR_TECH_BDPT, abi.DR_TYPE_GREEN),
                             (abi.DR_INTEGRATOR_PSSMLT, abi.DR_TECH_PATH, abi.DR_TYPE_MIRA)):
        cfg = _cfg(integrator=integ, technique=tech, type=typ, seed=5)
        lum, dep = cornell.bootstrap(cfg, 0, 3000)
        seeds = np.nonzero(lum > 0)[0][:48]
        rec, _, st = cornell.chain_steps(cfg, 1.0, seeds, dep[seeds], np.arange(len(seeds)), 40)
        r = np.ctypeslib.as_array(rec).view(np.recarray) if False else rec
        for c, s in enumerate(seeds):
            first = r[c * 40]
            assert first.L_x == pytest.approx(lum[s], rel=1e-6)
        a1 = np.array([x.a1 for x in r]); a2 = np.array([x.a2 for x in r])
        assert a1.min() >= 0 and a1.max() <= 1 and a2.min() >= 0 and a2.max() <= 1
        assert st.mutations == len(seeds) * 40
        # a second stage is only attempted after a rejected first stage
        assert not any(x.did_second and x.accept1 for x in r)
        assert not any(x.accept2 and not x.did_second for x in r)


def test_mmlt_depth_consistency(cornell):
    """pathsampler.cpp:176-185: an MMLT sample of depth d has s + t = d + 1 (light image on)."""
    cfg = _cfg(technique=abi.DR_TECH_MMLT, seed=9)
    n = 600
    rng = np.random.RandomState(0)
    depth = rng.randint(1, 7, n).astype(np.int32)
    us, ue, ud = rng.rand(n, 24).astype(np.float32), rng.rand(n, 24).astype(np.float32), rng.rand(n, 1).astype(np.float32)
    out, lum = cornell.eval_paths(cfg, us, ue, ud, depth)
    for i in range(n):
        assert out[i].s + out[i].t == depth[i] + 1
        assert out[i].n_splats in (0, 1)
        if out[i].n_splats:
            assert 0 < out[i].mis_weight <= 1.0 + 1e-9
            assert out[i].n_rays <= depth[i]          # a depth-d MMLT path costs at most d rays
    assert (lum > 0).sum() > n // 20


def test_chain_determinism_and_stats_bookkeeping(cornell):
    cfg = _cfg(technique=abi.DR_TECH_MMLT, type=abi.DR_TYPE_ORBITAL, seed=21)
    lum, dep = cornell.bootstrap(cfg, 0, 3000)
    seeds = np.nonzero(lum > 0)[0][:32]
    ids = np.arange(32)
    r1, f1, s1 = cornell.chain_steps(cfg, 1.0, seeds, dep[seeds], ids, 64, want_film=True, threads=1)
    r2, f2, s2 = cornell.chain_steps(cfg, 1.0, seeds, dep[seeds], ids, 64, want_film=True, threads=4)
    assert bytes(r1) == bytes(r2)
    assert np.allclose(f1, f2, rtol=1e-5, atol=1e-7)
    # Appendix A.6 bookkeeping identities
    assert s1.first_base == s1.mutations
    assert s1.large_base + s1.bold_base == s1.mutations
    assert s1.accept_base == s1.mutations + s1.second_base
    assert s1.accept == s1.first_accept + s1.second_accept
    # every mutation deposits exactly unit luminance (expectation weights sum to one, splats normalised)
    lum_film = (f1 * np.array([0.212671, 0.715160, 0.072169])).sum()
    assert lum_film == pytest.approx(s1.mutations, rel=0.12)   # gaussian filter mass 1 up to tabulation, border loss


# ------------------------------------------------------------------ two-stage MLT, crop windows
def _np_resample_axis(src, tgt_res):
    """Independent numpy restatement of Resampler (rfilter.h:123-177, 232-280) along axis 0, gaussian filter, EClamp."""
    src_res = src.shape[0]
    radius, scale = 2.0, 1.0
    if tgt_res < src_res:
        scale = src_res / tgt_res
        radius *= scale
    taps = int(math.ceil(radius * 2))
    out = np.zeros((tgt_res,) + src.shape[1:])
    for i in range(tgt_res):
        center = (i + 0.5) / tgt_res * src_res
        start = int(math.floor(center - radius + 0.5))
        pos = (start + np.arange(taps) + 0.5 - center) / scale
        w = np.maximum(0.0, np.exp(-2.0 * pos * pos) - math.exp(-2.0 * 4.0))
        w /= w.sum()
        idx = np.clip(start + np.arange(taps), 0, src_res - 1)
        out[i] = np.maximum(0.0, np.tensordot(w, src[idx], axes=(0, 0)))
    return out


@pytest.mark.parametrize("shape", [((20, 12), (160, 90)), ((64, 48), (16, 12)), ((9, 5), (9, 11)), ((3, 3), (3, 3))])
def test_importance_map_resampler(oracle, shape):
    (w, h), (W, H) = shape
    rng = np.random.RandomState(7)
    img = rng.rand(h, w, 3).astype(np.float32)
    m = oracle_lib.resample_luminance(img, (W, H))
    lum = (img.astype(np.float64) * [0.212671, 0.715160, 0.072169]).sum(-1)
    ref = lum
    if w != W:
        ref = _np_resample_axis(ref.T, W).T
    if h != H:
        ref = _np_resample_axis(ref, H)
    assert m.shape == (H, W)
    assert np.allclose(m, ref, rtol=1e-6, atol=1e-7)
    flat = oracle_lib.resample_luminance(np.full((h, w, 3), 2.0, np.float32), (W, H))
    assert np.allclose(flat, 2.0, rtol=1e-6)                  # taps are normalised (rfilter.h:171-176)


def test_importance_map_reweights_the_chain(cornell):
    """SplatList::normalize(importanceMap) (pathsampler.cpp:1001-1020): L -> L / map[pixel]; a constant map changes no
    decision; develop multiplies the map back (drmlt_proc.cpp:823-849)."""
    cfg = _cfg(technique=abi.DR_TECH_MMLT, type=abi.DR_TYPE_ORBITAL, seed=21)
    lum, dep = cornell.bootstrap(cfg, 0, 3000)
    seeds = np.nonzero(lum > 0)[0][:32]
    ids = np.arange(32)
    r0, f0, _ = cornell.chain_steps(cfg, 1.0, seeds, dep[seeds], ids, 32, want_film=True, threads=2)
    imp = np.full((64, 64), 0.5, np.float32)
    cfg2 = _cfg(technique=abi.DR_TECH_MMLT, type=abi.DR_TYPE_ORBITAL, seed=21, two_stage=1)
    cfg2.importance_map = oracle_lib.fptr(imp)
    r1, f1, _ = cornell.chain_steps(cfg2, 1.0, seeds, dep[seeds], ids, 32, want_film=True, threads=2)
    for a, b in zip(r0, r1):
        assert b.L_x == pytest.approx(2.0 * a.L_x, rel=1e-6) and b.L_y == pytest.approx(2.0 * a.L_y, rel=1e-6)
        assert (a.accept1, a.did_second, a.accept2) == (b.accept1, b.did_second, b.accept2)
    assert np.allclose(f0, f1, rtol=1e-5, atol=1e-7)          # splats are normalised to unit luminance either way
    # develop: the mean luminance of the developed image is b, with or without a map
    rng = np.random.RandomState(1)
    imp2 = (0.1 + rng.rand(64, 64)).astype(np.float32)
    for m in (None, imp2):
        img = oracle_lib.develop(f0, 0.37, False, m)
        l = (img.astype(np.float64) * [0.212671, 0.715160, 0.072169]).sum(-1)
        assert l.mean() == pytest.approx(0.37, rel=1e-5)
    # a non-constant map does change the chain
    cfg2.importance_map = oracle_lib.fptr(imp2)
    r2, _, _ = cornell.chain_steps(cfg2, 1.0, seeds, dep[seeds], ids, 32, threads=2)
    assert bytes(r2) != bytes(r1)


def test_two_stage_render_and_crop_window(cornell):
    """orc_render with twoStage: nested pass at size / firstStageSizeReduction with sampleCount x reduction (util.cpp:100-136);
    crop windows shift the sensor's sample space (perspective.cpp:132-157)."""
    cfg = _cfg(technique=abi.DR_TECH_PATH, type=abi.DR_TYPE_MIRA, seed=4, max_depth=5, two_stage=1, first_stage_size_reduction=4,
               rfilter=abi.DR_FILTER_BOX)
    nested = cornell.first_stage_config(cfg)
    assert (nested.film_width, nested.film_height, nested.crop_width, nested.crop_height) == (16, 16, 16, 16)
    assert nested.sample_count == cfg.sample_count * 4 and nested.first_stage == 1
    rc, img2, st2, _ = cornell.render(cfg, 20000, 256, 96, threads=8)
    assert rc == 0 and img2.shape == (64, 64, 3) and np.isfinite(img2).all()
    cfg1 = _cfg(technique=abi.DR_TECH_PATH, type=abi.DR_TYPE_MIRA, seed=4, max_depth=5, rfilter=abi.DR_FILTER_BOX)
    rc, img1, st1, _ = cornell.render(cfg1, 20000, 256, 96, threads=8)
    lum = lambda im: (im.astype(np.float64) * [0.212671, 0.715160, 0.072169]).sum(-1)
    assert lum(img2).mean() == pytest.approx(st2.luminance, rel=1e-4)
    assert st2.luminance == pytest.approx(st1.luminance, rel=1e-9)      # same bootstrap: b ignores the map
    a, b = lum(img1).reshape(8, 8, 8, 8).mean(axis=(1, 3)), lum(img2).reshape(8, 8, 8, 8).mean(axis=(1, 3))
    assert np.abs(a - b).sum() / a.sum() < 0.35
    # crop window: replayed paths land at (full-film pixel - crop offset); outside the window they contribute nothing
    n = 4000
    rng = np.random.RandomState(3)
    us = rng.rand(n, 40).astype(np.float32)
    z = np.zeros((n, 2), np.float32)
    full = _cfg(technique=abi.DR_TECH_PATH, seed=1, max_depth=5)
    crop = _cfg(technique=abi.DR_TECH_PATH, seed=1, max_depth=5, crop_offset_x=16, crop_offset_y=8, crop_width=32, crop_height=40)
    us_crop = us.copy()                                        # the same film points, expressed in crop sample space
    us_crop[:, 0] = (us[:, 0] * 64 - 16) / 32
    us_crop[:, 1] = (us[:, 1] * 64 - 8) / 40
    inside = (us_crop[:, 0] >= 0) & (us_crop[:, 0] < 1) & (us_crop[:, 1] >= 0) & (us_crop[:, 1] < 1)
    of, lf = cornell.eval_paths(full, us[inside], z[inside], z[inside], np.zeros(inside.sum(), np.int32))
    oc, lc = cornell.eval_paths(crop, us_crop[inside], z[inside], z[inside], np.zeros(inside.sum(), np.int32))
    nz = lf > 0
    assert nz.sum() > 100 and np.allclose(lc[nz], lf[nz], rtol=2e-4)
    for i in np.nonzero(nz)[0][:200]:
        assert oc[i].pos[0][0] == pytest.approx(of[i].pos[0][0] - 16, abs=2e-3)
        assert oc[i].pos[0][1] == pytest.approx(of[i].pos[0][1] - 8, abs=2e-3)


# ------------------------------------------------------------------ rough dielectric
def _sample3(oracle, m, wi, mode, u1, u2, u3):
    wo, w, pdf, ty, eta = (C.c_double * 3)(), (C.c_double * 3)(), C.c_double(), C.c_int(), C.c_double()
    oracle.orc_bsdf_sample3(C.byref(m), (C.c_double * 3)(*wi), mode, u1, u2, u3, wo, w, C.byref(pdf), C.byref(ty), C.byref(eta))
    return np.array(wo), np.array(w), pdf.value, ty.value, eta.value


@pytest.mark.parametrize("name,flags,alpha", [("ggx-vis", abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE, 0.2), ("ggx-all", abi.DR_MAT_GGX, 0.2),
                                              ("beckmann-vis", abi.DR_MAT_SAMPLE_VISIBLE, 0.3), ("beckmann-all", 0, 0.3)])
@pytest.mark.parametrize("side", [1, -1])
def test_rough_dielectric_sample_eval_pdf_consistency(oracle, name, flags, alpha, side):
    """roughdielectric.cpp:270-611: sample() = eval() / pdf() for the sampled direction, in both transport modes, from
    either side; reflection keeps the hemisphere (eta = 1), refraction changes it (eta = intIOR/extIOR or its inverse); the
    pdf returned by sample() passed through a float (sic, :543)."""
    mat = _mat(abi.DR_BSDF_ROUGHDIELECTRIC, flags, alpha=alpha)
    mat.reflectance[:] = (1, 1, 1)
    rng = np.random.RandomState(11)
    wi = np.array([0.3, -0.2, 0.0]); wi[2] = side * math.sqrt(1 - wi[0] ** 2 - wi[1] ** 2)
    n_r = n_t = 0
    for mode in (0, 1):
        for _ in range(300):
            u1, u2, u3 = rng.rand(3)
            wo, w, pdf, ty, eta = _sample3(oracle, mat, wi, mode, u1, u2, u3)
            if not w.any():
                continue
            f, p = _eval(oracle, mat, wi, wo, mode, 1)
            assert p == pytest.approx(pdf, rel=2e-6), name          # float rounding of temporaryPdf
            tol = 2e-2 if name == "beckmann-vis" else 1e-5
            assert np.allclose(f / p, w, rtol=tol, atol=1e-9), (name, mode)
            if wo[2] * wi[2] > 0:
                n_r += 1
                assert ty == 0x2 and eta == 1.0
            else:
                n_t += 1
                assert ty == 0x10 and eta == pytest.approx(1.5 if wi[2] > 0 else 1 / 1.5)
    assert n_r > 10 and n_t > 200


def test_rough_dielectric_reciprocity_and_smooth_limit(oracle):
    """f(wi, wo) in radiance mode and f(wo, wi) in importance mode differ by the eta^2 radiance scaling only
    (roughdielectric.cpp:339-345); for alpha -> 0 the lobe collapses onto the smooth dielectric's directions."""
    mat = _mat(abi.DR_BSDF_ROUGHDIELECTRIC, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE, alpha=0.25)
    rng = np.random.RandomState(2)
    wi = np.array([0.1, 0.4, 0.0]); wi[2] = math.sqrt(1 - 0.17)
    for _ in range(50):
        wo, w, pdf, ty, eta = _sample3(oracle, mat, wi, 1, *rng.rand(3))
        if not w.any():
            continue
        f_imp, _ = _eval(oracle, mat, wi, wo, 1, 1)           # importance transport: no eta^2 factor
        f_rad, _ = _eval(oracle, mat, wi, wo, 0, 1)
        if ty == 0x10:
            assert np.allclose(f_rad, f_imp * (1 / 1.5) ** 2, rtol=1e-9)
        else:
            assert np.allclose(f_rad, f_imp, rtol=1e-12)
    smooth = _mat(abi.DR_BSDF_DIELECTRIC)
    sharp = _mat(abi.DR_BSDF_ROUGHDIELECTRIC, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE, alpha=1e-4)
    for u3 in (0.001, 0.9):
        wo_r, w_r, _, _, _ = _sample3(oracle, sharp, wi, 0, 0.37, 0.61, u3)
        wo_s, w_s, _, _ = _sample(oracle, smooth, wi, 0, u3, 0.5)
        assert np.allclose(wo_r, wo_s, atol=2e-3)


def test_rough_dielectric_normalisation_agrees_across_techniques(oracle):
    """With a rough dielectric in the scene (one extra primary sample per BSDF sample, non-symmetric BSDF, eta^2 radiance
    scaling) the unidirectional, bidirectional and MMLT estimators of b still integrate the same path space."""
    data = scenes.glossy_scene(film=(64, 64), subdiv=2, rough_glass=(0.2, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE))
    orc = oracle_lib.OracleScene(data)
    b, err = {}, {}
    for name, tech in (("path", abi.DR_TECH_PATH), ("bdpt", abi.DR_TECH_BDPT), ("mmlt", abi.DR_TECH_MMLT)):
        cfg = _cfg(technique=tech, seed=9, direct_samples=0, max_depth=6)
        k = 6 if tech == abi.DR_TECH_MMLT else 1
        lum, _ = orc.bootstrap(cfg, 0, 150000 * k)
        b[name], err[name] = lum.mean() * k, lum.std() * k / math.sqrt(len(lum))
    for a_, c_ in (("path", "bdpt"), ("mmlt", "bdpt"), ("path", "mmlt")):
        assert abs(b[a_] - b[c_]) < 4 * math.hypot(err[a_], err[c_]), (b, err)
    assert err["bdpt"] < 0.02 * b["bdpt"]


# ------------------------------------------------------------------ smooth plastic
def _plastic(flags=0):
    m = _mat(abi.DR_BSDF_PLASTIC, flags)
    m.reflectance[:] = (0.5, 0.3, 0.2)        # diffuseReflectance
    m.transmittance[:] = (1.0, 0.9, 0.8)      # specularReflectance
    m.eta[:] = (1.49, 0, 0)
    return m


@pytest.mark.parametrize("flags", [0, abi.DR_MAT_NONLINEAR, abi.DR_MAT_TWOSIDED])
def test_plastic_components_and_consistency(oracle, flags):
    """plastic.cpp:240-412: a delta specular coating (discrete measure) over a diffuse base (solid angle); sample() = eval() /
    pdf() in the measure of the sampled component; the specular lobe is picked with the Fresnel-weighted probability."""
    mat = _plastic(flags)
    rng = np.random.RandomState(4)
    wi = np.array([0.5, -0.3, 0.0]); wi[2] = math.sqrt(1 - 0.34)
    n_s = n_d = 0
    for _ in range(600):
        u1, u2 = rng.rand(2)
        wo, w, pdf, ty = _sample(oracle, mat, wi, 0, u1, u2)
        assert w.any() and pdf > 0
        measure = 4 if ty == 0x4 else 1
        f, p = _eval(oracle, mat, wi, wo, 0, measure)
        assert p == pytest.approx(pdf, rel=1e-9)
        assert np.allclose(f / p, w, rtol=1e-9)
        other, _ = _eval(oracle, mat, wi, wo, 0, 1 if measure == 4 else 4)
        if ty == 0x4:
            n_s += 1
            assert np.allclose(wo, [-wi[0], -wi[1], wi[2]])
        else:
            n_d += 1
            assert ty == 0x1 and wo[2] > 0
            if np.dot(wo, [-wi[0], -wi[1], wi[2]]) < 0.99:             # (the delta test tolerates DeltaEpsilon = 1e-3, plastic.cpp:261)
                assert not other.any()                               # a generic direction has no specular component
    # specular sampling probability: Fi w / (Fi w + (1 - Fi)(1 - w)) (plastic.cpp:291-295)
    _, p_spec = _eval(oracle, mat, wi, [-wi[0], -wi[1], wi[2]], 0, 4)
    assert n_s / 600 == pytest.approx(p_spec, abs=0.06)
    assert n_s > 10 and n_d > 300
    below = np.array([0.5, -0.3, -wi[2]])
    if not flags & abi.DR_MAT_TWOSIDED:
        assert not _sample(oracle, mat, below, 0, 0.3, 0.3)[1].any()  # one-sided
    else:
        assert _sample(oracle, mat, below, 0, 0.3, 0.3)[1].any()


def test_plastic_energy_and_fresnel_diffuse_reflectance(oracle):
    """The diffuse base is normalised by 1 / (1 - fdrInt) with fdrInt = fresnelDiffuseReflectance(1 / eta)
    (util.cpp:822-867, approx. 0.6 for eta 1.5): a white base under a white coating reflects (almost) all light."""
    m = _plastic()
    m.reflectance[:] = (1, 1, 1); m.transmittance[:] = (1, 1, 1)
    rng = np.random.RandomState(6)
    wi = np.array([0.0, 0.0, 1.0])
    tot = np.zeros(3)
    n = 4000
    for _ in range(n):
        _, w, _, _ = _sample(oracle, m, wi, 0, *rng.rand(2))
        tot += w
    assert np.all(np.abs(tot / n - 1.0) < 0.03)
    # fdrInt through the diffuse value at normal incidence: f = diff / (1 - fdr) * cos/pi / eta^2 * (1 - F)^2
    F0 = ((1.49 - 1) / (1.49 + 1)) ** 2
    f, _ = _eval(oracle, m, wi, [0, 0, 1.0], 0, 1)
    fdr = 1 - (1 / math.pi) / 1.49 ** 2 * (1 - F0) ** 2 / f[0]
    assert fdr == pytest.approx(0.5925, abs=3e-3)
