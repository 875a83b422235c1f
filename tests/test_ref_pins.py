"""Pins of the oracle against the reference's OWN compiled code (CPU only).

oracle/ref/Makefile compiles the numerical leaves of the hot path from the sources where they lie under /root/reference
(libcore warp / util / triangle / quad / math, bsdfs/microfacet.h, core/pmf.h, render/triaccel.h,
drmlt/tools/transition.h) into oracle/_ref/libref_leaf.so; tools/make_ref_golden.py stored that library's outputs on
seeded inputs in tests/golden/ref_leaf.npz.  Here:
  * where oracle/_ref is present (the container that has /root/reference), oracle and reference run side by side on
    the same host and must agree BIT FOR BIT -- except 2 ulp on the GGX terms and the quadrature tolerance of
    fresnelDiffuseReflectance (the reference integrates to 1e-5, util.cpp:864; the oracle to ~1e-12);
  * everywhere, the oracle must reproduce the committed fixture: decisions (sampled index, hit / miss) exactly, values
    within 1e-9 (a host whose libm picks other exp / log / cos variants may round the last bits differently);
  * where oracle/_ref is present, the library reproduces the fixture to the same bound, i.e. the fixture is what the
    reference computes.
"""
import os

import numpy as np
import pytest

import ref_leaf_cases as R

# name -> relative tolerance (0 = bit-exact)
TOL = {"fresnelDiffuseReflectance": 1e-5, "microfacet_ggx_all": 4e-15, "microfacet_ggx_visible": 4e-15}


@pytest.fixture(scope="module", autouse=True)
def _oracle_built():
    import oracle_lib as _ol
    _ol.load()          # builds oracle/_build/liboracle.so if it is missing (fresh checkout)


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(R.GOLDEN))


def _compare(got, want, name, floor=0.0):
    tol = max(TOL.get(name, 0.0), floor)
    if name in ("pmf_index",) or name == "triAccel":
        flags = (got, want) if name == "pmf_index" else (got[:, 0], want[:, 0])
        assert np.array_equal(*flags), name + ": decisions differ"
    assert got.shape == want.shape
    if tol == 0.0:
        bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
        assert not bad.any(), "%s: %d of %d values differ, worst %g" % (name, bad.sum(), got.size, np.abs(got - want)[bad].max())
    else:
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)      # relative above 1e-3, absolute below
        rel[got == want] = 0
        assert rel.max() <= tol, "%s: worst relative difference %g > %g" % (name, rel.max(), tol)


needs_ref = pytest.mark.skipif(not os.path.exists(R.REF_LEAF), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
def test_oracle_equals_reference_leaves_bit_for_bit():
    ref = R.run_cases(R.load(R.REF_LEAF), "ref_")
    got = R.run_cases(R.load(R.ORACLE), "orc_")
    assert set(got) == set(ref)
    for name in sorted(ref):
        _compare(got[name], ref[name], name)


def test_oracle_reproduces_reference_fixture(golden):
    got = R.run_cases(R.load(R.ORACLE), "orc_")
    assert set(got) == set(golden)
    for name in sorted(golden):
        _compare(got[name], golden[name], name, floor=1e-9)


@needs_ref
def test_fixture_is_what_the_reference_computes(golden):
    got = R.run_cases(R.load(R.REF_LEAF), "ref_")
    for name in sorted(golden):
        _compare(got[name], golden[name], name, floor=1e-9)


def test_fixture_covers_edges(golden):
    # zero-probability entries are skipped (pmf.h:131-133), the reused sample stays in [0, 1)
    assert golden["pmf_pmf"][0] == 0 and (golden["pmf_pmf"][golden["pmf_index"].astype(int)] > 0).all()
    assert (golden["pmf_reused"] >= 0).all() and (golden["pmf_reused"] <= 1).all()
    # total internal reflection and normal incidence are among the Fresnel cases
    assert (golden["fresnelDielectricExt"][:, 0] == 1.0).any() and (golden["fresnelDielectricExt"][:, 0] < 0.05).any()
    # the triangle cases contain hits and misses
    assert 0.5 < golden["triAccel"][:, 0].mean() < 0.95
    # the hemisphere guard z = 1e-10f at the rim of the disk (warp.cpp:47-50)
    assert (golden["squareToCosineHemisphere"][:, 2] == np.float32(1e-10)).any()


# ================================================================ BSDF plugins and PathSampler::sampleSplats
# oracle/_ref/libref_path.so is the reference's OWN libcore + librender + libbidir + BSDF / emitter / sensor plugins
# (oracle/ref/Makefile, ref_path.cpp); tests/golden/ref_path.npz holds its outputs on the seeded cases of
# tests/ref_path_cases.py.  The bar is the north_star's: f(u) within 1e-4 relative for >= 99.9 % of the paths --
# measured: Cornell box 1e-12, all other cases >= 99.95 % within 1e-4 with medians below 1e-8.
import ctypes as C  # noqa: E402

import oracle_lib  # noqa: E402
import ref_path_cases as RP  # noqa: E402

needs_ref_path = pytest.mark.skipif(not os.path.exists(RP.REF_PATH), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def golden_path():
    return dict(np.load(RP.GOLDEN))


def _bsdf_tolerance(name):
    # delta, diffuse and Beckmann lobes: bit for bit.  GGX: 2 ulp.
    # Plastic: the reference integrates its internal diffuse reflectance to 1e-5 (util.cpp:864), and the material record
    # (dr_material.k) keeps the two derived constants as float: 1e-7 absolute in the re-scaled sample -> 1e-4 of 1e-3.
    if "roughplastic" in name:      # tables, Fdr and the sampling weight travel in double: only pow() and summation order differ
        return 1e-13
    if "plastic" in name:
        return 1e-4
    if "rough" in name:
        return 4e-15
    return 0.0


def _compare_bsdf(got, want, name):
    tol = _bsdf_tolerance(name)
    assert np.array_equal(got[..., 7], want[..., 7]), name + ": sampled lobe types differ"
    if tol == 0.0:
        assert np.array_equal(got, want), name
    else:
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
        assert err.max() <= tol, "%s: worst difference %g" % (name, err.max())


@needs_ref_path
def test_oracle_bsdfs_equal_reference_plugins():
    ref = RP.run_bsdf(C.CDLL(RP.REF_PATH), "ref_")
    got = RP.run_bsdf(C.CDLL(RP.ORACLE), "orc_")
    for name in sorted(ref):
        _compare_bsdf(got[name], ref[name], name)


def test_oracle_bsdfs_reproduce_reference_fixture(golden_path):
    got = RP.run_bsdf(C.CDLL(RP.ORACLE), "orc_")
    names = [k for k in golden_path if k.startswith("bsdf_")]
    assert len(names) == len(RP.bsdf_materials())
    for name in names:
        g, w = got[name], golden_path[name]
        assert np.array_equal(g[..., 7], w[..., 7]), name
        err = np.abs(g - w) / np.maximum(np.abs(w), 1e-3)
        assert err.max() <= max(_bsdf_tolerance(name), 1e-9), "%s: worst difference %g" % (name, err.max())


compare_paths = RP.compare_paths


_oracle_scenes = {}


def _oracle_paths(case):
    if case[0] not in _oracle_scenes:
        _oracle_scenes[case[0]] = oracle_lib.OracleScene(RP.SCENES[case[0]]())
    cfg = RP.case_config(case)
    cfg.ray_epsilon = cfg.shadow_epsilon = 0        # the oracle's defaults = the reference's double-build constants
    us, ue, ud, depth = RP.case_inputs(case)
    out, lum = _oracle_scenes[case[0]].eval_paths(cfg, us, ue, ud, depth)
    r = RP.unpack(out, len(lum))
    return lum, np.stack([r["s"], r["t"], r["n_splats"]], 1), r["pos0"], r["value0"]


@pytest.fixture(scope="module")
def golden_texture():
    return dict(np.load(RP.GOLDEN_TEXTURE))


# ---- bitmap textures: Texture2D::eval -> TMIPMap::evalBilinear / evalBox / evalTexel (texture.cpp:112-121, mipmap.h:503-596)
@needs_ref_path
def test_oracle_textures_equal_reference_mipmap_bit_for_bit():
    ref = RP.run_texture(C.CDLL(RP.REF_PATH), "ref_")
    got = RP.run_texture(C.CDLL(RP.ORACLE), "orc_")
    assert len(got) == len(RP.texture_leaf_cases())
    for k, v in got.items():
        assert np.array_equal(v, ref[k]), k


def test_oracle_textures_reproduce_reference_fixture(golden_texture):
    got = RP.run_texture(C.CDLL(RP.ORACLE), "orc_")
    for k, v in got.items():
        assert np.array_equal(v, golden_texture[k]), k
    # the fixture sees every boundary condition: lookups outside [0, 1)^2 differ between the wrap modes of one filter
    uv = RP.texture_leaf_uv()
    outside = ((uv < 0) | (uv >= 1)).any(axis=1)
    assert outside.sum() > 1000
    assert (golden_texture["tex_bilinear_zero"][outside] == 0).all(axis=1).sum() > 500 and (golden_texture["tex_bilinear_one"][outside] == 1).all(axis=1).sum() > 500
    # Texture::getAverage of the reference = the mean of the texels the host passes (what SceneData.add_texture stores in the material)
    for name, (t, arr) in RP.texture_leaf_cases().items():
        assert np.allclose(golden_texture["texavg_" + name], arr.reshape(-1, 3).astype(np.float64).mean(axis=0), rtol=1e-12), name


@needs_ref_path
def test_texture_fixture_is_what_the_reference_computes(golden_texture):
    k = RP.case_key(RP.TEXTURE_CASES[0])
    r = RP.run_paths_ref(C.CDLL(RP.REF_PATH), RP.TEXTURE_CASES[0])
    assert np.allclose(r["lum"], golden_texture[k + "_lum"], rtol=1e-9, atol=0)
    assert np.array_equal(np.stack([r["s"], r["t"], r["n_splats"]], 1), golden_texture[k + "_st"])


@pytest.mark.parametrize("case", RP.TEXTURE_CASES, ids=RP.case_key)
def test_oracle_textured_paths_reproduce_reference_fixture(case, golden_texture):
    """The BSDF plugins with <texture> children, TriMesh texture coordinates and UV tangents, through PathSampler::sampleSplats."""
    k = RP.case_key(case)
    lum, st, pos0, value0 = _oracle_paths(case)
    rel = compare_paths(lum, st, pos0, value0, golden_texture[k + "_lum"], golden_texture[k + "_st"].astype(np.int32),
                        golden_texture[k + "_pos0"], golden_texture[k + "_value0"], k, case[1] == "mmlt")
    assert np.median(rel) < 1e-12, k
    # the textures matter: the same vectors on the scene with its textures replaced by their averages give other contributions
    assert len(np.unique(np.round(value0[lum > 0], 3), axis=0)) > 200


@pytest.mark.parametrize("case", RP.PATH_CASES, ids=RP.case_key)
def test_oracle_paths_reproduce_reference_fixture(case, golden_path):
    k = RP.case_key(case)
    lum, st, pos0, value0 = _oracle_paths(case)
    rel = compare_paths(lum, st, pos0, value0, golden_path[k + "_lum"], golden_path[k + "_st"].astype(np.int32),
                        golden_path[k + "_pos0"], golden_path[k + "_value0"], k, case[1] == "mmlt")
    assert np.median(rel) < 1e-7, k
    if case[0] == "cornell":                        # diffuse walls only: nothing but rounding separates the two
        assert rel.max() < 1e-9, k


@needs_ref_path
@pytest.mark.parametrize("case", [RP.PATH_CASES[0], RP.PATH_CASES[4], RP.PATH_CASES[9]], ids=RP.case_key)
def test_fixture_is_what_the_reference_path_sampler_computes(case, golden_path):
    k = RP.case_key(case)
    r = RP.run_paths_ref(C.CDLL(RP.REF_PATH), case)
    assert np.allclose(r["lum"], golden_path[k + "_lum"], rtol=1e-9, atol=0)
    assert np.array_equal(np.stack([r["s"], r["t"], r["n_splats"]], 1), golden_path[k + "_st"])


# ================================================================ whole jobs: the reference's DRMLT / PSSMLT integrators
# tests/golden/ref_render.npz: three end-to-end runs per configuration of the reference's own integrators
# (DRMLT::render / PSSMLT::render through a RenderJob, oracle/ref/ref_path.cpp:ref_render) -- statistics counters, b,
# developed image -- plus a 16x longer render as converged image.
@pytest.fixture(scope="module")
def golden_render():
    return dict(np.load(RP.GOLDEN_RENDER))


@pytest.mark.parametrize("name", ["drmlt_mira_path", "pssmlt_mmlt"])
def test_oracle_job_statistics_match_reference_integrators(name, golden_render):
    params, spp = RP.RENDER_CASES[name]
    spp //= 4                                   # a quarter of the reference's mutations keeps the CPU suite short
    orc = oracle_lib.OracleScene(RP.RENDER_SCENE())
    cfg = RP.make_config(sampleCount=spp, seed=11, **params)
    cfg.ray_epsilon = cfg.shadow_epsilon = 0
    r, img, st, sec = orc.render(cfg, 300000, 4096, 64 * 64 * spp // 4096, 0)
    assert r == 0
    RP.check_rates_and_b(name, st, st.luminance, golden_render)


def test_reference_render_fixture_is_consistent(golden_render):
    # develop scales the film so that its mean luminance is b (drmlt_proc.cpp:823-849): every run's image carries its b
    for name in RP.RENDER_CASES:
        assert golden_render[name + "_stats"].shape[0] == 3
        assert RP.luminance(golden_render[name + "_image"]).mean() == pytest.approx(golden_render[name + "_b"][0], rel=1e-6)
    # path tracing, BDPT and MMLT estimate the same image: their b agree (the BDPT case has maxDepth 5 and a light image)
    assert golden_render["drmlt_mira_path_b"].mean() == pytest.approx(golden_render["pssmlt_mmlt_b"].mean(), rel=1e-2)
    runs = golden_render["drmlt_orbital_mmlt_relmse_runs"]
    assert (runs > 0).all() and runs.max() < 0.05


# ================================================================ the delayed-rejection samplers themselves
# drmlt_sampler.cpp (Green / Mira / Orbital) compiled into oracle/_ref (oracle/ref/ref_sampler.cpp).  The reference's Random is
# seeded explicitly and a twin generator records the uniforms the sampler consumes; the oracle's sampler, fed the same current
# state and the same stream in call order, must produce the same stage-1 and stage-2 proposals (Kelemen, Gaussian, the 2-D
# radial Kelemen and the pairwise orbital rotation with its wrapped-Cauchy angle), the same Green reverse state
# y* = z - (y - x) and the same Mira transition ratio -- bit for bit.
def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@needs_ref_path
def test_oracle_drmlt_samplers_equal_reference_samplers_bit_for_bit():
    ref = RP.run_sampler_ref(C.CDLL(RP.REF_PATH))
    got = RP.run_sampler_oracle(C.CDLL(RP.ORACLE), ref)
    for key in sorted(ref):
        assert _same(got[key], ref[key]), key


def test_oracle_drmlt_samplers_reproduce_reference_fixture():
    gold = dict(np.load(RP.GOLDEN_SAMPLER))
    assert len(gold) == 6
    got = RP.run_sampler_oracle(C.CDLL(RP.ORACLE), gold)
    md = RP.SAMPLER_DIM
    for key in sorted(gold):
        g, w = got[key], gold[key]
        assert g.shape == w.shape and np.isnan(g).tolist() == np.isnan(w).tolist(), key
        assert np.allclose(g, w, rtol=1e-9, atol=1e-12, equal_nan=True), key       # other libm variants may round the last bits
        # large steps redraw every coordinate uniformly; small steps stay within the kernels' reach of the current state
        p1 = w[:, 7 * md:8 * md]
        if key.endswith("_0"):
            assert np.abs(p1 - w[:, :md]).max() < 2 * 1.9 / 64 + 1e-12, key
        else:
            assert 0.2 < np.abs(p1 - w[:, :md]).mean() < 0.45, key


# ================================================================ the PSSMLT sampler over a sequence of mutations
# pssmlt_sampler.{h,cpp} compiled into oracle/_ref (oracle/ref/ref_pssmlt_sampler.cpp): seed replay, then 12 mutations with a
# fixed large-step / accept pattern.  The oracle's PSSMLTSampler, fed the same current state and the recorded stream in call
# order, must propose the same states in every mutation -- eager fill of all maxDim coordinates, Kelemen (one uniform) and
# Gaussian (two uniforms, Box-Muller) mutation with their wraps, uniform large steps, restore on reject -- bit for bit.
@needs_ref_path
def test_oracle_pssmlt_sampler_equals_reference_sampler_bit_for_bit():
    ref = RP.run_pss_sampler_ref(C.CDLL(RP.REF_PATH))
    got, _ = RP.run_pss_sampler_oracle(C.CDLL(RP.ORACLE), ref)
    for key in sorted(ref):
        assert _same(got[key], ref[key]), key


def test_oracle_pssmlt_sampler_reproduces_reference_fixture():
    gold = dict(np.load(RP.GOLDEN_PSS_SAMPLER))
    assert sorted(gold) == ["pss_0", "pss_1"]
    got, used = RP.run_pss_sampler_oracle(C.CDLL(RP.ORACLE), gold)
    md, nm = RP.PSS_DIM, RP.PSS_MUT
    for key in sorted(gold):
        g, w = got[key], gold[key]
        assert g.shape == w.shape
        assert np.allclose(g, w, rtol=1e-9, atol=1e-12), key       # other libm variants may round the last bits
        props = w[:, md + 2 * md * nm:].reshape(len(w), nm, md)
        assert props.min() >= 0.0 and props.max() <= 1.0, key
        for seed in range(len(w)):
            large, _ = RP.pss_pattern(seed)
            draws = 1 if key == "pss_1" else 2                     # uniforms per coordinate of a small step
            assert used[key][seed] == md * int(large.sum() + draws * (nm - large.sum())), (key, seed)


# ================================================================ the DRMLT samplers over a sequence of mutations
# ref_drmlt_sampler_seq (oracle/ref/ref_sampler.cpp) drives Green / Mira / Orbital the way DRMLTRenderer::process does
# (drmlt_proc.cpp:541-760) for 10 consecutive mutations with a fixed pattern of large steps and outcomes (first stage accepted,
# second stage accepted, second stage rejected), plain, after handleLightTracing() (the emitter sampler under fixEmitterPath,
# with light-tracing second stages) and after setStagesToIdentity() (the MMLT direct sampler).  Every proposal of every
# mutation, Green's reverse states and Mira's ratios must agree bit for bit: the state carried from mutation to mutation
# (accept's wrap, reject's reset) is then the reference's too.
@needs_ref_path
def test_oracle_drmlt_sampler_sequences_equal_reference_bit_for_bit():
    ref = RP.run_sampler_seq_ref(C.CDLL(RP.REF_PATH))
    got = RP.run_sampler_seq_oracle(C.CDLL(RP.ORACLE), ref)
    for key in sorted(ref):
        assert _same(got[key], ref[key]), key


def test_oracle_drmlt_sampler_sequences_reproduce_reference_fixture():
    gold = dict(np.load(RP.GOLDEN_SAMPLER_SEQ))
    assert len(gold) == 9
    got = RP.run_sampler_seq_oracle(C.CDLL(RP.ORACLE), gold)
    md, nm = RP.SEQ_DIM, RP.SEQ_MUT
    off = md + 6 * md * nm
    for key in sorted(gold):
        g, w = got[key], gold[key]
        assert g.shape == w.shape and np.isnan(g).tolist() == np.isnan(w).tolist(), key
        assert np.allclose(g, w, rtol=1e-9, atol=1e-12, equal_nan=True), key       # other libm variants may round the last bits
        type_, mode = int(key.split("_")[1]), int(key.split("_")[2])
        for seed in range(len(w)):
            _, outc, _ = RP.seq_pattern(seed)
            p1 = w[seed, off:off + md * nm].reshape(nm, md)
            p2 = w[seed, off + md * nm:off + 2 * md * nm].reshape(nm, md)
            rv = w[seed, off + 2 * md * nm:off + 3 * md * nm].reshape(nm, md)
            assert not np.isnan(p1).any()
            assert np.isnan(p2).all(axis=1).tolist() == (outc == 0).tolist(), (key, seed)      # a second stage exactly where one was taken
            assert np.isnan(rv).all(axis=1).tolist() == ((outc == 0) | (type_ != 0)).tolist(), (key, seed)
            large = RP.seq_pattern(seed)[0]
            if mode == 2 and not large[0]:       # identity stages: a small step proposes the current state itself
                assert np.array_equal(p1[0], w[seed, :md]), (key, seed)


# ================================================================ the film: ImageBlock::put with the reference's filter plugins
# oracle/ref/ref_film.cpp: a block created as DRMLTProcess::createWorkResult creates it, 4000 splats (inside, on pixel and
# half-pixel lattices, up to 3 pixels outside the film; NaN / inf / negative values) -> the oracle's Film::put accumulates
# the same doubles and rejects the same splats.  (Pinning found: the box radius is Float 0.5 + the FLOAT literal 1e-5f,
# box.cpp:38, and the filter table is normalised by multiplication, rfilter.cpp:52-54.)
@needs_ref_path
def test_oracle_film_equals_reference_imageblock_bit_for_bit():
    lib_ref = C.CDLL(RP.REF_PATH)
    ref, got = RP.run_film(lib_ref.ref_splat, True), RP.run_film(C.CDLL(RP.ORACLE).orc_splat_f64, False)
    for key in sorted(got):
        assert np.array_equal(got[key], ref[key]), key


def test_oracle_film_reproduces_reference_fixture():
    gold = dict(np.load(RP.GOLDEN_FILM))
    got = RP.run_film(C.CDLL(RP.ORACLE).orc_splat_f64, False)
    pos, rgb = RP.film_inputs()
    for name, _, _ in RP.FILM_FILTERS:
        assert np.array_equal(got["ok_" + name], gold["ok_" + name])
        assert gold["ok_" + name][5:8].tolist() == [0, 0, 0] and gold["ok_" + name].sum() == RP.FILM_N - 3
        assert np.allclose(got["film_" + name], gold["film_" + name], rtol=1e-12, atol=1e-14), name
    # a normalised filter keeps the energy of splats whose footprint lies inside the film
    inside = (pos[:, 0] > 3) & (pos[:, 0] < RP.FILM_W - 3) & (pos[:, 1] > 3) & (pos[:, 1] < RP.FILM_H - 3) & (gold["ok_box"] == 1)
    assert inside.sum() > 1000


# ================================================================ findMaxDimensions on the scene (pssmlt_utils.h:27-77)
# The reference derives the primary-sample space sizes from the scene it was given: a RoughDielectric BSDF on any shape adds
# one dimension per vertex, Russian roulette another.  Same scene description -> the reference's own function vs the oracle's.
@needs_ref_path
@pytest.mark.parametrize("scene_name", ["cornell", "roughglass"])
def test_oracle_max_dimensions_equal_reference_on_the_scene(scene_name):
    import oracle_lib
    from drmlt_mitsuba_b200 import abi
    from drmlt_mitsuba_b200.integrator import make_config
    P = C.POINTER
    lib = C.CDLL(RP.REF_PATH)
    lib.ref_scene_create.restype = C.c_void_p
    lib.ref_scene_create.argtypes = [P(abi.dr_scene_desc), C.c_int]
    lib.ref_max_dimensions.argtypes = [C.c_void_p] + [C.c_int] * 5 + [P(C.c_int)]
    data = RP.SCENES[scene_name]()
    d = data.desc()
    h = lib.ref_scene_create(C.byref(d), abi.DR_FILTER_GAUSSIAN)
    assert h
    orc = oracle_lib.OracleScene(data)
    orc.lib.orc_max_dimensions_scene.argtypes = [C.c_void_p, P(abi.dr_config), C.c_int, P(C.c_int)]
    seen = set()
    for tech_name, tech in (("path", abi.DR_TECH_PATH), ("bdpt", abi.DR_TECH_BDPT), ("mmlt", abi.DR_TECH_MMLT)):
        for max_depth in (3, 8, 11):
            for rr_depth in (5, 20):
                for direct in (False, True):
                    cfg = make_config(integrator="drmlt", type="mira", technique=tech_name, maxDepth=max_depth, rrDepth=rr_depth, directSampling=False)
                    cfg.direct_sampling = int(direct and tech_name == "bdpt")     # the product refuses it; the sizes are still defined
                    for depth in range(1, max_depth + 1):
                        a, b = (C.c_int * 3)(), (C.c_int * 3)()
                        assert lib.ref_max_dimensions(h, max_depth, rr_depth, depth, tech, int(cfg.direct_sampling), a) == 0
                        orc.lib.orc_max_dimensions_scene(orc.h, C.byref(cfg), depth, b)
                        assert list(a) == list(b), (tech_name, max_depth, rr_depth, direct, depth, list(a), list(b))
                        seen.add(tuple(a))
    assert len(seen) > 12
    if scene_name == "roughglass":
        assert (10 * 6, 0, 0) in seen          # path, maxDepth 8, RR and rough dielectric: (8 + 2) * (4 + 1 + 1)


# ================================================================ whole chains of DRMLTRenderer::process / PSSMLTRenderer::process
# ref_drmlt_chain / ref_pssmlt_chain (oracle/ref/ref_sampler.cpp, ref_pssmlt_sampler.cpp) compile drmlt_proc.cpp / pssmlt_proc.cpp into
# the driver's translation unit and run the renderer's prepare() + process() on one SeedWorkUnit with explicitly seeded generators;
# twin generators hand out the uniforms the chain consumes.  The oracle's chain step, fed those streams in call order, must consume
# exactly as many uniforms, take every decision alike (large step, first accepted, second stage, second accepted -- read off the
# reference's statistics counters after every prefix of the chain) and leave the same film after every mutation: the
# north_star's "given identical uniform streams, accept / reject decisions must be bit-exact", for mira / green / orbital x
# path / bdpt / mmlt, mixture, timidAfterLarge, fixEmitterPath, acceptanceMap and PSSMLT Kelemen / Gaussian / Veach weights.
@needs_ref_path
@pytest.mark.parametrize("case", RP.CHAIN_REPLAY_CASES, ids=lambda c: c[0])
def test_oracle_chain_equals_reference_chain(case):
    lib = C.CDLL(RP.REF_PATH)
    for pick in RP.CHAIN_PICKS:
        RP.check_chain_oracle_vs_ref(case, RP.run_chain_ref(lib, case, pick))


@pytest.fixture(scope="module")
def golden_chain():
    return dict(np.load(RP.GOLDEN_CHAIN))


@pytest.mark.parametrize("case", RP.CHAIN_REPLAY_CASES, ids=lambda c: c[0])
def test_oracle_chain_reproduces_reference_fixture(case, golden_chain):
    for pick in RP.CHAIN_PICKS:
        RP.check_chain_oracle_vs_ref(case, RP.chain_from_golden(golden_chain, case, pick))


@needs_ref_path
def test_reference_chain_fixture_is_current(golden_chain):
    live = RP.run_chain_ref_all(C.CDLL(RP.REF_PATH))
    assert sorted(live) == sorted(golden_chain)
    for k, v in live.items():
        assert np.array_equal(v, golden_chain[k]), k


# ================================================================ the drop-in plugin under the reference's plugin manager
# oracle/_ref/plugins/drmlt.so / pssmlt.so = drmlt-mitsuba_b200/shim/mts_plugin.cpp compiled against the reference (oracle/ref/Makefile).
# Without a GPU the job must fail LOUDLY and the reference's way: dr_scene_create reports "no CUDA device", the plugin raises it as
# Log(EError) and RenderJob::run catches the exception (renderjob.cpp:110-114) -- after the plugin was dlopen()ed, passed the
# plugin manager's class check (plugin.cpp:188-193) and flattened the mitsuba::Scene.  (The GPU suite renders through it.)
PLUGIN_SO = os.path.join(RP.ROOT, "oracle", "_ref", "plugins", "drmlt.so")


@pytest.mark.skipif(not os.path.exists(PLUGIN_SO), reason="oracle/_ref plugins not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", ["drmlt_orbital_mmlt", "pssmlt_path"])
def test_drop_in_plugin_loads_and_fails_loudly_without_a_gpu(name, tmp_path):
    import json
    import subprocess
    import sys
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by tests/test_gpu_parity.py")
    except ImportError:
        pass
    p = subprocess.run([sys.executable, os.path.join(RP.ROOT, "tools", "plugin_render.py"), name, str(tmp_path / "x.npy"), "4"],
                       capture_output=True, text=True, timeout=300)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("PLUGIN_RENDER ")]
    assert line and not json.loads(line[-1][len("PLUGIN_RENDER "):])["ok"], (p.stdout[-1000:], p.stderr[-1000:])
    log = p.stdout + p.stderr           # Mitsuba's logger writes to stdout
    assert "no CUDA device available (there is no CPU fallback)" in log and "caught exception" in log, log[-1500:]


def _read_scene_dump(path):
    """DRMLT_DUMP_SCENE of shim/mts_plugin.cpp -> dict of arrays."""
    from drmlt_mitsuba_b200 import abi
    raw = open(path, "rb").read()
    off = [0]

    def take(dtype, n):
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off[0])
        off[0] += a.nbytes
        return a
    nv, nt, nm, ne, ntex, has_uv = (int(x) for x in take("<u4", 6))
    d = dict(P=take("<f4", 3 * nv).reshape(nv, 3), UV=take("<f4", 2 * nv).reshape(nv, 2) if has_uv else None, I=take("<u4", 3 * nt).reshape(nt, 3),
             mat=take("<u4", nt), flags=take("<u4", nt))
    d["materials"] = [abi.dr_material.from_buffer_copy(take("u1", 64).tobytes()) for _ in range(nm)]
    d["textures"] = []
    for _ in range(ntex):
        w, h, wu, wv, nearest = (int(x) for x in take("<u4", 5))
        scale, offset = take("<f8", 2).copy(), take("<f8", 2).copy()
        d["textures"].append(dict(w=w, h=h, wrap=(wu, wv), nearest=nearest, scale=scale, offset=offset, texels=take("<f4", 3 * w * h).reshape(h, w, 3)))
    nrt = int(take("<u4", 1)[0])
    d["rough_tables"] = take("<f8", nrt * abi.DR_ROUGH_TABLE_DOUBLES).reshape(nrt, abi.DR_ROUGH_TABLE_DOUBLES)
    d["N"] = take("<f4", 3 * nv).reshape(nv, 3) if int(take("<u4", 1)[0]) else None
    d["emi"] = take("<i4", nt)
    d["emitters"] = [abi.dr_emitter.from_buffer_copy(take("u1", C.sizeof(abi.dr_emitter)).tobytes()) for _ in range(ne)]
    d["camera"] = abi.dr_camera.from_buffer_copy(take("u1", C.sizeof(abi.dr_camera)).tobytes())
    nb = int(take("<u4", 1)[0])
    assert nb == C.sizeof(abi.dr_config), "dr_config of the plugin and of the ctypes mirror differ in size"
    d["config"] = abi.dr_config.from_buffer_copy(take("u1", nb).tobytes())
    assert off[0] == len(raw)
    return d


# The plugin's flattening of a textured mitsuba::Scene, EXACTLY: the scene the reference driver assembles from the textured Cornell box
# (TriMesh objects with texture coordinates, BSDF plugins with <texture> children) goes through shim/mts_plugin.cpp, which dumps the
# dr_scene_desc it built (DRMLT_DUMP_SCENE) before it fails for want of a GPU.  Every triangle must come back with its positions, its
# texture coordinates, the UV-tangent flag, and a material whose textured parameters point at textures with the original texels, filter,
# wrap modes and uv transform -- i.e. the order in which a BSDF serializes its textures was matched to the right parameters.
@pytest.mark.skipif(not os.path.exists(PLUGIN_SO), reason="oracle/_ref plugins not built (needs /root/reference at build time)")
def test_drop_in_plugin_flattens_textured_scene_exactly(tmp_path):
    import subprocess
    import sys
    from drmlt_mitsuba_b200 import abi, scenes
    dump = str(tmp_path / "scene.bin")
    env = dict(os.environ, DRMLT_DUMP_SCENE=dump)
    subprocess.run([sys.executable, os.path.join(RP.ROOT, "tools", "plugin_render.py"), "drmlt_orbital_mmlt", str(tmp_path / "x.npy"), "4", "--textured"],
                   capture_output=True, text=True, timeout=300, env=env)
    assert os.path.exists(dump), "the plugin did not reach the end of its flattening"
    got = _read_scene_dump(dump)
    data = scenes.cornell_box_textured(film=(64, 64), tess=4)
    P, N, I, mat, emi, flg, mats, emis, rt = data.arrays()
    UV = data._uv
    assert got["I"].shape == I.shape and got["UV"] is not None and len(got["textures"]) == len(data.textures)

    def tex_key(t):
        return (t["w"], t["h"], t["wrap"], t["nearest"], tuple(t["scale"]), tuple(t["offset"]), t["texels"].tobytes())
    want_tex = [tex_key(dict(w=t.width, h=t.height, wrap=(t.wrap_u, t.wrap_v), nearest=t.nearest, scale=np.array(t.uv_scale[:]), offset=np.array(t.uv_offset[:]),
                             texels=arr)) for t, arr in zip(data.textures, data._tex_keep)]
    got_tex = [tex_key(t) for t in got["textures"]]
    assert sorted(map(hash, got_tex)) == sorted(map(hash, want_tex))          # the same textures (half-precision texels: exact), in whatever order

    def mat_key(m, texs):
        tr, tt = (m.flags >> 8) & 0xfff, m.flags >> 20
        return (m.type, m.flags & 0xff, hash(texs[tr - 1]) if tr else tuple(np.float32(m.reflectance[:]).round(5)),
                (hash(texs[tt - 1]) if tt else tuple(np.float32(m.transmittance[:]).round(5))) if m.type in (1, 4, 5, 6) else None,
                tuple(np.float32(m.eta[:]).round(4)) if m.type in (2, 3) else round(m.eta[0], 5) if m.type != 0 else None,
                tuple(np.float32(m.k[:]).round(4)) if m.type in (2, 3) else None, round(m.alpha, 6) if m.type in (3, 4, 6) else None)
    original = {}
    for t in range(len(I)):
        key = tuple(np.round(P[I[t]].astype(np.float64), 6).reshape(-1))
        original[key] = (UV[I[t]].copy(), int(flg[t]), mat_key(data.materials[mat[t]], want_tex))
    assert len(original) == len(I)
    for t in range(len(got["I"])):
        tri = got["I"][t]
        key = tuple(np.round(got["P"][tri].astype(np.float64), 6).reshape(-1))
        uv, fl, mk = original[key]
        # a mesh with texture coordinates has tangents (trimesh.cpp:400-402); one without is flagged so (the ceiling and the right wall)
        uvbits = abi.DR_TRI_UV_TANGENTS | abi.DR_TRI_NO_TEXCOORDS
        assert (got["flags"][t] & uvbits) == (fl & uvbits) and (fl & uvbits) in (abi.DR_TRI_UV_TANGENTS, abi.DR_TRI_NO_TEXCOORDS), (t, got["flags"][t], fl)
        if not (fl & abi.DR_TRI_NO_TEXCOORDS):
            assert np.array_equal(got["UV"][tri], uv), t
        assert mat_key(got["materials"][got["mat"][t]], got_tex) == mk, (t, mat_key(got["materials"][got["mat"][t]], got_tex), mk)
    # the constants of textured parameters are the textures' averages (Texture::getAverage)
    assert sum(1 for f in got["flags"] if f & abi.DR_TRI_NO_TEXCOORDS) == 2 * 2 * 4 * 4            # two walls of 4 x 4 quads
    for m in got["materials"]:
        tr = (m.flags >> 8) & 0xfff
        if tr:
            assert np.allclose(m.reflectance[:], got["textures"][tr - 1]["texels"].reshape(-1, 3).astype(np.float64).mean(axis=0), rtol=2e-3)


# ... and the same round trip for every BSDF model of the path: dr_scene_desc -> the reference's own objects (BSDF plugins, TriMesh, area
# emitters; oracle/ref/ref_path.cpp) -> shim/mts_plugin.cpp -> dr_scene_desc.  Constants nested in `twosided` are parsed from toString()
# (6 significant digits), the others come from the BSDF's Properties.
@pytest.mark.skipif(not os.path.exists(PLUGIN_SO), reason="oracle/_ref plugins not built (needs /root/reference at build time)")
@pytest.mark.parametrize("scene_name", ["cornell", "glossy", "caustic", "roughglass", "roughglass-beckmann", "plastic", "roughplastic"])
def test_drop_in_plugin_flattening_round_trip(scene_name, tmp_path):
    import subprocess
    import sys
    from drmlt_mitsuba_b200 import abi
    dump = str(tmp_path / "scene.bin")
    subprocess.run([sys.executable, os.path.join(RP.ROOT, "tools", "plugin_render.py"), "drmlt_orbital_mmlt", str(tmp_path / "x.npy"), "4", "--scene=" + scene_name],
                   capture_output=True, text=True, timeout=600, env=dict(os.environ, DRMLT_DUMP_SCENE=dump))
    assert os.path.exists(dump), "the plugin did not reach the end of its flattening"
    got = _read_scene_dump(dump)
    data = RP.SCENES[scene_name]()
    P, N, I, mat, emi, flg, mats, emis, rt = data.arrays()
    assert got["I"].shape == I.shape and got["UV"] is None and not got["textures"]

    def close(a, b):
        return np.allclose(np.float64(a[:]), np.float64(b[:]), rtol=2e-5, atol=1e-7)
    original = {}
    for t in range(len(I)):
        original.setdefault(tuple(np.round(P[I[t]].astype(np.float64), 6).reshape(-1)), []).append(t)
    for t in range(len(got["I"])):
        cands = original[tuple(np.round(got["P"][got["I"][t]].astype(np.float64), 6).reshape(-1))]
        g = got["materials"][got["mat"][t]]
        ok = False
        for o in cands:                      # (coincident triangles of different objects: any of them)
            w = data.materials[mat[o]]
            same = g.type == w.type and (g.flags & 0xff) == (w.flags & 0xff) and bool(got["flags"][t] & abi.DR_TRI_SMOOTH) == bool(flg[o] & abi.DR_TRI_SMOOTH)
            if w.type in (0, 2, 3, 5, 6) or w.type in (1, 4):
                same &= close(g.reflectance, w.reflectance)
            if w.type in (1, 4, 5, 6):
                same &= close(g.transmittance, w.transmittance)
            if w.type in (2, 3):
                same &= close(g.eta, w.eta) and close(g.k, w.k)
            elif w.type != 0:
                same &= abs(g.eta[0] - w.eta[0]) <= 2e-5 * w.eta[0]
            if w.type in (3, 4, 6):
                same &= abs(g.alpha - w.alpha) <= 2e-5 * w.alpha
            if w.type == 6:                  # the rough-transmittance table the host's own RoughTransmittance produced ([102]: filled by the library)
                same &= np.allclose(got["rough_tables"][g.table][:102], data.rough_tables[w.table][:102], rtol=1e-9, atol=1e-12)
            ok |= bool(same)
        assert ok, (scene_name, t, g.type, g.flags, g.reflectance[:], g.transmittance[:], g.eta[:], g.k[:], g.alpha)
        # smooth triangles come back with their vertex normals; emissive ones with their emitter
        o = cands[0]
        if flg[o] & abi.DR_TRI_SMOOTH:
            assert np.allclose(got["N"][got["I"][t]], N[I[o]], atol=1e-6)
        assert (got["emi"][t] >= 0) == (emi[o] >= 0)
        if emi[o] >= 0:
            ge, we = got["emitters"][got["emi"][t]], data.emitters[emi[o]]
            assert close(ge.radiance, we.radiance) and abs(ge.sampling_weight - we.sampling_weight) < 1e-6 and ge.n_tris == we.n_tris
            assert ge.first_tri <= t < ge.first_tri + ge.n_tris
    # the sensor and the film (perspective.cpp:125-187, film.cpp:30-48)
    gc, wc = got["camera"], data.camera
    assert np.allclose(gc.to_world[:], wc.to_world[:], atol=1e-6) and abs(gc.xfov_deg - wc.xfov_deg) < 1e-4
    assert abs(gc.near_clip - wc.near_clip) <= 1e-6 * wc.near_clip and abs(gc.far_clip - wc.far_clip) <= 1e-6 * wc.far_clip
    assert (gc.film_width, gc.film_height) == data.film
    # the job's configuration: the reference's parameter names forwarded verbatim, film / crop / sampler / filter read off the scene
    cfg, (params, _) = got["config"], RP.RENDER_CASES["drmlt_orbital_mmlt"]
    assert (cfg.integrator, cfg.technique, cfg.type, cfg.max_depth, cfg.direct_samples) == (abi.DR_INTEGRATOR_DRMLT, abi.DR_TECH_MMLT, abi.DR_TYPE_ORBITAL, params["maxDepth"], -1)
    assert cfg.sample_count == 4 and (cfg.film_width, cfg.film_height, cfg.crop_width, cfg.crop_height) == data.film + data.film
    assert (cfg.crop_offset_x, cfg.crop_offset_y) == (0, 0) and cfg.rr_depth == 5 and abs(cfg.p_large - 0.3) < 1e-7 and abs(cfg.sigma - 1 / 64) < 1e-9
    assert cfg.rfilter == abi.DR_FILTER_TABLE and abs(cfg.filter_radius - 2.0) < 1e-12            # gaussian: radius 2, table = evalDiscretized
    table = np.array(cfg.filter_table[:])
    assert table[0] > 0.9 * table.max() and (np.diff(table) <= 1e-12).all() and 0 <= table[-1] < 1e-2 * table[0]


# ... and analytic shapes: the room of Mitsuba `rectangle` / `sphere` shapes (ref_path.cpp, REF_ANALYTIC_SCENE) comes out of the plugin as the
# shapes' own tessellations (Shape::createTriMesh): six rectangles of two triangles, a sphere of 1 482 triangles whose vertices lie on the
# sphere, the emissive rectangle as the one area emitter -- with the texture coordinates and UV tangents those meshes carry.
@pytest.mark.skipif(not os.path.exists(PLUGIN_SO), reason="oracle/_ref plugins not built (needs /root/reference at build time)")
def test_drop_in_plugin_tessellates_analytic_shapes_into_the_description(tmp_path):
    import subprocess
    import sys
    from drmlt_mitsuba_b200 import abi
    dump = str(tmp_path / "scene.bin")
    subprocess.run([sys.executable, os.path.join(RP.ROOT, "tools", "plugin_render.py"), "drmlt_orbital_mmlt", str(tmp_path / "x.npy"), "4", "--analytic"],
                   capture_output=True, text=True, timeout=600, env=dict(os.environ, DRMLT_DUMP_SCENE=dump))
    assert os.path.exists(dump), "the plugin did not reach the end of its flattening"
    d = _read_scene_dump(dump)
    counts = np.bincount(d["mat"], minlength=7)
    assert sorted(counts.tolist()) == [2] * 6 + [1482] and len(d["materials"]) == 7
    sphere = np.unique(d["I"][d["mat"] == int(np.argmax(counts))])
    r = np.linalg.norm(d["P"][sphere] - np.array([0.2, -0.55, -0.1]), axis=1)
    assert np.abs(r - 0.45).max() < 1e-6
    assert len(d["emitters"]) == 1 and d["emitters"][0].n_tris == 2 and np.allclose(d["emitters"][0].radiance[:], 15.0)
    assert (d["emi"] >= 0).sum() == 2 and all(m.type == abi.DR_BSDF_DIFFUSE for m in d["materials"])
    assert d["UV"] is not None and (d["flags"] & abi.DR_TRI_UV_TANGENTS).all() and not (d["flags"] & abi.DR_TRI_NO_TEXCOORDS).any()
    reds = [m for m in d["materials"] if abs(m.reflectance[0] - 0.63) < 1e-5]
    assert len(reds) == 1 and abs(reds[0].reflectance[1] - 0.06) < 1e-5


# ... and the parameters: for every integrator configuration of RENDER_CASES the dr_config the plugin hands to dr_render equals the one
# dr_config_set builds from the same `-D key=value` names (integrator.make_config), field by field -- apart from what the plugin reads off
# the scene (film, crop, sample count, reconstruction filter).
@pytest.mark.skipif(not os.path.exists(PLUGIN_SO), reason="oracle/_ref plugins not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", list(RP.RENDER_CASES))
def test_drop_in_plugin_forwards_the_references_parameters(name, tmp_path):
    import subprocess
    import sys
    from drmlt_mitsuba_b200 import abi
    from drmlt_mitsuba_b200.integrator import make_config
    dump = str(tmp_path / "scene.bin")
    subprocess.run([sys.executable, os.path.join(RP.ROOT, "tools", "plugin_render.py"), name, str(tmp_path / "x.npy"), "4"],
                   capture_output=True, text=True, timeout=600, env=dict(os.environ, DRMLT_DUMP_SCENE=dump))
    assert os.path.exists(dump), "the plugin did not reach the end of its flattening"
    got = _read_scene_dump(dump)["config"]
    params, _ = RP.RENDER_CASES[name]
    want = make_config(**params)
    from_scene = {"sample_count", "film_width", "film_height", "crop_offset_x", "crop_offset_y", "crop_width", "crop_height", "rfilter", "filter_radius", "filter_table",
                  "importance_map", "seed"}
    for field, _ty in abi.dr_config._fields_:
        if field in from_scene:
            continue
        a, b = getattr(got, field), getattr(want, field)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), (name, field, a, b)
    assert got.sample_count == 4 and (got.film_width, got.film_height) == (64, 64)


# ================================================================ SURVEY 8f rank 1 / rank 3: the direct pass, the importance-map resampling
# ref_direct_image runs BidirectionalUtils::renderDirectComponent itself (src/libbidir/util.cpp:30-94: the `direct` integrator with
# the pixelSamples x shadingSamples split, an `ldsampler`, SamplingIntegrator::render on local workers, film->develop);
# ref_resample_luminance the up-sampling step of mltLuminancePass (util.cpp:175-196): Bitmap::resample with the gaussian plugin,
# clamped borders, clamped to [0, inf).  The resampling is deterministic -> bit for bit; the direct image is compared statistically.
@needs_ref_path
def test_oracle_resampling_equals_reference_bitmap_resample_bit_for_bit():
    lib = C.CDLL(RP.REF_PATH)
    lib.ref_resample_luminance.argtypes = [RP.PD, C.c_int, C.c_int, C.c_int, C.c_int, RP.PD]
    got = RP.run_resample_oracle(C.CDLL(RP.ORACLE))
    for i, shape in enumerate(RP.RESAMPLE_SHAPES):
        (w, h), (W, H) = shape
        src, dst = np.ascontiguousarray(RP.resample_input(shape)), np.zeros((H, W))
        assert lib.ref_resample_luminance(src.ctypes.data_as(RP.PD), w, h, W, H, dst.ctypes.data_as(RP.PD)) == 0
        assert np.array_equal(got["resample_%d" % i], dst), shape


def test_oracle_resampling_reproduces_reference_fixture():
    gold = dict(np.load(RP.GOLDEN_DIRECT))
    got = RP.run_resample_oracle(C.CDLL(RP.ORACLE))
    for k, v in got.items():
        assert np.array_equal(v, gold[k]), k
    assert (gold["resample_0"] >= 0).all()


@pytest.mark.parametrize("name", list(RP.DIRECT_SCENES))
def test_oracle_direct_image_matches_reference_direct_integrator(name):
    from drmlt_mitsuba_b200 import abi
    from drmlt_mitsuba_b200.integrator import make_config
    gold = dict(np.load(RP.GOLDEN_DIRECT))
    orc = oracle_lib.OracleScene(RP.DIRECT_SCENES[name]())
    cfg = make_config(integrator="drmlt", technique="mmlt", type="orbital", maxDepth=6, directSamples=RP.DIRECT_SAMPLES, seed=3)
    o = abi.dr_config.from_buffer_copy(cfg)
    o.ray_epsilon = o.shadow_epsilon = 0.0
    img = orc.direct_image(o)
    RP.check_direct_image(img[0] if isinstance(img, tuple) else img, gold["direct_" + name], name)


# ---------------------------------------------------------------- roughplastic: the rough-transmittance tables
MICROFACET_DIR = "/root/reference/data/microfacet"


@pytest.mark.skipif(not os.path.isdir(MICROFACET_DIR), reason="needs the reference's data/microfacet/*.dat")
@pytest.mark.parametrize("key", sorted(RP.ROUGH_TABLES))
def test_rough_table_reduction_equals_reference_rough_transmittance(key):
    """drmlt_mitsuba_b200/rough_tables.py (what a non-Mitsuba host uses to fill dr_scene_desc.rough_tables) against the reference's own
    RoughTransmittance::setEta / setAlpha / evalDiffuse (tests/golden/ref_rough_tables.npz): the 100 theta samples and both diffuse
    transmittances to 1e-15."""
    from drmlt_mitsuba_b200 import rough_tables
    ggx, eta, alpha = RP.ROUGH_TABLES[key]
    got = rough_tables.reduce(os.path.join(MICROFACET_DIR, "ggx.dat" if ggx else "beckmann.dat"), eta, alpha)
    want = RP.rough_table(key)
    assert np.abs(got[:102] - want[:102]).max() <= 1e-15
    with pytest.raises(ValueError):
        rough_tables.reduce(os.path.join(MICROFACET_DIR, "beckmann.dat"), 9.0, 0.1)       # checkEta (rtrans.h:356-364)


@pytest.mark.parametrize("key", sorted(RP.ROUGH_TABLES))
def test_oracle_rough_transmittance_equals_reference_eval(key):
    """RoughTransmittance::eval(cosTheta) of the reduced table (rtrans.h:136-146 + evalCubicInterp1D) at 260 angles incl. the
    out-of-range ones, through the oracle's roughplastic pdf: pdf(wi, wo) - the diffuse part isolates T(cos wi)."""
    gold = dict(np.load(RP.GOLDEN_ROUGH))
    lib = C.CDLL(RP.ORACLE)
    lib.orc_rough_transmittance.restype = C.c_double
    lib.orc_rough_transmittance.argtypes = [C.POINTER(C.c_double), C.c_double]
    t = RP.rough_table(key)
    got = np.array([lib.orc_rough_transmittance(t.ctypes.data_as(C.POINTER(C.c_double)), float(c)) for c in gold["probe"]])
    assert np.abs(got - gold[key + "_probe"]).max() <= 2e-16
