"""Seeded cases for the reference's own BSDF plugins and PathSampler (oracle/_ref/libref_path.so, built by
oracle/ref/Makefile from /root/reference) against the oracle restatement.  Test infrastructure only."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from drmlt_mitsuba_b200 import abi  # noqa: E402

REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libref_path.so")
ORACLE = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_path.npz")
D = C.c_double
PD = C.POINTER(D)

# reference BSDF type bits (bsdf.h:230-242) -> the oracle's compact encoding (orc_bsdf.hpp:16-19)
TYPE_MAP = {0: 0, 0x2: 0x1, 0x8: 0x2, 0x20: 0x4, 0x40: 0x8, 0x10: 0x10}


def material(type_, flags=0, reflectance=(1, 1, 1), transmittance=(1, 1, 1), eta=(1.5, 0, 0), k=(0, 0, 0), alpha=0.1):
    m = abi.dr_material()
    m.type, m.flags, m.alpha = type_, flags, alpha
    m.reflectance[:], m.transmittance[:], m.eta[:], m.k[:] = reflectance, transmittance, eta, k
    return m


# roughplastic: (ggx, eta, alpha) of the rough-transmittance tables in tests/golden/ref_rough_tables.npz -- made by the reference's own
# RoughTransmittance (tools/make_ref_roughplastic_golden.py); eta / alpha are the float values the materials carry
GOLDEN_ROUGH = os.path.join(ROOT, "tests", "golden", "ref_rough_tables.npz")
ROUGH_TABLES = {"beckmann_1.49_0.1": (0, 1.49, 0.1), "ggx_1.9_0.3": (1, 1.9, 0.3), "beckmann_1.33_0.05": (0, 1.33, 0.05)}
ROUGH_OF = {"roughplastic_beckmann": "beckmann_1.49_0.1", "roughplastic_ggx_visible_nonlinear": "ggx_1.9_0.3",
            "roughplastic_beckmann_visible_twosided": "beckmann_1.33_0.05"}
_rough_cache = {}


def rough_table(key):
    if not _rough_cache:
        _rough_cache.update(dict(np.load(GOLDEN_ROUGH)))
    return np.ascontiguousarray(_rough_cache[key], np.float64)


def bsdf_materials():
    T, G, V, N = abi.DR_MAT_TWOSIDED, abi.DR_MAT_GGX, abi.DR_MAT_SAMPLE_VISIBLE, abi.DR_MAT_NONLINEAR
    cu = dict(eta=(0.2004, 0.9240, 1.1022), k=(3.9129, 2.4528, 2.1421))
    return {
        "diffuse": material(0, reflectance=(0.5, 0.6, 0.7)),
        "diffuse_twosided": material(0, T, reflectance=(0.73, 0.2, 0.1)),
        "dielectric": material(1, reflectance=(1, 0.9, 0.8), transmittance=(0.7, 0.8, 0.9), eta=(1.5, 0, 0)),
        "conductor": material(2, reflectance=(1, 0.9, 0.8), **cu),
        "roughconductor_beckmann": material(3, 0, reflectance=(0.9, 0.9, 0.9), alpha=0.2, **cu),
        "roughconductor_beckmann_visible": material(3, V, alpha=0.15, **cu),
        "roughconductor_ggx": material(3, G, alpha=0.3, **cu),
        "roughconductor_ggx_visible": material(3, G | V, alpha=0.05, **cu),
        "roughdielectric_beckmann": material(4, 0, eta=(1.5, 0, 0), alpha=0.2),
        "roughdielectric_ggx_visible": material(4, G | V, eta=(1.33, 0, 0), alpha=0.1, transmittance=(0.9, 0.95, 1.0)),
        "roughdielectric_beckmann_visible": material(4, V, eta=(1.5, 0, 0), alpha=0.3),
        "plastic": material(5, 0, reflectance=(0.5, 0.3, 0.2), transmittance=(1, 1, 1), eta=(1.49, 0, 0)),
        "plastic_nonlinear": material(5, N, reflectance=(0.2, 0.3, 0.7), transmittance=(0.9, 0.9, 0.9), eta=(1.9, 0, 0)),
        "roughplastic_beckmann": material(6, 0, reflectance=(0.5, 0.3, 0.2), transmittance=(1, 1, 1), eta=(1.49, 0, 0), alpha=0.1),
        "roughplastic_ggx_visible_nonlinear": material(6, G | V | N, reflectance=(0.2, 0.3, 0.7), transmittance=(0.9, 0.9, 0.9), eta=(1.9, 0, 0), alpha=0.3),
        "roughplastic_beckmann_visible_twosided": material(6, V | T, reflectance=(0.4, 0.4, 0.1), transmittance=(1, 0.8, 0.6), eta=(1.33, 0, 0), alpha=0.05),
    }


def _dirs(rng, n, both=True):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if not both:
        v[:, 2] = np.abs(v[:, 2])
    return np.ascontiguousarray(v)


def run_bsdf(lib, prefix, n=600, seed=7):
    """sample() and eval()/pdf() of every material on seeded directions; -> dict name -> array."""
    is_ref = prefix == "ref_"
    smp = getattr(lib, "ref_bsdf_sample" if is_ref else "orc_bsdf_sample3")
    smp.restype = None
    smp.argtypes = [C.c_void_p, PD, C.c_int, D, D, D, PD, PD, PD, C.POINTER(C.c_int), PD]
    ev = getattr(lib, prefix + "bsdf_eval")
    ev.restype = None
    ev.argtypes = [C.c_void_p, PD, PD, C.c_int, C.c_int, PD, PD]
    out = {}
    for name, m in bsdf_materials().items():
        if m.type == abi.DR_BSDF_ROUGHPLASTIC and not is_ref:      # the oracle takes the table, the reference plugin reads its own data file
            lib.orc_rough_table_register.restype = C.c_uint32
            lib.orc_rough_table_register.argtypes = [C.c_void_p, PD]
            m.table = lib.orc_rough_table_register(C.byref(m), rough_table(ROUGH_OF[name]).ctypes.data_as(PD))
        rng = np.random.default_rng(seed)
        two_sided_ok = m.type in (1, 4) or (m.flags & abi.DR_MAT_TWOSIDED)
        wi = _dirs(rng, n, both=bool(two_sided_ok))
        wo_e = _dirs(rng, n, both=True)
        u = rng.random((n, 3))
        mp = C.byref(m)
        res = np.zeros((n, 2, 14))
        for i in range(n):
            for mode in (0, 1):
                wo, w, pdf, ty, eta = (D * 3)(), (D * 3)(), D(), C.c_int(), D()
                smp(mp, wi[i].ctypes.data_as(PD), mode, u[i, 0], u[i, 1], u[i, 2], wo, w, C.byref(pdf), C.byref(ty), C.byref(eta))
                t = TYPE_MAP[ty.value] if is_ref else ty.value
                zero = pdf.value == 0 or (w[0] == 0 and w[1] == 0 and w[2] == 0)      # a failed sample: outputs undefined
                res[i, mode, :9] = [0] * 9 if zero else [wo[0], wo[1], wo[2], w[0], w[1], w[2], pdf.value, t, eta.value]
                v, p = (D * 3)(), D()
                ev(mp, wi[i].ctypes.data_as(PD), wo_e[i].ctypes.data_as(PD), mode, 1, v, C.byref(p))      # ESolidAngle
                res[i, mode, 9:13] = [v[0], v[1], v[2], p.value]
                if not zero and not (t & (0x4 | 0x8)):      # pdf of the sampled direction (smooth lobes)
                    v2, p2 = (D * 3)(), D()
                    ev(mp, wi[i].ctypes.data_as(PD), wo, mode, 1, v2, C.byref(p2))
                    res[i, mode, 13] = p2.value
        out["bsdf_" + name] = res
    return out


# ---------------------------------------------------------------- PathSampler::sampleSplats on replayed vectors
from drmlt_mitsuba_b200 import scenes  # noqa: E402
from drmlt_mitsuba_b200.integrator import make_config  # noqa: E402

SCENES = {
    "cornell": lambda: scenes.cornell_box(film=(128, 128), tess=8),
    "glossy": lambda: scenes.glossy_scene(film=(128, 128), subdiv=3),
    "caustic": lambda: scenes.caustic_scene(film=(128, 128), grid=48),
    "roughglass": lambda: scenes.glossy_scene(film=(128, 128), subdiv=3, rough_glass=(0.15, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE)),
    "roughglass-beckmann": lambda: scenes.glossy_scene(film=(128, 128), subdiv=3, rough_glass=(0.3, 0)),
    "plastic": lambda: scenes.cornell_box(film=(128, 128), tess=8, plastic=True),
    "roughplastic": lambda: scenes.cornell_box(film=(128, 128), tess=8, rough_tables=(rough_table("beckmann_1.49_0.1"), rough_table("ggx_1.9_0.3"))),
    # the bench's C5 "door" scene at test size (occluded light, displaced floor, spheres)
    "door": lambda: scenes.door_scene(film=(160, 90), floor_grid=64, n_spheres=16, sphere_subdiv=2),
}
# (scene, technique, maxDepth): C1..C4-shaped configurations of SURVEY 8 plus the 8f BSDFs
PATH_CASES = [("cornell", "mmlt", 6), ("cornell", "bdpt", 5), ("cornell", "path", 8),
              ("glossy", "mmlt", 8), ("glossy", "bdpt", 6), ("glossy", "path", 6),
              ("caustic", "mmlt", 8), ("caustic", "path", 8),
              ("roughglass", "mmlt", 8), ("roughglass", "path", 6), ("roughglass-beckmann", "bdpt", 5),
              ("plastic", "mmlt", 8), ("plastic", "bdpt", 6), ("plastic", "path", 8),
              ("door", "mmlt", 8), ("door", "path", 8),
              ("roughplastic", "mmlt", 8), ("roughplastic", "bdpt", 6), ("roughplastic", "path", 8)]
# bitmap textures: their own fixture, tests/golden/ref_texture.npz (tools/make_ref_texture_golden.py)
SCENES["textured"] = lambda: scenes.cornell_box_textured(film=(128, 128), tess=8, uv_tangents=True)
TEXTURE_CASES = [("textured", "mmlt", 8), ("textured", "bdpt", 6), ("textured", "path", 8)]
ALL_CASES = PATH_CASES + TEXTURE_CASES
GOLDEN_TEXTURE = os.path.join(ROOT, "tests", "golden", "ref_texture.npz")
N_PATHS = 6000
N_PATHS_OF = {"door": 48000}      # the door scene's light is occluded: ~1.5 % (MMLT) / 6 % (PT) of uniform vectors contribute


def n_paths(case):
    return N_PATHS_OF.get(case[0], N_PATHS)


def case_key(case):
    return "path_%s_%s_%d" % case


def case_config(case):
    name, tech, md = case
    params = dict(integrator="drmlt", type="orbital", technique=tech, maxDepth=md, directSamples=-1)
    if tech == "bdpt":
        params["directSampling"] = False       # the reference overflows its direct sampler otherwise
    return make_config(seed=3, **params)


def case_inputs(case, n=None):
    """The replayed primary-sample vectors of a case: a pure function of (case, n)."""
    name, tech, md = case
    n = n or n_paths(case)
    rng = np.random.RandomState(1234 + 17 * ALL_CASES.index(case))
    depth = rng.randint(1, md + 1, n).astype(np.int32)
    ds, de, dd = (6 * (md + 2), 2, 2) if tech == "path" else (3 * (md + 2), 3 * (md + 2), 1)
    us, ue, ud = [rng.rand(n, k).astype(np.float32) for k in (ds, de, dd)]
    return us, ue, ud, depth


def unpack(results, n):
    """dr_path_result array -> dict of arrays (luminance f32, n_splats, s, t, pos0, value0)."""
    r = np.frombuffer(results, dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
    off = 20 + 8 * abi.DR_MAX_SPLATS
    return dict(n_splats=r[:, 4:8].copy().view("<i4")[:, 0], s=r[:, 8:12].copy().view("<i4")[:, 0], t=r[:, 12:16].copy().view("<i4")[:, 0],
                pos0=r[:, 20:28].copy().view("<f4"), value0=r[:, off:off + 12].copy().view("<f4"))


_ref_scenes = {}


def run_paths_ref(lib, case, n=None):
    """The reference's own PathSampler on the case's vectors -> dict of arrays (lum in double)."""
    P = C.POINTER
    lib.ref_scene_create.restype = C.c_void_p
    lib.ref_scene_create.argtypes = [P(abi.dr_scene_desc), C.c_int]
    lib.ref_eval_paths.argtypes = [C.c_void_p] + [C.c_int] * 5 + [P(C.c_float), C.c_int] * 3 + [P(C.c_int32), C.c_int64, P(abi.dr_path_result), P(C.c_double), P(C.c_int32)]
    cfg = case_config(case)
    if case[0] not in _ref_scenes:
        data = SCENES[case[0]]()
        d = data.desc()
        h = lib.ref_scene_create(C.byref(d), cfg.rfilter)
        assert h, "oracle/_ref could not build the scene"
        _ref_scenes[case[0]] = (h, data, d)
    h = _ref_scenes[case[0]][0]
    n = n or n_paths(case)
    us, ue, ud, depth = case_inputs(case, n)
    out = (abi.dr_path_result * n)()
    lum = np.zeros(n, np.float64)
    f = lambda a, ty=C.c_float: a.ctypes.data_as(P(ty))     # noqa: E731
    rc = lib.ref_eval_paths(h, cfg.technique, cfg.max_depth, cfg.rr_depth, cfg.direct_sampling, cfg.light_image,
                            f(us), us.shape[1], f(ue), ue.shape[1], f(ud), ud.shape[1], f(depth, C.c_int32), n, out, f(lum, C.c_double), None)
    assert rc == 0, "the reference asked for more primary samples than findMaxDimensions provides"
    res = unpack(out, n)
    res["lum"] = lum
    return res


NOISE = 1e-18      # contributions 18 orders below the brightest are rounding noise (tests/test_gpu_parity.py)


def compare_paths(lum, st, pos0, value0, want_lum, want_st, want_pos0, want_value0, what, mmlt):
    """f(u), strategy, splat count, pixel and RGB of one implementation against the reference's."""
    top = want_lum.max()
    a = np.where(np.abs(lum) < NOISE * top, 0.0, lum)
    b = np.where(np.abs(want_lum) < NOISE * top, 0.0, want_lum)
    support = (a > 0) == (b > 0)
    both = (a > 0) & (b > 0)
    assert both.sum() > 500, what
    rel = np.abs(a[both] - b[both]) / b[both]
    ok = support.copy()
    ok[np.nonzero(both)[0][rel >= 1e-4]] = False
    assert ok.mean() >= 0.999, "%s: %.5f of the paths within 1e-4 (support mismatches %d, worst %.3g)" % (what, ok.mean(), (~support).sum(), rel.max())
    if mmlt:
        assert np.array_equal(st[:, :2], want_st[:, :2]), what + ": MMLT strategies differ"
    assert (st[both, 2] == want_st[both, 2]).mean() >= 0.999, what + ": splat counts differ"
    assert (np.abs(pos0[both] - want_pos0[both]).max(axis=1) < 2e-2).mean() >= 0.999, what + ": pixels differ"
    scale = np.abs(want_value0[both]).max(axis=1, keepdims=True)
    okv = (np.abs(value0[both] - want_value0[both]) <= 1e-4 * np.abs(want_value0[both]) + 1e-6 * scale).all(axis=1)
    assert okv.mean() >= 0.999, what + ": splat RGB differs"
    return rel


# ---------------------------------------------------------------- bitmap texture lookups (Texture2D::eval without ray differentials)
def texture_leaf_cases():
    """name -> (dr_texture, keep-alive array): every wrap mode, both filters, non-square sizes, scale / offset, a 1x1 texture."""
    out = {}
    W = dict(repeat=abi.DR_WRAP_REPEAT, clamp=abi.DR_WRAP_CLAMP, mirror=abi.DR_WRAP_MIRROR, zero=abi.DR_WRAP_ZERO, one=abi.DR_WRAP_ONE)

    def add(name, w, h, seed, wu, wv, nearest, scale=(1.0, 1.0), offset=(0.0, 0.0)):
        arr = scenes.procedural_texels(w, h, seed, cell=2)
        t = abi.dr_texture()
        t.width, t.height, t.texels = w, h, arr.ctypes.data_as(C.POINTER(C.c_float))
        t.wrap_u, t.wrap_v, t.nearest = W[wu], W[wv], int(nearest)
        t.uv_scale[:], t.uv_offset[:] = scale, offset
        out[name] = (t, arr)
    for i, wu in enumerate(W):
        add("bilinear_" + wu, 7 + i, 5 + 2 * i, 40 + i, wu, wu, False)
        add("nearest_" + wu, 9 - i, 4 + i, 50 + i, wu, wu, True)
    add("bilinear_repeat_mirror_scaled", 16, 8, 60, "repeat", "mirror", False, scale=(3.7, 0.1), offset=(0.3, -2.25))
    add("nearest_zero_clamp_scaled", 5, 12, 61, "zero", "clamp", True, scale=(0.5, 2.0), offset=(-0.2, 0.6))
    add("bilinear_1x1", 1, 1, 62, "repeat", "clamp", False)
    return out


def texture_leaf_uv(n=4000, seed=99):
    rng = np.random.RandomState(seed)
    uv = rng.uniform(-2.5, 3.5, (n, 2))
    uv[:200] = rng.uniform(0, 1, (200, 2))
    uv[200:220] = np.array([[0, 0], [1, 1], [0.5, 0.5], [1, 0], [0, 1], [-1, -1], [2, 2], [1e-17, 1 - 1e-16], [0.25, 0.75], [-0.0, 1.0]] * 2)
    return np.ascontiguousarray(uv, np.float64)


def run_texture(lib, prefix):
    """RGB of every leaf case at the uv set; -> dict name -> [n, 3] (double)."""
    fn = getattr(lib, prefix + "texture_eval")
    uv = texture_leaf_uv()
    out = {}
    for name, (t, _keep) in texture_leaf_cases().items():
        rgb = np.zeros((len(uv), 3), np.float64)
        if prefix == "ref_":
            fn.argtypes = [C.c_void_p, PD, C.c_int, PD, PD]
            avg = np.zeros(3, np.float64)
            fn(C.byref(t), uv.ctypes.data_as(PD), len(uv), rgb.ctypes.data_as(PD), avg.ctypes.data_as(PD))
            out["texavg_" + name] = avg
        else:
            fn.argtypes = [C.c_void_p, PD, C.c_int, PD]
            fn(C.byref(t), uv.ctypes.data_as(PD), len(uv), rgb.ctypes.data_as(PD))
        out["tex_" + name] = rgb
    return out


# ---------------------------------------------------------------- the reference's own integrators, end to end
import re  # noqa: E402

GOLDEN_RENDER = os.path.join(ROOT, "tests", "golden", "ref_render.npz")
RENDER_SCENE = lambda: scenes.cornell_box(film=(64, 64), tess=4)     # noqa: E731
# name -> (integrator parameters, mutations per pixel).  The reference needs at least one work unit of 1e5 (2e5 for
# technique=path) mutations per core (drmlt.cpp:430-450, 498-546), hence the high sample counts on a small film.
RENDER_CASES = {
    "drmlt_orbital_mmlt": (dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6, directSamples=-1), 1024),
    "drmlt_mira_path": (dict(integrator="drmlt", type="mira", technique="path", maxDepth=6, directSamples=-1), 1024),
    "drmlt_green_bdpt": (dict(integrator="drmlt", type="green", technique="bdpt", maxDepth=5, directSamples=-1, directSampling=False), 512),
    "pssmlt_path": (dict(integrator="pssmlt", technique="path", maxDepth=6, directSamples=-1), 1024),
    "pssmlt_mmlt": (dict(integrator="pssmlt", technique="mmlt", maxDepth=6, directSamples=-1), 1024),
}


def parse_stats(text):
    """Statistics::getStats() text -> {counter name: percent} (StatsCounter EPercentage lines, statistics.cpp)."""
    out = {}
    for m in re.finditer(r"-\s+(.+?)\s*:\s*([0-9.]+) %", text):
        out[m.group(1).strip()] = float(m.group(2))
    return out


def run_render_ref(lib, params, spp, threads=8, data=None):
    P = C.POINTER
    lib.ref_render.argtypes = [P(abi.dr_scene_desc), P(abi.dr_config), C.c_int, C.c_int, P(C.c_float), P(C.c_double), P(C.c_double), C.c_char_p, C.c_int]
    data = data or RENDER_SCENE()
    cfg = make_config(sampleCount=spp, seed=1, **params)
    d = data.desc()
    W, H = data.film
    img = np.zeros((H, W, 3), np.float32)
    sec, ssec, buf = C.c_double(), C.c_double(), C.create_string_buffer(1 << 16)
    rc = lib.ref_render(C.byref(d), C.byref(cfg), spp, threads, img.ctypes.data_as(P(C.c_float)), C.byref(sec), C.byref(ssec), buf, len(buf))
    assert rc == 0, "oracle/_ref: the reference render failed"
    return img, sec.value, ssec.value, parse_stats(buf.value.decode())


def luminance(img):
    return img[..., 0] * np.float32(0.212671) + img[..., 1] * np.float32(0.715160) + img[..., 2] * np.float32(0.072169)


def rel_mse(img, ref, eps=1e-2):
    """relMSE = mean((I - R)^2 / (R^2 + eps)) on luminance."""
    a, b = luminance(img.astype(np.float64)), luminance(ref.astype(np.float64))
    return float(np.mean((a - b) ** 2 / (b ** 2 + eps)))

# the reference's statistics counters (drmlt_proc.cpp:34-49, pssmlt_proc.cpp:33-40) -> dr_stats numerator / denominator
STATS_MAP = {
    "Accepted 1st-stage mutations": ("first_accept", "first_base"),
    "Accepted 2nd-stage mutations": ("second_accept", "second_base"),
    "Accepted 2nd-stage mutations after bold mutation": ("second_bold_accept", "second_bold_base"),
    "Accepted bold mutation in the 1st-stage mutations": ("bold_accept", "bold_base"),
    "Accepted large mutations in the 1st-stage mutations": ("large_accept", "large_base"),
    "Overall acceptance rate": ("accept", "accept_base"),
    "Accepted large steps": ("large_accept", "large_base"),
    "Accepted small steps": ("bold_accept", "bold_base"),
}


def check_rates_and_b(name, st, b, golden):
    """north_star: per-stage acceptance rates within 1 % absolute and b within 0.5 % of the reference's.  The reference
    seeds from /dev/urandom; the fixture keeps three of its runs, and their own spread (up to 2 % absolute for MMLT,
    whose ~40 work units each stay at one path depth) is added to the bound."""
    names, runs = golden[name + "_stats_names"], golden[name + "_stats"]
    out = {}
    for k, vals in zip(names, runs.T):
        a, base = STATS_MAP[str(k)]
        ours = 100.0 * getattr(st, a) / max(1, getattr(st, base))
        out[str(k)] = (ours, vals.mean())
        assert abs(ours - vals.mean()) <= 1.0 + (vals.max() - vals.min()), (name, str(k), ours, vals)
    bs = golden[name + "_b"]
    assert abs(b - bs.mean()) <= 0.005 * bs.mean() + (bs.max() - bs.min()) / 2, (name, b, bs)
    return out


# ---------------------------------------------------------------- the reference's DRMLT samplers (Green / Mira / Orbital)
GOLDEN_SAMPLER = os.path.join(ROOT, "tests", "golden", "ref_sampler.npz")
SAMPLER_DIM, SAMPLER_SEEDS = 30, 24


def run_sampler_ref(lib):
    """ref_drmlt_sampler (oracle/ref/ref_sampler.cpp) for every type x {small, large} step x seed ->
    dict of arrays: current state, the uniform stream the sampler consumed, stage-1 / stage-2 proposals, Green's reverse
    state, Mira's transition ratio."""
    lib.ref_drmlt_sampler.argtypes = [C.c_int, C.c_int, D, D, C.c_int, C.c_uint64, PD, PD, C.c_int, PD, PD, PD, PD]
    md, ns = SAMPLER_DIM, 6 * SAMPLER_DIM
    p = lambda a: a.ctypes.data_as(PD)      # noqa: E731
    out = {}
    for type_ in (0, 1, 2):
        for large in (0, 1):
            rows = []
            for seed in range(SAMPLER_SEEDS):
                uc, st, p1, p2, rv, ra = np.zeros(md), np.zeros(ns), np.zeros(md), np.zeros(md), np.zeros(md), D()
                rc = lib.ref_drmlt_sampler(type_, md, 1.0 / 64, 0.1, large, 1000 * type_ + 100 * large + seed + 1, p(uc), p(st), ns, p(p1), p(p2), p(rv), C.byref(ra))
                assert rc == 0
                rows.append(np.concatenate([uc, st, p1, p2, rv, [ra.value]]))
            out["sampler_%d_%d" % (type_, large)] = np.array(rows)
    return out


def run_sampler_oracle(lib, ref_rows):
    """The oracle's DRMLTSampler on the current states and uniform streams recorded from the reference."""
    lib.orc_drmlt_sampler.argtypes = [C.c_int, C.c_int, D, D, C.c_int, PD, PD, PD, PD, PD, PD]
    md, ns = SAMPLER_DIM, 6 * SAMPLER_DIM
    p = lambda a: a.ctypes.data_as(PD)      # noqa: E731
    out = {}
    for key, rows in ref_rows.items():
        type_, large = int(key.split("_")[1]), int(key.split("_")[2])
        res = []
        for row in rows:
            uc, st = np.ascontiguousarray(row[:md]), np.ascontiguousarray(row[md:md + ns])
            p1, p2, rv, ra = np.zeros(md), np.zeros(md), np.zeros(md), D()
            lib.orc_drmlt_sampler(type_, md, 1.0 / 64, 0.1, large, p(uc), p(st), p(p1), p(p2), p(rv), C.byref(ra))
            res.append(np.concatenate([uc, st, p1, p2, rv, [ra.value]]))
        out[key] = np.array(res)
    return out


# ---------------------------------------------------------------- the reference's PSSMLT sampler over a sequence of mutations
GOLDEN_PSS_SAMPLER = os.path.join(ROOT, "tests", "golden", "ref_pssmlt_sampler.npz")
PSS_DIM, PSS_MUT, PSS_SEEDS = 50, 12, 16            # maxDim of C1 (pssmlt_utils.h:27-77: 50 for path, maxDepth 8)
PSS_S1, PSS_S2, PSS_SIGMA = 1.0 / 1024, 1.0 / 64, 1.0 / 64
PI32 = C.POINTER(C.c_int32)


def pss_pattern(seed):
    """Which mutations are large steps and which are accepted: a pure function of the seed (pLarge 0.3, ~half accepted)."""
    rng = np.random.RandomState(4242 + seed)
    return (rng.rand(PSS_MUT) < 0.3).astype(np.int32), (rng.rand(PSS_MUT) < 0.5).astype(np.int32)


def run_pss_sampler_ref(lib):
    """ref_pssmlt_sampler (oracle/ref/ref_pssmlt_sampler.cpp) for {Kelemen, Gaussian} x seed -> rows
    [current state | recorded uniform stream | proposals of every mutation]."""
    lib.ref_pssmlt_sampler.argtypes = [C.c_int, C.c_int, D, D, D, C.c_uint64, C.c_int, PI32, PI32, PD, PD, C.c_int, PD]
    md, ns = PSS_DIM, 2 * PSS_DIM * PSS_MUT
    p = lambda a: a.ctypes.data_as(PD)      # noqa: E731
    out = {}
    for kelemen in (1, 0):
        rows = []
        for seed in range(PSS_SEEDS):
            large, acc = pss_pattern(seed)
            uc, st, pr = np.zeros(md), np.zeros(ns), np.zeros(md * PSS_MUT)
            rc = lib.ref_pssmlt_sampler(kelemen, md, PSS_S1, PSS_S2, PSS_SIGMA, 7000 + 100 * kelemen + seed, PSS_MUT,
                                        large.ctypes.data_as(PI32), acc.ctypes.data_as(PI32), p(uc), p(st), ns, p(pr))
            assert rc == 0
            rows.append(np.concatenate([uc, st, pr]))
        out["pss_%d" % kelemen] = np.array(rows)
    return out


def run_pss_sampler_oracle(lib, ref_rows):
    """The oracle's PSSMLTSampler on the current states and uniform streams recorded from the reference."""
    lib.orc_pssmlt_sampler.argtypes = [C.c_int, C.c_int, D, D, D, C.c_int, PI32, PI32, PD, PD, PD]
    md, ns = PSS_DIM, 2 * PSS_DIM * PSS_MUT
    p = lambda a: a.ctypes.data_as(PD)      # noqa: E731
    out, used = {}, {}
    for key, rows in ref_rows.items():
        kelemen = int(key.split("_")[1])
        res, cnt = [], []
        for seed, row in enumerate(rows):
            large, acc = pss_pattern(seed)
            uc, st = np.ascontiguousarray(row[:md]), np.ascontiguousarray(row[md:md + ns])
            pr = np.zeros(md * PSS_MUT)
            cnt.append(lib.orc_pssmlt_sampler(kelemen, md, PSS_S1, PSS_S2, PSS_SIGMA, PSS_MUT, large.ctypes.data_as(PI32),
                                              acc.ctypes.data_as(PI32), p(uc), p(st), p(pr)))
            res.append(np.concatenate([uc, st, pr]))
        out[key], used[key] = np.array(res), cnt
    return out, used


# ---------------------------------------------------------------- the DRMLT samplers over a sequence of mutations
GOLDEN_SAMPLER_SEQ = os.path.join(ROOT, "tests", "golden", "ref_sampler_seq.npz")
SEQ_DIM, SEQ_MUT, SEQ_SEEDS = 30, 10, 8


def seq_pattern(seed):
    """Per mutation: large step?, outcome (0 first stage accepted / 1 second accepted / 2 second rejected), light-tracing stage?"""
    rng = np.random.RandomState(977 + seed)
    return ((rng.rand(SEQ_MUT) < 0.3).astype(np.int32), rng.randint(0, 3, SEQ_MUT).astype(np.int32), (rng.rand(SEQ_MUT) < 0.5).astype(np.int32))


def _seq_run(fn, is_ref, type_, mode, seed, uc=None, st=None):
    md, nm, ns = SEQ_DIM, SEQ_MUT, 6 * SEQ_DIM * SEQ_MUT
    p = lambda a: a.ctypes.data_as(PD)      # noqa: E731
    pi = lambda a: a.ctypes.data_as(PI32)   # noqa: E731
    large, outc, lt = seq_pattern(seed)
    if mode != 1:
        lt = np.zeros_like(lt)          # nextStage(true) only reaches a sampler after handleLightTracing() (drmlt_proc.cpp:566-571)
    p1, p2, rv, ra = np.zeros(md * nm), np.zeros(md * nm), np.zeros(md * nm), np.zeros(nm)
    if is_ref:
        uc, st = np.zeros(md), np.zeros(ns)
        rc = fn(type_, mode, md, 1.0 / 64, 0.1, 5000 + 1000 * type_ + 100 * mode + seed, nm, pi(large), pi(outc), pi(lt), p(uc), p(st), ns, p(p1), p(p2), p(rv), p(ra))
        assert rc == 0
    else:
        rc = fn(type_, mode, md, 1.0 / 64, 0.1, nm, pi(large), pi(outc), pi(lt), p(uc), p(st), p(p1), p(p2), p(rv), p(ra))
        assert 0 < rc <= ns
    return np.concatenate([uc, st, p1, p2, rv, ra])


def run_sampler_seq_ref(lib):
    """ref_drmlt_sampler_seq for every type x mode (0 plain, 1 handleLightTracing, 2 setStagesToIdentity) x seed."""
    lib.ref_drmlt_sampler_seq.argtypes = [C.c_int, C.c_int, C.c_int, D, D, C.c_uint64, C.c_int, PI32, PI32, PI32, PD, PD, C.c_int, PD, PD, PD, PD]
    return {"seq_%d_%d" % (t, m): np.array([_seq_run(lib.ref_drmlt_sampler_seq, True, t, m, s) for s in range(SEQ_SEEDS)])
            for t in (0, 1, 2) for m in (0, 1, 2)}


def run_sampler_seq_oracle(lib, ref_rows):
    lib.orc_drmlt_sampler_seq.argtypes = [C.c_int, C.c_int, C.c_int, D, D, C.c_int, PI32, PI32, PI32, PD, PD, PD, PD, PD, PD]
    md, ns = SEQ_DIM, 6 * SEQ_DIM * SEQ_MUT
    out = {}
    for key, rows in ref_rows.items():
        t, m = int(key.split("_")[1]), int(key.split("_")[2])
        out[key] = np.array([_seq_run(lib.orc_drmlt_sampler_seq, False, t, m, s, np.ascontiguousarray(r[:md]), np.ascontiguousarray(r[md:md + ns]))
                             for s, r in enumerate(rows)])
    return out


# ---------------------------------------------------------------- the reference's ImageBlock::put with its own filter plugins
GOLDEN_FILM = os.path.join(ROOT, "tests", "golden", "ref_film.npz")
FILM_W, FILM_H, FILM_N = 37, 23, 4000
PF32 = C.POINTER(C.c_float)


def film_inputs():
    """Splat positions (inside, on pixel / half-pixel lattices, up to 3 pixels outside) and values (incl. NaN, inf, negative)."""
    rng = np.random.RandomState(5)
    pos = (rng.rand(FILM_N, 2) * [FILM_W + 6, FILM_H + 6] - 3).astype(np.float32)
    pos[:50] = np.round(pos[:50] * 2) / 2
    rgb = (rng.rand(FILM_N, 3) ** 4 * 10).astype(np.float32)
    rgb[5, 1], rgb[6, 0], rgb[7, 2] = np.nan, np.inf, -1e-3
    return pos, rgb


FILM_FILTERS = (("gaussian", 1, abi.DR_FILTER_GAUSSIAN), ("box", 0, abi.DR_FILTER_BOX), ("tent", 3, abi.DR_FILTER_TENT),
                ("mitchell", 4, abi.DR_FILTER_MITCHELL), ("catmullrom", 5, abi.DR_FILTER_CATMULLROM), ("lanczos", 6, abi.DR_FILTER_LANCZOS))


def run_film(fn, is_ref):
    """Every filter plugin -> film [h][w][3] in double + the verdict of every put (+ radius and 32-entry table of the reference's)."""
    fn.argtypes = [C.c_int, C.c_int, C.c_int, PF32, PF32, C.c_int64, PD, PI32] + ([PD, PD] if is_ref else [])
    pos, rgb = film_inputs()
    out = {}
    for name, ref_id, orc_id in FILM_FILTERS:
        film, ok = np.zeros((FILM_H, FILM_W, 3)), np.zeros(FILM_N, np.int32)
        radius, table = D(), np.zeros(32)
        extra = [C.byref(radius), table.ctypes.data_as(PD)] if is_ref else []
        assert fn(FILM_W, FILM_H, ref_id if is_ref else orc_id, pos.ctypes.data_as(PF32), rgb.ctypes.data_as(PF32), FILM_N, film.ctypes.data_as(PD),
                  ok.ctypes.data_as(PI32), *extra) == 0
        out["film_" + name], out["ok_" + name] = film, ok
        if is_ref:
            out["table_" + name] = np.concatenate([[radius.value], table])
    return out


# ---------------------------------------------------------------- WHOLE CHAINS of the reference, replayed
# ref_drmlt_chain / ref_pssmlt_chain (oracle/ref/ref_sampler.cpp, ref_pssmlt_sampler.cpp) run the reference's own
# DRMLTRenderer::process / PSSMLTRenderer::process (drmlt_proc.cpp:386-771, :161-380; pssmlt_proc.cpp:110-285) on one work unit with
# explicitly seeded generators, for every prefix 0..K of the chain, and hand out the two uniform streams the chain consumes.
# The oracle's chain step (orc_chain_stream) consumes those streams in the reference's call order; the CUDA path replays the
# same chain from the table of keyed uniforms the oracle derives from the streams (dr_chain_replay).
GOLDEN_CHAIN = os.path.join(ROOT, "tests", "golden", "ref_chain.npz")
CHAIN_K = 24                       # mutations per replayed chain
CHAIN_PICKS = (0, 5)               # two seeds per case
CHAIN_NBOOT, CHAIN_NSEEDS = 3000, 8
CHAIN_STREAM = (512, 16384)        # recorded lengths: seed-replay stream, worker stream
CHAIN_TABLE_DIM = 64               # coordinates per sampler in the replay table (>= findMaxDimensions of every case, even)
CHAIN_SCENES = {
    "cornell": lambda: scenes.cornell_box(film=(24, 24), tess=2),
    "glossy": lambda: scenes.glossy_scene(film=(24, 24), subdiv=1, rough_glass=(0.15, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE)),
}
# (name, scene, parameters): every delayed-rejection type x technique, the flags of drmlt_proc.cpp, PSSMLT Kelemen / Gaussian
CHAIN_REPLAY_CASES = [
    ("mira_path", "cornell", dict(integrator="drmlt", type="mira", technique="path", maxDepth=6)),
    ("mira_bdpt", "cornell", dict(integrator="drmlt", type="mira", technique="bdpt", maxDepth=5, directSampling=False)),
    ("mira_mmlt", "cornell", dict(integrator="drmlt", type="mira", technique="mmlt", maxDepth=6)),
    ("green_path", "cornell", dict(integrator="drmlt", type="green", technique="path", maxDepth=6)),
    ("green_bdpt", "cornell", dict(integrator="drmlt", type="green", technique="bdpt", maxDepth=5, directSampling=False)),
    ("green_mmlt", "cornell", dict(integrator="drmlt", type="green", technique="mmlt", maxDepth=6)),
    ("orbital_path", "cornell", dict(integrator="drmlt", type="orbital", technique="path", maxDepth=6)),
    ("orbital_bdpt", "cornell", dict(integrator="drmlt", type="orbital", technique="bdpt", maxDepth=5, directSampling=False)),
    ("orbital_mmlt", "cornell", dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=8)),
    ("orbital_mmlt_glossy", "glossy", dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6)),
    ("mira_path_glossy", "glossy", dict(integrator="drmlt", type="mira", technique="path", maxDepth=5)),
    ("mixture_mmlt", "cornell", dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6, useMixture=True)),
    ("mixture_path", "cornell", dict(integrator="drmlt", type="mira", technique="path", maxDepth=6, useMixture=True)),
    ("timid_orbital_mmlt", "cornell", dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6, timidAfterLarge=True)),
    ("timid_green_mmlt", "cornell", dict(integrator="drmlt", type="green", technique="mmlt", maxDepth=6, timidAfterLarge=True)),
    ("fixemitter_orbital_mmlt", "cornell", dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6, fixEmitterPath=True)),
    ("fixemitter_mira_mmlt", "cornell", dict(integrator="drmlt", type="mira", technique="mmlt", maxDepth=6, fixEmitterPath=True)),
    ("accmap_mira_path", "cornell", dict(integrator="drmlt", type="mira", technique="path", maxDepth=6, acceptanceMap=True, rfilter="box")),
    ("pssmlt_kelemen_path", "cornell", dict(integrator="pssmlt", technique="path", maxDepth=6)),
    ("pssmlt_kelemen_mmlt", "cornell", dict(integrator="pssmlt", technique="mmlt", maxDepth=6)),
    ("pssmlt_kelemen_bdpt", "cornell", dict(integrator="pssmlt", technique="bdpt", maxDepth=5, directSampling=False)),
    ("pssmlt_gaussian_path", "cornell", dict(integrator="pssmlt", technique="path", maxDepth=6, kelemenStyleMutation=False)),
    ("pssmlt_veach_weights_path", "glossy", dict(integrator="pssmlt", technique="path", maxDepth=5, kelemenStyleWeights=False)),
]
CHAIN_B = 0.37                     # m_config.luminance of the replayed work units (PSSMLT's Kelemen weights use it)
PU64 = C.POINTER(C.c_uint64)


def chain_case_config(params):
    return make_config(seed=1, sampleCount=1, directSamples=-1, **params)


def chain_projections(shape):
    """Three fixed weight images: a film is pinned per prefix through its projections on them (a few doubles instead of a frame)."""
    rng = np.random.RandomState(77)
    return rng.rand(3, *shape)


def run_chain_ref(lib, case, pick):
    """One replayed chain of the reference -> dict: seed, the two streams, per-prefix counters and film projections, final film."""
    name, scene_name, params = case
    P = C.POINTER
    cfg = chain_case_config(params)
    data = CHAIN_SCENES[scene_name]()
    d = data.desc()
    W, H = data.film
    K = CHAIN_K
    pssmlt = cfg.integrator == abi.DR_INTEGRATOR_PSSMLT
    nc = 6 if pssmlt else 14
    films, ctr = np.zeros((K + 1, H, W, 3)), np.zeros((K + 1, nc), np.uint64)
    boot, work = np.zeros(CHAIN_STREAM[0]), np.zeros(CHAIN_STREAM[1])
    dep, sidx, slum, nxt = C.c_int32(), C.c_uint64(), D(), D()
    seeds = (1000 + 7 * CHAIN_REPLAY_CASES.index(case) + pick, 2000 + 13 * CHAIN_REPLAY_CASES.index(case) + pick)
    common = [seeds[0], seeds[1], CHAIN_NBOOT, CHAIN_NSEEDS, pick, K, C.byref(dep), C.byref(sidx), C.byref(slum),
              boot.ctypes.data_as(PD), len(boot), work.ctypes.data_as(PD), len(work), films.ctypes.data_as(PD), ctr.ctypes.data_as(PU64), C.byref(nxt)]
    tail = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_int32), PU64, PD, PD, C.c_int, PD, C.c_int, PD, PU64, PD]
    if pssmlt:
        lib.ref_pssmlt_chain.argtypes = [P(abi.dr_scene_desc), P(abi.dr_config), D] + tail
        rc = lib.ref_pssmlt_chain(C.byref(d), C.byref(cfg), CHAIN_B, *common)
    else:
        # timidAfterLarge trips the reference's (always live) assertions at the first rejected large step: the assertion-free
        # twin of the same translation unit replays the arithmetic behind them (oracle/ref/ref_sampler.cpp)
        fn = lib.ref_drmlt_chain_ndebug if cfg.timid_after_large else lib.ref_drmlt_chain
        fn.argtypes = [P(abi.dr_scene_desc), P(abi.dr_config)] + tail
        rc = fn(C.byref(d), C.byref(cfg), *common)
    assert rc == 0, "oracle/_ref: the reference chain failed (%s)" % name
    used = np.nonzero(work == nxt.value)[0]
    assert len(used) == 1, "worker stream too short for %s" % name
    proj = np.einsum("khwc,jhw->kjc", films, chain_projections((H, W)))
    return dict(depth=np.int32(dep.value), sample_index=np.uint64(sidx.value), seed_lum=np.float64(slum.value), used=np.int64(used[0]),
                boot=boot, work=work[:used[0] + 8].copy(), counters=ctr, proj=proj, film=films[K].copy())


def chain_decisions_from_counters(ctr, pssmlt):
    """Per mutation (large step, first accepted, second stage ran, second accepted) from the per-prefix statistics counters
    (value, base pairs: drmlt_proc.cpp:34-49 in order; pssmlt_proc.cpp:33-40)."""
    d = np.diff(ctr.astype(np.int64), axis=0)
    if pssmlt:      # largeStepRatio, smallStepRatio, acceptanceRate
        return np.stack([d[:, 1], d[:, 4], np.zeros_like(d[:, 0]), np.zeros_like(d[:, 0])], axis=1)
    # firstLevel, largeStep, boldStep, secondLevel, secondLevelLarge, secondLevelBold, acceptanceRate
    return np.stack([d[:, 3], d[:, 0], d[:, 7], d[:, 6]], axis=1)


def oracle_chain_fn(lib):
    P = C.POINTER
    lib.orc_chain_stream.restype = C.c_longlong
    lib.orc_chain_stream.argtypes = [C.c_void_p, P(abi.dr_config), D, C.c_int, PD, C.c_longlong, PD, C.c_longlong, C.c_longlong,
                                     P(abi.dr_step_record), PD, P(abi.dr_stats), PD, PD, C.c_int]
    return lib.orc_chain_stream


def chain_table_size(steps=CHAIN_K, dim=CHAIN_TABLE_DIM):
    return 3 * dim + steps * (4 + 12 * dim)


def run_chain_oracle(orc, case, ref, steps=CHAIN_K, table_in=None, want_table=False):
    """The oracle's chain step on the recorded streams of `ref` (or from a replay table) -> dict like run_chain_ref's."""
    name, scene_name, params = case
    cfg = chain_case_config(params)
    cfg.ray_epsilon = cfg.shadow_epsilon = 0.0          # the reference's own epsilons (double build)
    W, H = orc.data.film
    fn = oracle_chain_fn(orc.lib)
    pssmlt = cfg.integrator == abi.DR_INTEGRATOR_PSSMLT
    film, st = np.zeros((H, W, 3)), abi.dr_stats()
    rec = (abi.dr_step_record * max(steps, 1))()
    table = np.zeros(chain_table_size(steps)) if (want_table and table_in is None) else None
    boot, work = np.ascontiguousarray(ref["boot"]), np.ascontiguousarray(ref["work"])
    used = fn(orc.h, C.byref(cfg), CHAIN_B if pssmlt else 1.0, int(ref["depth"]),
              boot.ctypes.data_as(PD) if table_in is None else None, len(boot), work.ctypes.data_as(PD) if table_in is None else None, len(work),
              steps, rec, film.ctypes.data_as(PD), C.byref(st), table.ctypes.data_as(PD) if table is not None else None,
              table_in.ctypes.data_as(PD) if table_in is not None else None, CHAIN_TABLE_DIM)
    r = np.frombuffer(rec, dtype=np.uint8).reshape(-1, C.sizeof(abi.dr_step_record))[:steps]
    dec = r[:, 20:24].astype(np.int64)                   # large_step, accept1, did_second, accept2
    return dict(used=used, film=film, decisions=dec, stats=st, table=table, records=rec,
                L=r[:, 0:20].copy().view("<f4").reshape(-1, 5))


def chain_key(case, pick):
    return "chain_%s_%d" % (case[0], pick)


def run_chain_ref_all(lib):
    """Every replayed chain of the reference -> flat dict of arrays (tests/golden/ref_chain.npz)."""
    out = {}
    for case in CHAIN_REPLAY_CASES:
        for pick in CHAIN_PICKS:
            r = run_chain_ref(lib, case, pick)
            k = chain_key(case, pick)
            out[k + "_seed"] = np.array([float(r["depth"]), float(r["sample_index"]), float(r["seed_lum"]), float(r["used"])])
            out[k + "_boot"], out[k + "_work"] = r["boot"], r["work"]
            out[k + "_counters"], out[k + "_proj"], out[k + "_film"] = r["counters"], r["proj"], r["film"]
    return out


def chain_from_golden(gold, case, pick):
    k = chain_key(case, pick)
    seed = gold[k + "_seed"]
    return dict(depth=np.int32(seed[0]), sample_index=np.uint64(seed[1]), seed_lum=seed[2], used=np.int64(seed[3]),
                boot=gold[k + "_boot"], work=gold[k + "_work"], counters=gold[k + "_counters"], proj=gold[k + "_proj"], film=gold[k + "_film"])


_chain_orcs = {}


def chain_oracle_scene(scene_name):
    import oracle_lib
    if scene_name not in _chain_orcs:
        _chain_orcs[scene_name] = oracle_lib.OracleScene(CHAIN_SCENES[scene_name]())
    return _chain_orcs[scene_name]


def check_chain_oracle_vs_ref(case, ref):
    """The oracle's chain step on the recorded streams against the reference's chain: the number of uniforms consumed, every
    decision (from the reference's per-prefix statistics counters), and the work unit's film after EVERY mutation (projections;
    pins the splat weights a1, (1 - a1) a2 and the splat positions of each step), plus the final film pixel by pixel."""
    orc = chain_oracle_scene(case[1])
    pssmlt = case[2]["integrator"] == "pssmlt"
    o = run_chain_oracle(orc, case, ref, want_table=True)
    assert o["used"] == ref["used"], (case[0], "uniforms consumed", o["used"], int(ref["used"]))
    want = chain_decisions_from_counters(ref["counters"], pssmlt)
    got = o["decisions"]
    if pssmlt:
        got = np.stack([got[:, 0], got[:, 1], 0 * got[:, 0], 0 * got[:, 0]], axis=1)
    assert np.array_equal(want, got), (case[0], "decisions", np.nonzero((want != got).any(axis=1))[0][:4])
    scale = max(np.abs(ref["film"]).max(), 1e-30)
    assert np.abs(o["film"] - ref["film"]).max() <= 1e-12 * scale, (case[0], "final film")
    W, H = orc.data.film
    w = chain_projections((H, W))
    for k in range(CHAIN_K + 1):
        fk = run_chain_oracle(orc, case, ref, steps=k)["film"]
        pk = np.einsum("hwc,jhw->jc", fk, w)
        assert np.abs(pk - ref["proj"][k]).max() <= 1e-11 * max(np.abs(ref["proj"][k]).max(), 1e-30), (case[0], "film after mutation", k)
    # the same chain from the replay table (keyed address space) the CUDA path is fed with
    t = run_chain_oracle(orc, case, ref, table_in=o["table"])
    assert np.array_equal(t["decisions"], o["decisions"]) and np.abs(t["film"] - o["film"]).max() <= 1e-13 * scale, (case[0], "replay table")
    return o


# ---------------------------------------------------------------- the direct pass and the importance-map resampling
GOLDEN_DIRECT = os.path.join(ROOT, "tests", "golden", "ref_direct.npz")
DIRECT_SCENES = {"cornell": lambda: scenes.cornell_box(film=(64, 64), tess=4), "glossy": lambda: scenes.glossy_scene(film=(64, 64), subdiv=2)}
DIRECT_SAMPLES = 64
RESAMPLE_SHAPES = [((20, 12), (160, 90)), ((64, 48), (16, 12)), ((33, 17), (33, 40)), ((8, 8), (128, 8)), ((5, 7), (5, 7)), ((1, 1), (16, 16)), ((20, 11), (320, 180))]


def resample_input(shape):
    (w, h), (W, H) = shape
    rng = np.random.RandomState(w * 131 + H)
    return (rng.rand(h, w) ** 3) * 4.0


def run_direct_ref(lib):
    """BidirectionalUtils::renderDirectComponent of the reference (ref_direct_image) on the test scenes, and Bitmap::resample
    (ref_resample_luminance) on the test maps -> flat dict of arrays (tests/golden/ref_direct.npz)."""
    P = C.POINTER
    lib.ref_direct_image.argtypes = [P(abi.dr_scene_desc), C.c_int, C.c_int, C.c_int, PF32]
    lib.ref_resample_luminance.argtypes = [PD, C.c_int, C.c_int, C.c_int, C.c_int, PD]
    out = {}
    for name, make in DIRECT_SCENES.items():
        data = make()
        d = data.desc()
        W, H = data.film
        img = np.zeros((H, W, 3), np.float32)
        assert lib.ref_direct_image(C.byref(d), abi.DR_FILTER_GAUSSIAN, DIRECT_SAMPLES, 4, img.ctypes.data_as(PF32)) == 0
        out["direct_" + name] = img
    for i, shape in enumerate(RESAMPLE_SHAPES):
        (w, h), (W, H) = shape
        src, dst = np.ascontiguousarray(resample_input(shape)), np.zeros((H, W))
        assert lib.ref_resample_luminance(src.ctypes.data_as(PD), w, h, W, H, dst.ctypes.data_as(PD)) == 0
        out["resample_%d" % i] = dst
    return out


def run_resample_oracle(lib):
    lib.orc_resample_map_f64.argtypes = [PD, C.c_int, C.c_int, C.c_int, C.c_int, PD]
    out = {}
    for i, shape in enumerate(RESAMPLE_SHAPES):
        (w, h), (W, H) = shape
        src, dst = np.ascontiguousarray(resample_input(shape)), np.zeros((H, W))
        lib.orc_resample_map_f64(src.ctypes.data_as(PD), w, h, W, H, dst.ctypes.data_as(PD))
        out["resample_%d" % i] = dst
    return out


def check_direct_image(img, ref, what):
    """The separate direct-illumination image against the reference's (util.cpp:30-94), statistically: the reference draws its
    pixel and shading samples from an ldsampler scrambled from /dev/urandom, the product from keyed uniforms.  The pixels that see
    the area light (only 8 PIXEL samples land there, whatever directSamples: util.cpp:44-54) carry most of the image mean and all
    of its noise; they are compared loosely, everything else within 1 % in the mean and a small relMSE."""
    from scipy.ndimage import binary_dilation
    a, b = luminance(np.asarray(img, np.float64)), luminance(np.asarray(ref, np.float64))
    lit = binary_dilation(b > 1.0, iterations=3)
    m = ~lit
    assert m.sum() > 0.8 * m.size
    assert abs(a[m].mean() / b[m].mean() - 1.0) < 0.01, (what, a[m].mean(), b[m].mean())
    err = float(np.mean((a[m] - b[m]) ** 2 / (b[m] ** 2 + 1e-2)))
    assert err < 1e-2, (what, err)     # both images carry their own sampling noise (measured 1e-3 .. 5e-3)
    assert abs(a[lit].mean() / b[lit].mean() - 1.0) < 0.08, (what, a[lit].mean(), b[lit].mean())
    return err


# ---------------------------------------------------------------- two-stage MLT, whole job
GOLDEN_TWOSTAGE = os.path.join(ROOT, "tests", "golden", "ref_twostage.npz")
TWOSTAGE_PARAMS = dict(integrator="drmlt", type="orbital", technique="mmlt", maxDepth=6, directSamples=-1, twoStage=True, firstStageSizeReduction=4)
TWOSTAGE_SPP = 512
