"""ctypes loader of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from drmlt_mitsuba_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_LIB):
        build()
    lib = C.CDLL(ORACLE_LIB)
    P = C.POINTER
    lib.orc_scene_create.argtypes = [P(abi.dr_scene_desc)]
    lib.orc_scene_create.restype = C.c_void_p
    lib.orc_scene_destroy.argtypes = [C.c_void_p]
    lib.orc_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_uint32)]
    lib.orc_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
    lib.orc_uniform.restype = C.c_float
    lib.orc_max_dimensions.argtypes = [P(abi.dr_config), C.c_int, P(C.c_int), P(C.c_int), P(C.c_int)]
    lib.orc_trace_rays.argtypes = [C.c_void_p, P(abi.dr_ray), C.c_int64, C.c_int, C.c_float, P(abi.dr_hit)]
    lib.orc_eval_paths.argtypes = [C.c_void_p, P(abi.dr_config), P(C.c_float), C.c_int, P(C.c_float), C.c_int,
                                   P(C.c_float), C.c_int, P(C.c_int32), C.c_int64, P(abi.dr_path_result), P(C.c_double)]
    lib.orc_bootstrap_luminance.argtypes = [C.c_void_p, P(abi.dr_config), C.c_uint64, C.c_int64, P(C.c_float),
                                            P(C.c_int32), P(C.c_double)]
    lib.orc_chain_steps.argtypes = [C.c_void_p, P(abi.dr_config), C.c_double, P(C.c_uint64), P(C.c_int32), P(C.c_uint64),
                                    C.c_int64, C.c_int64, P(abi.dr_step_record), P(C.c_float), P(abi.dr_stats), C.c_int]
    lib.orc_render.argtypes = [C.c_void_p, P(abi.dr_config), C.c_int64, C.c_int64, C.c_int64, C.c_int,
                               P(C.c_float), P(abi.dr_stats), P(C.c_double)]
    lib.orc_direct_image.argtypes = [C.c_void_p, P(abi.dr_config), P(C.c_float), P(C.c_double)]
    lib.orc_first_stage_config.argtypes = [C.c_void_p, P(abi.dr_config), P(abi.dr_config)]
    lib.orc_first_stage_config.restype = None
    lib.orc_resample_luminance.argtypes = [P(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_float)]
    lib.orc_resample_luminance.restype = None
    lib.orc_develop.argtypes = [P(C.c_float), C.c_int, C.c_int, C.c_double, C.c_int, P(C.c_float), P(C.c_float)]
    lib.orc_develop.restype = None
    lib.orc_splat.argtypes = [C.c_int, C.c_int, C.c_int, P(C.c_float), P(C.c_float), C.c_int64, P(C.c_float)]
    lib.orc_bsdf_sample.argtypes = [P(abi.dr_material), P(C.c_double), C.c_int, C.c_double, C.c_double,
                                    P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_int)]
    lib.orc_bsdf_sample3.argtypes = [P(abi.dr_material), P(C.c_double), C.c_int, C.c_double, C.c_double, C.c_double,
                                     P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_int), P(C.c_double)]
    lib.orc_bsdf_eval.argtypes = [P(abi.dr_material), P(C.c_double), P(C.c_double), C.c_int, C.c_int,
                                  P(C.c_double), P(C.c_double)]
    for name in ("orc_kelemen_sample", "orc_kelemen_pdf", "orc_gaussian_sample"):
        getattr(lib, name).argtypes = [C.c_double] * 3
        getattr(lib, name).restype = C.c_double
    lib.orc_cauchy_sample.argtypes = [C.c_double, C.c_double]
    lib.orc_cauchy_sample.restype = C.c_double
    lib.orc_wrap.argtypes = [C.c_double]
    lib.orc_wrap.restype = C.c_double
    _lib = lib
    return lib


def fptr(a, ty=C.c_float):
    return a.ctypes.data_as(C.POINTER(ty))


class OracleScene:
    def __init__(self, scene_data):
        self.lib = load()
        self.data = scene_data
        d = scene_data.desc()
        self.h = self.lib.orc_scene_create(C.byref(d))

    def __del__(self):
        try:
            self.lib.orc_scene_destroy(self.h)
        except Exception:
            pass

    def film_size(self, cfg):
        """Crop window of the (possibly overridden) film: Film::Film, src/librender/film.cpp:30-48."""
        fw = cfg.film_width if cfg.film_width > 0 else self.data.film[0]
        fh = cfg.film_height if cfg.film_height > 0 else self.data.film[1]
        return (cfg.crop_width if cfg.crop_width > 0 else fw, cfg.crop_height if cfg.crop_height > 0 else fh)

    def first_stage_config(self, cfg):
        nested = abi.dr_config()
        self.lib.orc_first_stage_config(self.h, C.byref(cfg), C.byref(nested))
        return nested

    def bootstrap(self, cfg, first, n):
        lum = np.zeros(n, np.float32)
        dep = np.zeros(n, np.int32)
        lum64 = np.zeros(n, np.float64)
        self.lib.orc_bootstrap_luminance(self.h, C.byref(cfg), first, n, fptr(lum), fptr(dep, C.c_int32), fptr(lum64, C.c_double))
        return lum64, dep

    def eval_paths(self, cfg, us, ue, ud, depth):
        n = us.shape[0]
        out = (abi.dr_path_result * n)()
        lum = np.zeros(n, np.float64)
        us, ue, ud = [np.ascontiguousarray(x, np.float32) for x in (us, ue, ud)]
        depth = np.ascontiguousarray(depth, np.int32)
        self.lib.orc_eval_paths(self.h, C.byref(cfg), fptr(us), us.shape[1], fptr(ue), ue.shape[1], fptr(ud), ud.shape[1],
                                fptr(depth, C.c_int32), n, out, fptr(lum, C.c_double))
        return out, lum

    def chain_steps(self, cfg, b, seed_index, depth, chain_id, steps, want_film=False, threads=0, want_records=True):
        n = len(seed_index)
        seed_index = np.ascontiguousarray(seed_index, np.uint64)
        chain_id = np.ascontiguousarray(chain_id, np.uint64)
        depth = np.ascontiguousarray(depth, np.int32)
        rec = (abi.dr_step_record * (n * steps))() if want_records else None
        W, H = self.film_size(cfg)
        film = np.zeros((H, W, 3), np.float32) if want_film else None
        st = abi.dr_stats()
        self.lib.orc_chain_steps(self.h, C.byref(cfg), b, fptr(seed_index, C.c_uint64), fptr(depth, C.c_int32),
                                 fptr(chain_id, C.c_uint64), n, steps, rec, fptr(film) if want_film else None,
                                 C.byref(st), threads)
        return rec, film, st

    def direct_image(self, cfg, want_li=False):
        W, H = self.film_size(cfg)
        ps = max(int(cfg.direct_samples), 1)
        while ps > 8:
            ps //= 2
        img = np.zeros((H, W, 3), np.float32)
        li = np.zeros((H, W, ps, 3), np.float64) if want_li else None
        self.lib.orc_direct_image(self.h, C.byref(cfg), fptr(img), fptr(li, C.c_double) if want_li else None)
        return (img, li) if want_li else img

    def render(self, cfg, n_boot, n_chains, steps, threads=0):
        W, H = self.film_size(cfg)
        img = np.zeros((H, W, 3), np.float32)
        st = abi.dr_stats()
        sec = C.c_double(0)
        r = self.lib.orc_render(self.h, C.byref(cfg), n_boot, n_chains, steps, threads, fptr(img), C.byref(st), C.byref(sec))
        return r, img, st, sec.value

    def trace(self, rays, shadow=False, eps=0.0):
        n = len(rays)
        hits = (abi.dr_hit * n)()
        self.lib.orc_trace_rays(self.h, rays, n, int(shadow), eps, hits)
        return hits


def resample_luminance(image_rgb, size):
    """Oracle of dr_resample_luminance: (h, w, 3) developed first-stage image -> (H, W) importance map."""
    lib = load()
    img = np.ascontiguousarray(image_rgb, np.float32)
    h, w = img.shape[:2]
    W, H = size
    out = np.zeros((H, W), np.float32)
    lib.orc_resample_luminance(fptr(img), w, h, W, H, fptr(out))
    return out


def develop(film_rgb, b, acceptance_map=False, importance=None):
    lib = load()
    film = np.ascontiguousarray(film_rgb, np.float32)
    H, W = film.shape[:2]
    out = np.zeros((H, W, 3), np.float32)
    imp = np.ascontiguousarray(importance, np.float32) if importance is not None else None
    lib.orc_develop(fptr(film), W, H, b, int(acceptance_map), fptr(imp) if imp is not None else None, fptr(out))
    return out


def default_config(**kw):
    """dr_config_default restated in Python for oracle-only tests (defaults: drmlt.cpp:193-349)."""
    c = abi.dr_config()
    c.integrator, c.technique, c.type = abi.DR_INTEGRATOR_DRMLT, abi.DR_TECH_PATH, abi.DR_TYPE_MIRA
    c.max_depth, c.rr_depth = -1, 5
    c.direct_sampling, c.direct_samples, c.luminance_samples = 1, 16, 100000
    c.p_large, c.work_units, c.kelemen_style_weights = 0.3, -1, 1
    c.two_stage, c.timeout, c.average_luminance, c.light_image = 0, 0, -1.0, 1
    c.acceptance_map = c.timid_after_large = c.fix_emitter_path = c.use_mixture = 0
    c.sigma, c.scale_second = 1.0 / 64.0, 0.1
    c.kelemen_style_mutation, c.mutation_size_low, c.mutation_size_high = 1, 1.0 / 1024.0, 1.0 / 64.0
    c.sample_count, c.rfilter = 64, abi.DR_FILTER_GAUSSIAN
    c.n_chains, c.seed, c.rank, c.world_size = 0, 1234, 0, 1
    c.ray_epsilon = c.shadow_epsilon = 0.0
    c.first_stage, c.first_stage_size_reduction = 0, 16
    for k, v in kw.items():
        setattr(c, k, v)
    return c
