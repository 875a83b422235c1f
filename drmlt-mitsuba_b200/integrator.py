"""Python-side handles over the C ABI (include/drmlt_b200.h) used by tests and bench.py.

`make_config` takes the reference's own parameter names (the `-D key=value` strings of
`mitsuba scene.xml -D integrator=drmlt -D technique=mmlt ...`, drmlt.cpp:178-351) and goes through
`dr_config_set` / `dr_config_validate`, so unknown keys and illegal combinations fail exactly where
the reference's constructor logs EError.  Everything here is plumbing: all computation happens in
csrc/libdrmlt_b200.so on the GPU; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import abi


def _fp(a, ty=C.c_float):
    return a.ctypes.data_as(C.POINTER(ty))


def _fmt(v):
    if isinstance(v, bool):
        return b"true" if v else b"false"
    if isinstance(v, float):
        return repr(float(v)).encode()
    return str(v).encode()


def set_importance_map(cfg, importance):
    """Attach a (H, W) float32 importance map (m_config.importanceMap) to a configuration; None detaches it.
    The array is kept alive by the configuration object."""
    if importance is None:
        cfg.importance_map = None
        cfg._importance_keep = None
        return cfg
    arr = np.ascontiguousarray(importance, np.float32)
    cfg._importance_keep = arr
    cfg.importance_map = _fp(arr)
    return cfg


def make_config(**params):
    """dr_config from reference-style parameters, e.g. make_config(integrator="drmlt",
    technique="mmlt", type="orbital", maxDepth=8, sigma=1/64, acceptanceMap=True)."""
    lib = abi.load_library()
    cfg = abi.dr_config()
    lib.dr_config_default(C.byref(cfg))
    for k, v in params.items():
        abi.check(lib, lib.dr_config_set(C.byref(cfg), k.encode(), _fmt(v)))
    abi.check(lib, lib.dr_config_validate(C.byref(cfg)))
    return cfg


class Scene:
    """GPU-resident flattened scene (dr_scene)."""

    def __init__(self, data, device=0, gpu_bvh=None):
        """gpu_bvh: True = LBVH built on the device (dr_scene_create_ex, DR_SCENE_BVH_GPU), False = the host's binned-SAH
        build, None = dr_scene_create's default (host)."""
        self.lib = abi.load_library()
        self.data = data
        self.device = device
        desc = data.desc()
        h = C.c_void_p()
        if gpu_bvh is None:
            abi.check(self.lib, self.lib.dr_scene_create(C.byref(desc), device, C.byref(h)))
        else:
            abi.check(self.lib, self.lib.dr_scene_create_ex(C.byref(desc), device, 1 if gpu_bvh else 0, C.byref(h)))
        self.h = h

    def bvh_info(self):
        """dict(builder="gpu"|"host", nodes, stack_bound, build_ms) of dr_scene_bvh_info."""
        b, n, d, ms = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        self.lib.dr_scene_bvh_info(self.h, C.byref(b), C.byref(n), C.byref(d), C.byref(ms))
        return dict(builder="gpu" if b.value == 1 else "host", nodes=n.value, stack_bound=d.value, build_ms=ms.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dr_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def film(self):
        return self.data.film

    def film_size(self, cfg):
        """(W, H) of the image a job with this configuration renders: the crop window of the (possibly overridden) film."""
        w, h = C.c_int32(0), C.c_int32(0)
        abi.check(self.lib, self.lib.dr_film_size(self.h, C.byref(cfg), C.byref(w), C.byref(h)))
        return w.value, h.value

    # ---- two-stage MLT (BidirectionalUtils::mltLuminancePass, src/libbidir/util.cpp:96-199)
    def first_stage_config(self, cfg):
        nested = abi.dr_config()
        abi.check(self.lib, self.lib.dr_first_stage_config(self.h, C.byref(cfg), C.byref(nested)))
        return nested

    def resample_luminance(self, image_rgb, size):
        """Developed first-stage image (h, w, 3) -> importance map (H, W) for size = (W, H)."""
        img = np.ascontiguousarray(image_rgb, np.float32)
        h, w = img.shape[:2]
        W, H = size
        out = np.zeros((H, W), np.float32)
        abi.check(self.lib, self.lib.dr_resample_luminance(self.h, _fp(img), w, h, W, H, _fp(out)))
        return out

    def importance_map(self, cfg):
        W, H = self.film_size(cfg)
        out = np.zeros((H, W), np.float32)
        st = abi.dr_stats()
        abi.check(self.lib, self.lib.dr_importance_map(self.h, C.byref(cfg), _fp(out), C.byref(st)))
        return out, st

    def texture_eval(self, texture, uv):
        """Texture `texture` of the scene at intersection uv pairs [n, 2] (double) -> RGB [n, 3] (double), on the device."""
        uv = np.ascontiguousarray(uv, np.float64)
        rgb = np.zeros((len(uv), 3), np.float64)
        abi.check(self.lib, self.lib.dr_texture_eval(self.h, texture, _fp(uv, C.c_double), len(uv), _fp(rgb, C.c_double)))
        return rgb

    def reupload(self):
        n = C.c_int64(0)
        abi.check(self.lib, self.lib.dr_scene_reupload(self.h, C.byref(n)))
        return n.value

    def clone(self, device):
        """A replica of this scene on another GPU (dr_scene_clone: no second BVH build)."""
        other = Scene.__new__(Scene)
        other.lib, other.data, other.device = self.lib, self.data, device
        h = C.c_void_p()
        abi.check(self.lib, self.lib.dr_scene_clone(self.h, device, C.byref(h)))
        other.h = h
        return other

    def render_multi(self, cfg, replicas):
        """dr_render_multi: one job on this scene's GPU and on the GPUs of `replicas` (Scene.clone), NCCL inside the library."""
        W, H = self.film_size(cfg)
        img = np.zeros((H, W, 3), np.float32)
        st = abi.dr_stats()
        scenes_ = [self] + list(replicas)
        arr = (C.c_void_p * len(scenes_))(*[s.h for s in scenes_])
        abi.check(self.lib, self.lib.dr_render_multi(arr, len(scenes_), C.byref(cfg), _fp(img), C.byref(st)))
        return img, st

    # ---- whole job, host buffers (DRMLT::render / PSSMLT::render)
    def render(self, cfg, out=None):
        W, H = self.film_size(cfg)
        img = out if out is not None else np.zeros((H, W, 3), np.float32)
        st = abi.dr_stats()
        abi.check(self.lib, self.lib.dr_render(self.h, C.byref(cfg), _fp(img), C.byref(st)))
        return img, st

    def render_progressive(self, cfg, refresh_seconds, on_refresh):
        """dr_render with a develop of the partial image about every `refresh_seconds`:
        on_refresh(image (H, W, 3) view, seconds, stats) -> truthy cancels the job."""
        W, H = self.film_size(cfg)
        img = np.zeros((H, W, 3), np.float32)
        st = abi.dr_stats()

        def _cb(ptr, w, h, seconds, stats, user):
            view = np.ctypeslib.as_array(ptr, shape=(h, w, 3))
            return 1 if on_refresh(view, seconds, stats.contents) else 0

        cb = abi.dr_refresh_fn(_cb)
        abi.check(self.lib, self.lib.dr_render_progressive(self.h, C.byref(cfg), _fp(img), C.byref(st), float(refresh_seconds), cb, None))
        return img, st

    # ---- replay entry points
    def trace(self, rays, shadow=False):
        n = len(rays)
        hits = (abi.dr_hit * n)()
        abi.check(self.lib, self.lib.dr_trace_rays(self.h, rays, n, int(shadow), hits))
        return hits

    def eval_paths(self, cfg, us, ue, ud, depth):
        us, ue, ud = [np.ascontiguousarray(x, np.float32) for x in (us, ue, ud)]
        n = us.shape[0]
        depth = np.ascontiguousarray(depth, np.int32)
        out = (abi.dr_path_result * n)()
        abi.check(self.lib, self.lib.dr_eval_paths(self.h, C.byref(cfg), _fp(us), us.shape[1], _fp(ue), ue.shape[1],
                                                   _fp(ud), ud.shape[1], _fp(depth, C.c_int32), n, out))
        return out

    def bootstrap_luminance(self, cfg, first, n):
        lum = np.zeros(n, np.float32)
        dep = np.zeros(n, np.int32)
        abi.check(self.lib, self.lib.dr_bootstrap_luminance(self.h, C.byref(cfg), first, n, _fp(lum), _fp(dep, C.c_int32)))
        return lum, dep

    def direct_image(self, cfg, want_li=False):
        """The separate direct-illumination image (directSamples > 0); optionally the radiance of every pixel sample."""
        W, H = self.film_size(cfg)
        ps = max(int(cfg.direct_samples), 1)
        while ps > 8:
            ps //= 2
        img = np.zeros((H, W, 3), np.float32)
        li = np.zeros((H, W, ps, 3), np.float64) if want_li else None
        abi.check(self.lib, self.lib.dr_direct_image(self.h, C.byref(cfg), _fp(img), _fp(li, C.c_double) if want_li else None))
        return (img, li) if want_li else img

    def chain_steps(self, cfg, b, seed_index, depth, chain_id, steps, want_film=False):
        n = len(seed_index)
        seed_index = np.ascontiguousarray(seed_index, np.uint64)
        chain_id = np.ascontiguousarray(chain_id, np.uint64)
        depth = np.ascontiguousarray(depth, np.int32)
        rec = (abi.dr_step_record * (n * steps))()
        W, H = self.film_size(cfg)
        film = np.zeros((H, W, 3), np.float32) if want_film else None
        abi.check(self.lib, self.lib.dr_chain_steps(self.h, C.byref(cfg), b, _fp(seed_index, C.c_uint64), _fp(depth, C.c_int32),
                                                    _fp(chain_id, C.c_uint64), n, steps, rec, _fp(film) if want_film else None))
        return (rec, film) if want_film else rec


    def chain_replay(self, cfg, b, depth, tables, dim, steps, want_film=False):
        """Chains of the reference replayed from tables of keyed uniforms (dr_chain_replay): tables is [n_chains][3 dim + steps (4 + 12 dim)] float64."""
        tables = np.ascontiguousarray(tables, np.float64)
        n = tables.shape[0]
        assert tables.shape[1] == 3 * dim + steps * (4 + 12 * dim)
        depth = np.ascontiguousarray(depth, np.int32)
        rec = (abi.dr_step_record * (n * max(steps, 1)))()
        W, H = self.film_size(cfg)
        film = np.zeros((H, W, 3), np.float32) if want_film else None
        abi.check(self.lib, self.lib.dr_chain_replay(self.h, C.byref(cfg), b, _fp(depth, C.c_int32), n, steps, _fp(tables, C.c_double), dim,
                                                     rec, _fp(film) if want_film else None))
        return (rec, film) if want_film else rec


class Job:
    """One rank's share of a render (dr_job): bootstrap -> [all-reduce] -> seed -> run* -> develop."""

    def __init__(self, scene, cfg):
        self.lib = scene.lib
        self.scene = scene
        self.cfg = cfg
        h = C.c_void_p()
        abi.check(self.lib, self.lib.dr_job_create(scene.h, C.byref(cfg), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.dr_job_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bootstrap(self):
        s, c = C.c_double(0), C.c_double(0)
        abi.check(self.lib, self.lib.dr_job_bootstrap(self.h, C.byref(s), C.byref(c)))
        return s.value, c.value

    def normalization(self, lum_sum, count):
        """b = mean luminance (x maxDepth for MMLT): pathsampler.cpp:922-934."""
        b = lum_sum / count if count > 0 else 0.0
        if self.cfg.technique == abi.DR_TECH_MMLT:
            b *= self.cfg.max_depth
        return b

    def seed_chains(self, b):
        abi.check(self.lib, self.lib.dr_job_seed_chains(self.h, b))

    def run(self, mutations_per_chain):
        abi.check(self.lib, self.lib.dr_job_run(self.h, int(mutations_per_chain)))

    def direct(self):
        """Render the separate direct-illumination image (no-op unless directSamples > 0); develop() adds it."""
        abi.check(self.lib, self.lib.dr_job_direct(self.h))

    def profile(self, on=True):
        """Stage profiling: CUDA events around every stage of every round; sums appear in stats()."""
        self.lib.dr_job_profile(self.h, 1 if on else 0)

    def film_device(self):
        p, n = C.c_void_p(), C.c_int64(0)
        abi.check(self.lib, self.lib.dr_job_film_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def develop(self):
        W, H = self.scene.film_size(self.cfg)
        img = np.zeros((H, W, 3), np.float32)
        abi.check(self.lib, self.lib.dr_job_develop(self.h, _fp(img)))
        return img

    def stats(self):
        st = abi.dr_stats()
        abi.check(self.lib, self.lib.dr_job_stats(self.h, C.byref(st)))
        return st

    @property
    def num_chains(self):
        return self.lib.dr_job_num_chains(self.h)

    @property
    def total_mutations(self):
        return self.lib.dr_job_total_mutations(self.h)


class DeviceFilm:
    """Zero-copy view of a job's accumulation film for torch (`torch.as_tensor(DeviceFilm(job), device=...)`):
    the buffer a multi-GPU driver hands to the NCCL reduce."""

    def __init__(self, job):
        ptr, n = job.film_device()
        self._keep = job
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
