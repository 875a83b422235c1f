"""ctypes mirror of include/drmlt_b200.h and loader of the CUDA library.

The product path is the shared library `csrc/libdrmlt_b200.so` (hand-written sm_100a kernels
behind a C ABI).  There is no CPU fallback: if the library is missing, or no CUDA device is
visible, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRMLT_B200_LIB: another build of the same library (tools/build_variant.sh, A/B tuning runs)
LIB_PATH = os.environ.get("DRMLT_B200_LIB") or os.path.join(_HERE, "csrc", "libdrmlt_b200.so")

DR_OK = 0
DR_BSDF_DIFFUSE, DR_BSDF_DIELECTRIC, DR_BSDF_CONDUCTOR, DR_BSDF_ROUGHCONDUCTOR, DR_BSDF_ROUGHDIELECTRIC, DR_BSDF_PLASTIC, DR_BSDF_ROUGHPLASTIC = 0, 1, 2, 3, 4, 5, 6
DR_ROUGH_TABLE_THETA, DR_ROUGH_TABLE_DOUBLES = 100, 104
DR_MAT_TWOSIDED, DR_MAT_GGX, DR_MAT_SAMPLE_VISIBLE, DR_MAT_NONLINEAR = 1, 2, 4, 8
DR_TRI_SMOOTH, DR_TRI_UV_TANGENTS, DR_TRI_NO_TEXCOORDS = 1, 2, 4
DR_WRAP_REPEAT, DR_WRAP_CLAMP, DR_WRAP_MIRROR, DR_WRAP_ZERO, DR_WRAP_ONE = range(5)
DR_MAX_TEXTURES = 4095


def DR_MAT_TEX_REFLECTANCE(i):
    return ((i + 1) & 0xfff) << 8


def DR_MAT_TEX_TRANSMITTANCE(i):
    return ((i + 1) & 0xfff) << 20
DR_INTEGRATOR_PSSMLT, DR_INTEGRATOR_DRMLT = 0, 1
DR_TECH_PATH, DR_TECH_BDPT, DR_TECH_MMLT = 0, 1, 2
DR_TYPE_GREEN, DR_TYPE_MIRA, DR_TYPE_ORBITAL = 0, 1, 2
DR_FILTER_GAUSSIAN, DR_FILTER_BOX, DR_FILTER_TABLE, DR_FILTER_TENT, DR_FILTER_MITCHELL, DR_FILTER_CATMULLROM, DR_FILTER_LANCZOS = range(7)
DR_MAX_SPLATS = 12


class dr_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("flags", C.c_uint32),
                ("reflectance", C.c_float * 3), ("transmittance", C.c_float * 3),
                ("eta", C.c_float * 3), ("k", C.c_float * 3),
                ("alpha", C.c_float), ("table", C.c_uint32)]


class dr_texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("texels", C.POINTER(C.c_float)),
                ("wrap_u", C.c_uint32), ("wrap_v", C.c_uint32), ("nearest", C.c_uint32), ("pad", C.c_uint32),
                ("uv_scale", C.c_double * 2), ("uv_offset", C.c_double * 2)]


class dr_emitter(C.Structure):
    _fields_ = [("first_tri", C.c_uint32), ("n_tris", C.c_uint32),
                ("radiance", C.c_float * 3), ("sampling_weight", C.c_float)]


class dr_camera(C.Structure):
    _fields_ = [("to_world", C.c_float * 16), ("xfov_deg", C.c_float),
                ("near_clip", C.c_float), ("far_clip", C.c_float),
                ("film_width", C.c_int32), ("film_height", C.c_int32)]


class dr_scene_desc(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_triangles", C.c_uint32),
                ("n_materials", C.c_uint32), ("n_emitters", C.c_uint32),
                ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("indices", C.POINTER(C.c_uint32)), ("tri_material", C.POINTER(C.c_uint32)),
                ("tri_emitter", C.POINTER(C.c_int32)), ("tri_flags", C.POINTER(C.c_uint32)),
                ("materials", C.POINTER(dr_material)), ("emitters", C.POINTER(dr_emitter)),
                ("camera", dr_camera), ("rough_tables", C.POINTER(C.c_double)), ("n_rough_tables", C.c_uint32),
                ("n_textures", C.c_uint32), ("texcoords", C.POINTER(C.c_float)), ("textures", C.POINTER(dr_texture))]


class dr_config(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("technique", C.c_int32), ("type", C.c_int32),
                ("max_depth", C.c_int32), ("rr_depth", C.c_int32), ("direct_sampling", C.c_int32),
                ("direct_samples", C.c_int32), ("luminance_samples", C.c_int32), ("p_large", C.c_float),
                ("work_units", C.c_int32), ("kelemen_style_weights", C.c_int32), ("two_stage", C.c_int32),
                ("timeout", C.c_int32), ("average_luminance", C.c_float), ("light_image", C.c_int32),
                ("acceptance_map", C.c_int32), ("timid_after_large", C.c_int32), ("fix_emitter_path", C.c_int32),
                ("use_mixture", C.c_int32), ("sigma", C.c_float), ("scale_second", C.c_float),
                ("kelemen_style_mutation", C.c_int32), ("mutation_size_low", C.c_float),
                ("mutation_size_high", C.c_float), ("sample_count", C.c_int32), ("rfilter", C.c_int32),
                ("n_chains", C.c_int32), ("seed", C.c_uint64), ("rank", C.c_int32), ("world_size", C.c_int32),
                ("ray_epsilon", C.c_float), ("shadow_epsilon", C.c_float),
                ("first_stage", C.c_int32), ("first_stage_size_reduction", C.c_int32),
                ("film_width", C.c_int32), ("film_height", C.c_int32),
                ("crop_offset_x", C.c_int32), ("crop_offset_y", C.c_int32),
                ("crop_width", C.c_int32), ("crop_height", C.c_int32), ("n_lanes", C.c_int32), ("depth_balance", C.c_int32),
                ("importance_map", C.POINTER(C.c_float)), ("filter_radius", C.c_double), ("filter_table", C.c_double * 32)]


class dr_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "mutations", "first_accept", "first_base", "large_accept", "large_base", "bold_accept", "bold_base",
        "second_accept", "second_base", "second_large_accept", "second_large_base",
        "second_bold_accept", "second_bold_base", "accept", "accept_base", "paths", "rays",
        "bootstrap_paths", "bootstrap_rays")] + [
        ("luminance", C.c_double), ("bootstrap_ms", C.c_double), ("chains_ms", C.c_double),
        ("total_ms", C.c_double), ("kernel_launches", C.c_uint64), ("rounds", C.c_uint64),
        ("trace_ms", C.c_double), ("walk_ms", C.c_double), ("chain_ms", C.c_double),
        ("trace_launches", C.c_uint64), ("walk_launches", C.c_uint64), ("chain_launches", C.c_uint64), ("direct_ms", C.c_double),
        ("first_stage_ms", C.c_double)]


class dr_ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("mint", C.c_float), ("d", C.c_float * 3), ("maxt", C.c_float)]


class dr_hit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("prim", C.c_int32)]


class dr_path_result(C.Structure):
    _fields_ = [("luminance", C.c_float), ("n_splats", C.c_int32), ("s", C.c_int32), ("t", C.c_int32),
                ("mis_weight", C.c_float), ("pos", (C.c_float * 2) * DR_MAX_SPLATS),
                ("value", (C.c_float * 3) * DR_MAX_SPLATS), ("n_rays", C.c_int32)]


class dr_step_record(C.Structure):
    _fields_ = [("L_x", C.c_float), ("L_y", C.c_float), ("L_z", C.c_float), ("a1", C.c_float), ("a2", C.c_float),
                ("large_step", C.c_uint8), ("accept1", C.c_uint8), ("did_second", C.c_uint8), ("accept2", C.c_uint8)]


dr_refresh_fn = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_double, C.POINTER(dr_stats), C.c_void_p)

# every symbol include/drmlt_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "dr_abi_version", "dr_last_error", "dr_device_count", "dr_config_default", "dr_config_set",
    "dr_config_validate", "dr_scene_create", "dr_scene_create_ex", "dr_scene_bvh_info", "dr_scene_destroy", "dr_scene_reupload", "dr_scene_clone", "dr_render_multi", "dr_render", "dr_cancel",
    "dr_job_create", "dr_job_destroy", "dr_job_bootstrap", "dr_job_seed_chains", "dr_job_run",
    "dr_job_film_device", "dr_job_develop", "dr_job_stats", "dr_job_profile", "dr_job_direct", "dr_direct_image", "dr_job_num_chains", "dr_job_total_mutations",
    "dr_trace_rays", "dr_texture_eval", "dr_eval_paths", "dr_chain_steps", "dr_chain_replay", "dr_splat_points", "dr_bootstrap_luminance", "dr_max_dimensions",
    "dr_render_progressive", "dr_film_size", "dr_first_stage_config", "dr_resample_luminance", "dr_importance_map",
]

_lib = None


class DrmltError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("drmlt_b200 status %d: %s" % (status, message))
        self.status = status


def load_library(path=None):
    """dlopen the CUDA library and declare argument types.  Raises if it was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError("CUDA library not built: %s missing (run __graft_entry__.build()); "
                          "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    P = C.POINTER
    lib.dr_abi_version.restype = C.c_int
    lib.dr_last_error.restype = C.c_char_p
    lib.dr_device_count.restype = C.c_int
    lib.dr_config_default.argtypes = [P(dr_config)]
    lib.dr_config_default.restype = None
    lib.dr_config_set.argtypes = [P(dr_config), C.c_char_p, C.c_char_p]
    lib.dr_config_validate.argtypes = [P(dr_config)]
    lib.dr_scene_create.argtypes = [P(dr_scene_desc), C.c_int, P(C.c_void_p)]
    lib.dr_scene_create_ex.argtypes = [P(dr_scene_desc), C.c_int, C.c_uint32, P(C.c_void_p)]
    lib.dr_scene_bvh_info.argtypes = [C.c_void_p, P(C.c_int32), P(C.c_int32), P(C.c_int32), P(C.c_double)]
    lib.dr_scene_bvh_info.restype = None
    lib.dr_scene_destroy.argtypes = [C.c_void_p]
    lib.dr_scene_destroy.restype = None
    lib.dr_scene_clone.argtypes = [C.c_void_p, C.c_int, P(C.c_void_p)]
    lib.dr_render_multi.argtypes = [P(C.c_void_p), C.c_int32, P(dr_config), P(C.c_float), P(dr_stats)]
    lib.dr_scene_reupload.argtypes = [C.c_void_p, P(C.c_int64)]
    lib.dr_render.argtypes = [C.c_void_p, P(dr_config), P(C.c_float), P(dr_stats)]
    lib.dr_cancel.argtypes = [C.c_void_p]
    lib.dr_cancel.restype = None
    lib.dr_job_create.argtypes = [C.c_void_p, P(dr_config), P(C.c_void_p)]
    lib.dr_job_destroy.argtypes = [C.c_void_p]
    lib.dr_job_destroy.restype = None
    lib.dr_job_bootstrap.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double)]
    lib.dr_job_seed_chains.argtypes = [C.c_void_p, C.c_double]
    lib.dr_job_run.argtypes = [C.c_void_p, C.c_int64]
    lib.dr_job_film_device.argtypes = [C.c_void_p, P(C.c_void_p), P(C.c_int64)]
    lib.dr_job_develop.argtypes = [C.c_void_p, P(C.c_float)]
    lib.dr_job_stats.argtypes = [C.c_void_p, P(dr_stats)]
    lib.dr_job_direct.argtypes = [C.c_void_p]
    lib.dr_direct_image.argtypes = [C.c_void_p, P(dr_config), P(C.c_float), P(C.c_double)]
    lib.dr_job_profile.argtypes = [C.c_void_p, C.c_int]
    lib.dr_job_profile.restype = None
    lib.dr_job_num_chains.argtypes = [C.c_void_p]
    lib.dr_job_num_chains.restype = C.c_int64
    lib.dr_job_total_mutations.argtypes = [C.c_void_p]
    lib.dr_job_total_mutations.restype = C.c_int64
    lib.dr_trace_rays.argtypes = [C.c_void_p, P(dr_ray), C.c_int64, C.c_int, P(dr_hit)]
    lib.dr_texture_eval.argtypes = [C.c_void_p, C.c_uint32, P(C.c_double), C.c_int64, P(C.c_double)]
    lib.dr_eval_paths.argtypes = [C.c_void_p, P(dr_config), P(C.c_float), C.c_int, P(C.c_float), C.c_int,
                                  P(C.c_float), C.c_int, P(C.c_int32), C.c_int64, P(dr_path_result)]
    lib.dr_chain_steps.argtypes = [C.c_void_p, P(dr_config), C.c_double, P(C.c_uint64), P(C.c_int32), P(C.c_uint64),
                                   C.c_int64, C.c_int64, P(dr_step_record), P(C.c_float)]
    lib.dr_chain_replay.argtypes = [C.c_void_p, P(dr_config), C.c_double, P(C.c_int32), C.c_int64, C.c_int64, P(C.c_double), C.c_int32,
                                    P(dr_step_record), P(C.c_float)]
    lib.dr_splat_points.argtypes = [C.c_int, P(dr_config), C.c_int32, C.c_int32, P(C.c_float), P(C.c_float), C.c_int64, P(C.c_float)]
    lib.dr_bootstrap_luminance.argtypes = [C.c_void_p, P(dr_config), C.c_uint64, C.c_int64, P(C.c_float), P(C.c_int32)]
    lib.dr_max_dimensions.argtypes = [P(dr_config), C.c_int, P(C.c_int), P(C.c_int), P(C.c_int)]
    lib.dr_max_dimensions.restype = None
    lib.dr_render_progressive.argtypes = [C.c_void_p, P(dr_config), P(C.c_float), P(dr_stats), C.c_double, dr_refresh_fn, C.c_void_p]
    lib.dr_film_size.argtypes = [C.c_void_p, P(dr_config), P(C.c_int32), P(C.c_int32)]
    lib.dr_first_stage_config.argtypes = [C.c_void_p, P(dr_config), P(dr_config)]
    lib.dr_resample_luminance.argtypes = [C.c_void_p, P(C.c_float), C.c_int32, C.c_int32, C.c_int32, C.c_int32, P(C.c_float)]
    lib.dr_importance_map.argtypes = [C.c_void_p, P(dr_config), P(C.c_float), P(dr_stats)]
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(lib, status):
    if status != DR_OK:
        msg = lib.dr_last_error()
        raise DrmltError(status, msg.decode() if msg else "")
