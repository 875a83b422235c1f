"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/drmlt_b200.h
declares, parses the reference's parameter names, applies its constructor-time rules, and refuses
to compute without a GPU (no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from drmlt_mitsuba_b200 import abi
from drmlt_mitsuba_b200.integrator import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "drmlt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dr_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = _header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libdrmlt_b200.so does not export %s" % name
    assert sorted(abi.EXPORTED_SYMBOLS) == declared
    assert lib.dr_abi_version() == 6


def test_struct_layouts_match_header(lib):
    assert C.sizeof(abi.dr_material) == 64
    assert C.sizeof(abi.dr_ray) == 32 and C.sizeof(abi.dr_hit) == 16
    assert C.sizeof(abi.dr_step_record) == 24
    assert C.sizeof(abi.dr_emitter) == 24
    assert C.sizeof(abi.dr_texture) == 64
    assert abi.DR_MAT_TEX_REFLECTANCE(0) == 0x100 and abi.DR_MAT_TEX_TRANSMITTANCE(2) == 3 << 20      # 1 + index in bits 8-19 / 20-31
    # the library and the ctypes mirror agree on dr_config: defaults land in the right fields
    cfg = abi.dr_config()
    lib.dr_config_default(C.byref(cfg))
    assert cfg.rr_depth == 5 and cfg.direct_samples == 16 and cfg.luminance_samples == 100000
    assert cfg.p_large == pytest.approx(0.3) and cfg.sigma == pytest.approx(1 / 64) and cfg.scale_second == pytest.approx(0.1)
    assert cfg.mutation_size_low == pytest.approx(1 / 1024) and cfg.mutation_size_high == pytest.approx(1 / 64)
    assert cfg.kelemen_style_weights == 1 and cfg.kelemen_style_mutation == 1 and cfg.light_image == 1
    assert cfg.max_depth == -1 and cfg.average_luminance == -1.0 and cfg.world_size == 1


def test_config_parameter_names_of_the_reference(lib):
    cfg = make_config(integrator="drmlt", technique="mmlt", type="orbital", maxDepth=8, sigma=1 / 64, scaleSecond=0.1,
                      timidAfterLarge=True, fixEmitterPath=True, useMixture=False, acceptanceMap=False, directSamples=-1)
    assert (cfg.integrator, cfg.technique, cfg.type) == (abi.DR_INTEGRATOR_DRMLT, abi.DR_TECH_MMLT, abi.DR_TYPE_ORBITAL)
    assert cfg.timid_after_large == 1 and cfg.fix_emitter_path == 1
    # MMLT forces directSampling = false and kelemenStyleWeights = false (drmlt.cpp:229-231, 266-268)
    assert cfg.direct_sampling == 0 and cfg.kelemen_style_weights == 0
    cfg = make_config(integrator="pssmlt", technique="path", maxDepth=8, kelemenStyleMutation=False, sigma=0.02)
    assert cfg.integrator == abi.DR_INTEGRATOR_PSSMLT and cfg.kelemen_style_mutation == 0
    assert make_config(integrator="drmlt", technique="path", type="mirasym", maxDepth=5).type == abi.DR_TYPE_ORBITAL
    # two-stage MLT (drmlt.cpp:278-293) and the film plugin's window (film.cpp:30-48)
    cfg = abi.dr_config()
    lib.dr_config_default(C.byref(cfg))
    assert cfg.two_stage == 0 and cfg.first_stage == 0 and cfg.first_stage_size_reduction == 16 and not cfg.importance_map
    cfg = make_config(integrator="drmlt", technique="path", type="mira", maxDepth=5, twoStage=True, firstStageSizeReduction=8,
                      width=640, height=360, cropOffsetX=10, cropOffsetY=20, cropWidth=100, cropHeight=50)
    assert (cfg.two_stage, cfg.first_stage_size_reduction, cfg.film_width, cfg.film_height) == (1, 8, 640, 360)
    assert (cfg.crop_offset_x, cfg.crop_offset_y, cfg.crop_width, cfg.crop_height) == (10, 20, 100, 50)


@pytest.mark.parametrize("params,needle", [
    (dict(integrator="drmlt", technique="vcm", type="mira", maxDepth=8), "Unknown technique"),
    (dict(integrator="drmlt", technique="path", type="tierney", maxDepth=8), "Unknown implementation"),
    (dict(integrator="drmlt", technique="mmlt", type="mira"), "MMLT with no max depth"),
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, fixEmitterPath=True), "fixEmitterPath without MMLT"),
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, scaleSecond=1.5), "scaleSecond is bigger"),
    (dict(integrator="drmlt", technique="path", maxDepth=8), "Unknown implementation"),      # type is required
    (dict(integrator="drmlt", type="mira", maxDepth=8), "Unknown technique"),                # technique is required
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, acceptanceMap=True), "Box filter required for acceptance map!"),      # the reference's own text (drmlt_proc.cpp:78)
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, bogus=1), "Unknown parameter"),
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, twoStage=True, firstStageSizeReduction=0), "firstStageSizeReduction"),
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, cropOffsetX=-1), "Invalid crop window"),
    (dict(integrator="drmlt", technique="path", type="mira", maxDepth=8, sigma="abc"), "not a number"),
])
def test_config_errors_mirror_the_reference(lib, params, needle):
    with pytest.raises(abi.DrmltError) as e:
        make_config(**params)
    assert needle in str(e.value)
    assert e.value.status in (1, 6)


def test_max_dimensions_matches_oracle(lib, oracle):
    for tech, md, ds in ((abi.DR_TECH_PATH, 8, 1), (abi.DR_TECH_BDPT, 8, 0), (abi.DR_TECH_BDPT, 8, 1), (abi.DR_TECH_MMLT, 8, 0),
                         (abi.DR_TECH_PATH, 4, 1)):
        for depth in range(1, md + 1):
            cfg = abi.dr_config()
            lib.dr_config_default(C.byref(cfg))
            cfg.technique, cfg.max_depth, cfg.direct_sampling = tech, md, ds
            a = [C.c_int() for _ in range(3)]
            b = [C.c_int() for _ in range(3)]
            lib.dr_max_dimensions(C.byref(cfg), depth, *[C.byref(x) for x in a])
            oracle.orc_max_dimensions(C.byref(cfg), depth, *[C.byref(x) for x in b])
            assert [x.value for x in a] == [x.value for x in b]


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly (DR_ERR_NO_DEVICE), never compute."""
    if lib.dr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from drmlt_mitsuba_b200 import scenes
    from drmlt_mitsuba_b200.integrator import Scene
    with pytest.raises(abi.DrmltError) as e:
        Scene(scenes.cornell_box(film=(16, 16), tess=1))
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)


def test_argument_validation_needs_no_device(lib):
    h = C.c_void_p()
    assert lib.dr_scene_create(None, 0, C.byref(h)) == 1
    assert b"null" in lib.dr_last_error()
    empty = abi.dr_scene_desc()
    assert lib.dr_scene_create(C.byref(empty), 0, C.byref(h)) == 1
    assert lib.dr_trace_rays(None, None, 1, 0, None) == 1
    assert lib.dr_job_run(None, 1) == 1
    # roughplastic (roughplastic.cpp:210-222, rtrans.h): needs its rough-transmittance table and eta != 1
    from drmlt_mitsuba_b200 import scenes
    for eta, table, needle in ((1.5, None, b"rough-transmittance table"), (1.0, np.zeros(abi.DR_ROUGH_TABLE_DOUBLES), b"must be positive and differ")):
        data = scenes.SceneData("rp", (16, 16))
        m = data.add_material(abi.DR_BSDF_ROUGHPLASTIC, eta=(eta, 0, 0), rough_table=np.zeros(abi.DR_ROUGH_TABLE_DOUBLES))
        data.add_quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0), m)
        data.set_camera((0, 0, 3), (0, 0, 0), (0, 1, 0), 40.0)
        desc = data.desc()
        if table is None:
            desc.rough_tables = None
            desc.n_rough_tables = 0
        assert lib.dr_scene_create(C.byref(desc), 0, C.byref(h)) == 1 and needle in lib.dr_last_error(), lib.dr_last_error()
    assert lib.dr_scene_create_ex(None, 0, 1, C.byref(h)) == 1
    # bitmap textures (ABI 6): texture indices, texel pointers, wrap modes, UV tangents need texture coordinates
    def textured():
        data = scenes.SceneData("tex", (16, 16))
        t = data.add_texture(scenes.procedural_texels(4, 4, 1))
        m = data.add_material(abi.DR_BSDF_DIFFUSE, reflectance_tex=t)
        data.add_quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0), m, uv=True, uv_tangents=True)
        data.set_camera((0, 0, 3), (0, 0, 0), (0, 1, 0), 40.0)
        return data
    data = textured(); desc = data.desc(); desc.n_textures = 0
    assert lib.dr_scene_create(C.byref(desc), 0, C.byref(h)) == 1 and b"texture index out of range" in lib.dr_last_error()
    data = textured(); desc = data.desc(); desc.texcoords = None
    assert lib.dr_scene_create(C.byref(desc), 0, C.byref(h)) == 1 and b"DR_TRI_UV_TANGENTS without texcoords" in lib.dr_last_error()
    data = textured(); data.textures[0].wrap_u = 7; desc = data.desc()
    assert lib.dr_scene_create(C.byref(desc), 0, C.byref(h)) == 1 and b"wrap mode" in lib.dr_last_error()
    data = textured(); data.textures[0].texels = None; desc = data.desc()
    assert lib.dr_scene_create(C.byref(desc), 0, C.byref(h)) == 1 and b"missing texels" in lib.dr_last_error()
    data = textured(); desc = data.desc()                       # a valid textured scene passes validation (and then needs the device)
    assert lib.dr_scene_create(C.byref(desc), 0, C.byref(h)) in (0, 2)


# The header is C (not C++), and a plain C99 host drives the library through it: examples/render_box.c compiles with -std=c99 -pedantic,
# links against libdrmlt_b200.so and -- without a GPU -- gets DR_ERR_NO_DEVICE from dr_scene_create after its textured scene passed validation.
def build_c_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "render_box")
    libdir = os.path.join(root, "drmlt-mitsuba_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O2", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "render_box.c"), "-L" + libdir, "-ldrmlt_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    return exe


def test_c99_host_example_builds_and_fails_loudly_without_a_gpu(lib, tmp_path):
    import subprocess
    exe = build_c_example(tmp_path)
    if lib.dr_device_count() > 0:
        pytest.skip("a CUDA device is present: tests/test_gpu_parity.py runs the example")
    p = subprocess.run([exe, str(tmp_path / "o.ppm"), "4"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr, (p.returncode, p.stderr)
