"""Generate tests/golden/vectors.npz: known-answer vectors of the hot path on small procedural scenes.

PROVENANCE.  The reference ships no golden vectors for this path and cannot be built or imported in this image
(SURVEY.md F5/F6), so these vectors are produced by the CPU oracle (oracle/, a restatement of the reference's
algorithm) -- they pin the oracle against regressions and give the CUDA path a fixture that travels to the GPU box;
they are NOT outputs of the reference.  Outputs of the reference's own compiled code are in tests/golden/ref_*.npz
(tools/make_ref_golden.py), and the oracle is pinned to those by tests/test_ref_pins.py.

    python tools/make_golden.py            # rewrites tests/golden/vectors.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from drmlt_mitsuba_b200 import abi, scenes  # noqa: E402

# (name, scene factory, oracle config overrides, primary-sample dims (sensor, emitter, direct))
CASES = [
    ("cornell_pssmlt_path", lambda: scenes.cornell_box(film=(64, 64), tess=4),
     dict(integrator=abi.DR_INTEGRATOR_PSSMLT, technique=abi.DR_TECH_PATH, max_depth=8, direct_samples=-1), (50, 2, 2)),
    ("cornell_drmlt_mira_path", lambda: scenes.cornell_box(film=(64, 64), tess=4),
     dict(integrator=abi.DR_INTEGRATOR_DRMLT, type=abi.DR_TYPE_MIRA, technique=abi.DR_TECH_PATH, max_depth=8, direct_samples=-1, scale_second=0.1), (50, 2, 2)),
    ("glossy_drmlt_green_bdpt", lambda: scenes.glossy_scene(film=(64, 64), subdiv=2),
     dict(integrator=abi.DR_INTEGRATOR_DRMLT, type=abi.DR_TYPE_GREEN, technique=abi.DR_TECH_BDPT, max_depth=6, direct_samples=-1, direct_sampling=0), (24, 24, 1)),
    ("caustic_drmlt_orbital_mmlt_fix", lambda: scenes.caustic_scene(film=(64, 64), grid=24),
     dict(integrator=abi.DR_INTEGRATOR_DRMLT, type=abi.DR_TYPE_ORBITAL, technique=abi.DR_TECH_MMLT, max_depth=8, direct_samples=-1, direct_sampling=0,
          kelemen_style_weights=0, fix_emitter_path=1), (30, 30, 1)),
    ("door_drmlt_orbital_mmlt", lambda: scenes.door_scene(film=(80, 45), floor_grid=32, n_spheres=9, sphere_subdiv=2),
     dict(integrator=abi.DR_INTEGRATOR_DRMLT, type=abi.DR_TYPE_ORBITAL, technique=abi.DR_TECH_MMLT, max_depth=8, direct_samples=-1, direct_sampling=0,
          kelemen_style_weights=0), (30, 30, 1)),
    # rough dielectric (extra primary sample per BSDF sample) and a crop window on an overridden film size
    ("roughglass_drmlt_mira_mmlt", lambda: scenes.glossy_scene(film=(64, 64), subdiv=2, rough_glass=(0.15, abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE)),
     dict(integrator=abi.DR_INTEGRATOR_DRMLT, type=abi.DR_TYPE_MIRA, technique=abi.DR_TECH_MMLT, max_depth=6, direct_samples=-1, direct_sampling=0,
          kelemen_style_weights=0), (24, 24, 1)),
    ("roughglass_pssmlt_path_crop", lambda: scenes.glossy_scene(film=(64, 64), subdiv=2, rough_glass=(0.3, 0)),
     dict(integrator=abi.DR_INTEGRATOR_PSSMLT, technique=abi.DR_TECH_PATH, max_depth=6, direct_samples=-1,
          film_width=96, film_height=64, crop_offset_x=16, crop_offset_y=8, crop_width=64, crop_height=48), (48, 2, 2)),
    ("plastic_drmlt_orbital_mmlt", lambda: scenes.cornell_box(film=(64, 64), tess=4, plastic=True),
     dict(integrator=abi.DR_INTEGRATOR_DRMLT, type=abi.DR_TYPE_ORBITAL, technique=abi.DR_TECH_MMLT, max_depth=6, direct_samples=-1, direct_sampling=0,
          kelemen_style_weights=0), (24, 24, 1)),
]
N_PATHS, N_BOOT, N_CHAINS, N_STEPS = 640, 3000, 6, 24


def config(over):
    cfg = oracle_lib.default_config(seed=20261018, ray_epsilon=1e-4, shadow_epsilon=1e-3, **over)
    return cfg


def generate(name, make, over, dims):
    data = make()
    orc = oracle_lib.OracleScene(data)
    cfg = config(over)
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    md = cfg.max_depth
    depth = rng.randint(1, md + 1, N_PATHS).astype(np.int32)
    us, ue, ud = [rng.rand(N_PATHS, k).astype(np.float32) for k in dims]
    res, lum64 = orc.eval_paths(cfg, us, ue, ud, depth)
    res = np.frombuffer(res, dtype=np.uint8).reshape(N_PATHS, C.sizeof(abi.dr_path_result)).copy()
    lum, dep = orc.bootstrap(cfg, 0, N_BOOT)
    seeds = np.nonzero(lum > 1e-9 * np.nanmax(lum))[0][:N_CHAINS].astype(np.uint64)
    ids = np.arange(len(seeds), dtype=np.uint64) + 77
    rec, film, st = orc.chain_steps(cfg, 0.25, seeds, dep[seeds.astype(np.int64)], ids, N_STEPS, want_film=True)
    rec = np.frombuffer(rec, dtype=np.uint8).reshape(len(seeds) * N_STEPS, C.sizeof(abi.dr_step_record)).copy()
    return {name + "/depth": depth, name + "/us": us, name + "/ue": ue, name + "/ud": ud, name + "/path_results": res,
            name + "/lum64": lum64, name + "/boot_lum": lum.astype(np.float32), name + "/seeds": seeds,
            name + "/seed_depth": dep[seeds.astype(np.int64)], name + "/chain_ids": ids, name + "/records": rec,
            name + "/film_sum": np.array([film.sum()], np.float64)}


def main():
    oracle_lib.build()
    out = {}
    for name, make, over, dims in CASES:
        out.update(generate(name, make, over, dims))
    path = os.path.join(ROOT, "tests", "golden", "vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
