"""Known-answer fixtures (tests/golden/vectors.npz, made by tools/make_golden.py).

The reference has no golden vectors for this path and cannot run here, so the fixtures come from the CPU oracle (see
the generator's docstring): the CPU test pins the oracle to them bit-for-bit where it must be (decisions, integer
fields) and to 1e-12 elsewhere; the GPU test runs the CUDA path, through the C ABI, against the same committed
fixtures -- they travel to the GPU box, /root/reference does not."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

import oracle_lib
from drmlt_mitsuba_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
REC = np.dtype([("L_x", "<f4"), ("L_y", "<f4"), ("L_z", "<f4"), ("a1", "<f4"), ("a2", "<f4"),
                ("large", "u1"), ("acc1", "u1"), ("did2", "u1"), ("acc2", "u1")])


@pytest.fixture(scope="module")
def vectors():
    return np.load(os.path.join(ROOT, "tests", "golden", "vectors.npz"))


def _res_fields(buf, n):
    b = np.frombuffer(np.ascontiguousarray(buf), dtype=np.uint8).reshape(n, C.sizeof(abi.dr_path_result))
    K = abi.DR_MAX_SPLATS
    off = 20 + 8 * K
    return dict(lum=b[:, 0:4].copy().view("<f4")[:, 0], n=b[:, 4:8].copy().view("<i4")[:, 0], s=b[:, 8:12].copy().view("<i4")[:, 0],
                t=b[:, 12:16].copy().view("<i4")[:, 0], mis=b[:, 16:20].copy().view("<f4")[:, 0],
                pos=b[:, 20:off].copy().view("<f4").reshape(n, K, 2), val=b[:, off:off + 12 * K].copy().view("<f4").reshape(n, K, 3),
                rays=b[:, -4:].copy().view("<i4")[:, 0])


@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_oracle_reproduces_golden_vectors(oracle, vectors, case):
    name, make, over, dims = case
    fresh = G.generate(name, make, over, dims)
    for key, val in fresh.items():
        ref = vectors[key]
        if key.endswith("/path_results") or key.endswith("/records"):
            assert np.array_equal(np.asarray(val), ref), key          # every byte: floats are deterministic on one ISA
        elif val.dtype.kind == "f":
            np.testing.assert_allclose(val, ref, rtol=1e-12, atol=0, err_msg=key)
        else:
            assert np.array_equal(val, ref), key


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_cuda_path_matches_golden_vectors(lib, vectors, case):
    from drmlt_mitsuba_b200.integrator import Scene
    name, make, over, dims = case
    v = {k.split("/", 1)[1]: vectors[k] for k in vectors.files if k.startswith(name + "/")}
    gpu = Scene(make())
    cfg = G.config(over)
    abi.check(lib, lib.dr_config_validate(C.byref(cfg)))
    n = len(v["depth"])
    # ---- f(u) on the replayed primary-sample vectors
    got = _res_fields(gpu.eval_paths(cfg, v["us"], v["ue"], v["ud"], v["depth"]), n)
    ref = _res_fields(v["path_results"], n)
    lum = v["lum64"]
    noise = 1e-18 * lum.max()
    gl, rl = np.where(got["lum"] < noise, 0, got["lum"]).astype(np.float64), np.where(lum < noise, 0, lum)
    assert ((gl > 0) == (rl > 0)).mean() >= 0.995
    both = (gl > 0) & (rl > 0)
    assert both.sum() >= 8
    assert (np.abs(gl[both] - rl[both]) <= 1e-4 * rl[both]).mean() >= 0.999
    if cfg.technique != abi.DR_TECH_PATH:
        assert np.array_equal(got["s"][both], ref["s"][both]) and np.array_equal(got["t"][both], ref["t"][both])
    assert (got["n"][both] == ref["n"][both]).mean() >= 0.999
    same = both & (got["n"] == ref["n"])
    scale = np.abs(ref["val"][same]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(got["val"][same] - ref["val"][same]) <= 1e-4 * np.abs(ref["val"][same]) + 1e-6 * scale).all(axis=(1, 2)).mean() >= 0.999
    assert (np.abs(got["pos"][same] - ref["pos"][same]).max(axis=(1, 2)) < 2e-2).mean() >= 0.999
    # ---- bootstrap luminances of the keyed samples
    bl, _ = gpu.bootstrap_luminance(cfg, 0, len(v["boot_lum"]))
    rb = v["boot_lum"].astype(np.float64)
    nb = 1e-18 * np.nanmax(rb)
    a, b = np.where(bl < nb, 0, bl).astype(np.float64), np.where(rb < nb, 0, rb)
    ok = ((a > 0) == (b > 0)) & (np.abs(a - b) <= 1e-4 * np.maximum(b, 1e-300))
    assert ok.mean() >= 0.999
    # ---- accept / reject decisions of the recorded chains under identical uniforms
    steps = len(v["records"]) // len(v["seeds"])
    rg = np.frombuffer(gpu.chain_steps(cfg, 0.25, v["seeds"], v["seed_depth"], v["chain_ids"], steps), dtype=REC)
    rr = np.frombuffer(np.ascontiguousarray(v["records"]), dtype=REC)
    agree = np.stack([rg[k] == rr[k] for k in ("large", "acc1", "did2", "acc2")]).all(0)
    assert agree.mean() > 0.9                     # a near-threshold flip changes one chain's tail, nothing else
    first = agree.reshape(len(v["seeds"]), steps)
    assert first[:, 0].all()                      # the first mutation of every chain agrees
