"""Multi-rank host logic on CPU (gloo, world_size 2): the two exchanges of the sharded job
 -- one all-reduce for the normalisation b, one reduce of the films -- and
their equivalence with a single-rank job computed by the oracle.  No GPU is needed: the ranks'
bootstrap sums and films come from the CPU oracle on the rank's own shard."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from drmlt_mitsuba_b200 import abi, distributed, scenes  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg(rank=0, world=1):
    import oracle_lib
    return oracle_lib.default_config(integrator=abi.DR_INTEGRATOR_DRMLT, technique=abi.DR_TECH_MMLT, type=abi.DR_TYPE_ORBITAL,
                                     max_depth=5, direct_samples=-1, direct_sampling=0, kelemen_style_weights=0, seed=11,
                                     rank=rank, world_size=world)


def _rank_job(rank, world, n_boot, n_chains, steps):
    """What one rank of the sharded job computes: its bootstrap shard, then (given b) its chains' film."""
    import oracle_lib
    data = scenes.cornell_box(film=(32, 32), tess=2)
    orc = oracle_lib.OracleScene(data)
    cfg = _cfg(rank, world)
    first, n = distributed.shard_range(n_boot, world, rank)
    lum, dep = orc.bootstrap(cfg, first, n)
    ok = ~np.isnan(lum)
    return orc, cfg, first, lum, dep, float(lum[ok].sum()), float(ok.sum())


def _worker(rank, world, port, n_boot, n_chains, steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc, cfg, first, lum, dep, s, c = _rank_job(rank, world, n_boot, n_chains, steps)
        b = distributed.all_reduce_normalization(s, c, True, cfg.max_depth, dist, "cpu")     # exchange 1
        # chains of this rank: seeds from its own pool, chain ids offset by rank (as dr_job_seed_chains does)
        seeds = np.nonzero(lum > 1e-12)[0][:n_chains]
        ids = np.arange(len(seeds), dtype=np.uint64) + rank * n_chains
        _, film, st = orc.chain_steps(cfg, b, (seeds + first).astype(np.uint64), dep[seeds], ids, steps, want_film=True, want_records=False)
        t = torch.from_numpy(np.ascontiguousarray(film, dtype=np.float32))
        distributed.reduce_film(t, dist, 0)                                                    # exchange 2
        if rank == 0:
            out["b"] = b
            out["film"] = t.numpy().copy()
        out["sum%d" % rank] = (s, c, int(st.mutations))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for total, world in [(10, 3), (1280 * 720 * 64, 8), (7, 8), (0, 2)]:
        parts = [distributed.shard_range(total, world, r) for r in range(world)]
        assert sum(n for _, n in parts) == total
        pos = 0
        for first, n in parts:
            assert first == pos
            pos += n


def test_normalization_matches_reference_formula():
    # mean over all samples, x maxDepth for MMLT (pathsampler.cpp:922-934)
    assert distributed.normalization_from_sums(6.0, 3.0, False, 8) == pytest.approx(2.0)
    assert distributed.normalization_from_sums(6.0, 3.0, True, 8) == pytest.approx(16.0)
    assert distributed.normalization_from_sums(0.0, 0.0, True, 8) == 0.0


def test_two_ranks_gloo_all_reduce_b_and_film_reduce(oracle):
    world, n_boot, n_chains, steps = 2, 4000, 16, 24
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_boot, n_chains, steps, out), nprocs=world, join=True)
    # single-process recomputation of the same shards
    parts = [_rank_job(r, world, n_boot, n_chains, steps) for r in range(world)]
    s_all = sum(p[5] for p in parts)
    c_all = sum(p[6] for p in parts)
    b = distributed.normalization_from_sums(s_all, c_all, True, parts[0][1].max_depth)
    assert out["b"] == pytest.approx(b, rel=1e-12)
    for r in range(world):
        assert out["sum%d" % r][0] == pytest.approx(parts[r][5], rel=1e-12)
        assert out["sum%d" % r][2] == n_chains * steps
    # b of the union equals b of one rank that bootstraps the whole range (keyed samples: sharding is invisible)
    orc, cfg = parts[0][0], _cfg(0, 1)
    lum, _ = orc.bootstrap(cfg, 0, n_boot)
    assert b == pytest.approx(distributed.normalization_from_sums(float(np.nansum(lum)), float((~np.isnan(lum)).sum()), True, cfg.max_depth), rel=1e-9)
    # the reduced film is the sum of the ranks' films
    film = np.zeros_like(out["film"])
    for r, (orc, cfg, first, lum, dep, s, c) in enumerate(parts):
        seeds = np.nonzero(lum > 1e-12)[0][:n_chains]
        ids = np.arange(len(seeds), dtype=np.uint64) + r * n_chains
        _, f, _ = orc.chain_steps(cfg, b, (seeds + first).astype(np.uint64), dep[seeds], ids, steps, want_film=True, want_records=False)
        film += f.astype(np.float32)
    assert film.sum() > 0
    np.testing.assert_allclose(out["film"], film, rtol=1e-5, atol=1e-7)


def _worker_two_stage(rank, world, port, out):
    """First stage of two-stage MLT on two ranks: each rank's nested film shard is ALL-reduced, then every rank derives
    the importance map on its own (no broadcast)."""
    import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(100 + rank)
        film = torch.from_numpy(rng.rand(6, 8, 3).astype(np.float32))
        distributed.reduce_film(film, dist, 0, all_ranks=True)
        out["map%d" % rank] = oracle_lib.resample_luminance(film.numpy(), (32, 24))
    finally:
        dist.destroy_process_group()


def test_two_ranks_first_stage_film_all_reduce(oracle):
    import oracle_lib
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_two_stage, args=(2, _free_port(), out), nprocs=2, join=True)
    total = sum(np.random.RandomState(100 + r).rand(6, 8, 3).astype(np.float32) for r in range(2))
    ref = oracle_lib.resample_luminance(total, (32, 24))
    assert np.array_equal(out["map0"], out["map1"])
    np.testing.assert_allclose(out["map0"], ref, rtol=1e-6)
