"""Multi-GPU driver: one process per GPU, chains sharded across ranks.

The path shards into independent units (chains never interact, drmlt_proc.cpp:856-883), so the
data path has no collective.  Exactly two exchanges exist, both mirroring what the reference does
between its init threads / work units:
  1. b: every rank bootstraps its own sample range; {sum luminance, sample count} are summed with
     ONE all-reduce (the reference averages the per-thread means, drmlt.cpp:530-545);
  2. film: the per-rank accumulation films are summed onto rank 0 with ONE reduce at the end
     (the reference's processResult adds full-frame blocks under a mutex, drmlt_proc.cpp:856-867).
Two-stage MLT (twoStage=true, src/libbidir/util.cpp:96-199) runs the nested low-resolution job first, sharded the same
way; its (small) film is ALL-reduced so that every rank develops the same first-stage image and derives the same
importance map without a broadcast.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is only the transport.
"""
import numpy as np


def shard_range(total, world_size, rank):
    """Contiguous share [first, first + n) of `total` units for `rank` (remainder to the low ranks)."""
    base, rem = divmod(int(total), int(world_size))
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def normalization_from_sums(lum_sum, count, is_mmlt, max_depth):
    """b = mean luminance over all non-NaN bootstrap samples, x maxDepth for MMLT (pathsampler.cpp:922-934)."""
    b = lum_sum / count if count > 0 else 0.0
    return b * max_depth if is_mmlt else b


def all_reduce_normalization(lum_sum, count, is_mmlt, max_depth, dist=None, device="cpu"):
    """One all-reduce of {sum, count} -> the global b (identical on every rank)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.tensor([lum_sum, count], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        lum_sum, count = float(t[0].item()), float(t[1].item())
    return normalization_from_sums(lum_sum, count, is_mmlt, max_depth)


def reduce_film(film_tensor, dist=None, dst=0, all_ranks=False):
    """One reduce(sum) of the accumulation film onto rank `dst` (or onto every rank: the first-stage film of
    two-stage MLT), in place."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if all_ranks:
            dist.all_reduce(film_tensor, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor


def render(scene, params, dist=None, rank=0, world_size=1, mutations_per_chain=None):
    """A whole job on `world_size` GPUs: returns (image on rank 0 else None, job stats, b).
    `params` are reference-style parameters (see integrator.make_config)."""
    from .integrator import make_config, set_importance_map
    import time
    cfg = make_config(rank=rank, worldSize=world_size, **params)
    first_stage_ms = 0.0
    if cfg.two_stage and not cfg.first_stage:
        # mltLuminancePass (util.cpp:96-199): nested job on all ranks, all-reduced film, identical map everywhere
        t0 = time.perf_counter()
        nested = scene.first_stage_config(cfg)
        img_n, _, _ = _render_cfg(scene, nested, dist, rank, world_size, None, all_ranks=True)
        set_importance_map(cfg, scene.resample_luminance(img_n, scene.film_size(cfg)))
        first_stage_ms = (time.perf_counter() - t0) * 1e3
    out = _render_cfg(scene, cfg, dist, rank, world_size, mutations_per_chain, all_ranks=False)
    render.last_timing = dict(_render_cfg.last_timing, first_stage_ms=first_stage_ms)
    return out


def _render_cfg(scene, cfg, dist, rank, world_size, mutations_per_chain, all_ranks):
    import torch
    from . import abi
    from .integrator import DeviceFilm, Job

    import time
    t0 = time.perf_counter()
    timing = {}

    def lap(name):
        nonlocal t0
        t1 = time.perf_counter()
        timing[name] = timing.get(name, 0.0) + (t1 - t0) * 1e3
        t0 = t1

    job = Job(scene, cfg)
    lap("job_create_ms")
    s, c = job.bootstrap()
    lap("bootstrap_ms")
    dev = "cuda:%d" % scene.device
    b = all_reduce_normalization(s, c, cfg.technique == abi.DR_TECH_MMLT, cfg.max_depth, dist, dev)
    lap("allreduce_b_ms")
    # a rank whose own seed pool is empty (or whose chains fail) must not leave the others waiting in the film reduce: every
    # rank learns of the failure through one small all-reduce and raises together
    failure = None
    try:
        job.seed_chains(b)
        lap("seed_chains_ms")
        per = mutations_per_chain if mutations_per_chain is not None else max(1, job.total_mutations // job.num_chains)
        job.run(per)
        lap("chains_ms")
    except Exception as e:              # noqa: BLE001 -- re-raised below, on every rank
        failure = e
    if world_size > 1:
        flag = torch.tensor([1.0 if failure is not None else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        if float(flag.item()) > 0:
            job.close()
            raise failure if failure is not None else RuntimeError("another rank failed before the film reduce")
        film = torch.as_tensor(DeviceFilm(job), device=dev)
        reduce_film(film, dist, 0, all_ranks)
        torch.cuda.synchronize(dev)      # the job's streams are non-blocking: this host sync orders the reduce before develop
    elif failure is not None:
        job.close()
        raise failure
    lap("film_reduce_ms")
    if rank == 0:
        job.direct()                       # separate direct-illumination image (directSamples > 0), added by develop
    lap("direct_ms")
    img = job.develop() if (rank == 0 or all_ranks) else None
    lap("develop_ms")
    st = job.stats()
    job.close()
    lap("job_destroy_ms")
    _render_cfg.last_timing = timing
    return img, st, b


def relmse(img, ref, eps=1e-2):
    """relMSE = mean((I - R)^2 / (R^2 + eps))."""
    img = np.asarray(img, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))
