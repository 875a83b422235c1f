"""Writes tests/golden/ref_leaf.npz and tests/golden/ref_path.npz: outputs of the REFERENCE'S OWN leaf code (oracle/_ref/libref_leaf.so, compiled by
oracle/ref/Makefile from the sources under /root/reference) on the seeded inputs of tests/ref_leaf_cases.py.
ref_path.npz holds the same for oracle/_ref/libref_path.so: the reference's BSDF plugins (sample / eval / pdf) and
PathSampler::sampleSplats (MMLT / BDPT / PT) on the replayed vectors of tests/ref_path_cases.py.
Run in the container that has /root/reference; the fixture travels, the reference does not."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_leaf_cases as R  # noqa: E402

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref")])
out = R.run_cases(R.load(R.REF_LEAF), "ref_")
np.savez_compressed(R.GOLDEN, **out)
print("wrote", R.GOLDEN, {k: v.shape for k, v in out.items()})

import ctypes as C  # noqa: E402
import ref_path_cases as RP  # noqa: E402

lib = C.CDLL(RP.REF_PATH)
out = RP.run_bsdf(lib, "ref_")
for case in RP.PATH_CASES:
    r = RP.run_paths_ref(lib, case)
    k = RP.case_key(case)
    out[k + "_lum"] = r["lum"]
    out[k + "_st"] = np.stack([r["s"], r["t"], r["n_splats"]], 1).astype(np.int8)
    out[k + "_pos0"] = r["pos0"]
    out[k + "_value0"] = r["value0"]
    print(k, "contributing", int((r["lum"] > 0).sum()), "of", len(r["lum"]))
np.savez_compressed(RP.GOLDEN, **out)
print("wrote", RP.GOLDEN, os.path.getsize(RP.GOLDEN), "bytes")

# ---- the reference's own DRMLT samplers (Green / Mira / Orbital) on recorded uniform streams
np.savez_compressed(RP.GOLDEN_SAMPLER, **RP.run_sampler_ref(lib))
print("wrote", RP.GOLDEN_SAMPLER, os.path.getsize(RP.GOLDEN_SAMPLER), "bytes")
np.savez_compressed(RP.GOLDEN_PSS_SAMPLER, **RP.run_pss_sampler_ref(lib))
print("wrote", RP.GOLDEN_PSS_SAMPLER, os.path.getsize(RP.GOLDEN_PSS_SAMPLER), "bytes")
np.savez_compressed(RP.GOLDEN_SAMPLER_SEQ, **RP.run_sampler_seq_ref(lib))
print("wrote", RP.GOLDEN_SAMPLER_SEQ, os.path.getsize(RP.GOLDEN_SAMPLER_SEQ), "bytes")

# ---- the reference's own ImageBlock::put with its gaussian / box filter plugins
np.savez_compressed(RP.GOLDEN_FILM, **RP.run_film(lib.ref_splat, True))
print("wrote", RP.GOLDEN_FILM, os.path.getsize(RP.GOLDEN_FILM), "bytes")

# ---- the reference's own DRMLT / PSSMLT integrators end to end (statistics counters, b, images)
rout = {}
RUNS = 3            # the reference seeds from /dev/urandom: keep its own run-to-run spread next to the values
for name, (params, spp) in RP.RENDER_CASES.items():
    runs = [RP.run_render_ref(lib, params, spp) for _ in range(RUNS)]
    names = sorted(runs[0][3])
    rout[name + "_image"] = runs[0][0]
    rout[name + "_stats_names"] = np.array(names)
    rout[name + "_stats"] = np.array([[r[3][k] for k in names] for r in runs])
    rout[name + "_b"] = np.array([RP.luminance(r[0]).mean() for r in runs])
    rout[name + "_mutations_per_s"] = np.array([64 * 64 * spp / r[1] for r in runs])
    print(name, "b =", rout[name + "_b"], dict(zip(names, rout[name + "_stats"].T.round(2).tolist())))
# a long render of the first case as the converged image of the equal-mutation relMSE test
params, spp = RP.RENDER_CASES["drmlt_orbital_mmlt"]
img, sec, _, _ = RP.run_render_ref(lib, params, 16 * spp)
rout["converged_drmlt_orbital_mmlt"] = img
# ... and two more runs at the test's sample count: the reference's own run-to-run spread of relMSE
rout["drmlt_orbital_mmlt_relmse_runs"] = np.array([RP.rel_mse(rout["drmlt_orbital_mmlt_image"], img)] +
                                                  [RP.rel_mse(RP.run_render_ref(lib, params, spp)[0], img) for _ in range(2)])
print("converged in %.1f s; relMSE of the reference's own runs at %d spp:" % (sec, spp), rout["drmlt_orbital_mmlt_relmse_runs"])
np.savez_compressed(RP.GOLDEN_RENDER, **rout)
print("wrote", RP.GOLDEN_RENDER, os.path.getsize(RP.GOLDEN_RENDER), "bytes")
os._exit(0)        # the reference's scheduler threads are not torn down
