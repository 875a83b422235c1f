"""Writes tests/golden/ref_chain.npz: whole chains of the REFERENCE'S OWN DRMLTRenderer::process / PSSMLTRenderer::process
(oracle/_ref/libref_path.so: ref_drmlt_chain, ref_pssmlt_chain; compiled by oracle/ref/Makefile from the sources under
/root/reference) on explicitly seeded generators -- the two uniform streams each chain consumes, its per-prefix statistics
counters and film projections, and its final film -- for the cases of tests/ref_path_cases.py (CHAIN_REPLAY_CASES).
Run in the container that has /root/reference; the fixture travels, the reference does not."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_path_cases as RP  # noqa: E402

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref")])
out = RP.run_chain_ref_all(C.CDLL(RP.REF_PATH))
np.savez_compressed(RP.GOLDEN_CHAIN, **out)
print("wrote", RP.GOLDEN_CHAIN, os.path.getsize(RP.GOLDEN_CHAIN), "bytes,", len(RP.CHAIN_REPLAY_CASES) * len(RP.CHAIN_PICKS), "chains")
