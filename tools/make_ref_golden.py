"""Writes tests/golden/ref_leaf.npz: outputs of the REFERENCE'S OWN leaf code (oracle/_ref/libref_leaf.so, compiled by
oracle/ref/Makefile from the sources under /root/reference) on the seeded inputs of tests/ref_leaf_cases.py.
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
