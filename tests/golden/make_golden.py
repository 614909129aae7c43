"""Generates tests/golden/oracle_checksums.json from the CPU oracle (exact build).

These are regression pins of the ORACLE, not reference outputs: the Fortran reference cannot be
built in this image and stores no vectors for this path.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cice_b200 import abi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {
    "tiny_s1": dict(config="tiny"),
    "tiny_s2": dict(config="tiny", seed=7),
    "gx3_s1_ndte120": dict(config="gx3"),
    "gx3_s2_blocks": dict(config="gx3", seed=20260101, ndte=12, block_size=(25, 29)),
    "gx3_revised": dict(config="gx3", ndte=12, revised_evp=True),
}

out = {}
for name, kw in CASES.items():
    c = synth.make_case(**kw)
    f = c.copy_fields()
    oracle.evp_run_bgrid(c.grid, c.params, f, nthreads=1)
    sums = {n: hashlib.sha256(np.ascontiguousarray(synth.gather(f[n], c.blocks)).tobytes()).hexdigest()
            for n in ("uvel", "vvel", "stressp_1", "stressm_3", "stress12_4", "strintxU", "strintyU")}
    out[name] = dict(case=kw, sums=sums)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_checksums.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", path)
