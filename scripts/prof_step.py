"""Tiny driver for ncu: uploads one synthetic state and runs a few subcycle loops.
usage: python scripts/prof_step.py [workload] [kernel] [mode] [ndte] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cice_b200 import abi, dyn_evp, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gx1"
kernel = sys.argv[2] if len(sys.argv) > 2 else "fused"
mode = sys.argv[3] if len(sys.argv) > 3 else "exact"
ndte = int(sys.argv[4]) if len(sys.argv) > 4 else 8
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
c = synth.make_case(wl, ndte=ndte)
p = dict(c.params, mode=abi.MODE_EXACT if mode == "exact" else abi.MODE_FAST, kernel=abi.KERNEL_NAMES[kernel])
dyn_evp.dyn_evp_b200_init(c.grid)
# the metric arrays (what the Fortran shim hands over in its init): derived geometry for the streaming forms, required by tstream
bad = dyn_evp.set_metric(synth.scatter(c.X["HTN"], c.blocks), synth.scatter(c.X["HTE"], c.blocks), c.params["deltaminEVP"])
f = c.copy_fields()
dyn_evp.upload(f)
for _ in range(reps):
    dyn_evp.subcycle(p)
    print("loop ms", dyn_evp.last_loop_ms(), "launches", dyn_evp.last_launches())
print("metric mismatches", bad, "|", dyn_evp.describe()[-160:])
dyn_evp.dyn_evp_b200_finalize()
