import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cice_b200 import abi, synth, dyn_evp
from oracle import oracle
kw = dict(seed=7, ew="cyclic", ns="cyclic", kmt="none")
for ndte in (1, 2):
    c = synth.make_ccase("tiny", ndte=ndte, **kw)
    ref = c.copy_fields(); oracle.evp_run_cgrid(c.grid, c.cgrid, c.params, ref, 1)
    got = c.copy_fields()
    dyn_evp.dyn_evp_b200_init(c.grid); dyn_evp.dyn_evp_b200_init_cgrid(c.cgrid)
    print(dyn_evp.describe())
    dyn_evp.dyn_evp_b200_run_cgrid(dict(c.params, mode=0), got); dyn_evp.dyn_evp_b200_finalize()
    print("ndte", ndte)
    for n in abi.CFIELDS_INOUT + abi.CFIELDS_OUT:
        bad = np.argwhere(got[n][0] != ref[n][0])
        if len(bad):
            rows = sorted(set(bad[:, 0])); cols = sorted(set(bad[:, 1]))
            print(f"  {n}: {len(bad)} cells; rows {rows[:6]}..{rows[-3:]} cols {cols[:6]}..{cols[-3:]}")
