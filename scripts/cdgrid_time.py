"""time the CD-grid loop (gx1, ndte=600, 1 GPU) through the C ABI."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cice_b200 import abi, synth, dyn_evp  # noqa: E402

ndte = int(sys.argv[1]) if len(sys.argv) > 1 else 600
c = synth.make_cdcase("gx1", ndte=ndte)
f = c.copy_fields()
dyn_evp.dyn_evp_b200_init(c.grid)
dyn_evp.dyn_evp_b200_init_cgrid(c.cgrid)
for it in range(3):
    t0 = time.perf_counter()
    dyn_evp.dyn_evp_b200_run_cdgrid(dict(c.params, mode=abi.MODE_EXACT), f)
    t = time.perf_counter() - t0
    ms = dyn_evp.last_loop_ms()
    print(f"cdgrid gx1 ndte={ndte}: loop {ms:.3f} ms ({ms / ndte * 1e3:.2f} us/subcycle), call {t * 1e3:.2f} ms, "
          f"launches {dyn_evp.last_launches()}, {320 * 384 * ndte / (ms * 1e-3):.3e} cell-subcycles/s")
dyn_evp.dyn_evp_b200_finalize()
