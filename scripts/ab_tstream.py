"""(As run in round 2, job Y -> profiles/r2_tstream_ab.txt; the issue-mode and L2-promotion switches it drove were removed from the
library with the variants that lost, EVP_B200_TSTREAM_ROWS is what is left.)
A/B of the tile-streaming kernel's variants on ONE box in ONE process (box-to-box differences of a few per cent -- power capping --
are larger than the differences between the variants): rows per block x who issues the box loads, with the launch-per-subcycle
streaming kernel as the control, interleaved over several rounds.  usage: python scripts/ab_tstream.py [workload] [ndte] [rounds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cice_b200 import abi, dyn_evp, synth  # noqa: E402

try:
    import pynvml
    pynvml.nvmlInit()
    H = pynvml.nvmlDeviceGetHandleByIndex(0)
    clock = lambda: pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)
    power = lambda: pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0
except Exception:  # noqa: BLE001
    clock = power = lambda: -1

wl = sys.argv[1] if len(sys.argv) > 1 else "p1deg"
ndte = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
c = synth.make_case(wl, ndte=ndte)
dyn_evp.dyn_evp_b200_init(c.grid)
bad = dyn_evp.set_metric(synth.scatter(c.X["HTN"], c.blocks), synth.scatter(c.X["HTE"], c.blocks), c.params["deltaminEVP"])
dyn_evp.upload(c.copy_fields())
# (kernel, rows, issue mode, L2 promotion of the box loads: 0 none, 2 128 B, 3 256 B)
configs = [("stream", None, None, None), ("tstream", 12, 0, 2), ("tstream", 12, 1, 2), ("tstream", 6, 0, 2), ("tstream", 6, 1, 2),
           ("tstream", 12, 1, 0), ("tstream", 12, 1, 3), ("tstream", 6, 1, 0)]
best = {}
for r in range(rounds):
    for kern, rows, issue, l2 in configs:
        if rows:
            os.environ["EVP_B200_TSTREAM_ROWS"] = str(rows)
            os.environ["EVP_B200_TSTREAM_ISSUE"] = str(issue)
            os.environ["EVP_B200_TSTREAM_L2"] = str(l2)
        p = dict(c.params, mode=abi.MODE_EXACT, kernel=abi.KERNEL_NAMES[kern])
        ms = []
        for _ in range(4):
            dyn_evp.subcycle(p)
            ms.append(dyn_evp.last_loop_ms())
        key = (kern, rows, issue, l2)
        best[key] = min(best.get(key, 1e9), min(ms[1:]))
        print("round %d %-8s rows %-4s issue %-4s L2 %-4s loop ms %s   sm %s MHz %.0f W" % (r, kern, rows, issue, l2, " ".join("%.3f" % m for m in ms), clock(), power()), flush=True)
print("best per configuration (ms per %d subcycles):" % ndte)
for k, v in best.items():
    print("  %-8s rows %-4s issue %-4s L2 %-4s %.3f   (%.1f us per subcycle)" % (k[0], k[1], k[2], k[3], v, v * 1e3 / ndte))
dyn_evp.dyn_evp_b200_finalize()
