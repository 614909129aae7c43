"""Does a TMA tensor-descriptor load issued by an independent stack (Triton's own cuTensorMapEncodeTiled + cp.async.bulk.tensor) run on
this box?  (scripts/micro/tma_probe*.cu are this repo's own attempts.)  Not part of the product."""
import torch
import triton
import triton.language as tl
from triton.tools.tensor_descriptor import TensorDescriptor


@triton.jit
def k(desc, out_ptr, BM: tl.constexpr, BN: tl.constexpr):
    x = desc.load([2, 64])
    offs = tl.arange(0, BM)[:, None] * BN + tl.arange(0, BN)[None, :]
    tl.store(out_ptr + offs, x)


for rows, dt in ((64, torch.float32), (4096, torch.float32), (64, torch.float64)):
    a = torch.arange(rows * 1024, device="cuda", dtype=dt).reshape(rows, 1024)
    out = torch.zeros(16 * 32, device="cuda", dtype=dt)
    try:
        d = TensorDescriptor.from_tensor(a, [16, 32])
        h = k[(1,)](d, out, 16, 32)
        torch.cuda.synchronize()
        print(rows, dt, "ok", out[:3].tolist(), "expect", [2 * 1024 + 64, 2 * 1024 + 65, 2 * 1024 + 66],
              "UTMALDG in sass:", "UTMALDG" in h.asm.get("sass", "") if hasattr(h, "asm") else "?")
    except Exception as e:  # noqa: BLE001
        print(rows, dt, "FAILED", type(e).__name__, str(e)[:300])
        break
