"""Time the WENO variant on an n x n grid: python tools/weno_time.py [n] [random|example]
(example = the fields of the reference's own program, examples/src/2d_xyWENOADV_p.cu:97-101)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custen_b200 as cs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
field = sys.argv[2] if len(sys.argv) > 2 else "random"
if field == "random":
    inp = torch.rand((n, n), device="cuda", dtype=torch.float64)
    u = torch.rand((n, n), device="cuda", dtype=torch.float64) * 2 - 1
    v = torch.rand((n, n), device="cuda", dtype=torch.float64) * 2 - 1
else:
    x = torch.arange(n, device="cuda", dtype=torch.float64) * (2 * torch.pi / n)
    inp = (torch.cos(x)[None, :] * torch.sin(x)[:, None]).contiguous()
    u = torch.sin(x)[:, None].expand(n, n).contiguous()
    v = (-torch.sin(x))[None, :].expand(n, n).contiguous()
out = torch.zeros_like(inp)
h = cs.cuSten_t()
cs.cuStenCreate2DXYWENOADVp(h, 0, 1, n, n, 32, 32, 1.0 / n, 1.0 / n, u, v, out, inp)
lib = cs.load()
hp = ctypes.addressof(h)
for _ in range(3):
    cs.cuStenCompute2DXYWENOADVp(h, 0)
cs.device_synchronize()
e0, e1 = lib.custen_event_create(), lib.custen_event_create()
lib.custen_event_record(e0, hp, 0)
for _ in range(10):
    cs.cuStenCompute2DXYWENOADVp(h, 0)
lib.custen_event_record(e1, hp, 0)
lib.custen_event_synchronize(e1)
ms = lib.custen_event_elapsed_ms(e0, e1) / 10
print("WENO", n, field, "ms", ms, "Gpt/s", n * n / ms / 1e6, cs.last_path(h))
cs.cuStenDestroy2DXYWENOADVp(h)
