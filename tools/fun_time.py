"""Time the Fun variants on an n x n grid, registered (inlined) and through the opaque device pointer (what an
unmodified cuSten program gets): python tools/fun_time.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import custen_b200 as cs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
inp = torch.rand((n, n), device="cuda", dtype=torch.float64) * 0.2 - 0.1
out = torch.zeros_like(inp)
res = {}
for v in ("XpFun", "YpFun", "XYpFun", "XYnpFun"):
    coef, kw = bench.stencil_args(v, n)
    tc = torch.from_numpy(np.ascontiguousarray(coef)).cuda()
    st = cs.Stencil2D(v, n, n, out, inp, tc, **kw)
    ms = bench.time_resident(cs, st, 10, 3) / 10
    cs.set_tuning(force_opaque=1)
    mso = bench.time_resident(cs, st, 10, 3) / 10
    cs.set_tuning()
    res[v] = (round(n * n / ms / 1e6, 1), round(n * n / mso / 1e6, 1))
    st.destroy()
print("fun (inlined, opaque) Gpt/s", n, res)
