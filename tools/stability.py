"""Run-to-run spread of a few shapes on 16384^2: python tools/stability.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402
import cases  # noqa: E402
import custen_b200 as cs  # noqa: E402

n = 16384
shapes = {
    "XYp3x3": ("XYp", cases.weights_cross_xy(1e-3, 1e-3), dict(H=3, L=1, R=1, V=3, T=1, B=1)),
    "XYp5x5": ("XYp", cases.weights_biharmonic(0.01), dict(H=5, L=2, R=2, V=5, T=2, B=2)),
    "Xp9": ("Xp", cases.weights_d2_8th(1e-3), dict(H=9, L=4, R=4)),
    "Yp9": ("Yp", cases.weights_d2_8th(1e-3), dict(V=9, T=4, B=4)),
    "XYpFun": ("XYpFun", cases.weights_laplace5(0.2), dict(H=3, L=1, R=1, V=3, T=1, B=1, fun="cubic_xy")),
}
for name, (v, coef, kw) in [(k, shapes[k]) for k in (sys.argv[1:] or shapes)]:
    res = []
    for rep in range(4):
        inp = torch.rand((n, n), device="cuda", dtype=torch.float64)
        out = torch.zeros_like(inp)
        tc = torch.from_numpy(np.ascontiguousarray(coef)).cuda()
        st = cs.Stencil2D(v, n, n, out, inp, tc, **kw)
        for sub in range(3):
            ms = bench.time_resident(cs, st, 15, 3) / 15
            res.append(n * n / ms / 1e6)
        st.destroy()
        del inp, out
    print(name, "min %.1f max %.1f mean %.1f Gpt/s" % (min(res), max(res), sum(res) / len(res)), [round(r) for r in res])
