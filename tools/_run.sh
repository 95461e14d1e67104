mkdir -p gpurun_out
for g in 0 1 2 3 0 1 2 3; do CUSTEN_BIG_GEOM=$g timeout 60 python - <<P
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench, custen_b200 as cs
out = {}
for n in (16384, 32768):
    inp = torch.rand((n, n), device="cuda", dtype=torch.float64) * 0.2 - 0.1
    o = torch.zeros_like(inp)
    coef, kw = bench.stencil_args("XYpFun", n)
    st = cs.Stencil2D("XYpFun", n, n, o, inp, torch.from_numpy(np.ascontiguousarray(coef)).cuda(), **kw)
    out[n] = round(n * n / (bench.time_resident(cs, st, 10, 3) / 10) / 1e6, 1)
    st.destroy(); del inp, o
print("big geom $g XYpFun", out)
P
done 2>&1 | grep "big geom" | tee gpurun_out/r2_tilebig_geom.log
