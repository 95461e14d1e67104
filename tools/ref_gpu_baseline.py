"""R-GPU rows of BASELINE.md: the reference's own CUDA kernels (sm_100 rebuild, oracle/_ref/libcusten_ref.so) timed on
device-resident managed buffers, next to the new engine on the same shapes.  Needs a GPU; writes JSON.

    python tools/ref_gpu_baseline.py gpurun_out/ref_gpu.json [n]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle_lib as ol  # noqa: E402

# block sizes: the examples' own where the sm_100 build can launch them (Fun kernels need > 64 registers, so a
# 32x32 block is "too many resources" on sm_100; 32x16 is used instead, 16x32 for XYpFun as in 2d_xy_p_fun.cu:39-40)
BLOCKS = {"Xp": (32, 32), "Xnp": (32, 32), "XnpFun": (32, 16), "Yp": (32, 32), "Ynp": (32, 32), "YpFun": (32, 16),
          "YnpFun": (32, 16), "XYp": (32, 32), "XYnp": (32, 32), "XYpFun": (16, 32), "XYnpFun": (16, 32)}


def main(out, n):
    res = {"n": n, "unit": "Gpoints/s", "rows": {}}
    for v, blk in BLOCKS.items():
        coef, kw = bench.stencil_args(v, n)
        kw.pop("numCoe", None)
        ms = ol.ref_time(v, n, n, coef, tiles=1, block=blk, warmup=3, iters=10, **kw)
        res["rows"][v] = {"block": blk, "ms": ms, "gpoints_per_s": (n * n / ms / 1e6) if ms > 0 else None}
        print(v, res["rows"][v], flush=True)
    # the solver's launch shape for the nonlinear term: 8x8 blocks (cuPentCahnADI.cu:44-45)
    coef, kw = bench.stencil_args("XYpFun", n)
    ms = ol.ref_time("XYpFun", n, n, coef, tiles=1, block=(8, 8), warmup=3, iters=10, **kw)
    res["rows"]["XYpFun_8x8"] = {"block": (8, 8), "ms": ms, "gpoints_per_s": n * n / ms / 1e6}
    print("XYpFun_8x8", res["rows"]["XYpFun_8x8"], flush=True)
    # 13th variant: WENO advection (32x32 blocks as examples/src/2d_xyWENOADV_p.cu:38-39; 32x16 if that cannot launch)
    lib = ol.ref_gpu()
    import ctypes
    lib.ref_weno_time.argtypes = [ctypes.c_int] * 6
    lib.ref_weno_time.restype = ctypes.c_double
    ms = lib.ref_weno_time(n, n, 32, 16, 3, 10)
    res["rows"]["XYWENOADVp"] = {"block": (32, 16), "ms": ms, "gpoints_per_s": n * n / ms / 1e6}
    print("XYWENOADVp", res["rows"]["XYWENOADVp"], flush=True)
    # config 5: the reference's GPU Cahn-Hilliard solver, ms per step
    for ncahn, steps in ((512, 50), (4096, 10)):
        c0 = np.random.default_rng(0).uniform(-0.1, 0.1, (ncahn, ncahn))
        r = ol.ref_cahn_run(c0, steps, 16 * np.pi, warm=3)  # 3 untimed steps: unified-memory pages settle on the GPU
        if r is not None:
            res["rows"][f"cahn_hilliard_{ncahn}"] = {"ms_per_step": r[1], "mpoint_steps_per_s": ncahn * ncahn / r[1] / 1e3}
            print("cahn", ncahn, res["rows"][f"cahn_hilliard_{ncahn}"], flush=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16384)
