#!/usr/bin/env python
"""Unified-memory probe (BASELINE.json config 3 the way the reference's callers run it: cudaMallocManaged grids).

Times cuStenCompute2DXYnp(DEVICE) on a 16384^2 grid, numTiles = 4, for combinations of
  who touched the pages first (CPU fill like the reference's examples / a GPU kernel), and
  what was done to the ranges before the first Compute (nothing, advice, one whole-range prefetch, re-homing).
Prints one JSON line per combination.  Not part of the product.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import custen_b200 as cs  # noqa: E402
import bench  # noqa: E402

ADV_SET_PREF, ADV_UNSET_PREF, ADV_SET_ACC, ADV_UNSET_ACC = 3, 4, 5, 6
CPU = -1


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    lib = cs.load()
    variant, tiles = "XYnp", 4
    coef, kw = bench.stencil_args(variant, n)
    cnt = n * n
    for first_touch in ("cpu", "gpu"):
        for prep in ("none", "coef_in_device_memory", "advise_pref_gpu", "prefetch_whole", "policy_pipeline"):
            m_in, m_out, m_w = lib.custen_managed_alloc(cnt * 8), lib.custen_managed_alloc(cnt * 8), lib.custen_managed_alloc(9 * 8)
            np.ctypeslib.as_array((ctypes.c_double * 9).from_address(m_w))[:] = coef
            w_arg = m_w
            if prep == "coef_in_device_memory":
                import torch
                w_dev = torch.from_numpy(coef).cuda()
                w_arg = w_dev
            if prep in ("advise_pref_gpu", "advise+prefetch"):
                for p in (m_in, m_out):
                    lib.custen_mem_advise(p, cnt * 8, ADV_SET_PREF, 0)
                    lib.custen_mem_advise(p, cnt * 8, ADV_SET_ACC, 0)
            t0 = time.perf_counter()
            if first_touch == "cpu":
                a = np.ctypeslib.as_array((ctypes.c_double * cnt).from_address(m_in))
                a[:] = 0.25
                a[::4097] = -0.5
                np.ctypeslib.as_array((ctypes.c_double * cnt).from_address(m_out))[:] = 0.0
                del a
            else:
                lib.custen_fill_hash(m_in, 0, n, n, 1, -1.0, 1.0)
                lib.custen_fill_hash(m_out, 0, n, n, 2, 0.0, 0.0)
                cs.device_synchronize()
            t_fill = time.perf_counter() - t0
            t0 = time.perf_counter()
            if prep in ("prefetch_whole", "advise+prefetch"):
                for p in (m_in, m_out):
                    lib.custen_mem_prefetch(p, cnt * 8, 0)
                cs.device_synchronize()
            if prep == "rehome_via_cpu":
                for p in (m_in, m_out):
                    lib.custen_mem_prefetch(p, cnt * 8, CPU)
                cs.device_synchronize()
                for p in (m_in, m_out):
                    lib.custen_mem_prefetch(p, cnt * 8, 0)
                cs.device_synchronize()
            t_prep = time.perf_counter() - t0
            cs.set_managed_policy(1 if prep == "policy_pipeline" else 0)
            st = cs.Stencil2D(variant, n, n, m_out, m_in, w_arg, numTiles=tiles, **kw)
            t0 = time.perf_counter()
            st.compute(cs.DEVICE)
            cs.device_synchronize()
            t_first = time.perf_counter() - t0
            st.compute(cs.DEVICE)
            cs.device_synchronize()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                st.compute(cs.DEVICE)
            cs.device_synchronize()
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"first_touch": first_touch, "prep": prep, "gpoints_per_s": round(cnt / dt / 1e9, 1),
                              "ms": round(dt * 1e3, 3), "mode": st.mode, "first_call_ms": round(t_first * 1e3, 1),
                              "fill_s": round(t_fill, 2), "prep_ms": round(t_prep * 1e3, 1)}), flush=True)
            st.destroy()
            cs.set_managed_policy(0)
            cs.device_synchronize()
            for p in (m_in, m_out, m_w):
                lib.custen_managed_free(p)


if __name__ == "__main__":
    main()
