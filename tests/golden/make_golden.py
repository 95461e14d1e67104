"""Generate tests/golden/*.npz: outputs of the REFERENCE's own CUDA kernels (rebuilt for sm_100,
oracle/_ref/libcusten_ref.so) on seeded inputs.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy the files into tests/golden/

The inputs are stored next to the outputs, so the CPU tests can replay the oracle on them without a GPU.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import oracle_lib as ol  # noqa: E402


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for c in cases.GOLDEN_CASES:
        inp = np.random.default_rng(sum(map(ord, c["name"]))).uniform(-1, 1, size=(c["ny"], c["nx"]))
        out = np.full_like(inp, cases.SENTINEL)
        res = ol.ref_sweep(c["variant"], inp, out, c["coef"], tiles=c["tiles"], block=c["block"], **cases.case_kwargs(c))
        assert res is not None, c["name"]
        np.savez_compressed(os.path.join(outdir, c["name"] + ".npz"), variant=c["variant"], inp=inp, out=res,
                            coef=c["coef"], H=c["H"], L=c["L"], R=c["R"], V=c["V"], T=c["T"], B=c["B"],
                            fun=c["fun"] or "", tiles=c["tiles"], block=np.array(c["block"]))
        print("wrote", c["name"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE))
