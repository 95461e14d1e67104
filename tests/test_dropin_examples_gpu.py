"""Drop-in proof on the GPU: the reference's OWN example programs (examples/src/*.cu, compiled by oracle/Makefile from
/root/reference against the reference's own cuSten.h) linked once against the reference library (<name>.ref) and once
against this repo's libcuSten.a (<name>.new) must print the same thing.  The programs print every grid point
(result, analytic answer, input), so this is an end-to-end comparison through the C++ API, unified memory, numTiles and
the HOST offload path, including user __device__ functions device-linked across translation units."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXDIR = os.path.join(ROOT, "oracle", "_ref", "examples")
EXAMPLES = ["2d_x_p", "2d_x_np", "2d_x_np_fun", "2d_y_p", "2d_y_np", "2d_y_p_fun", "2d_y_np_fun", "2d_xy_np", "2d_xy_np_fun",
            "2d_xy_p", "2d_xy_p_fun", "2d_xyWENOADV_p"]   # all the reference ships except 2d_x_p_fun (broken upstream)


@pytest.mark.parametrize("name", EXAMPLES)
def test_reference_example_prints_the_same_with_the_new_library(name):
    ref, new = os.path.join(EXDIR, name + ".ref"), os.path.join(EXDIR, name + ".new")
    if not (os.path.exists(ref) and os.path.exists(new)):
        pytest.skip("example binaries not built (need /root/reference at build time)")
    a = subprocess.run([ref], capture_output=True, text=True, timeout=300)
    b = subprocess.run([new], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0, a.stdout[-500:]
    assert b.returncode == 0, b.stdout[-500:]
    la, lb = a.stdout.splitlines(), b.stdout.splitlines()
    assert len(la) == len(lb)
    diff = [i for i, (x, y) in enumerate(zip(la, lb)) if x != y]
    if name == "2d_y_np" and diff:
        # reference defect: the bottom block row of kernel2DYnp is racy (2d_y_np_kernel.cu:229-241); lines are
        # printed row by row, 512 per row, so only the last 8 rows (BLOCK_Y) may differ
        assert min(diff) >= len(la) - 8 * 512
        return
    assert not diff, f"{len(diff)} of {len(la)} lines differ, first: {la[diff[0]]!r} vs {lb[diff[0]]!r}"


def test_user_program_with_registered_function():
    """examples/registered_fun.cu: a user translation unit linked against libcuSten.a; its __device__ function once as
    an opaque pointer, once registered with CUSTEN_REGISTER_FUN_XY.  The program exits 0 iff both give the same bits."""
    exe = os.path.join(ROOT, "examples", "bin", "registered_fun")
    if not os.path.exists(exe):
        pytest.skip("examples not built (make examples)")
    r = subprocess.run([exe, "2048"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-identical" in r.stdout


@pytest.mark.parametrize("n,steps,slabs", [(512, 7, 2), (1024, 10, 4), (2048, 5, 8)])
def test_cpp_program_moves_its_time_stepper_to_slabs(n, steps, slabs):
    """examples/multi_gpu_stencil.cu: a C++ translation unit with its own __device__ function, linked against libcuSten.a
    only.  It time-steps a field with the reference API on one GPU, then runs the same steps through custen_mg_* on
    y-slabs (over all GPUs of the box; several slabs share a GPU when there are fewer) and exits 0 iff no double differs."""
    exe = os.path.join(ROOT, "examples", "bin", "multi_gpu_stencil")
    if not os.path.exists(exe):
        pytest.skip("examples not built (make examples)")
    r = subprocess.run([exe, str(n), str(steps), str(slabs)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 of %d doubles differ" % (n * n) in r.stdout and "time-outs 0" in r.stdout


@pytest.mark.parametrize("n,steps", [(512, 20), (4096, 3)])
def test_reference_cahn_hilliard_driver_with_the_new_library(n, steps):
    """Zero-source-change proof for BASELINE config 5: the reference's GPU Cahn-Hilliard program (timing twin: driver,
    unregistered user function nonLinRHS, managed memory, cuBLAS transposes, BatchHyper + cuPentBatch solver) device-
    linked against this repo's libcuSten.a instead of the reference's library gives the same final field bit for bit.
    Only the two cuStenCompute2DXYp / XYpFun calls per step run through the new engine here; the solver is the reference's."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    c0 = np.random.default_rng(n).uniform(-0.1, 0.1, (n, n))
    ref = ol.ref_cahn_run(c0, steps, 16.0 * np.pi, warm=1)
    new = ol.ref_cahn_run(c0, steps, 16.0 * np.pi, warm=1, engine="new")
    if ref is None or new is None:
        pytest.skip("reference GPU solver builds not present (need /root/reference at build time)")
    assert ol.count_diff(new[0], ref[0]) == 0
    print(f"config 5 driver, n = {n}: {ref[1]:.3f} ms/step with the reference library, {new[1]:.3f} with the new one")
