mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_weno_gpu.py -q -m gpu -x 2>&1 | tail -4
