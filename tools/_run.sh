mkdir -p gpurun_out
examples/bin/multi_gpu_stencil 1024 10 2
examples/bin/multi_gpu_stencil 4096 10 4
timeout 900 python -m pytest tests/test_dropin_examples_gpu.py -q -m gpu -x 2>&1 | tail -3
