mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_weno_gpu.py -q -m gpu -x 2>&1 | tail -2
for rep in 1 2; do for f in example random; do timeout 120 python tools/weno_time.py 16384 $f; done; done 2>&1 | grep WENO | tee gpurun_out/r2_weno_carry.log
