mkdir -p gpurun_out
CUSTEN_TILE_RELOAD=2 timeout 200 compute-sanitizer --tool memcheck python tools/weno_time.py 1024 example 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame" | head -20
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_weno_gpu.py tests/test_slab_c_gpu.py -q -m gpu -x 2>&1 | tail -4
for r in 0 2; do
CUSTEN_TILE_RELOAD=$r timeout 200 python tools/fun_time.py 16384
CUSTEN_TILE_RELOAD=$r timeout 200 python tools/fun_time.py 32768
CUSTEN_TILE_RELOAD=$r timeout 120 python tools/weno_time.py 16384 example
CUSTEN_TILE_RELOAD=$r timeout 120 python tools/weno_time.py 16384 random
done 2>&1 | grep "WENO\|fun" | tee gpurun_out/r2q_reload.log
