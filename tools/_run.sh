mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_weno_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "weno or opaque" 2>&1 | tail -5
for g in 0 1 2 3 4; do for f in random example; do CUSTEN_WENO_GEOM=$g timeout 120 python tools/weno_time.py 16384 $f; done; done 2>&1 | grep WENO | tee gpurun_out/r2n_weno_geom.log
