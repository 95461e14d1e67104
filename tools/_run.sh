mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_weno_gpu.py -q -m gpu -x 2>&1 | tail -1
for rep in 1 2; do
for g in 0 1; do CUSTEN_WENO_GEOM=$g timeout 120 python tools/weno_time.py 16384 example | sed "s/^/geom $g /"; CUSTEN_WENO_GEOM=$g timeout 120 python tools/weno_time.py 16384 random | sed "s/^/geom $g /"; done
done 2>&1 | grep "WENO" | tee gpurun_out/r2y_geom.log
