mkdir -p gpurun_out
for rep in 1 2; do
for g in 0 1 2 3; do CUSTEN_WENO_GEOM=$g timeout 120 python tools/weno_time.py 16384 example | sed "s/^/geom $g /"; done
for m in 1 2; do timeout 120 python tools/cahn_steps.py 4096 40 2 0 $m; timeout 120 python tools/cahn_steps.py 2048 40 2 0 $m; timeout 120 python tools/cahn_steps.py 1024 40 2 0 $m; done
done 2>&1 | grep "WENO\|ms/step" | tee gpurun_out/r2x_geom.log
timeout 300 python -m pytest tests/test_cahn_gpu.py -q -m gpu -x -k "row_streaming or tolerance or full_size" 2>&1 | tail -2
