mkdir -p gpurun_out
tools/weno_pow_probe.bin | tee gpurun_out/r2o_weno_pow_probe.log
timeout 600 python -m pytest tests/test_weno_gpu.py -q -m gpu -x 2>&1 | tail -5
for g in 0 1 4; do for f in random example; do CUSTEN_WENO_GEOM=$g timeout 120 python tools/weno_time.py 16384 $f; done; done 2>&1 | grep WENO | tee gpurun_out/r2o_weno_geom.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_tile -c 1 -s 3 -o gpurun_out/r2o_weno python tools/weno_time.py 4096 example > gpurun_out/r2o_weno_ncu.log 2>&1; tail -2 gpurun_out/r2o_weno_ncu.log
