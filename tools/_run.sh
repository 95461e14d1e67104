mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cahn_gpu.py tests/test_cahn_slab_gpu.py -q -m gpu > gpurun_out/r2f_cahn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_cahn_tests.log
grep -n "^E   \|passed\|failed\|^FAILED" gpurun_out/r2f_cahn_tests.log | cut -c1-300 | head -40
for np in 64 128; do python tools/cahn_steps.py 4096 40 2 $np; done 2>&1 | tee gpurun_out/r2f_cahn_time.log
