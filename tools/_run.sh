mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pent_part_gpu.py tests/test_cahn_gpu.py tests/test_cahn_slab_gpu.py -q -m gpu > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.log
grep -n "^E   \|passed\|failed\|^FAILED" gpurun_out/r2h_tests.log | cut -c1-300 | head -40
for np in 32 64 128 256; do python tools/cahn_steps.py 4096 40 2 $np; done 2>&1 | tee gpurun_out/r2h_cahn_time.log
for np in 64 128; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2h_launches_cahn4096_np$np.csv python tools/cahn_steps.py 4096 3 2 $np > /dev/null 2>&1
done
