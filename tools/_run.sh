mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cahn_gpu.py tests/test_cahn_slab_gpu.py -q -m gpu > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log
grep -n "^E   \|passed\|failed\|^FAILED" gpurun_out/r2j_tests.log | cut -c1-300 | head -40
for np in 64 128; do python tools/cahn_steps.py 4096 40 2 $np; done 2>&1 | tee gpurun_out/r2j_cahn_time.log
python tools/cahn_steps.py 512 200 2 2>&1 | tee -a gpurun_out/r2j_cahn_time.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2j_launches_cahn4096.csv python tools/cahn_steps.py 4096 3 2 128 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rhs_stream -s 4 -c 1 -o gpurun_out/r2j_k_rhs_stream -f python tools/cahn_steps.py 4096 3 2 128 > /dev/null 2>&1
