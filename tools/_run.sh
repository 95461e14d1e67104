mkdir -p gpurun_out
python -m pytest tests/test_cahn_gpu.py -x -q -m gpu > gpurun_out/r2c_cahn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_cahn_tests.log
tail -15 gpurun_out/r2c_cahn_tests.log
for np in 32 64 128 256; do python tools/cahn_steps.py 4096 40 2 $np; done 2>&1 | tee gpurun_out/r2c_cahn_time.log
python tools/cahn_steps.py 4096 40 0 2>&1 | tee -a gpurun_out/r2c_cahn_time.log
python tools/cahn_steps.py 512 200 2 2>&1 | tee -a gpurun_out/r2c_cahn_time.log
python tools/cahn_steps.py 512 200 0 2>&1 | tee -a gpurun_out/r2c_cahn_time.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches_cahn4096.csv python tools/cahn_steps.py 4096 3 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2c_launches_cahn4096.csv 2>&1 | tail -30
