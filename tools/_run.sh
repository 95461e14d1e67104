mkdir -p gpurun_out
free -g | head -2
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/r2l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_tests.log
grep -n "^E   \|passed\|failed\|^FAILED\|s call\|s setup" gpurun_out/r2l_tests.log | cut -c1-250 | head -40
timeout 300 python -m pytest tests/test_dropin_examples_gpu.py -q -m gpu -s -k cahn_hilliard_driver 2>&1 | grep "config 5" | tee gpurun_out/r2l_dropin_cfg5.log
timeout 900 python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2l_bench_n1.err
