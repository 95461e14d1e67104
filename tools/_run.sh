mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_weno_gpu.py tests/test_slab_c_gpu.py tests/test_cahn_gpu.py tests/test_cahn_slab_gpu.py -q -m gpu -x 2>&1 | tail -3
filt() { grep -v "Host Frame" | grep "Error: \|SUMMARY\|passed\|failed\|smoke ok\|ms/step" | cut -c1-200 | head -8; }
{
for tool in memcheck racecheck; do
  echo "## $tool: python __graft_entry__.py --smoke"; timeout 420 compute-sanitizer --tool $tool --print-limit 4 python __graft_entry__.py --smoke 2>&1 | filt
  echo "## $tool: python tools/cahn_steps.py 256 3"; timeout 420 compute-sanitizer --tool $tool --print-limit 4 python tools/cahn_steps.py 256 3 2>&1 | filt
  for t in "tests/test_weno_gpu.py -k windows" "tests/test_slab_c_gpu.py -k static_input" "tests/test_cahn_slab_gpu.py -k 256-2-64-7" "tests/test_parity_gpu.py -k opaque_function_pointer_road"; do
    echo "## $tool: pytest $t"
    timeout 420 compute-sanitizer --tool $tool --print-limit 4 python -m pytest $t -q -m gpu -x 2>&1 | filt
  done
done
} > gpurun_out/r2w_sanitizer.log 2>&1
cat gpurun_out/r2w_sanitizer.log
