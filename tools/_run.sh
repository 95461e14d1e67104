mkdir -p gpurun_out
filt() { grep -v "Host Frame" | grep "Error\|SUMMARY\|smoke ok\|ms/step" | cut -c1-200 | head -6; }
{
echo "## synccheck: python __graft_entry__.py --smoke"; timeout 60 compute-sanitizer --tool synccheck python __graft_entry__.py --smoke 2>&1 | filt
echo "## synccheck: python tools/cahn_steps.py 256 3"; timeout 60 compute-sanitizer --tool synccheck python tools/cahn_steps.py 256 3 2>&1 | filt
} > gpurun_out/r2_synccheck.log 2>&1
cat gpurun_out/r2_synccheck.log
