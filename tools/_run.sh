mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/r2s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_tests.log
grep -n "^E   \|passed\|failed\|^FAILED\|rc=" gpurun_out/r2s_tests.log | cut -c1-250 | head -20
timeout 900 python bench.py > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2s_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
print({k:(v['gpoints_per_s'], v.get('opaque_pointer_gpoints_per_s'), v.get('random_fields_gpoints_per_s')) for k,v in d['variants_16384'].items()})
P
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
