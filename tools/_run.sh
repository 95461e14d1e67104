mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/r2f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests.log
grep -n "^E   \|passed\|failed\|^FAILED\|rc=" gpurun_out/r2f_tests.log | cut -c1-250 | head -20
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2f_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['link_frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
print({k:(v['gpoints_per_s'], v.get('opaque_pointer_gpoints_per_s'), v.get('example_fields_gpoints_per_s')) for k,v in d['variants_16384'].items()})
print(d['cahn_hilliard_4096']['ms_per_step'], d['cahn_hilliard_512']['ms_per_step'])
P
python -c "import __graft_entry__ as g; g.smoke()"
