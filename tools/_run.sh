mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bits_differing'], d['e2e']['value'], d['e2e']['link_frac'], d['clocks']['sm_mhz'])
c=d.get('cahn_hilliard_4096'); print(c.get('ms_per_step'), c.get('parity'))
P
