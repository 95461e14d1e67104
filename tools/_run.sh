mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cahn_slab_gpu.py tests/test_slab_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python tools/cahn_mg_bench.py 2>&1 | grep gpus | tee gpurun_out/r2z_cahn_mg.jsonl
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 --no-e2e > gpurun_out/r2z_bench_n$n.json 2> gpurun_out/r2z_bench_n$n.err; echo "bench$n rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2z_bench_n$n.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bits_differing'], d['clocks'])
c=d.get('cahn_hilliard_4096'); print(c.get('ms_per_step'), c.get('parity'), c.get('error'))
P
done
