mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2r_bench_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2r_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['link_frac'], d['parity'])
print(d.get('cahn_hilliard_4096'))
print(d.get('halo_exchange'))
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2r_bench_n4.json 2> gpurun_out/r2r_bench_n4.err; echo "bench4 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2r_bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'])
print(d.get('cahn_hilliard_4096'))
P
