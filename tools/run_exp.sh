timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -2 gpurun_out/r02_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','clocks')}); print('e2e', d['e2e']['value']); print('cfg4', d.get('cfg4'))
PY
