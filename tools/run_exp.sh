timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "--- bench"; timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['families'])
print('gpu_baseline', json.dumps(d.get('gpu_baseline'), indent=1)[:2500])
print('dataset', d.get('dataset_path')); print('latent', d.get('latent_kernel')); print('mask', {k:v for k,v in d.get('mask_iou',{}).items() if k!='note'})
print('cpu', d['cpu_baseline'])
PY
echo "--- reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
