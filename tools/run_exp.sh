timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "--- bench"; timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -3 gpurun_out/r02_bench_f.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['families'])
g=d.get('gpu_baseline',{}); print('gpu_baseline', {k:(v.get('ms_per_step') or v.get('error')) for k,v in g.get('variants',{}).items()}, g.get('speedup_vs_best_stock_pytorch'))
print('cpu', d['cpu_baseline']); print('dataset', d.get('dataset_path')); print('latent', d.get('latent_kernel',{}).get('frac')); print('mask', d.get('mask_iou',{}).get('frac'))
PY
echo "--- reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
