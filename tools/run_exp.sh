timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "--- wgrad bench"; timeout 300 python tools/wgrad_bench.py 256 2>&1 | tail -11
echo "--- bench"; timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -3 gpurun_out/r02_bench_e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_e.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['families'])
g=d.get('gpu_baseline',{}); print('gpu_baseline', {k:(v.get('ms_per_step') or v.get('error')) for k,v in g.get('variants',{}).items()}, g.get('speedup_vs_best_stock_pytorch'))
PY
