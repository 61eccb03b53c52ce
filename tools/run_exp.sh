timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_j.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['families'])
g=d.get('gpu_baseline',{}); print('gpu_baseline best', g.get('best_ms_per_step'), g.get('speedup_vs_best_stock_pytorch'))
PY
