timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { timeout 600 python bench.py --steps 300 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; }
echo "bench"; run; run
timeout 300 python tools/conv_bench.py 2>&1 | grep "E0f\|E1f\|E2f\|E3f"
