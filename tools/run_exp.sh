timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "--- conv"; CVAE_COUNTERS=1 timeout 300 python tools/conv_bench.py --only E0f 2>&1 | tail -3; timeout 300 python tools/conv_bench.py --only D4g 2>&1 | tail -2
echo "--- wgrad"; timeout 300 python tools/wgrad_bench.py 256 2>&1 | tail -11
echo "--- bench"; timeout 600 python bench.py --steps 200 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families'])"
