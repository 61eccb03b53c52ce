timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "--- bench"; timeout 600 python bench.py --steps 200 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "--- timeline"; timeout 300 python tools/step_timeline.py 256 2>&1 | grep "loss"
