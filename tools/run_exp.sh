timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { timeout 600 python bench.py --steps 300 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
echo "pack first"; run
echo "critic first"; CVAE_CRITIC_FIRST=1 run
echo "pack first"; run
