CVAE_DEBUG=1 timeout 300 python tools/wgrad_bench.py 256 --only E0w 2>&1 | grep -v "^conv_wgrad kind" | sort | uniq -c | tail -4
echo "--- sweep"; for cfg in "1 8" "3 4" "1 16" "7 2" "3 8" "1 24"; do set -- $cfg; echo "pw+$1 RA=$2"; CVAE_WG_STACK_PW=$1 CVAE_WG_STACK_RA=$2 timeout 120 python tools/wgrad_bench.py 256 --only E0w 2>&1 | tail -2 | head -1; done
echo "--- all"; timeout 300 python tools/wgrad_bench.py 256 2>&1 | tail -11
echo "--- tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "--- bench"; timeout 600 python bench.py --steps 200 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families'])"
