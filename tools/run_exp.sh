run() { timeout 600 python bench.py --steps 300 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; }
echo "default"; run
echo "p1=8"; CVAE_BN_P1_BLOCKS_PER_SM=8 run
echo "fwd=16"; CVAE_BN_FWD_BLOCKS_PER_SM=16 run
echo "default"; run
