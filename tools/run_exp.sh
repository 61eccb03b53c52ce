for b in 3 2 3 2; do CVAE_DP_BUCKETS=$b timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 300 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('buckets=$b', d['value'], d['ms_per_step'])"; done
