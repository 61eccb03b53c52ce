timeout 200 python -m pytest tests/test_dp_gpu.py -m gpu -q -x 2>&1 | tail -15
for OV in 1 0; do
echo "--- bench 2 overlap=$OV"; CVAE_DP_OVERLAP=$OV timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$OV bench.py --gpus 2 --steps 200 --warmup 5 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])"
done
echo "--- bench 1"; timeout 200 python bench.py --steps 200 --warmup 5 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})"
