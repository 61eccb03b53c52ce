timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "--- bench fused"; timeout 600 python bench.py --steps 200 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
echo "--- bench separate"; CVAE_NO_FUSED_BOTTLENECK=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
echo "--- timeline"; timeout 300 python tools/step_timeline.py 256 2>&1 | grep -v Warning | grep "bottleneck\|adam\|pack\|conv_gemm  *height=64\|sum of\|cvae_bn\|loss" | head -30
