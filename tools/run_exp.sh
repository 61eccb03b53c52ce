timeout 900 python -m pytest tests/test_conv_gemm.py tests/test_zz_fullsize.py tests/test_vae_module.py tests/test_guard_bands.py -m gpu -q -x 2>&1 | tail -3
CVAE_COUNTERS=1 timeout 300 python tools/conv_bench.py --only E0f 2>&1 | tail -3
run() { timeout 600 python bench.py --steps 300 --warmup 10 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; }
echo "bench"; run; run
