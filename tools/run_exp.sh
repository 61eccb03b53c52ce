timeout 600 python -m pytest tests/test_conv_gemm.py -q -x -k "wa_" 2>&1 | grep -E "^FAILED|^E  |passed|failed" | cut -c1-200 | head -30
echo "--- default"; CVAE_COUNTERS=1 timeout 200 python tools/conv_bench.py 256 --wa 2>&1 | grep -E "wa\]|^[A-Z][0-9][fg]:|sum"
