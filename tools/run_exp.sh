timeout 600 python -m pytest tests/test_conv_gemm.py -q -x -k "wa_" 2>&1 | grep -E "^FAILED|^E  |passed|failed" | cut -c1-200 | head
echo "--- default"; CVAE_DEBUG=1 CVAE_COUNTERS=1 CVAE_WA_ONLY=E2f,E3f,E3g,E2g,D0f,D2f,D2g,D1g,D0g timeout 200 python tools/conv_bench.py 256 --wa 2>&1 | grep -E "^conv_wa E|wa\]|^[A-Z][0-9][fg]:|sum"
