timeout 300 python tools/bottleneck_phases.py 2>&1 | tail -7
