"""Oracle against the LIVE, unmodified reference on inputs outside the committed fixtures.  Runs only where the
reference checkout exists (the build container); everywhere else the committed fixtures (test_oracle_vs_golden.py)
are the pin."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "vae_nets.py")), reason="reference checkout not present")
def test_oracle_matches_live_reference():
    run = lambda: subprocess.run([sys.executable, os.path.join(HERE, "live_reference_check.py"), REF], capture_output=True, text=True, timeout=600)
    r = run()
    if r.returncode != 0:      # seen once in a while when the box is busy pushing a snapshot: keep the evidence, try once more
        with open("/tmp/cvae_live_reference_first_failure.log", "w") as fh:
            fh.write(f"rc={r.returncode}\n{r.stdout[-4000:]}\n{r.stderr[-8000:]}")
        r = run()
    assert r.returncode == 0 and "LIVE REFERENCE CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
