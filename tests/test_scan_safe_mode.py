"""First hardware run of the scan's safe mode (ticket tile ids, vrenb200_exclusive_scan_u32_ex; ADVICE r1: the look-back of the
chained scan must not depend on the order in which CTAs are dispatched).  Written after the round's GPU budget had been spent:
the kernel text is verified on the host for forward, reverse and shuffled CTA orders (tests/test_scan_emulation.py) and the
default instantiations are byte-identical in SASS to the ones the GPU suite verified, but the two TICKET instances have never
run on a B200.  Hence nothing selects them by default, the run happens in a subprocess and its outcome is recorded without
gating the suite (xfail, non-strict)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
SCRIPT = Path(__file__).with_name("run_scan_safe_mode.py")


@pytest.mark.xfail(strict=False, reason="never executed on hardware before this run (GPU budget of round 2 spent); verified on the host CTA emulator only")
def test_scan_safe_mode_first_hardware_run(vren):
    r = subprocess.run([sys.executable, str(SCRIPT)], capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ok 8 cases" in r.stdout
