"""GPU checks of the opt-in "two_sided" contrastive backend and its grouped kernels (tests/two_sided_checks.py), run in a child process
with a timeout. Written after round 2's GPU budget was spent: these kernels have NOT run on hardware yet — the file sorts last, the check
is a non-strict expected-pass, and the child process isolates the verified suite from a crash or a hang in them. The default backend
("gathered_grad") does not use these kernels."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(strict=False, reason="first hardware run at round end (written after the round's GPU budget was spent)")
def test_two_sided_backend_and_grouped_kernels_in_child_process():
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "two_sided_checks.py")], capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"two_sided checks timed out (child killed): {(e.stdout or b'')[-1500:]}")
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "TWO_SIDED ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
