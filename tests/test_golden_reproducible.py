"""The committed fixtures under tests/golden/ ARE outputs of the unmodified reference: regenerate them with
tests/golden/make_golden.py (which imports /root/reference through oracle/ref_shim.py) and compare array by array
(bit for bit for the MSM and chain fixtures and every integer array, 1e-5 for the torch model outputs).
Runs only where the reference tree exists (the build container); skipped on the GPU box."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.timeout(600)
@pytest.mark.skipif(not os.path.isdir("/root/reference/cy2path"), reason="the reference tree is not on this machine")
def test_committed_fixtures_are_reference_outputs(tmp_path):
    env = dict(os.environ, CY2P_GOLDEN_OUT=str(tmp_path))
    subprocess.run([sys.executable, os.path.join(GOLDEN, "make_golden.py")], check=True, env=env, cwd=ROOT,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    committed = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert committed and sorted(os.path.basename(f) for f in committed) == sorted(os.listdir(tmp_path))
    for f in committed:
        a, b = np.load(f), np.load(os.path.join(tmp_path, os.path.basename(f)))
        assert set(a.files) == set(b.files), f
        for k in a.files:
            x, y = a[k], b[k]
            assert x.dtype == y.dtype and x.shape == y.shape, (f, k)
            name = os.path.basename(f)
            if x.dtype.kind != "f" or name.startswith(("msm_", "chains_")):
                assert x.tobytes() == y.tobytes(), (name, k)  # bit for bit: scipy / numpy arithmetic, integer paths
            else:  # torch CPU reductions may reorder sums with the thread count (bit-equal on this container)
                assert np.allclose(x, y, rtol=1e-5, atol=1e-6, equal_nan=True), (name, k)
