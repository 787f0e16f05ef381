"""When baseline/run_reference.py has been run on a box with the real pancax + jax (it cannot run in the build image),
its dump tests/golden/loss_grad_jax.npz pins parity against a RUNNING jax: same keys as loss_grad.npz.  Skipped until then."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
JAX = os.path.join(G, "loss_grad_jax.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(JAX), reason="no jax dump (baseline/run_reference.py needs the real pancax + jax)")


def test_stand_in_goldens_match_the_running_jax_reference():
    a, b = np.load(os.path.join(G, "loss_grad.npz")), np.load(JAX)
    common = [k for k in b.files if k in a.files and not k.endswith("__meta")]
    assert common
    for k in common:
        x, y = np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)
        if not np.all(np.isfinite(y)):
            continue          # jax's d||f||/df is NaN at f = 0; the documented convention differs there
        assert np.linalg.norm((x - y).ravel()) <= 1e-10 * max(np.linalg.norm(y.ravel()), 1e-300), k
