"""The example scripts under examples/ (ports of the reference's examples/forward_problems scripts to the host
mirror) run end to end on a B200 with small sizes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,args,expect", [
    ("poisson_variational.py", ["60"], "final energy"),
    ("hyperelastic_3d.py", ["8", "10"], "wrote"),
    ("viscoelastic_3d.py", ["6", "11"], "10 "),
    ("inverse_path_dependent.py", ["8", "3"], "done 3"),
    ("poisson_collocation.py", ["12", "30"], "residual loss"),
])
def test_example_runs(cuda_device, tmp_path, script, args, expect):
    env = dict(os.environ, PANCAX_WORKDIR=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)] + args, capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert expect in out.stdout, out.stdout[-1000:]
