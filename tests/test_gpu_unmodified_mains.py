"""The reference's own example programs (c++/ex_laplace.cpp, c++/ex_laplace_parallel.cpp), compiled UNMODIFIED against
rchol_b200/cxx (`make refmains`, run by build() where /root/reference is mounted; the binaries travel in baseline/_ref),
solving on the GPU through the reference's constructor `pcg(A, b, tol, maxit, G, x, relres, itr)` (pcg.hpp:13-16) - no
`part` argument exists there, so this is the generic (partition-free) schedule of the blocked solve."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exe,args", [("ref_ex_laplace", ["-n", "16"]), ("ref_ex_laplace_parallel", ["-n", "16", "-t", "4"])])
def test_reference_example_runs_unchanged_on_the_gpu_path(exe, args):
    path = os.path.join(ROOT, "baseline", "_ref", exe)
    if not os.path.exists(path):
        pytest.skip("reference mains not built (make refmains needs /root/reference)")
    out = subprocess.run([path] + args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    itr = int(re.search(r"# CG iterations: (\d+)", out.stdout).group(1))
    relres = float(re.search(r"Relative residual: ([0-9.eE+-]+)", out.stdout).group(1))
    # the examples' own settings: tol 1e-6, maxit 200 (ex_laplace.cpp:36-37); rhs from the reference's unseeded rand
    assert 3 <= itr < 200 and relres <= 1e-6, out.stdout[-300:]
