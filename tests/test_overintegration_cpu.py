"""The numpy oracle with n_q_points_1d = fe_degree + 2 against the reference's own output for that quadrature
(tests/golden/ref_nq, ref_dump compiled with -DREF_NQ_EXTRA=1): pins the checker the GPU tests of
tests/test_overintegration_gpu.py use.  CPU only."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest

from oracle.mf_oracle import MatrixFreeOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_nq")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if "hanging" not in p)   # the hanging-node cases are the GPU tests' (engine vs deal.II directly)


@pytest.mark.parametrize("name", CASES)
def test_oracle_with_more_quadrature_points_matches_deal_ii(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    dim, p, Q, nc, nd = int(g["dim"]), int(g["degree"]), int(g["n_q_points_1d"]), int(g["n_cells"]), int(g["n_dofs"])
    mesh = SimpleNamespace(dim=dim, degree=p, n_cells=nc, n_dofs=nd,
                           l2g=g["local_to_global"].reshape(nc, (p + 1) ** dim).astype(np.int64),
                           cell_vertices=g["cell_vertices"].reshape(nc, 2 ** dim, dim))
    mass = g["coefficient"].reshape(nc, Q ** dim) if str(g["op"]) == "helmholtz_var" else 10.0
    o = MatrixFreeOracle(mesh, n_q_points_1d=Q, mass_coefficient=mass, constrained_dofs=g["constrained_dofs"])
    # geometry as Portable::MatrixFree stored it
    assert np.abs(o.JxW.ravel() - g["JxW"]).max() < 1e-14 * np.abs(g["JxW"]).max()
    ref = g["dst_portable_matrixfree"]
    got = o.vmult(g["src"])
    scale = np.maximum(np.abs(ref), 1e-2 * np.abs(ref).max())
    assert (np.abs(got - ref) / scale).max() < 1e-12
    refd = g["diagonal_portable"]
    assert (np.abs(o.compute_diagonal() - refd) / np.abs(refd)).max() < 1e-12
