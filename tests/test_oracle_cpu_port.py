"""The C restatement of the reference's CPU MatrixFree path (oracle/mf_cpu.c, the thing
bench.py times as cpu_baseline) against the numpy oracle, which is pinned on the reference's
golden vectors (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle.mesh import HyperCubeMesh
from oracle.mf_cpu import MatrixFreeCPU
from oracle.mf_oracle import MatrixFreeOracle


def _sine(x):
    return x + 0.05 * np.prod(np.sin(np.pi * x), axis=1, keepdims=True)


@pytest.mark.parametrize("dim,degree,refinements,deform", [
    (2, 1, 3, False), (2, 2, 3, False), (2, 5, 2, True), (2, 8, 1, False),
    (3, 1, 3, False), (3, 2, 2, True), (3, 3, 2, False), (3, 4, 2, False), (3, 4, 1, True),
    (3, 5, 1, False), (3, 6, 1, True), (3, 7, 1, False), (3, 8, 1, False)])
def test_c_port_matches_numpy_oracle(dim, degree, refinements, deform):
    mesh = HyperCubeMesh(dim, degree, refinements=refinements, deformation=_sine if deform else None)
    ref = MatrixFreeOracle(mesh)
    port = MatrixFreeCPU(dim, degree, mesh.l2g, mesh.cell_vertices, mesh.n_dofs)
    assert port.cartesian == (not deform)
    x = np.random.default_rng(7).random(mesh.n_dofs)
    a, b = ref.vmult(x), port.vmult(x)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


def test_c_port_ragged_batch_and_linearity():
    # 27 cells: not a multiple of the 8 SIMD lanes (the reference pads the last batch)
    mesh = HyperCubeMesh(3, 2, subdivisions=3)
    port = MatrixFreeCPU(3, 2, mesh.l2g, mesh.cell_vertices, mesh.n_dofs)
    rng = np.random.default_rng(3)
    x, y = rng.random(mesh.n_dofs), rng.random(mesh.n_dofs)
    lhs = port.vmult(2.0 * x - 3.0 * y)
    rhs = 2.0 * port.vmult(x) - 3.0 * port.vmult(y)
    assert np.abs(lhs - rhs).max() <= 1e-12 * np.abs(rhs).max()
    # constants are in the kernel of the Laplacian
    assert np.abs(port.vmult(np.ones(mesh.n_dofs))).max() < 1e-12
