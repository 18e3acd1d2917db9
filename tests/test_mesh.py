"""Host mesh/DoF generator of the engine (b200mf_mesh_*) vs the oracle: bit-exact."""
import numpy as np
import pytest

import dealii_b200
from oracle.mesh import HyperCubeMesh as OracleMesh

CASES = [(2, 1, 3, None), (2, 2, 2, None), (2, 3, 3, None), (2, 5, 1, None), (2, 8, 1, None),
         (3, 1, 2, None), (3, 2, 2, None), (3, 3, 2, None), (3, 4, 2, None), (3, 5, 1, None),
         (3, 8, 1, None), (3, 4, 0, None), (2, 2, None, 3), (3, 2, None, 3), (3, 4, None, 5),
         (3, 3, None, 1)]


@pytest.mark.parametrize("dim,degree,refinements,subdivisions", CASES)
def test_numbering_matches_oracle(dim, degree, refinements, subdivisions):
    ours = dealii_b200.HyperCubeMesh(dim, degree, refinements=refinements,
                                     subdivisions=subdivisions, left=-0.5, right=1.5,
                                     dirichlet_boundary=True)
    ref = OracleMesh(dim, degree, refinements=refinements, subdivisions=subdivisions,
                     left=-0.5, right=1.5)
    assert ours.n_cells == ref.n_cells and ours.n_dofs == ref.n_dofs
    np.testing.assert_array_equal(ours.l2g.astype(np.int64), ref.l2g)
    np.testing.assert_array_equal(ours.boundary_dofs.astype(np.int64), ref.boundary_dofs)
    np.testing.assert_allclose(ours.cell_vertices, ref.cell_vertices, rtol=0, atol=1e-15)


def test_constrained_bit_marks_boundary():
    m = dealii_b200.HyperCubeMesh(3, 2, refinements=1, dirichlet_boundary=True,
                                  mark_constrained_l2g=True)
    l2g = m.l2g
    flagged = np.unique(l2g[(l2g & 0x80000000) != 0] & 0x7FFFFFFF)
    np.testing.assert_array_equal(flagged, m.boundary_dofs)


def test_deformation_matches_oracle_formula():
    amp = 0.07
    ours = dealii_b200.HyperCubeMesh(3, 2, refinements=2, deformation_amplitude=amp)

    def deform(x):
        s = amp * np.prod(np.sin(np.pi * x), axis=1)
        return x + s[:, None]
    ref = OracleMesh(3, 2, refinements=2, deformation=deform)
    np.testing.assert_allclose(ours.cell_vertices, ref.cell_vertices, rtol=0, atol=1e-15)


def test_invalid_arguments_fail_loudly():
    with pytest.raises(dealii_b200.B200MFError):
        dealii_b200.HyperCubeMesh(4, 2, refinements=1)
    with pytest.raises(dealii_b200.B200MFError):
        dealii_b200.HyperCubeMesh(3, 9, refinements=1)


@pytest.mark.parametrize("name,degree,refinements,dirichlet", [("lex_q4", 4, 2, True), ("lex_q2", 2, 3, False)])
def test_lexicographic_numbering_equals_dof_renumbering_lexicographic(name, degree, refinements, dirichlet):
    """numbering="lexicographic" reproduces DoFRenumbering::lexicographic of the reference itself
    (tests/golden/ref/lex_*.npz were dumped by deal.II after that call)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref", name + ".npz"))
    m = dealii_b200.HyperCubeMesh(3, degree, refinements=refinements, numbering="lexicographic",
                                  dirichlet_boundary=dirichlet)
    assert np.array_equal(m.l2g, z["local_to_global"].reshape(-1, (degree + 1) ** 3))
    if dirichlet:
        assert np.array_equal(np.sort(m.boundary_dofs), z["constrained_dofs"])
