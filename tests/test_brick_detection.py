"""Host logic of the brick detection (csrc/brick_setup.cpp) through b200mf_brick_probe: no GPU."""
import ctypes as C

import numpy as np
import pytest

import dealii_b200
from dealii_b200 import _lib as L


def probe(degree, l2g, n_dofs, mask=None):
    lib = L.load()
    l2g = np.ascontiguousarray(l2g, dtype=np.uint32)
    d = L.SetupDesc()
    d.dim, d.degree, d.n_q_points_1d, d.number = 3, degree, degree + 1, L.F64
    d.n_cells, d.n_owned_dofs, d.n_ghost_dofs = l2g.shape[0], n_dofs, 0
    d.local_to_global = l2g.ctypes.data_as(C.c_void_p)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint16)
        d.constraint_mask = mask.ctypes.data_as(C.c_void_p)
    nb, cpb, ncomp = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.check(lib.b200mf_brick_probe(C.byref(d), C.byref(nb), C.byref(cpb), C.byref(ncomp)))
    return nb.value, cpb.value, ncomp.value


@pytest.mark.parametrize("degree,refinements,b", [(1, 4, 16), (2, 3, 8), (3, 2, 4), (4, 3, 4), (5, 2, 2),
                                                  (6, 2, 2), (8, 1, 2)])
def test_every_window_of_a_morton_hyper_cube_is_a_brick(degree, refinements, b):
    mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=refinements)
    nb, cpb, ncomp = probe(degree, mesh.l2g, mesh.n_dofs)
    assert cpb == b ** 3 and nb * cpb == mesh.n_cells
    # complete nodes = lattice nodes that are not on a face shared with another brick
    per_dir = 2 ** refinements // b                   # bricks per direction
    L1 = b * degree + 1
    interior_1d = [L1 - (1 if i > 0 else 0) - (1 if i < per_dir - 1 else 0) for i in range(per_dir)]
    assert ncomp == sum(interior_1d) ** 3


def test_constrained_entries_and_dirichlet_flags_do_not_break_bricks():
    mesh = dealii_b200.HyperCubeMesh(3, 4, refinements=2, dirichlet_boundary=True, mark_constrained_l2g=True)
    nb, cpb, ncomp = probe(4, mesh.l2g, mesh.n_dofs)
    assert nb == 1 and ncomp == 15 ** 3               # the boundary nodes carry the "constrained" bit


def test_no_bricks_without_morton_blocks():
    mesh = dealii_b200.HyperCubeMesh(3, 4, subdivisions=4)          # lexicographic cell order
    assert probe(4, mesh.l2g, mesh.n_dofs)[0] == 0
    mesh = dealii_b200.HyperCubeMesh(3, 4, refinements=2)
    shuffled = mesh.l2g[np.random.default_rng(0).permutation(mesh.n_cells)]
    assert probe(4, shuffled, mesh.n_dofs)[0] == 0
    assert probe(4, mesh.l2g, mesh.n_dofs, mask=np.eye(1, mesh.n_cells, 3, dtype=np.uint16)[0] * 9)[0] == 0


def test_partial_coverage_and_foreign_references():
    """A second, separately numbered cube appended after the first: both are bricks; a cell that
    re-uses interior dofs of the first brick turns those dofs incomplete."""
    mesh = dealii_b200.HyperCubeMesh(3, 3, refinements=2)
    l2g = np.concatenate([mesh.l2g, mesh.l2g + mesh.n_dofs])
    nb, cpb, ncomp = probe(3, l2g, 2 * mesh.n_dofs)
    assert nb == 2 and ncomp == 2 * 13 ** 3
    extra = np.concatenate([l2g, mesh.l2g[:1]])      # one more cell = a copy of cell 0 of brick 0
    nb, cpb, ncomp2 = probe(3, extra, 2 * mesh.n_dofs)
    assert nb == 2 and ncomp2 == ncomp - 4 ** 3
