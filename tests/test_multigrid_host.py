"""Host-side pieces of the multigrid that need no GPU: the 1D embedding matrix against a numpy Lagrange
oracle, and the nesting of the partitioned level meshes the transfer relies on (a rank's cells of level l+1
are the children of its own cells of level l, found through cell_morton_position)."""
import ctypes as C

import numpy as np
import pytest

from dealii_b200 import _lib as L
from oracle.shape import ShapeInfo


def lagrange(nodes, i, x):
    out = 1.0
    for j, xj in enumerate(nodes):
        if j != i:
            out *= (x - xj) / (nodes[i] - xj)
    return out


@pytest.mark.parametrize("degree", range(1, 9))
def test_prolongation_matrix_1d(degree):
    n, M = degree + 1, 2 * degree + 1
    P = np.zeros(M * n)
    L.check(L.load().b200mf_mg_prolongation_matrix_1d(degree, P.ctypes.data_as(C.POINTER(C.c_double))))
    P = P.reshape(M, n)
    gl = np.asarray(ShapeInfo(degree).support_points, dtype=np.float64)
    x = np.concatenate([0.5 * gl, 0.5 + 0.5 * gl[1:]])
    ref = np.array([[lagrange(gl, i, xx) for i in range(n)] for xx in x])
    assert np.abs(P - ref).max() < 1e-13
    # partition of unity, and the embedding is exact on the polynomials of the coarse space
    assert np.abs(P.sum(axis=1) - 1.0).max() < 1e-13
    for k in range(n):
        assert np.abs(P @ gl ** k - x ** k).max() < 1e-12
    # parent nodes that are child nodes are copied
    assert np.array_equal(P[0], np.eye(n)[0]) and np.array_equal(P[-1], np.eye(n)[-1])


@pytest.mark.parametrize("world,coarse", [(1, (2, 2, 2)), (2, (2, 2, 2)), (4, (2, 2, 2)), (8, (2, 2, 2)), (2, (2, 1, 1))])
def test_partitioned_levels_are_nested(world, coarse):
    from dealii_b200.distributed import PartitionedHyperCubeMesh
    dim, degree = 3, 2
    for rank in range(world):
        levels = [PartitionedHyperCubeMesh(dim, degree, l, world, rank, coarse=coarse, dirichlet_boundary=True,
                                           ghost_mode="touched") for l in range(0, 3)]
        for lc, lf in zip(levels[:-1], levels[1:]):
            assert lf.n_cells == lc.n_cells << dim
            pos_c, pos_f = lc.cell_morton_position.astype(np.int64), lf.cell_morton_position.astype(np.int64)
            assert sorted(pos_c) == list(range(lc.n_cells)) and sorted(pos_f) == list(range(lf.n_cells))
            inv_f = np.empty(lf.n_cells, dtype=np.int64)
            inv_f[pos_f] = np.arange(lf.n_cells)
            kids = inv_f[(pos_c[:, None] << dim) + np.arange(1 << dim)[None, :]]
            vc = lc.cell_vertices.reshape(lc.n_cells, 1 << dim, dim)
            vf = lf.cell_vertices.reshape(lf.n_cells, 1 << dim, dim)
            lo, hi = vc[:, 0, :], vc[:, -1, :]
            mid = 0.5 * (lo + hi)
            for k in range(1 << dim):
                off = np.array([(k >> d) & 1 for d in range(dim)])
                # child k (x fastest) occupies the k-th octant of its parent
                exp_lo = np.where(off[None, :] == 1, mid, lo)
                exp_hi = np.where(off[None, :] == 1, hi, mid)
                assert np.allclose(vf[kids[:, k], 0, :], exp_lo, atol=1e-14)
                assert np.allclose(vf[kids[:, k], -1, :], exp_hi, atol=1e-14)
