"""The adaptive mesh generator (csrc/mesh_adaptive.cpp) against the reference ITSELF: on one rank
the numbering, the index lists with hanging-node redirection, the ConstraintKinds masks and the
constrained-dof list must be bit-identical to what deal.II built for the same mesh
(tests/golden/ref/c4_*.npz, dumped from Portable::MatrixFree by oracle/ref_drivers/ref_dump.cc:
hyper_cube refine_global(3), cells with centre within 0.3 of (0.5, 0.5, 0.5) refined once).
No GPU."""
import os

import numpy as np
import pytest

from dealii_b200.distributed import AdaptiveHyperCubeMesh

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name,degree", [("c4_q3_ball", 3), ("c4_q1_ball", 1)])
@pytest.mark.parametrize("friendly", [False, True])
def test_generator_reproduces_deal_ii_on_one_rank(name, degree, friendly):
    g = load(name)
    m = AdaptiveHyperCubeMesh(3, degree, 3, ball_radius=0.3, dirichlet_boundary=True,
                              brick_friendly_order=friendly)
    npc = (degree + 1) ** 3
    assert m.n_cells == int(g["n_cells"]) and m.n_dofs == int(g["n_dofs"])
    order = m.active_cell_index                       # emitted cell -> the reference's active cell
    assert sorted(order.tolist()) == list(range(m.n_cells))
    if not friendly:
        assert np.array_equal(order, np.arange(m.n_cells))
    ref_l2g = g["local_to_global"].reshape(-1, npc)[order]
    ref_mask = g["constraint_mask"][order]
    ref_vert = g["cell_vertices"].reshape(-1, 8, 3)[order]
    assert np.array_equal(m.constraint_mask, ref_mask)
    assert np.array_equal(m.l2g, ref_l2g)
    assert np.allclose(m.cell_vertices, ref_vert, atol=1e-15)
    assert np.array_equal(np.sort(m.constrained_dofs), np.sort(g["constrained_dofs"]))
    assert m.n_hanging_dofs == len(g["hanging_dofs"])
    assert m.n_masked_cells == int((g["constraint_mask"] != 0).sum())


def test_brick_friendly_order_groups_unmasked_blocks():
    m = AdaptiveHyperCubeMesh(3, 3, 4, ball_radius=0.3, brick_friendly_order=True)
    mask = m.constraint_mask
    masked = np.nonzero(mask)[0]
    assert len(masked) > 0 and masked.max() - masked.min() + 1 == len(masked), "masked cells are contiguous"
    # the cells in front of the masked range start with whole 4^3 blocks of equal-sized cells
    v = m.cell_vertices
    size = v[:, 1, 0] - v[:, 0, 0]
    first_fine = int(np.argmax(size < size[0]))
    assert first_fine % 64 == 0 and first_fine > 0


def test_two_rank_partition_is_consistent():
    """Interface dofs: owned by the lower rank, ghosts on the higher one, same support points."""
    m0 = AdaptiveHyperCubeMesh(3, 2, 2, n_ranks=2, rank=0, coarse=(2, 1, 1), ball_radius=0.3, want_coords=True)
    m1 = AdaptiveHyperCubeMesh(3, 2, 2, n_ranks=2, rank=1, coarse=(2, 1, 1), ball_radius=0.3, want_coords=True)
    assert m0.n_ghost == 0 and m1.n_ghost == (2 * 4 + 1) ** 2
    assert m0.n_owned + m1.n_owned == m0.n_global_dofs == m1.n_global_dofs
    assert int(m1.first_owned_global) == m0.n_owned
    # ghost k of rank 1 is global index ghost_global[k] on rank 0: same point in space
    c0, c1 = m0.dof_coords, m1.dof_coords
    gl = m1.ghost_global.astype(np.int64)
    assert np.all(gl < m0.n_owned)
    assert np.allclose(c1[m1.n_owned:], c0[gl], atol=1e-14)
    assert np.allclose(c1[m1.n_owned:, 0], 1.0)
