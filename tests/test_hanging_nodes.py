"""Hanging nodes (BASELINE configs[3] ingredient): the device kernel's
resolve_hanging_nodes (ConstraintKinds masks + redirected indices,
matrix_free/portable_hanging_nodes_internal.h:124-459) against the conforming operator
C^T A C built geometrically by oracle/hanging.py."""
import numpy as np
import pytest

from oracle.hanging import HangingNodeMesh
from oracle.mf_oracle import MatrixFreeOracle

CASES_2D = [((2, 2), [(0, 0)]), ((3, 2), [(1, 0), (2, 1)]), ((2, 2), [(0, 0), (1, 1), (0, 1)])]
CASES_3D = [((2, 1, 1), [(1, 0, 0)]),                               # faces only
            ((2, 2, 1), [(1, 0, 0), (0, 1, 0), (1, 1, 0)]),         # faces + an edge-only neighbour
            ((2, 2, 2), [(0, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1)]),
            ((2, 2, 2), [(1, 1, 1)]),
            ((2, 2, 2), [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1)])]


def conforming_apply(hm, oracle, u):
    return hm.C.T @ oracle.cell_loop(hm.C @ u)


@pytest.mark.parametrize("dim,case", [(2, c) for c in CASES_2D] + [(3, c) for c in CASES_3D])
def test_oracle_hanging_mesh_is_consistent(dim, case):
    """Partition of unity of the constraints, Laplace annihilates constants, symmetry, and
    every mask is a valid ConstraintKinds state (hanging_nodes_internal.h: check())."""
    shape, refined = case
    hm = HangingNodeMesh(dim, 3, shape, refined)
    assert len(hm.hanging) > 0
    assert np.allclose(hm.C @ np.ones(hm.n_dofs), 1.0, atol=1e-13)
    lap = MatrixFreeOracle(hm.all_nodes_mesh())
    assert np.abs(conforming_apply(hm, lap, np.ones(hm.n_dofs))).max() < 1e-12
    rng = np.random.default_rng(0)
    u, v = rng.random(hm.n_dofs), rng.random(hm.n_dofs)
    assert abs(v @ conforming_apply(hm, lap, u) - u @ conforming_apply(hm, lap, v)) < 1e-11
    for m in hm.constraint_mask:
        faces, edges = (m >> 3) & 7, (m >> 6) & 7
        assert (m == 0) or faces or edges
        for e in range(3):            # an edge is never flagged next to a flagged face containing it
            if edges & (1 << e):
                assert not faces & (1 << ((e + 1) % 3)) and not faces & (1 << ((e + 2) % 3))
    assert (hm.constraint_mask != 0).sum() >= 2 ** (dim - 1)


@pytest.mark.gpu
@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("degree", [1, 2, 3, 4])
@pytest.mark.parametrize("dim,case", [(2, c) for c in CASES_2D] + [(3, c) for c in CASES_3D])
def test_device_resolve_hanging_nodes(dim, case, degree, number):
    import torch
    import dealii_b200
    shape, refined = case
    if dim == 3 and degree == 4 and len(refined) > 4:
        pytest.skip("oracle size")
    hm = HangingNodeMesh(dim, degree, shape, refined)
    oracle = MatrixFreeOracle(hm.all_nodes_mesh(), mass_coefficient=1.5)
    mf = dealii_b200.MatrixFree(number)
    mf.reinit(dim, degree, hm.l2g, cell_vertices=hm.cell_vertices, constraint_mask=hm.constraint_mask,
              n_owned_dofs=hm.n_dofs)
    op = dealii_b200.MatrixFreeOperator(mf, grad_constant=1.0, mass_constant=1.5)
    u = np.random.default_rng(degree).random(hm.n_dofs)
    x = torch.from_numpy(u.astype(mf.np_dtype)).cuda()
    y = mf.initialize_dof_vector()
    op.vmult(y, x)
    torch.cuda.synchronize()
    ref = conforming_apply(hm, oracle, u)
    err = np.abs(y.cpu().numpy().astype(np.float64) - ref).max() / np.abs(ref).max()
    assert err < (1e-12 if number == "f64" else 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("number", ["f64", "f32"])
def test_device_kernel_reproduces_reference_golden_vectors(number):
    """The reference's own known-answer test for this kernel
    (tests/matrix_free/hanging_node_kernels_01.output, all 234 blocks) through the C ABI."""
    import ctypes as C
    import os
    from dealii_b200 import _lib as L
    from oracle.hanging_kernel import golden_cases, parse_golden
    lib = L.load()
    golden = os.path.join(os.path.dirname(__file__), "golden", "hanging_node_kernels_01.output")
    groups = parse_golden(golden)
    k = 0
    for dim, degree, mask in golden_cases():
        for transpose in (0, 1):
            inp, ref, _ = groups[k]
            k += 1
            v = np.ascontiguousarray(inp, dtype=np.float64).copy()
            L.check(lib.b200mf_debug_resolve_hanging_nodes(dim, degree, L.F64 if number == "f64" else L.F32,
                                                           mask, transpose, v.ctypes.data_as(C.c_void_p)))
            assert np.allclose(v, ref, rtol=2e-5, atol=2e-5), (dim, degree, mask, transpose)
    assert k == 234
