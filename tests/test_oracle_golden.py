"""Pins the oracle against the golden vectors the reference ships (tests/golden/, copied
by tests/golden/make_golden.py from the reference's test outputs / tutorial docs)."""
import os
import re

import numpy as np
import pytest

from oracle.mesh import HyperCubeMesh, hierarchic_to_lexicographic
from oracle.mf_oracle import MatrixFreeOracle
from oracle.partitioner import PartitionerOracle
from oracle.shape import ShapeInfo, dealii_testing_rand
from oracle.solvers import DiagonalMatrix, PreconditionChebyshev
from oracle import step64

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _read(name):
    with open(os.path.join(GOLD, name)) as f:
        return f.read()


def test_testing_rand_is_glibc_type3():
    # tests/tests.h Testing::rand() == glibc rand() with seed 1
    assert dealii_testing_rand(3) == [1804289383, 846930886, 1681692777]


def test_partitioner_03_four_ranks():
    """tests/mpi/parallel_partitioner_03.cc with mpirun -np 4."""
    nproc, s = 4, 200
    owned, start = [], 0
    for r in range(nproc):
        owned.append((start, start + s - r))
        start += s - r
    ghosts = [1, 2, 13, s - 2, s - 1, s, s + 1, 2 * s, 2 * s + 1, 2 * s + 3]
    part = PartitionerOracle(owned, [ghosts] * nproc)
    text = "".join(part.format_like_reference_test(r) + "\n" for r in range(nproc))
    golden = _read("parallel_partitioner_03.mpirun=4.output")
    golden = golden[golden.index("**** proc 0"):]
    assert text.split() == golden.split()


def test_precondition_chebyshev_01():
    """tests/lac/precondition_chebyshev_01.cc: diagonal matrix diag(1..10), degree 4,
    smoothing_range 20, default Jacobi inner preconditioner, then identity."""
    size = 10
    diag = np.arange(1, size + 1, dtype=float)
    rmax = 2147483647
    vin = np.array(dealii_testing_rand(size), dtype=float) / rmax
    A = lambda x: diag * x
    lines = {}
    for line in _read("precondition_chebyshev_01.output").splitlines():
        m = re.match(r"DEAL::(.*?):\s+(.*)", line)
        if m:
            lines[m.group(1).strip()] = np.array([float(t) for t in m.group(2).split()])
    np.testing.assert_allclose(vin / diag, lines["Exact inverse"], atol=5.01e-3)
    p1 = PreconditionChebyshev(A, 1.0 / diag, degree=4, smoothing_range=2 * size)
    np.testing.assert_allclose(p1.vmult(vin), lines["Check  vmult orig"], atol=5.01e-3)
    p2 = PreconditionChebyshev(A, np.ones(size), degree=4, smoothing_range=2 * size)
    np.testing.assert_allclose(p2.vmult(vin), lines["Check  vmult diag"], atol=5.01e-3)


@pytest.mark.parametrize("cycle", [0, 1, 2])
def test_step64_golden_cycles(cycle):
    """examples/step-64/doc/results.dox:6-30: cells, DoFs, CG iterations, solution norm."""
    blocks = re.findall(
        r"Cycle (\d+)\s+Number of active cells:\s+(\d+)\s+Number of degrees of freedom:\s+(\d+)"
        r"\s+Solved in (\d+) iterations\.\s+solution norm:\s+([0-9.]+)",
        _read("step-64.results.txt"))
    gold = {int(b[0]): (int(b[1]), int(b[2]), int(b[3]), float(b[4])) for b in blocks}
    assert set(gold) == {0, 1, 2, 3}
    res = step64.run_cycle(cycle + 1)
    cells, dofs, its, norm = gold[cycle]
    assert (res["n_cells"], res["n_dofs"], res["iterations"]) == (cells, dofs, its)
    assert f"{norm:.6g}" == f"{res['norm']:.6g}"


@pytest.mark.slow
def test_step64_golden_cycle3():
    res = step64.run_cycle(4)
    assert (res["n_cells"], res["n_dofs"], res["iterations"]) == (4096, 117649, 58)
    assert f"{res['norm']:.6g}" == "0.0205261"


def test_step37_dof_counts():
    """tests/matrix_free/step-37.with_lapack=true.output: Q2, hyper_cube refined
    3-dim+cycle times... DoF counts 81/289/1089 (2D) and 125/729/4913 (3D)."""
    gold = re.findall(r"DEAL:(\d)d::Number of degrees of freedom: (\d+)", _read("step-37.output"))
    seen = {}
    for d, n in gold:
        seen.setdefault(int(d), []).append(int(n))
    for dim, counts in seen.items():
        ours = []
        r = 1
        while len(ours) < len(counts):
            nd = HyperCubeMesh(dim, 2, refinements=r).n_dofs
            if nd >= counts[0]:
                ours.append(nd)
            r += 1
        assert ours == counts


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 3)])
def test_matrix_vector_01_strategy(dim, degree):
    """tests/matrix_free_kokkos/matrix_free_device_matrix_vector_01.cc (+ _02 with
    Dirichlet constraints): matrix-free Helmholtz vmult vs assembled sparse matrix,
    hyper_cube refine_global(5-dim), coefficient 10, random src on unconstrained dofs;
    the golden files record differences of 1e-15..2e-14."""
    gold = re.findall(r"Norm of difference: ([0-9.e+-]+)",
                      _read("matrix_free_device_matrix_vector_01.output"))
    assert max(float(g) for g in gold) < 1e-13
    m = HyperCubeMesh(dim, degree, refinements=5 - dim)
    for constrained in (None, m.boundary_dofs):
        o = MatrixFreeOracle(m, mass_coefficient=10.0, constrained_dofs=constrained)
        rmax = 2147483647
        src = np.array(dealii_testing_rand(m.n_dofs), dtype=float) / rmax
        if constrained is not None:
            src[constrained] = 0.0
        ref = o.assemble_sparse() @ src
        out = o.vmult(src)
        diff = np.linalg.norm(out - ref) / np.abs(ref).max()
        assert diff < 1e-13
        np.testing.assert_allclose(o.vmult_cpu_matrixfree(src), out, rtol=0, atol=1e-14)


def test_h2l_is_permutation_and_vertices_first():
    for dim in (2, 3):
        for p in range(1, 7):
            h2l = hierarchic_to_lexicographic(dim, p)
            assert sorted(h2l) == list(range((p + 1) ** dim))
            assert h2l[0] == 0 and h2l[1] == p


def test_shape_info_properties():
    for p in range(1, 9):
        s = ShapeInfo(p)
        np.testing.assert_allclose(s.shape_values.sum(0), 1.0, atol=1e-14)
        np.testing.assert_allclose(s.shape_gradients.sum(0), 0.0, atol=1e-12)
        # collocation derivative differentiates polynomials of degree p exactly
        x = s.q_points
        np.testing.assert_allclose(s.shape_gradients_collocation.T @ x ** p,
                                   p * x ** (p - 1), atol=1e-12)
        # shape_gradients = shape_values * collocation derivative
        np.testing.assert_allclose(s.shape_values @ s.shape_gradients_collocation,
                                   s.shape_gradients, atol=1e-11)
        np.testing.assert_allclose(s.subface_interpolation_matrix.sum(1), 1.0, atol=1e-13)


def test_hanging_node_kernels_01():
    """tests/matrix_free/hanging_node_kernels_01.output: the reference applies its hanging-node
    interpolation (and the transpose) to values[i] = i for every face / edge mask, degrees 1-3,
    2D and 3D; the numpy restatement of the device kernel reproduces all 234 blocks."""
    from oracle.hanging_kernel import golden_cases, parse_golden, resolve_hanging_nodes
    groups = parse_golden(os.path.join(GOLD, "hanging_node_kernels_01.output"))
    cases = golden_cases()
    assert len(groups) == 2 * len(cases) == 234
    k = 0
    for dim, degree, mask in cases:
        W = ShapeInfo(degree).subface_interpolation_matrix
        for transpose in (False, True):
            inp, ref, opt = groups[k]
            k += 1
            assert np.array_equal(inp, np.arange((degree + 1) ** dim))
            out = resolve_hanging_nodes(inp, mask, dim, degree, W, transpose)
            assert np.allclose(out, ref, rtol=2e-5, atol=2e-5), (dim, degree, mask, transpose)
            assert np.allclose(out, opt, rtol=2e-5, atol=2e-5)


def solver_cg_interleave_golden():
    """(dim, degree, ||solution||_2, iterations) of tests/matrix_free/solver_cg_interleave.cc:
    test<2>(3), test<3>(4), test<3>(3); the l2 norm and the iteration count do not depend on the
    numbering or on the three ranks the reference ran on."""
    txt = _read("solver_cg_interleave.mpirun=3.output")
    norms = [float(x) for x in re.findall(r"CG solver with interleaving support\nDEAL::Norm of the solution: ([0-9.]+)", txt)]
    calls = [int(x) for x in re.findall(r"CG solver with interleaving support\nDEAL::Norm of the solution: [0-9.]+\nDEAL::Number of calls to special vmult: (\d+)", txt)]
    assert len(norms) == 3 and len(calls) == 3
    return [(2, 3, norms[0], calls[0]), (3, 4, norms[1], calls[1]), (3, 3, norms[2], calls[2])]


@pytest.mark.parametrize("case", [0, 1, 2])
def test_solver_cg_interleave_norms_and_iterations(case):
    """SolverCG + DiagonalMatrix on (grad u, grad v) + 10 (u, v), hyper_cube refine_global(6 - dim),
    no constraints, rhs = 1/sqrt(N), preconditioner = 1 ./ (A rhs), tolerance 1e-2 ||rhs||."""
    from oracle.solvers import solver_cg
    dim, degree, norm, its = solver_cg_interleave_golden()[case]
    m = HyperCubeMesh(dim, degree, refinements=6 - dim)
    o = MatrixFreeOracle(m, mass_coefficient=10.0)
    rhs = np.full(m.n_dofs, 1.0 / np.sqrt(m.n_dofs))
    d = o.cell_loop(rhs)
    r = solver_cg(o.cell_loop, rhs, DiagonalMatrix(np.where(d != 0, 1.0 / d, 0.0)),
                  tol=1e-2 * np.linalg.norm(rhs), max_steps=200)
    assert r["iterations"] == its
    assert abs(np.linalg.norm(r["x"]) - norm) < 1e-7 * norm
