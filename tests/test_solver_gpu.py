"""CG / Jacobi / Chebyshev on the device vs the oracle (which reproduces the reference's
golden step-64 output): iteration counts within +-1, solutions to solver tolerance."""
import numpy as np
import pytest
import torch

import dealii_b200
from oracle import step64
from oracle.solvers import DiagonalMatrix as ODiag, solver_cg

pytestmark = pytest.mark.gpu

GOLDEN = {1: (343, 10, "0.0205439"), 2: (2197, 14, "0.0205269"), 3: (15625, 29, "0.0205261")}


def build(ref, number="f64"):
    o, m = ref["oracle"], ref["oracle"].mesh
    mf = dealii_b200.MatrixFree(number)
    mf.reinit(m.dim, m.degree, m.l2g.astype(np.uint32), cell_vertices=m.cell_vertices,
              constrained_dofs=m.boundary_dofs, n_owned_dofs=m.n_dofs)
    coef = mf.evaluate_coefficients(step64.helmholtz_coefficient)
    A = dealii_b200.HelmholtzOperator(mf, coef)
    return mf, A


@pytest.mark.parametrize("refinements", [1, 2, 3])
def test_step64_chebyshev_cg_matches_golden(refinements):
    """examples/step-64 (Q3): DoFs, CG iterations and solution norm of doc/results.dox."""
    ref = step64.run_cycle(refinements)
    n_dofs, its, norm = GOLDEN[refinements]
    assert ref["n_dofs"] == n_dofs and ref["iterations"] == its
    mf, A = build(ref)
    inv_diag = A.compute_diagonal()
    np.testing.assert_allclose(inv_diag.vector.cpu().numpy(), ref["inv_diag"], rtol=1e-12)
    prec = dealii_b200.PreconditionChebyshev(degree=5, smoothing_range=15.0,
                                             eig_cg_n_iterations=10, preconditioner=inv_diag)
    b = torch.from_numpy(ref["b"]).cuda()
    x = mf.initialize_dof_vector()
    control = dealii_b200.SolverControl(n_dofs, 1e-12 * float(np.linalg.norm(ref["b"])))
    res = dealii_b200.SolverCG(control).solve(A, x, b, prec)
    assert abs(control.last_step() - its) <= 1
    cheb = ref["chebyshev"].info
    assert abs(res.chebyshev_max_eigenvalue - cheb["max_eigenvalue"]) < 1e-8 * cheb["max_eigenvalue"]
    xs = x.cpu().numpy()
    assert np.abs(xs - ref["x"]).max() < 1e-9 * np.abs(ref["x"]).max()
    assert f"{step64.l2_norm_of_solution(ref['oracle'], xs):.6g}" == norm


@pytest.mark.parametrize("prec_kind", ["jacobi", "none"])
def test_cg_jacobi_and_identity_iteration_counts(prec_kind):
    ref = step64.run_cycle(2, preconditioner=prec_kind)
    mf, A = build(ref)
    inv_diag = A.compute_diagonal()
    b = torch.from_numpy(ref["b"]).cuda()
    x = mf.initialize_dof_vector()
    control = dealii_b200.SolverControl(5000, 1e-12 * float(np.linalg.norm(ref["b"])))
    dealii_b200.SolverCG(control).solve(A, x, b, inv_diag if prec_kind == "jacobi" else None)
    assert abs(control.last_step() - ref["iterations"]) <= 1
    assert np.abs(x.cpu().numpy() - ref["x"]).max() < 1e-9 * np.abs(ref["x"]).max()


def test_cg_nonzero_start_vector_and_host_entry_point():
    ref = step64.run_cycle(1, preconditioner="jacobi")
    mf, A = build(ref)
    inv_diag = A.compute_diagonal()
    o = ref["oracle"]
    x0 = np.random.default_rng(0).random(o.mesh.n_dofs)
    x0[o.mesh.boundary_dofs] = 0.0
    tol = 1e-12 * float(np.linalg.norm(ref["b"]))
    want = solver_cg(o.vmult, ref["b"], ODiag(ref["inv_diag"]), x0=x0, tol=tol, max_steps=1000)
    x = torch.from_numpy(x0).cuda()
    control = dealii_b200.SolverControl(1000, tol)
    dealii_b200.SolverCG(control).solve(A, x, torch.from_numpy(ref["b"]).cuda(), inv_diag)
    assert abs(control.last_step() - want["iterations"]) <= 1
    assert np.abs(x.cpu().numpy() - want["x"]).max() < 1e-9 * np.abs(want["x"]).max()


def test_no_convergence_is_reported():
    ref = step64.run_cycle(2, preconditioner="none")
    mf, A = build(ref)
    x = mf.initialize_dof_vector()
    control = dealii_b200.SolverControl(3, 1e-30)
    with pytest.raises(dealii_b200.B200MFError) as e:
        dealii_b200.SolverCG(control).solve(A, x, torch.from_numpy(ref["b"]).cuda(), None)
    assert e.value.code == dealii_b200._lib.ERR_NOCONVERGENCE and control.last_step() == 3


def test_cg_fp32():
    ref = step64.run_cycle(1, preconditioner="jacobi")
    mf, A = build(ref, "f32")
    inv_diag = A.compute_diagonal()
    b = torch.from_numpy(ref["b"].astype(np.float32)).cuda()
    x = mf.initialize_dof_vector()
    control = dealii_b200.SolverControl(1000, 1e-5 * float(np.linalg.norm(ref["b"])))
    dealii_b200.SolverCG(control).solve(A, x, b, inv_diag)
    assert np.abs(x.cpu().numpy() - ref["x"]).max() < 1e-4 * np.abs(ref["x"]).max()


@pytest.mark.parametrize("case", [0, 1, 2])
def test_solver_cg_interleave_golden(case):
    """The reference's known answers of tests/matrix_free/solver_cg_interleave.cc (golden output
    solver_cg_interleave.with_p4est=true.mpirun=3.output): CG + DiagonalMatrix on
    (grad u, grad v) + 10 (u, v), hyper_cube refine_global(6 - dim), rhs = 1/sqrt(N),
    preconditioner = 1 ./ (A rhs), tolerance 1e-2 ||rhs||: solution norms 240.33305 / 3609.9220 /
    1572.3941 and 51 / 61 / 39 operator applications inside the solver (3D cases: brick kernel)."""
    import os
    import re
    with open(os.path.join(os.path.dirname(__file__), "golden", "solver_cg_interleave.mpirun=3.output")) as f:
        txt = f.read()
    found = re.findall(r"CG solver with interleaving support\nDEAL::Norm of the solution: ([0-9.]+)\n"
                       r"DEAL::Number of calls to special vmult: (\d+)", txt)
    assert len(found) == 3
    dim, degree = [(2, 3), (3, 4), (3, 3)][case]
    norm, its = float(found[case][0]), int(found[case][1])
    mesh = dealii_b200.HyperCubeMesh(dim, degree, refinements=6 - dim)
    mf = dealii_b200.MatrixFree("f64").reinit_from_mesh(mesh)
    A = dealii_b200.MatrixFreeOperator(mf, grad_constant=1.0, mass_constant=10.0)
    n = mf.n_owned
    b = torch.full((n,), 1.0 / np.sqrt(n), dtype=torch.float64, device="cuda")
    d = mf.initialize_dof_vector()
    A.vmult(d, b)
    inv = torch.where(d != 0, 1.0 / d, torch.zeros_like(d))
    x = mf.initialize_dof_vector()
    control = dealii_b200.SolverControl(200, 1e-2 * float(b.norm()))
    dealii_b200.SolverCG(control).solve(A, x, b, dealii_b200.DiagonalMatrix(inv))
    assert abs(control.last_step() - its) <= 1
    assert abs(float(x.norm()) - norm) < 2e-5 * norm   # stopped at 1e-2: +-1 iteration moves the norm
    if control.last_step() == its:
        assert abs(float(x.norm()) - norm) < 1e-7 * norm


def test_chebyshev_with_a_prescribed_maximal_eigenvalue():
    """PreconditionChebyshev::AdditionalData::eig_cg_n_iterations = 0: no Lanczos estimate, the interval is
    [max_eigenvalue / smoothing_range, max_eigenvalue] (lac/precondition.h:2563-2568).  Handing the estimate
    of a first solve back as max_eigenvalue must reproduce that solve."""
    ref = step64.run_cycle(2)
    mf, A = build(ref)
    inv_diag = A.compute_diagonal()
    b = torch.from_numpy(ref["b"]).cuda()
    tol = 1e-12 * float(np.linalg.norm(ref["b"]))
    est = dealii_b200.PreconditionChebyshev(degree=4, smoothing_range=12.0, eig_cg_n_iterations=10, preconditioner=inv_diag)
    x0 = mf.initialize_dof_vector()
    c0 = dealii_b200.SolverControl(1000, tol)
    r0 = dealii_b200.SolverCG(c0).solve(A, x0, b, est)
    fixed = dealii_b200.PreconditionChebyshev(degree=4, smoothing_range=12.0, eig_cg_n_iterations=0,
                                              preconditioner=inv_diag, max_eigenvalue=r0.chebyshev_max_eigenvalue)
    x1 = mf.initialize_dof_vector()
    c1 = dealii_b200.SolverControl(1000, tol)
    r1 = dealii_b200.SolverCG(c1).solve(A, x1, b, fixed)
    assert r1.chebyshev_max_eigenvalue == pytest.approx(r0.chebyshev_max_eigenvalue, rel=1e-14)
    assert r1.chebyshev_min_eigenvalue == pytest.approx(r0.chebyshev_max_eigenvalue / 12.0, rel=1e-14)
    assert c1.last_step() == c0.last_step()
    assert torch.allclose(x1, x0, rtol=1e-9, atol=1e-12 * float(x0.abs().max()))
    # a different interval is a different preconditioner
    other = dealii_b200.PreconditionChebyshev(degree=4, smoothing_range=12.0, eig_cg_n_iterations=0,
                                              preconditioner=inv_diag, max_eigenvalue=3.0 * r0.chebyshev_max_eigenvalue)
    c2 = dealii_b200.SolverControl(1000, tol)
    dealii_b200.SolverCG(c2).solve(A, mf.initialize_dof_vector(), b, other)
    assert c2.last_step() != c0.last_step()
