"""step-64 (examples/step-64/step-64.cc) restated on top of the oracle pieces: Helmholtz
operator (grad u, grad v) + (a(x) u, v), a = 10/(0.05+2|x|^2) (:92-105), rhs = 1 (:560-575),
zero Dirichlet values, CG with tolerance 1e-12*|b| preconditioned by a degree-5
Chebyshev polynomial over Jacobi (smoothing_range 15, 10 Lanczos iterations, :595-622),
L2 norm of the solution by QGauss(p+2) accumulated in a Vector<float> (:660-680).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np

from .mesh import HyperCubeMesh
from .mf_oracle import MatrixFreeOracle
from .shape import gauss_points_weights, lagrange_values_and_derivatives
from .solvers import DiagonalMatrix, PreconditionChebyshev, solver_cg


def helmholtz_coefficient(x):
    return 10.0 / (0.05 + 2.0 * (x ** 2).sum(1))


def unit_rhs(oracle):
    """(1, phi_i) assembled cell by cell, constrained entries zero
    (constraints.distribute_local_to_global with homogeneous constraints)."""
    m, dim, n = oracle.mesh, oracle.dim, oracle.n
    S = oracle.shape.shape_values
    nq = oracle.n_q
    li = np.array([[(i // n ** d) % n for d in range(dim)] for i in range(n ** dim)])
    qi = np.array([[(q // nq ** d) % nq for d in range(dim)] for q in range(nq ** dim)])
    phi = np.ones((nq ** dim, n ** dim))
    for d in range(dim):
        phi *= S[li[:, d]][:, qi[:, d]].T
    loc = oracle.JxW @ phi                                   # (cells, npc)
    b = np.zeros(m.n_dofs)
    np.add.at(b, m.l2g.ravel(), loc.ravel())
    b[oracle.constrained] = 0.0
    return b


def l2_norm_of_solution(oracle, x, n_q=None):
    m, dim, n = oracle.mesh, oracle.dim, oracle.n
    n_q = n_q or m.degree + 2
    xq, wq = gauss_points_weights(n_q)
    V, _ = lagrange_values_and_derivatives(oracle.shape.support_points, xq)
    u = x[m.l2g].reshape((m.n_cells,) + (n,) * dim)
    for ax in range(1, dim + 1):
        u = np.moveaxis(np.tensordot(u, V, axes=([ax], [0])), -1, ax)
    h = (m.right - m.left) / m.N
    W = wq
    for _ in range(dim - 1):
        W = np.multiply.outer(W, wq)
    cell = np.sqrt((u ** 2 * (W * h ** dim)[None]).sum(tuple(range(1, dim + 1))))
    cell = cell.astype(np.float32).astype(np.float64)        # Vector<float> cellwise_norm
    return float(np.sqrt((cell ** 2).sum()))


def run_cycle(refinements, degree=3, dim=3, preconditioner="chebyshev"):
    m = HyperCubeMesh(dim, degree, refinements=refinements)
    o = MatrixFreeOracle(m, mass_coefficient=helmholtz_coefficient,
                         constrained_dofs=m.boundary_dofs)
    b = unit_rhs(o)
    inv_diag = 1.0 / o.compute_diagonal()
    if preconditioner == "chebyshev":
        P = PreconditionChebyshev(o.vmult, inv_diag, degree=5, smoothing_range=15.0,
                                  eig_cg_n_iterations=10, constrained_dofs=m.boundary_dofs)
    elif preconditioner == "jacobi":
        P = DiagonalMatrix(inv_diag)
    else:
        P = None
    out = solver_cg(o.vmult, b, P, tol=1e-12 * np.linalg.norm(b), max_steps=m.n_dofs)
    return dict(n_cells=m.n_cells, n_dofs=m.n_dofs, iterations=out["iterations"],
                norm=l2_norm_of_solution(o, out["x"]), x=out["x"], b=b, oracle=o,
                inv_diag=inv_diag, chebyshev=P if preconditioner == "chebyshev" else None)
