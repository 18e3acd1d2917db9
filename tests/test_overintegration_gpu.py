"""n_q_points_1d > fe_degree + 1 (over-integration; Portable::MatrixFree only asserts n_q_points_1d >= fe_degree + 1,
matrix_free/portable_matrix_free.templates.h:1243): the engine's non-collocation cell kernel (csrc/overint.cu)
against the numpy oracle with the same quadrature -- Cartesian, deformed (general cells from vertices), cells
given by Jacobians, variable coefficients, 2D and 3D, FP64 and FP32 -- plus compute_diagonal and CG."""
import numpy as np
import pytest
import torch

import dealii_b200
from oracle.mesh import HyperCubeMesh as OracleMesh
from oracle.mf_oracle import MatrixFreeOracle
from oracle.solvers import DiagonalMatrix as ODiag, solver_cg

pytestmark = pytest.mark.gpu


def coefficient(x):
    return 1.0 / (0.05 + 2.0 * (x * x).sum(axis=1))


def mass(x):
    return 10.0 / (0.05 + 2.0 * (x * x).sum(axis=1))


def per_entry(a, ref, tol):
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * (1e-2 if tol < 1e-9 else 0.1))
    err = np.abs(a - ref) / scale
    assert err.max() < tol, f"per-entry error {err.max():.3e} at {err.argmax()}"


CASES = [  # dim, degree, extra points, refinements, deformation, variable, geometry input, number
    (3, 2, 1, 2, 0.0, False, "vertices", "f64"),
    (3, 2, 2, 2, 0.0, True, "vertices", "f64"),
    (3, 4, 1, 1, 0.0, True, "vertices", "f64"),
    (3, 3, 1, 2, 0.05, True, "vertices", "f64"),
    (3, 1, 2, 3, 0.05, False, "vertices", "f64"),
    (3, 2, 1, 2, 0.05, True, "jacobians", "f64"),
    (2, 3, 2, 3, 0.0, True, "vertices", "f64"),
    (2, 5, 1, 2, 0.04, True, "vertices", "f64"),
    (2, 2, 3, 3, 0.04, True, "jacobians", "f64"),
    (3, 3, 1, 2, 0.05, True, "vertices", "f32"),
    (3, 8, 1, 0, 0.0, False, "vertices", "f64"),
]


@pytest.mark.parametrize("dim,degree,extra,refinements,amp,variable,geometry,number", CASES)
def test_overintegrated_operator_matches_the_oracle(dim, degree, extra, refinements, amp, variable, geometry, number):
    Q = degree + 1 + extra
    defo = (lambda v: v + amp * np.prod(np.sin(np.pi * v), axis=1, keepdims=True)) if amp else None
    om = OracleMesh(dim, degree, refinements=refinements, deformation=defo)
    o = MatrixFreeOracle(om, n_q_points_1d=Q, grad_coefficient=coefficient if variable else None,
                         mass_coefficient=mass if variable else 3.0, constrained_dofs=om.boundary_dofs)
    mf = dealii_b200.MatrixFree(number)
    kw = dict(constrained_dofs=om.boundary_dofs, n_owned_dofs=om.n_dofs, n_q_points_1d=Q)
    if geometry == "vertices":
        mf.reinit(dim, degree, om.l2g.astype(np.uint32), cell_vertices=om.cell_vertices, **kw)
    else:
        mf.reinit(dim, degree, om.l2g.astype(np.uint32), inv_jacobian=o.inv_jacobian, JxW=o.JxW, **kw)
    assert mf.n_q_points == Q ** dim
    tdt = torch.float64 if number == "f64" else torch.float32
    gc = torch.from_numpy(o.grad_coef.reshape(-1)).to("cuda", tdt) if variable else None
    mc = torch.from_numpy(o.mass_coef.reshape(-1)).to("cuda", tdt) if variable else None
    A = dealii_b200.MatrixFreeOperator(mf, grad_coefficient=gc, mass_coefficient=mc,
                                       mass_constant=0.0 if variable else 3.0)
    if geometry == "vertices" and variable:
        # the engine's own quadrature points (evaluate_coefficients) are the oracle's
        q = mf.get_quadrature_points().reshape(-1, dim)
        assert np.abs(q - o.q_points.reshape(-1, dim)).max() < 1e-13
    src = np.random.default_rng(3).random(om.n_dofs) - 0.5
    x = torch.from_numpy(src).to("cuda", tdt)
    y = torch.zeros_like(x)
    A.vmult(y, x)
    tol = 1e-12 if number == "f64" else 1e-5
    per_entry(y.cpu().numpy().astype(np.float64), o.vmult(src), tol)
    diag = torch.zeros_like(x)
    mf.compute_diagonal(A.op, diag)
    per_entry(diag.cpu().numpy().astype(np.float64), o.compute_diagonal(), tol)
    if number == "f64" and om.n_dofs < 3000:
        # CG + Jacobi: iteration count of the oracle's CG on the same operator
        b = np.ones(om.n_dofs)
        b[om.boundary_dofs] = 0.0
        ref = solver_cg(o.vmult, b, ODiag(1.0 / o.compute_diagonal()), tol=1e-10 * np.linalg.norm(b), max_steps=2000)
        inv = A.compute_diagonal()
        xs = mf.initialize_dof_vector()
        control = dealii_b200.SolverControl(2000, 1e-10 * float(np.linalg.norm(b)))
        dealii_b200.SolverCG(control).solve(A, xs, torch.from_numpy(b).cuda(), inv)
        assert abs(control.last_step() - ref["iterations"]) <= 1
        assert np.abs(xs.cpu().numpy() - ref["x"]).max() < 1e-8 * np.abs(ref["x"]).max()


# ---- against the reference itself: ref_dump compiled with more quadrature points (tests/golden/ref_nq)
import glob
import os

GOLDEN_NQ = os.path.join(os.path.dirname(__file__), "golden", "ref_nq")
NQ_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_NQ, "*.npz")))


def load_nq(name):
    z = np.load(os.path.join(GOLDEN_NQ, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("dim", "degree", "n_dofs", "n_cells", "n_q_points_1d"):
        g[k] = int(g[k])
    return g


@pytest.mark.parametrize("name", NQ_CASES)
@pytest.mark.parametrize("geometry", ["vertices", "jacobians"])
def test_overintegrated_operator_matches_deal_ii(name, geometry):
    """The arrays Portable::MatrixFree built with QGauss(fe_degree + 2) go into the engine; vmult, diagonal and
    CG + Jacobi are compared with what Portable::MatrixFree / the CPU MatrixFree computed."""
    g = load_nq(name)
    dim, p, Q, nc = g["dim"], g["degree"], g["n_q_points_1d"], g["n_cells"]
    l2g = g["local_to_global"].reshape(nc, (p + 1) ** dim).astype(np.uint32)
    mf = dealii_b200.MatrixFree("f64")
    kw = dict(constrained_dofs=g["constrained_dofs"].astype(np.uint32), n_owned_dofs=g["n_dofs"], n_q_points_1d=Q)
    if g["constraint_mask"].any():
        # hanging nodes: ConstraintKinds masks and redirected index lists as Portable::MatrixFree built them
        kw["constraint_mask"] = g["constraint_mask"].astype(np.uint16)
    if geometry == "vertices":
        mf.reinit(dim, p, l2g, cell_vertices=g["cell_vertices"].reshape(nc, 2 ** dim, dim), **kw)
    else:
        mf.reinit(dim, p, l2g, inv_jacobian=g["inv_jacobian"].reshape(nc, Q ** dim, dim, dim),
                  JxW=g["JxW"].reshape(nc, Q ** dim), **kw)
    op = str(g["op"])
    coef = torch.from_numpy(g["coefficient"]).cuda() if op == "helmholtz_var" else None
    A = dealii_b200.MatrixFreeOperator(mf, mass_coefficient=coef, mass_constant=0.0 if coef is not None else 10.0)
    x = torch.from_numpy(g["src"]).cuda()
    y = torch.zeros_like(x)
    A.vmult(y, x)
    per_entry(y.cpu().numpy(), g["dst_portable_matrixfree"], 1e-12)
    diag = torch.zeros_like(x)
    mf.compute_diagonal(A.op, diag)
    per_entry(diag.cpu().numpy(), g["diagonal_portable"], 1e-12)
    if "cg_jacobi_iterations" in g:
        b = torch.ones_like(x)
        mf.set_constrained_values(0.0, b)
        inv = A.compute_diagonal()
        xs = mf.initialize_dof_vector()
        control = dealii_b200.SolverControl(2000, float(g["cg_tolerance"]))
        dealii_b200.SolverCG(control).solve(A, xs, b, inv)
        assert abs(control.last_step() - int(g["cg_jacobi_iterations"])) <= 1
