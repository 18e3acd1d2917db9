"""Geometric multigrid on the GPU.

Parity half: tests/golden/ref_gmg/*.npz were produced by the reference itself (oracle/ref_drivers/ref_gmg.cc:
step-37's Multigrid + MGTransferMatrixFree + PreconditionChebyshev smoothers run by the unmodified deal.II of
oracle/build_ref.sh).  The engine's hierarchy on the same problem must reproduce, per entry, the level
numbering, the level diagonals, one prolongation and one restriction, one V-cycle, and the CG iteration count /
solution.  Property half (no reference needed): restriction is the transpose of prolongation, prolongation
reproduces polynomials, the V-cycle is symmetric, iteration counts do not grow with the mesh.
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_gmg")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def step37_coefficient(x):
    return 1.0 / (0.05 + 2.0 * (x * x).sum(axis=1))


def hierarchy(dim, degree, refinements, number, variable, deformation=0.0, **kw):
    from dealii_b200 import GeometricMultigrid
    return GeometricMultigrid.for_hyper_cube(dim, degree, refinements, number=number, deformation_amplitude=deformation,
                                             coefficient=step37_coefficient if variable else None, **kw)


def system_operator(dim, degree, refinements, variable, deformation=0.0):
    from dealii_b200 import HyperCubeMesh, LaplaceOperator, MatrixFree
    mesh = HyperCubeMesh(dim, degree, refinements=refinements, dirichlet_boundary=True, mark_constrained_l2g=True,
                         deformation_amplitude=deformation)
    mesh.deformation = deformation
    mf = MatrixFree("f64")
    mf.reinit_from_mesh(mesh)
    coef = mf.evaluate_coefficients(step37_coefficient) if variable else None
    return mesh, mf, LaplaceOperator(mf, coef)


def unit_rhs(mesh, mf):
    """rhs_i = (phi_i, 1) with constrained entries zero: the mass operator applied to the constant 1."""
    from dealii_b200 import HyperCubeMesh, MatrixFree, MatrixFreeOperator
    plain = MatrixFree("f64")
    plain.reinit_from_mesh(HyperCubeMesh(mesh.dim, mesh.degree, refinements=int(round(np.log2(mesh.n_cells) / mesh.dim)),
                                         deformation_amplitude=getattr(mesh, "deformation", 0.0)))
    mass = MatrixFreeOperator(plain, grad_constant=0.0, mass_constant=1.0)
    one = torch.ones(mesh.n_dofs, dtype=torch.float64, device="cuda")
    b = torch.zeros_like(one)
    mass.vmult(b, one)
    mf.set_constrained_values(0.0, b)
    return b


def assert_per_entry(a, ref, tol, what):
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * 1e-2)
    err = np.abs(a - ref) / scale
    assert err.max() < tol, f"{what}: per-entry error {err.max():.3e} at entry {err.argmax()}"


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("dim", "degree", "refinements", "n_levels", "n_dofs", "cg_iterations", "level_number_bytes"):
        g[k] = int(g[k])
    g["variable"] = bool(int(g["variable_coefficient"]))
    g["deformation"] = float(g["deformation"]) if "deformation" in g else 0.0
    g["number"] = "f32" if g["level_number_bytes"] == 4 else "f64"
    return g


@pytest.mark.parametrize("name", CASES)
def test_multigrid_matches_reference(name):
    from dealii_b200 import HyperCubeMesh, SolverCG, SolverControl
    g = load(name)
    dim, p, r, f32 = g["dim"], g["degree"], g["refinements"], g["number"] == "f32"
    tol = 2e-4 if f32 else 1e-10
    mg = hierarchy(dim, p, r, g["number"], g["variable"], g["deformation"])
    assert mg.n_levels() == g["n_levels"]
    # ---- level numbering: the engine's mesh of level l is deal.II's level l
    for level in range(g["n_levels"]):
        key = f"level_l2g_{level}"
        if key in g:
            mesh = HyperCubeMesh(dim, p, refinements=level)
            assert np.array_equal(mesh.l2g.reshape(-1), g[key]), f"level {level} numbering"
    # ---- smoother data: eigenvalue estimates, degrees, level diagonals
    for level in range(g["n_levels"]):
        info = mg.level_info(level)
        # a Lanczos process that ran into convergence (fewer iterations than asked for: tiny levels) has
        # round-off-sized last coefficients; otherwise the estimates agree to solver precision
        early = level == 0 or int(g[f"eig_cg_iterations_{level}"]) < 10
        etol = 5e-3 if f32 else (1e-3 if early else 1e-6)
        assert abs(info.eig_cg_iterations - int(g[f"eig_cg_iterations_{level}"])) <= (2 if early else 0), f"level {level} Lanczos iterations"
        if float(g[f"eig_max_{level}"]) == 1.0 and float(g[f"eig_min_{level}"]) == 1.0:
            # one unknown: the reference's Lanczos CG stopped after one step without an estimate; a second
            # step on a round-off-sized residual gives the single eigenvalue 1 (times the safety factor)
            assert 1.0 - 1e-6 <= info.eig_min <= info.eig_max <= 1.2 + 1e-6
        else:
            assert info.eig_max == pytest.approx(float(g[f"eig_max_{level}"]), rel=etol), f"level {level} eig_max"
            assert info.eig_min == pytest.approx(float(g[f"eig_min_{level}"]), rel=max(etol, 1e-4)), f"level {level} eig_min"
        assert info.degree == int(g[f"cheb_degree_{level}"]), f"level {level} degree"
        key = f"level_inverse_diagonal_{level}"
        if key in g:
            assert_per_entry(mg.inverse_diagonal(level).cpu().numpy().astype(np.float64), g[key], tol, key)
    dtype = torch.float32 if f32 else torch.float64
    # ---- transfer between the two finest levels
    if "prolongate_src" in g:
        top = g["n_levels"] - 1
        src = torch.from_numpy(g["prolongate_src"]).to("cuda", dtype)
        dst = torch.zeros(g["prolongate_dst"].size, dtype=dtype, device="cuda")
        mg.prolongate(top, dst, src)
        assert_per_entry(dst.cpu().numpy().astype(np.float64), g["prolongate_dst"], tol, "prolongate")
        src = torch.from_numpy(g["restrict_src"]).to("cuda", dtype)
        dst = torch.zeros(g["restrict_dst"].size, dtype=dtype, device="cuda")
        mg.restrict_and_add(top, dst, src)
        assert_per_entry(dst.cpu().numpy().astype(np.float64), g["restrict_dst"], tol, "restrict_and_add")
    # ---- the solve of step-37: rhs = (phi_i, 1), CG preconditioned by one V-cycle
    mesh, mf, A = system_operator(dim, p, r, g["variable"], g["deformation"])
    b = unit_rhs(mesh, mf)
    if "rhs" in g:
        assert_per_entry(b.cpu().numpy(), g["rhs"], 1e-12, "rhs")
        z = torch.zeros_like(b)
        mg.vmult(z, b)
        assert_per_entry(z.cpu().numpy(), g["vcycle_of_rhs"], 5e-4 if f32 else 1e-8, "one V-cycle")
    assert float(b.norm()) == pytest.approx(float(g["rhs_l2"]), rel=1e-12)
    x = torch.zeros_like(b)
    control = SolverControl(100, 1e-12 * float(b.norm()))
    SolverCG(control).solve(A, x, b, mg)
    assert abs(control.last_step() - g["cg_iterations"]) <= (1 if f32 else 0), (control.last_step(), g["cg_iterations"])
    assert float(x.norm()) == pytest.approx(float(g["solution_l2"]), rel=1e-9)
    if "solution" in g:
        assert_per_entry(x.cpu().numpy(), g["solution"], 1e-8, "solution")


@pytest.mark.parametrize("dim,degree,refinements", [(3, 2, 3), (3, 4, 2), (3, 7, 1), (2, 3, 4), (2, 8, 2), (3, 8, 1)])
def test_restriction_is_transposed_prolongation(dim, degree, refinements):
    mg = hierarchy(dim, degree, refinements, "f64", False)
    top = mg.n_levels() - 1
    fine, coarse = mg.level_operators[top].mf, mg.level_operators[top - 1].mf
    gen = torch.Generator(device="cuda").manual_seed(5)
    u = torch.randn(coarse.n_owned, dtype=torch.float64, device="cuda", generator=gen)
    v = torch.randn(fine.n_owned, dtype=torch.float64, device="cuda", generator=gen)
    coarse.set_constrained_values(0.0, u)
    fine.set_constrained_values(0.0, v)
    Pu = torch.zeros_like(v)
    mg.prolongate(top, Pu, u)
    Rv = torch.zeros_like(u)
    mg.restrict_and_add(top, Rv, v)
    coarse.set_constrained_values(0.0, Rv)
    lhs, rhs = float(Pu @ v), float(u @ Rv)
    assert lhs == pytest.approx(rhs, rel=1e-12)
    # restrict_and_add adds
    Rv2 = Rv.clone()
    mg.restrict_and_add(top, Rv2, v)
    coarse.set_constrained_values(0.0, Rv2)
    assert torch.allclose(Rv2, 2 * Rv, rtol=1e-12, atol=1e-12)


def dof_coordinates(mesh):
    """Support point of every dof from the cell vertices (Cartesian cells) and the Gauss-Lobatto nodes."""
    from oracle.shape import ShapeInfo
    n, dim = mesh.degree + 1, mesh.dim
    gl = np.asarray(ShapeInfo(mesh.degree).support_points, dtype=np.float64)
    l2g = mesh.l2g.reshape(mesh.n_cells, -1).astype(np.int64)
    v = mesh.cell_vertices.reshape(mesh.n_cells, 2 ** dim, dim)
    lo, hi = v[:, 0, :], v[:, -1, :]
    idx = np.arange(n ** dim)
    coords = np.zeros((mesh.n_dofs, dim))
    for d in range(dim):
        t = gl[(idx // n ** d) % n]
        coords[l2g, d] = lo[:, None, d] + (hi - lo)[:, None, d] * t[None, :]
    return coords


@pytest.mark.parametrize("dim,degree", [(3, 2), (3, 4), (2, 5)])
def test_prolongation_reproduces_polynomials(dim, degree):
    from dealii_b200 import HyperCubeMesh
    r = 2
    mg = hierarchy(dim, degree, r, "f64", False)
    coarse_mesh, fine_mesh = HyperCubeMesh(dim, degree, refinements=r - 1), HyperCubeMesh(dim, degree, refinements=r)

    def poly(x):
        # degree p in every variable, zero on the boundary of (0,1)^dim like the constrained dofs
        out = np.ones(x.shape[0])
        for d in range(dim):
            out = out * x[:, d] * (1 - x[:, d]) * (1 + x[:, d]) ** max(degree - 2, 0)
        return out

    u = torch.from_numpy(poly(dof_coordinates(coarse_mesh))).cuda()
    ref = poly(dof_coordinates(fine_mesh))
    dst = torch.zeros(fine_mesh.n_dofs, dtype=torch.float64, device="cuda")
    mg.prolongate(r, dst, u)
    assert np.abs(dst.cpu().numpy() - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("number", ["f64", "f32"])
def test_vcycle_is_symmetric_and_iterations_are_mesh_independent(number):
    from dealii_b200 import SolverCG, SolverControl
    its = []
    for r in (2, 3, 4):
        mg = hierarchy(3, 2, r, number, True)
        mesh, mf, A = system_operator(3, 2, r, True)
        if r == 3:
            gen = torch.Generator(device="cuda").manual_seed(11)
            a = torch.randn(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=gen)
            c = torch.randn(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=gen)
            mf.set_constrained_values(0.0, a)
            mf.set_constrained_values(0.0, c)
            Ma, Mc = torch.zeros_like(a), torch.zeros_like(c)
            mg.vmult(Ma, a)
            mg.vmult(Mc, c)
            assert float(Ma @ c) == pytest.approx(float(a @ Mc), rel=1e-4 if number == "f32" else 1e-10)
            assert float(Ma @ a) > 0
        b = unit_rhs(mesh, mf)
        x = torch.zeros_like(b)
        control = SolverControl(100, 1e-10 * float(b.norm()))
        SolverCG(control).solve(A, x, b, mg)
        its.append(control.last_step())
        # the same answer as Jacobi-CG
        y = torch.zeros_like(b)
        SolverCG(SolverControl(5000, 1e-10 * float(b.norm()))).solve(A, y, b, A.compute_diagonal())
        assert float((x - y).norm()) < 1e-7 * float(y.norm())
    assert max(its) <= 8 and max(its) - min(its) <= 2, its


def test_chebyshev_estimate_with_the_reference_default_constraints():
    """PreconditionChebyshev in b200mf_cg_solve with AdditionalData::constraints left empty (the reference's
    default): the Lanczos estimate of the finest-level operator must be the one deal.II's level smoother got."""
    from dealii_b200 import PreconditionChebyshev, SolverCG, SolverControl
    g = load("gmg_q2_r2_f64")
    mesh, mf, A = system_operator(3, 2, 2, False)
    inv = A.compute_diagonal()
    b = unit_rhs(mesh, mf)
    out = {}
    for constraints in (False, True):
        prec = PreconditionChebyshev(degree=5, smoothing_range=15.0, eig_cg_n_iterations=10, preconditioner=inv,
                                     constraints=constraints)
        x = torch.zeros_like(b)
        control = SolverControl(200, 1e-10 * float(b.norm()))
        out[constraints] = SolverCG(control).solve(A, x, b, prec)
    assert out[False].chebyshev_max_eigenvalue == pytest.approx(float(g["eig_max_2"]), rel=1e-6)
    assert out[False].chebyshev_min_eigenvalue == pytest.approx(float(g["eig_min_2"]), rel=1e-4)
    # with the constrained entries zeroed the Krylov space differs: another (valid) estimate
    assert out[True].chebyshev_max_eigenvalue != out[False].chebyshev_max_eigenvalue
    assert abs(out[True].iterations - out[False].iterations) <= 2
