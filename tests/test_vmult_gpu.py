"""Parity of the CUDA path (through the C ABI) against the oracle on identical meshes,
coefficients and vectors.  Tolerance: FP64 1e-12, FP32 1e-5, relative to ||ref||_inf
(the measure tests/matrix_free_kokkos/matrix_vector_device_common.h:175-180 prints)."""
import numpy as np
import pytest
import torch

import dealii_b200
from dealii_b200 import _lib as L
from oracle.mesh import HyperCubeMesh as OracleMesh
from oracle.mf_oracle import MatrixFreeOracle
from oracle import step64

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}


def varying(x):
    return 10.0 / (0.05 + 2.0 * (x ** 2).sum(1))


def step37_coef(x):
    return 1.0 / (0.05 + 2.0 * (x ** 2).sum(1))


def sine_deformation(amp):
    def f(x):
        return x + (amp * np.prod(np.sin(np.pi * x), axis=1))[:, None]
    return f


def shear(dim):
    A = np.eye(dim) + 0.15 * np.triu(np.ones((dim, dim)), 1)
    A[dim - 1, 0] = -0.1
    return lambda x: x @ A.T


def make_pair(dim, degree, refinements, number, geometry="cartesian", mass=None, grad=None,
              dirichlet=False, cpu_mf_semantics=False):
    deform = {"cartesian": None, "affine": shear(dim), "general": sine_deformation(0.08)}[geometry]
    om = OracleMesh(dim, degree, refinements=refinements, deformation=deform)
    constrained = om.boundary_dofs if dirichlet else None
    oracle = MatrixFreeOracle(om, grad_coefficient=grad, mass_coefficient=mass,
                              constrained_dofs=constrained)
    mf = dealii_b200.MatrixFree(number)
    l2g = om.l2g.astype(np.uint32)
    if cpu_mf_semantics and dirichlet:
        flag = np.zeros(om.n_dofs, dtype=bool)
        flag[om.boundary_dofs] = True
        l2g = np.where(flag[om.l2g], l2g | np.uint32(0x80000000), l2g).astype(np.uint32)
    mf.reinit(dim, degree, l2g, cell_vertices=om.cell_vertices, constrained_dofs=constrained,
              n_owned_dofs=om.n_dofs)
    gc = mf.evaluate_coefficients(grad) if callable(grad) else None
    mc = mf.evaluate_coefficients(mass) if callable(mass) else None
    op = dealii_b200.MatrixFreeOperator(
        mf, grad_coefficient=gc, mass_coefficient=mc,
        grad_constant=float(grad) if isinstance(grad, (int, float)) else 1.0,
        mass_constant=float(mass) if isinstance(mass, (int, float)) else 0.0)
    return om, oracle, mf, op


def run_vmult(op, mf, src):
    x = torch.from_numpy(src.astype(mf.np_dtype)).cuda()
    y = mf.initialize_dof_vector()
    y.fill_(123.0)  # vmult must overwrite
    op.vmult(y, x)
    torch.cuda.synchronize()
    return y.cpu().numpy().astype(np.float64)


def rel_err(a, ref):
    return np.abs(a - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("dim", [2, 3])
def test_laplace_cartesian_all_degrees(dim, degree, number):
    """BASELINE config 2 family (tests/performance/timing_matrix_free_kokkos.cc operator)."""
    r = {2: 3, 3: 2 if degree <= 4 else 1}[dim]
    om, oracle, mf, op = make_pair(dim, degree, r, number)
    assert mf.info.cell_kind == L.CELLS_CARTESIAN and mf.info.n_distinct_geometries == 1
    src = np.random.default_rng(degree).random(om.n_dofs)
    assert rel_err(run_vmult(op, mf, src), oracle.vmult(src)) < TOL[number]


@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("geometry", ["affine", "general"])
@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (2, 4), (3, 1), (3, 2), (3, 3), (3, 4),
                                        (3, 5), (3, 8)])
def test_helmholtz_varying_coefficient_deformed(dim, degree, geometry, number):
    """matrix_free_device_matrix_vector_0x family: (grad u, grad v) + (a(x) u, v) on affine
    (parallelepiped) and general (deformed) cells, zero Dirichlet boundary."""
    r = {2: 3, 3: 2 if degree <= 3 else 1}[dim]
    om, oracle, mf, op = make_pair(dim, degree, r, number, geometry=geometry, mass=varying,
                                   dirichlet=True)
    kind = {"affine": L.CELLS_AFFINE, "general": L.CELLS_GENERAL}[geometry]
    assert mf.info.cell_kind == kind
    src = np.random.default_rng(7).random(om.n_dofs)
    src[om.boundary_dofs] = 0.0
    got = run_vmult(op, mf, src)
    assert rel_err(got, oracle.vmult(src)) < TOL[number]
    # PMF and CPU-MatrixFree semantics coincide on such vectors (SURVEY note P1)
    assert rel_err(got, oracle.vmult_cpu_matrixfree(src)) < TOL[number]


@pytest.mark.parametrize("dim,degree", [(2, 2), (3, 2), (3, 4)])
def test_step37_variable_coefficient_laplace(dim, degree):
    """BASELINE config 1 operator: coefficient 1/(0.05+2|x|^2) on the gradient term
    (examples/step-37/step-37.cc:123), zero Dirichlet, CPU-MatrixFree semantics with
    non-zero src on constrained entries."""
    om, oracle, mf, op = make_pair(dim, degree, 2, "f64", grad=step37_coef, dirichlet=True,
                                   cpu_mf_semantics=True)
    src = np.random.default_rng(3).random(om.n_dofs)      # non-zero on the boundary
    x = torch.from_numpy(src).cuda()
    y = mf.initialize_dof_vector()
    mf.vmult(op.op, y, x)   # dst = 0; loop (constrained skipped); dst_c = src_c
    torch.cuda.synchronize()
    assert rel_err(y.cpu().numpy(), oracle.vmult_cpu_matrixfree(src)) < 1e-12


def test_constant_coefficients_and_cell_loop_adds():
    om, oracle, mf, op = make_pair(3, 3, 2, "f64", mass=10.0, grad=2.5)
    src = np.random.default_rng(5).random(om.n_dofs)
    ref = oracle.vmult(src)
    assert rel_err(run_vmult(op, mf, src), ref) < 1e-12
    x = torch.from_numpy(src).cuda()
    y = torch.ones(om.n_dofs, dtype=torch.float64, device="cuda")
    mf.cell_loop(op.op, x, y)      # Portable::MatrixFree::cell_loop adds into dst
    torch.cuda.synchronize()
    assert rel_err(y.cpu().numpy() - 1.0, ref) < 1e-12


def test_jacobian_input_equals_vertex_input():
    """Geometry handed over as the arrays Portable::MatrixFree stores (inv_jacobian, JxW)."""
    om = OracleMesh(3, 3, refinements=1, deformation=sine_deformation(0.1))
    oracle = MatrixFreeOracle(om, mass_coefficient=varying)
    mf = dealii_b200.MatrixFree("f64")
    mf.reinit(3, 3, om.l2g.astype(np.uint32), inv_jacobian=oracle.inv_jacobian, JxW=oracle.JxW,
              n_owned_dofs=om.n_dofs)
    mc = torch.from_numpy(oracle.mass_coef.reshape(-1)).cuda()
    op = dealii_b200.HelmholtzOperator(mf, mc)
    src = np.random.default_rng(11).random(om.n_dofs)
    assert rel_err(run_vmult(op, mf, src), oracle.vmult(src)) < 1e-12


@pytest.mark.parametrize("dim,degree,geometry", [(2, 2, "cartesian"), (3, 1, "general"),
                                                 (3, 3, "cartesian"), (3, 4, "affine")])
def test_compute_diagonal(dim, degree, geometry):
    """tests/matrix_free_kokkos/compute_diagonal_01.cc strategy: diagonal vs the assembled
    matrix (here: the oracle's diagonal and its sparse matrix)."""
    om, oracle, mf, op = make_pair(dim, degree, 1 if dim == 3 else 2, "f64", geometry=geometry,
                                   mass=varying, dirichlet=True)
    op.compute_diagonal()
    torch.cuda.synchronize()
    diag = op.diagonal.cpu().numpy()
    ref = oracle.compute_diagonal()
    assert rel_err(diag, ref) < 1e-12
    A = oracle.assemble_sparse()
    free = np.setdiff1d(np.arange(om.n_dofs), om.boundary_dofs)
    assert np.abs(diag[free] - A.diagonal()[free]).max() < 1e-11 * np.abs(ref).max()
    assert np.all(diag[om.boundary_dofs] == 1.0)


def test_linearity_and_symmetry_large():
    """Size-independent properties at a size the oracle does not need to touch."""
    mesh = dealii_b200.HyperCubeMesh(3, 4, refinements=4, deformation_amplitude=0.05)
    mf = dealii_b200.MatrixFree("f64").reinit_from_mesh(mesh)
    op = dealii_b200.LaplaceOperator(mf)
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.rand(mf.n_owned, dtype=torch.float64, device="cuda", generator=g)
    v = torch.rand(mf.n_owned, dtype=torch.float64, device="cuda", generator=g)
    Au, Av, Auv = (mf.initialize_dof_vector() for _ in range(3))
    op.vmult(Au, u)
    op.vmult(Av, v)
    op.vmult(Auv, 2.0 * u - 3.0 * v)
    scale = Au.abs().max().item()
    assert (Auv - (2.0 * Au - 3.0 * Av)).abs().max().item() < 1e-12 * scale
    assert abs((v @ Au).item() - (u @ Av).item()) < 1e-11 * abs((u @ Au).item())
    # Laplace annihilates constants (no Dirichlet rows here)
    one = torch.ones_like(u)
    op.vmult(Au, one)
    assert Au.abs().max().item() < 1e-11 * scale


def test_vmult_host_batch_matches_single_calls():
    """b200mf_vmult_host_batch: pipelined host-to-host vmults of several vectors == one by one."""
    om, oracle, mf, op = make_pair(3, 3, 2, "f64", mass=2.0)
    rng = np.random.default_rng(21)
    srcs = [torch.from_numpy(rng.random(om.n_dofs)).pin_memory() for _ in range(5)]
    dsts = [torch.empty(om.n_dofs, dtype=torch.float64).pin_memory() for _ in range(5)]
    op.vmult_host_batch([d.numpy() for d in dsts], [s.numpy() for s in srcs])
    for s, d in zip(srcs, dsts):
        assert rel_err(d.numpy(), oracle.vmult(s.numpy())) < 1e-12
