"""Parity of the bulk brick kernel (csrc/bulk_kernel.cuh: own ranges moved by bulk async copies,
pattern tables instead of per-node index maps, first-toucher-stores instead of memset + atomics)
against the oracle (per entry) and against the index-map brick kernel on the same inputs.
FP64: |a - ref| <= 1e-12 * max(|ref_i|, 1e-2 ||ref||_inf) per entry (north star: 1e-12 per entry;
entries that cancel to ~0 are measured against a fraction of the vector's scale);
FP32: 1e-5 * max(|ref_i|, 0.1 ||ref||_inf)."""
import numpy as np
import pytest
import torch

import dealii_b200
from oracle.mesh import HyperCubeMesh as OracleMesh
from oracle.mf_oracle import MatrixFreeOracle

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-12, "f32": 1e-5}


def assert_per_entry(a, ref, tol):
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * (1e-2 if tol < 1e-9 else 0.1))
    err = np.abs(a - ref) / scale
    assert err.max() < tol, f"per-entry error {err.max():.3e} at {err.argmax()}"


def brick_refinements(degree):
    return 4 if degree == 1 else 3 if degree == 2 else 2 if degree <= 4 else 1


def make(degree, refinements, number, dirichlet=False, cpu_mf=False, mass=0.0, grad=1.0):
    om = OracleMesh(3, degree, refinements=refinements)
    constrained = om.boundary_dofs if dirichlet else None
    oracle = MatrixFreeOracle(om, grad_coefficient=grad, mass_coefficient=mass if mass else None,
                              constrained_dofs=constrained)
    l2g = om.l2g.astype(np.uint32)
    if cpu_mf:
        flag = np.zeros(om.n_dofs, dtype=bool)
        flag[om.boundary_dofs] = True
        l2g = np.where(flag[om.l2g], l2g | np.uint32(0x80000000), l2g).astype(np.uint32)
    mf = dealii_b200.MatrixFree(number)
    mf.reinit(3, degree, l2g, cell_vertices=om.cell_vertices, constrained_dofs=constrained,
              n_owned_dofs=om.n_dofs)
    op = dealii_b200.MatrixFreeOperator(mf, grad_constant=grad, mass_constant=mass)
    return om, oracle, mf, op


@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("degree,extra", [(1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1), (4, 0), (4, 1),
                                          (5, 0), (5, 1), (6, 1), (7, 1), (8, 1)])
def test_bulk_vmult_matches_oracle_per_entry(degree, extra, number):
    om, oracle, mf, op = make(degree, brick_refinements(degree) + extra, number)
    info = mf.bulk_info()
    assert info["usable"] == 1 and info["n_zero"] == 0 and info["n_general_cells"] == 0
    src = np.random.default_rng(degree).random(om.n_dofs)
    x = torch.from_numpy(src.astype(mf.np_dtype)).cuda()
    y = mf.initialize_dof_vector()
    ref = oracle.vmult(src)
    # every brick path, whatever the setup's own measurement chose: 2 = bulk tables, 1 = coloured
    # launches (both without a memset of dst), 0 = index maps + memset + atomics
    for path in (2, 1, 0):
        assert mf.select_brick_path(path) == path
        results = []
        for rep in range(3):                    # repeated launches: flag epochs, ticket re-arming
            y.fill_(float("nan"))               # vmult must write every entry
            op.vmult(y, x)
            torch.cuda.synchronize()
            results.append(y.clone())
            assert_per_entry(y.cpu().numpy().astype(np.float64), ref, TOL[number])
        if path == 1:                           # coloured launches are bit-reproducible
            assert torch.equal(results[0], results[1]) and torch.equal(results[1], results[2])


@pytest.mark.parametrize("degree,extra", [(4, 1), (3, 1), (2, 1), (6, 1)])
def test_bulk_helmholtz_dirichlet_both_constraint_semantics(degree, extra):
    r = brick_refinements(degree) + extra
    for cpu_mf in (False, True):
        om, oracle, mf, op = make(degree, r, "f64", dirichlet=True, cpu_mf=cpu_mf, mass=10.0, grad=2.5)
        assert mf.bulk_info()["usable"] == 1
        src = np.random.default_rng(3).random(om.n_dofs)
        if not cpu_mf:
            src[om.boundary_dofs] = 0.0
        x = torch.from_numpy(src).cuda()
        y = mf.initialize_dof_vector()
        ref = oracle.vmult_cpu_matrixfree(src) if cpu_mf else oracle.vmult(src)
        for path in (2, 1):
            assert mf.select_brick_path(path) == path
            y.fill_(float("nan"))
            op.vmult(y, x)
            torch.cuda.synchronize()
            assert_per_entry(y.cpu().numpy(), ref, 1e-12)


@pytest.mark.parametrize("number", ["f64", "f32"])
def test_bulk_equals_index_map_kernel_on_a_large_mesh(number):
    """2.1 M dofs (512 bricks, ~130 patterns): both kernels apply the same macro-element operator, so
    they agree to round-off of the summation order at brick faces; plus linearity of the bulk path."""
    mesh = dealii_b200.HyperCubeMesh(3, 4, refinements=5)
    mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
    op = dealii_b200.LaplaceOperator(mf)
    g = torch.Generator(device="cuda").manual_seed(1)
    x1 = torch.rand(mesh.n_dofs, dtype=mf.torch_dtype, device="cuda", generator=g)
    x2 = torch.rand(mesh.n_dofs, dtype=mf.torch_dtype, device="cuda", generator=g)
    y1, y2, y12, yl = (mf.initialize_dof_vector() for _ in range(4))
    assert mf.enable_bulk(True)
    op.vmult(y1, x1)
    op.vmult(y2, x2)
    op.vmult(y12, 2.0 * x1 - 3.0 * x2)
    mf.enable_bulk(False)
    op.vmult(yl, x1)
    mf.enable_bulk(True)
    torch.cuda.synchronize()
    tol = TOL[number]
    scale = yl.abs().max().item()
    assert (y1 - yl).abs().max().item() < tol * scale
    assert (y12 - (2.0 * y1 - 3.0 * y2)).abs().max().item() < 20 * tol * scale


def test_bulk_fused_dot_and_cg():
    """src.(A src) accumulated by the bulk kernel (the p.Ap of CG) and a CG solve through it."""
    om, oracle, mf, op = make(4, 3, "f64", dirichlet=True)
    src = np.random.default_rng(5).random(om.n_dofs)
    src[om.boundary_dofs] = 0.0
    inv_diag = op.compute_diagonal()
    b = torch.from_numpy(src).cuda()
    res = {}
    for bulk in (2, 1, 0):
        mf.select_brick_path(bulk)
        x = mf.initialize_dof_vector()
        solver = dealii_b200.SolverCG(dealii_b200.SolverControl(200, 1e-10 * float(np.linalg.norm(src))))
        r = solver.solve(op, x, b, inv_diag)
        res[bulk] = (r.iterations, x.clone())
    for path in (2, 1):
        assert abs(res[path][0] - res[0][0]) <= 1
        assert (res[path][1] - res[0][1]).abs().max().item() < 1e-8 * res[0][1].abs().max().item()
