"""Parity of the brick kernel (csrc/brick_kernel.cuh: Cartesian cells, constant coefficients,
whole b^3-cell blocks applied as one macro element) against the oracle and against the per-cell
kernels on the same inputs.  Tolerances as in test_vmult_gpu.py (1e-12 / 1e-5 of ||ref||_inf)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import dealii_b200
from dealii_b200 import _lib as L
from oracle.mesh import HyperCubeMesh as OracleMesh
from oracle.mf_oracle import MatrixFreeOracle

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-12, "f32": 1e-5}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a, ref):
    return np.abs(a - ref).max() / np.abs(ref).max()


def brick_refinements(degree):
    # smallest refine_global() that holds bricks: 16^3 cells for degree 1, 8^3 for 2, 4^3 up to 4, then 2^3
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
@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_brick_laplace_all_degrees(degree, number):
    r = brick_refinements(degree)
    om, oracle, mf, op = make(degree, r, number)
    assert mf.info.n_bricks * mf.info.cells_per_brick == om.n_cells, "every cell must sit in a brick"
    src = np.random.default_rng(degree).random(om.n_dofs)
    x = torch.from_numpy(src.astype(mf.np_dtype)).cuda()
    y = mf.initialize_dof_vector()
    y.fill_(7.0)
    op.vmult(y, x)
    torch.cuda.synchronize()
    assert rel_err(y.cpu().numpy().astype(np.float64), oracle.vmult(src)) < TOL[number]


@pytest.mark.parametrize("degree,extra", [(4, 1), (3, 1), (2, 0), (6, 1)])
def test_brick_several_bricks_helmholtz_dirichlet(degree, extra):
    """More than one brick (shared brick faces go through atomics), constant Helmholtz
    coefficients, zero Dirichlet boundary in both constraint semantics."""
    r = brick_refinements(degree) + extra
    for cpu_mf in (False, True):
        om, oracle, mf, op = make(degree, r, "f64", dirichlet=True, cpu_mf=cpu_mf, mass=10.0, grad=2.5)
        assert mf.info.n_bricks * mf.info.cells_per_brick == om.n_cells
        src = np.random.default_rng(3).random(om.n_dofs)
        if not cpu_mf:
            src[om.boundary_dofs] = 0.0
        x = torch.from_numpy(src).cuda()
        y = mf.initialize_dof_vector()
        op.vmult(y, x)
        torch.cuda.synchronize()
        ref = oracle.vmult_cpu_matrixfree(src) if cpu_mf else oracle.vmult(src)
        assert rel_err(y.cpu().numpy(), ref) < 1e-12


def test_brick_cell_ranges_and_adds():
    """cell_loop adds into dst; ranges that cut through bricks fall back to the per-cell kernels
    for the partial bricks (the interior/boundary split of the distributed cell loop)."""
    om, oracle, mf, op = make(4, 3, "f64")       # 512 cells = 8 bricks of 64
    assert mf.info.n_bricks == 8
    src = np.random.default_rng(5).random(om.n_dofs)
    ref = oracle.vmult(src)
    x = torch.from_numpy(src).cuda()
    lib = L.load()
    for cuts in ([0, 512], [0, 64, 512], [0, 100, 130, 300, 512], [0, 1, 511, 512], [0, 192, 193, 512]):
        y = torch.ones(om.n_dofs, dtype=torch.float64, device="cuda")
        for b, e in zip(cuts[:-1], cuts[1:]):
            L.check(lib.b200mf_cell_loop_range(mf._h, C.byref(op.op), y.data_ptr(), x.data_ptr(), b, e,
                                               torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert rel_err(y.cpu().numpy() - 1.0, ref) < 1e-12, cuts


def test_brick_fused_dot():
    """src . (A src) accumulated by the brick kernel (the p.Ap of CG)."""
    om, oracle, mf, op = make(4, 3, "f64", dirichlet=True, cpu_mf=True, mass=1.0)
    src = np.random.default_rng(9).random(om.n_dofs)
    ref = oracle.vmult_cpu_matrixfree(src)
    x = torch.from_numpy(src).cuda()
    y = torch.zeros(om.n_dofs, dtype=torch.float64, device="cuda")
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200mf_cell_loop_range_dot(mf._h, C.byref(op.op), y.data_ptr(), x.data_ptr(), 0,
                                           om.n_cells, acc.data_ptr(), st))
    L.check(lib.b200mf_copy_constrained_values_dot(mf._h, y.data_ptr(), x.data_ptr(),
                                                   acc.data_ptr(), st))
    torch.cuda.synchronize()
    assert rel_err(y.cpu().numpy(), ref) < 1e-12
    assert abs(acc.item() - src @ ref) < 1e-12 * abs(src @ ref)


def test_brick_matches_per_cell_kernels_large():
    """At a size the oracle does not touch: brick kernel vs the plane kernel (B200MF_KERNEL=plane
    in a child process) on the same seeded vector, and linearity / symmetry / constants."""
    mesh = dealii_b200.HyperCubeMesh(3, 4, refinements=5)
    mf = dealii_b200.MatrixFree("f64").reinit_from_mesh(mesh)
    assert mf.info.n_bricks == 8 ** 5 // 64
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
    one = torch.ones_like(u)
    op.vmult(Av, one)
    assert Av.abs().max().item() < 1e-11 * scale
    code = (
        "import torch, dealii_b200\n"
        "mesh = dealii_b200.HyperCubeMesh(3, 4, refinements=5)\n"
        "mf = dealii_b200.MatrixFree('f64').reinit_from_mesh(mesh)\n"
        "op = dealii_b200.LaplaceOperator(mf)\n"
        "g = torch.Generator(device='cuda').manual_seed(1)\n"
        "u = torch.rand(mf.n_owned, dtype=torch.float64, device='cuda', generator=g)\n"
        "y = mf.initialize_dof_vector(); op.vmult(y, u); torch.cuda.synchronize()\n"
        "torch.save(y.cpu(), '/tmp/b200mf_plane_ref.pt')\n")
    env = dict(os.environ, B200MF_KERNEL="plane", PYTHONPATH=ROOT)
    subprocess.check_call([sys.executable, "-c", code], env=env, cwd=ROOT)
    ref = torch.load("/tmp/b200mf_plane_ref.pt")
    assert (Au.cpu() - ref).abs().max().item() < 1e-12 * scale


def test_no_bricks_on_lexicographic_cell_order():
    """Cells that are not Morton ordered do not form bricks; the per-cell kernels serve them."""
    mesh = dealii_b200.HyperCubeMesh(3, 4, subdivisions=4)
    mf = dealii_b200.MatrixFree("f64").reinit_from_mesh(mesh)
    assert mf.info.n_bricks == 0


@pytest.mark.parametrize("selective", [False, True])
def test_vmult_pieces_and_selective_zeroing(selective, monkeypatch):
    """b200mf_vmult_prepare + brick-aligned b200mf_vmult_range pieces == vmult, starting from a dst
    full of garbage.  With B200MF_SELECTIVE_ZERO only the dofs no brick stores are zeroed, and a
    piece that cuts a brick is refused."""
    if selective:
        monkeypatch.setenv("B200MF_SELECTIVE_ZERO", "1")
    else:
        monkeypatch.delenv("B200MF_SELECTIVE_ZERO", raising=False)
    om, oracle, mf, op = make(4, 3, "f64", dirichlet=True, cpu_mf=True, mass=2.0)
    src = np.random.default_rng(17).random(om.n_dofs)
    ref = oracle.vmult_cpu_matrixfree(src)
    x = torch.from_numpy(src).cuda()
    y = torch.full((om.n_dofs,), 1.0e300, dtype=torch.float64, device="cuda")
    mf.vmult_prepare(op.op, y)
    for a, b in ((0, 128), (320, 512), (128, 320)):
        mf.vmult_range(op.op, y, x, a, b)
    mf.copy_constrained_values(x, y)
    torch.cuda.synchronize()
    assert rel_err(y.cpu().numpy(), ref) < 1e-12
    if selective:
        with pytest.raises(L.B200MFError):
            mf.vmult_range(op.op, y, x, 0, 100)
