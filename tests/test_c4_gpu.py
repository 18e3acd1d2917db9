"""BASELINE configs[3] on one GPU: Q3 (and Q1, Q2, Q4) Poisson on the adaptively refined mesh of
csrc/mesh_adaptive.cpp (hanging nodes on a ball inside the cube), brick-friendly cell order.
Checked per entry against what deal.II itself computed on the same mesh (tests/golden/ref/c4_*.npz),
against the numpy oracle at a second size, and through size-independent properties."""
import os

import numpy as np
import pytest
import torch

import dealii_b200
from dealii_b200.distributed import AdaptiveHyperCubeMesh
from test_reference_parity import GoldenMesh, assert_per_entry, load, oracle_vmult_portable
from oracle.mf_oracle import MatrixFreeOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("name,degree", [("c4_q3_ball", 3), ("c4_q1_ball", 1)])
def test_generated_mesh_through_the_engine_matches_deal_ii(name, degree, number):
    g = load(name)
    mesh = AdaptiveHyperCubeMesh(3, degree, 3, ball_radius=0.3, dirichlet_boundary=True)
    mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
    op = dealii_b200.LaplaceOperator(mf)
    x = torch.from_numpy(g["src"].astype(mf.np_dtype)).cuda()
    y = mf.initialize_dof_vector()
    y.fill_(float("nan"))
    op.vmult(y, x)
    torch.cuda.synchronize()
    tol = 1e-12 if number == "f64" else 1e-5
    assert_per_entry(y.cpu().numpy().astype(np.float64), g["dst_portable_matrixfree"], tol, "vs Portable::MatrixFree")
    assert_per_entry(y.cpu().numpy().astype(np.float64), g["dst_cpu_matrixfree"], tol, "vs CPU MatrixFree")


def test_cg_jacobi_on_the_generated_mesh_matches_deal_ii():
    g = load("c4_q3_ball")
    mesh = AdaptiveHyperCubeMesh(3, 3, 3, ball_radius=0.3, dirichlet_boundary=True)
    mf = dealii_b200.MatrixFree("f64").reinit_from_mesh(mesh)
    op = dealii_b200.LaplaceOperator(mf)
    b = np.ones(mesh.n_dofs)
    b[g["constrained_dofs"]] = 0.0
    inv_diag = op.compute_diagonal()
    x = mf.initialize_dof_vector()
    r = dealii_b200.SolverCG(dealii_b200.SolverControl(10000, float(g["cg_tolerance"]))).solve(
        op, x, torch.from_numpy(b).cuda(), inv_diag)
    assert abs(r.iterations - int(g["cg_jacobi_iterations"])) <= 1
    assert np.abs(x.cpu().numpy() - g["cg_jacobi_solution"]).max() < 1e-9 * np.abs(g["cg_jacobi_solution"]).max()


@pytest.mark.parametrize("degree,refinements", [(2, 3), (3, 4), (4, 3)])
def test_generated_mesh_matches_the_oracle(degree, refinements):
    mesh = AdaptiveHyperCubeMesh(3, degree, refinements, ball_radius=0.33, dirichlet_boundary=True)
    g = {"dim": 3, "degree": degree, "l2g": mesh.l2g.copy(), "vertices": mesh.cell_vertices.copy(),
         "n_cells": mesh.n_cells, "n_dofs": mesh.n_dofs, "constraint_mask": mesh.constraint_mask.copy(),
         "constrained_dofs": mesh.constrained_dofs.copy()}
    oracle = MatrixFreeOracle(GoldenMesh(g), constrained_dofs=g["constrained_dofs"])
    src = np.random.default_rng(1).random(mesh.n_dofs)
    src[g["constrained_dofs"]] = 0.0
    ref = oracle_vmult_portable(g, oracle, src.copy())
    mf = dealii_b200.MatrixFree("f64").reinit_from_mesh(mesh)
    if (degree, refinements) == (3, 4):
        assert mf.info.n_bricks > 0, "unmasked 4^3 blocks must stay on the brick path"
    op = dealii_b200.LaplaceOperator(mf)
    y = mf.initialize_dof_vector()
    for path in (0, 1):              # atomics on zeroed dst / coloured launches
        if mf.select_brick_path(path) != path:
            continue
        y.fill_(float("nan"))
        op.vmult(y, torch.from_numpy(src).cuda())
        torch.cuda.synchronize()
        assert_per_entry(y.cpu().numpy(), ref, 1e-12, f"engine vs oracle (brick path {path})")


@pytest.mark.parametrize("number", ["f64", "f32"])
def test_properties_at_a_larger_size(number):
    """4.3 M dofs, no Dirichlet boundary: the Laplacian annihilates constants on a mesh with hanging
    nodes only if every interpolation and its transpose are right; the operator is symmetric."""
    mesh = AdaptiveHyperCubeMesh(3, 3, 5, ball_radius=0.35)
    mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
    assert mf.info.n_bricks * mf.info.cells_per_brick > 0.5 * mesh.n_cells
    op = dealii_b200.LaplaceOperator(mf)
    cons = torch.from_numpy(mesh.constrained_dofs.astype(np.int64)).cuda()    # hanging-node dofs only
    one = torch.ones(mesh.n_dofs, dtype=mf.torch_dtype, device="cuda")
    one[cons] = 0.0      # Portable::MatrixFree semantics: constrained entries are not read through l2g
    y = mf.initialize_dof_vector()
    op.vmult(y, one)
    y[cons] = 0.0
    diag = mf.initialize_dof_vector()
    mf.compute_diagonal(op.op, diag)
    tol = 1e-11 if number == "f64" else 2e-4
    assert float((y.abs() / diag.abs()).max()) < tol
    g = torch.Generator(device="cuda").manual_seed(3)
    u = torch.rand(mesh.n_dofs, dtype=mf.torch_dtype, device="cuda", generator=g)
    v = torch.rand(mesh.n_dofs, dtype=mf.torch_dtype, device="cuda", generator=g)
    u[cons] = 0.0
    v[cons] = 0.0
    au, av = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    op.vmult(au, u)
    op.vmult(av, v)
    a, b = float(torch.dot(v.double(), au.double())), float(torch.dot(u.double(), av.double()))
    assert abs(a - b) < (1e-11 if number == "f64" else 1e-4) * abs(a)
