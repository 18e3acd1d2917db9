"""The numpy restatement of step-37's multigrid (oracle/multigrid.py) against what deal.II's own Multigrid
produced (tests/golden/ref_gmg/*.npz, oracle/ref_drivers/ref_gmg.cc): level numbering, level diagonals and
eigenvalue estimates, one prolongation / restriction, one V-cycle, CG iteration count and solution.  CPU only."""
import os

import numpy as np
import pytest

from oracle.multigrid import MultigridOracle, prolongation_matrix_1d
from oracle.mf_oracle import MatrixFreeOracle
from oracle.solvers import solver_cg

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_gmg")
CASES = ["gmg_q2_r2_f64", "gmg_q2_r3_f64_step37", "gmg_q4_r2_f64", "gmg_q1_r3_f64", "gmg_d2_q3_r4_f64",
         "gmg_q2_r3_f64_deformed"]


def step37_coefficient(x):
    return 1.0 / (0.05 + 2.0 * (x * x).sum(axis=1))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def per_entry(a, ref, tol, what):
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * 1e-2)
    err = np.abs(a - ref) / scale
    assert err.max() < tol, f"{what}: {err.max():.3e} at {err.argmax()}"


@pytest.mark.parametrize("name", CASES)
def test_multigrid_oracle_matches_deal_ii(name):
    g = load(name)
    dim, degree, r = int(g["dim"]), int(g["degree"]), int(g["refinements"])
    amp = float(g["deformation"]) if "deformation" in g else 0.0
    defo = (lambda v: v + amp * np.prod(np.sin(np.pi * v), axis=1, keepdims=True)) if amp else None
    coef = step37_coefficient if int(g["variable_coefficient"]) else None
    mg = MultigridOracle(dim, degree, r, coefficient=coef, deformation=defo)
    for level in range(r + 1):
        assert np.array_equal(mg.meshes[level].l2g.ravel(), g[f"level_l2g_{level}"]), f"level {level} numbering"
        per_entry(mg.smoothers[level].P.diagonal, g[f"level_inverse_diagonal_{level}"], 1e-11, f"diagonal {level}")
        info = mg.smoothers[level].info
        early = level == 0 or int(g[f"eig_cg_iterations_{level}"]) < 10
        tol = 1e-3 if early else 1e-7
        if not (float(g[f"eig_max_{level}"]) == 1.0 and float(g[f"eig_min_{level}"]) == 1.0):
            assert info["max_eigenvalue"] == pytest.approx(float(g[f"eig_max_{level}"]), rel=tol)
            assert info["min_eigenvalue"] == pytest.approx(float(g[f"eig_min_{level}"]), rel=max(tol, 1e-4))
        assert mg.smoothers[level].degree == int(g[f"cheb_degree_{level}"])
    per_entry(mg.prolongate(r, g["prolongate_src"]), g["prolongate_dst"], 1e-12, "prolongate")
    per_entry(mg.restrict_and_add(r, np.zeros(g["restrict_dst"].size), g["restrict_src"]), g["restrict_dst"], 1e-12,
              "restrict_and_add")
    per_entry(mg.vmult(g["rhs"]), g["vcycle_of_rhs"], 1e-8, "one V-cycle")
    A = mg.ops[r].vmult_cpu_matrixfree
    out = solver_cg(A, g["rhs"], mg, tol=1e-12 * float(np.linalg.norm(g["rhs"])), max_steps=100)
    assert out["iterations"] == int(g["cg_iterations"])
    per_entry(out["x"], g["solution"], 1e-8, "solution")


@pytest.mark.parametrize("degree", [1, 3, 6])
def test_embedding_reproduces_the_coarse_space(degree):
    P = prolongation_matrix_1d(degree)
    assert np.abs(P.sum(axis=1) - 1).max() < 1e-13 and P.shape == (2 * degree + 1, degree + 1)
