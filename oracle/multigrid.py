"""Geometric multigrid of step-37 -- oracle restatement in numpy (TEST INFRASTRUCTURE ONLY, see
oracle/__init__.py): levels = hyper_cube refined 0..r times, the level operators of oracle/mf_oracle.py with
the CPU MatrixFree treatment of constrained dofs, and

  Multigrid::level_v_step                 multigrid/multigrid.templates.h:112-171
  MGTransferMatrixFree::prolongate /
    restrict_and_add                      multigrid/mg_transfer_matrix_free.templates.h (cell-wise embedding of the
                                          parent's FE_Q in the children's, fine residual weighted by 1/valence)
  mg::SmootherRelaxation<PreconditionChebyshev>: apply = vmult, smooth = step   multigrid/mg_smoother.h
  MGCoarseGridApplySmoother with the Chebyshev "solver" (degree from the error estimate,
                                          lac/precondition.h:3952-3982) on level 0
  parameters of examples/step-37/step-37.cc:965-984

It is pinned on tests/golden/ref_gmg/*.npz (produced by deal.II's own Multigrid) in tests/test_oracle_multigrid.py.
"""
import numpy as np

from .mesh import HyperCubeMesh
from .mf_oracle import MatrixFreeOracle
from .shape import ShapeInfo
from .solvers import PreconditionChebyshev


def prolongation_matrix_1d(degree):
    """P[X, i] = l_i(x_X): the 2p+1 nodes of both children in the parent's unit cell."""
    gl = np.asarray(ShapeInfo(degree).support_points, dtype=np.float64)
    x = np.concatenate([0.5 * gl, 0.5 + 0.5 * gl[1:]])
    P = np.ones((len(x), len(gl)))
    for i in range(len(gl)):
        for j in range(len(gl)):
            if j != i:
                P[:, i] *= (x - gl[j]) / (gl[i] - gl[j])
    return P


class ChebyshevSmoother(PreconditionChebyshev):
    """PreconditionChebyshev with step() (nonzero initial guess, apply_internal with zero_out_dst = false,
    lac/precondition.h:4029-4121) and the solver mode of level 0."""

    def step(self, x, rhs):
        if not self.initialized:
            self.estimate_eigenvalues(len(rhs))
        sol_old = x
        sol = x + (1.0 / self.theta) * self.P.vmult(rhs - self.A(x))
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return sol
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for _ in range(self.degree - 1):
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            new = (1.0 + f1) * sol - f1 * sol_old + f2 * self.P.vmult(rhs - self.A(sol))
            sol_old, sol = sol, new
        return sol

    def make_solver(self, n, tolerance):
        """degree == numbers::invalid_unsigned_int: smoothing_range < 1 is the relative tolerance."""
        self.smoothing_range = tolerance
        self.estimate_eigenvalues(n)
        lmin, lmax = self.info["min_eigenvalue"], self.info["max_eigenvalue"]
        alpha = min(0.9 * lmax, lmin)
        actual_range = lmax / alpha
        sigma = (1.0 - np.sqrt(1.0 / actual_range)) / (1.0 + np.sqrt(1.0 / actual_range))
        eps = tolerance
        self.degree = 1 + int(np.log(1.0 / eps + np.sqrt(1.0 / eps / eps - 1.0)) / np.log(1.0 / sigma))


class MultigridOracle:
    def __init__(self, dim, degree, refinements, coefficient=None, deformation=None, smoother_degree=5,
                 smoothing_range=15.0, eig_cg_n_iterations=10, coarse_tolerance=1e-3):
        self.dim, self.degree, self.n = dim, degree, degree + 1
        self.P1 = prolongation_matrix_1d(degree)
        self.meshes, self.ops, self.smoothers, self.inv_valence = [], [], [], []
        for level in range(refinements + 1):
            mesh = HyperCubeMesh(dim, degree, refinements=level, deformation=deformation)
            op = MatrixFreeOracle(mesh, grad_coefficient=coefficient, constrained_dofs=mesh.boundary_dofs)
            inv_diag = 1.0 / op.compute_diagonal()
            # AdditionalData::constraints stays empty in step-37: the start vector keeps its constrained entries
            sm = ChebyshevSmoother(op.vmult_cpu_matrixfree, inv_diag, degree=smoother_degree,
                                   smoothing_range=smoothing_range, constrained_dofs=None,
                                   eig_cg_n_iterations=eig_cg_n_iterations if level > 0 else mesh.n_dofs)
            if level == 0:
                sm.make_solver(mesh.n_dofs, coarse_tolerance)
            else:
                sm.estimate_eigenvalues(mesh.n_dofs)
            valence = np.zeros(mesh.n_dofs)
            np.add.at(valence, mesh.l2g.ravel(), 1.0)
            self.meshes.append(mesh)
            self.ops.append(op)
            self.smoothers.append(sm)
            self.inv_valence.append(1.0 / valence)

    # -------------------------------------------------------------- transfer
    def _child_l2g(self, level):
        """(n_coarse_cells, (2n-1)^dim) fine dof of every node of the children's lattice, x fastest."""
        n, dim, M = self.n, self.dim, 2 * self.n - 1
        fine = self.meshes[level].l2g
        nodes = np.array([[(o // M ** d) % M for d in range(dim)] for o in range(M ** dim)])
        child = (nodes > n - 1).astype(np.int64)
        local = nodes - (n - 1) * child
        k = (child * (1 << np.arange(dim))).sum(axis=1)
        loc = (local * (n ** np.arange(dim))).sum(axis=1)
        shared = (nodes == n - 1).sum(axis=1)
        nc = self.meshes[level - 1].n_cells
        cells = (np.arange(nc)[:, None] << dim) + k[None, :]
        return fine[cells, loc[None, :]], shared

    def _sweeps(self, u, transpose):
        """apply P1 (or its transpose) along every direction of (cells, n or M per direction) arrays."""
        dim = self.dim
        for d in range(dim):
            axis = dim - d          # x is the last axis
            M = self.P1.T if transpose else self.P1
            u = np.moveaxis(np.tensordot(u, M, axes=([axis], [1])), -1, axis)
        return u

    def prolongate(self, to_level, src):
        coarse, fine = self.meshes[to_level - 1], self.meshes[to_level]
        n, dim, M = self.n, self.dim, 2 * self.n - 1
        u = np.asarray(src)[coarse.l2g].reshape((coarse.n_cells,) + (n,) * dim)
        u[np.isin(coarse.l2g, coarse.boundary_dofs).reshape(u.shape)] = 0.0
        v = self._sweeps(u, False).reshape(coarse.n_cells, M ** dim)
        idx, _ = self._child_l2g(to_level)
        dst = np.zeros(fine.n_dofs)
        dst[idx.ravel()] = v.ravel()
        dst[fine.boundary_dofs] = 0.0
        return dst

    def restrict_and_add(self, from_level, dst, src):
        coarse, fine = self.meshes[from_level - 1], self.meshes[from_level]
        n, dim, M = self.n, self.dim, 2 * self.n - 1
        idx, shared = self._child_l2g(from_level)
        w = np.asarray(src) * self.inv_valence[from_level]
        w[fine.boundary_dofs] = 0.0
        g = (w[idx] * (2.0 ** shared)[None, :]).reshape((coarse.n_cells,) + (M,) * dim)
        r = self._sweeps(g, True).reshape(coarse.n_cells, n ** dim)
        r[np.isin(coarse.l2g, coarse.boundary_dofs)] = 0.0
        out = np.array(dst, dtype=np.float64)
        np.add.at(out, coarse.l2g.ravel(), r.ravel())
        return out

    # -------------------------------------------------------------- cycle
    def level_v_step(self, level, defect):
        sm, A = self.smoothers[level], self.ops[level].vmult_cpu_matrixfree
        if level == 0:
            return sm.vmult(defect)
        sol = sm.vmult(defect)
        t = defect - A(sol)
        coarse_defect = self.restrict_and_add(level, np.zeros(self.meshes[level - 1].n_dofs), t)
        sol = sol + self.prolongate(level, self.level_v_step(level - 1, coarse_defect))
        return sm.step(sol, defect)

    def vmult(self, src):
        """PreconditionMG::vmult: one V-cycle from the finest level."""
        return self.level_v_step(len(self.meshes) - 1, np.asarray(src, dtype=np.float64))
