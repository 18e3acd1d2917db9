"""Cell-wise matrix-free operator application -- oracle restatement in numpy.

The operator is  (c_grad(x) grad u, grad v) + (c_mass(x) u, v)  on FE_Q(p) with
QGauss(n_q)^dim: Laplace (c_mass absent), step-37 variable-coefficient Laplace, and the
step-64 / matrix_free_kokkos Helmholtz operator (c_grad = 1, c_mass = a(x)).

Follows, per cell (all citations relative to the deal.II tree):
  read_dof_values             matrix_free/portable_fe_evaluation.h:363-390
  evaluate (collocation route) matrix_free/portable_evaluation_kernels.h:504-560
  get_gradient / submit_gradient / get_value / submit_value
                              matrix_free/portable_fe_evaluation.h:544,653-745
  integrate                   matrix_free/portable_evaluation_kernels.h:562-650
  distribute_local_to_global  matrix_free/portable_fe_evaluation.h:398-435
  vmult = (dst = 0; cell_loop; copy_constrained_values)   examples/step-64/step-64.cc:313-325
  CPU MatrixFree semantics (constrained entries skipped on read/write, identity
  added on them): matrix_free/fe_evaluation.h:3059-3171, matrix_free/operators.h:1487-1642
  geometry: FEValues JxW / inverse_jacobian as stored by
  matrix_free/portable_matrix_free.templates.h:267-338

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np

from .mesh import jacobians_q1, map_q1
from .shape import ShapeInfo


def _apply_1d(M, u, axis):
    """out[..., q, ...] = sum_i M[i, q] u[..., i, ...] along ``axis``."""
    return np.moveaxis(np.tensordot(u, M, axes=([axis], [0])), -1, axis)


class MatrixFreeOracle:
    def __init__(self, mesh, n_q_points_1d=None, grad_coefficient=None,
                 mass_coefficient=None, constrained_dofs=None, dtype=np.float64):
        """``grad_coefficient`` / ``mass_coefficient``: None, a scalar, a callable
        f(points[(m, dim)]) -> (m,), or an array (n_cells, n_q^dim) in the cell /
        q-point order (q lexicographic, x fastest)."""
        self.mesh = mesh
        self.dim = mesh.dim
        self.shape = ShapeInfo(mesh.degree, n_q_points_1d)
        self.n, self.n_q = self.shape.n, self.shape.n_q
        self.dtype = dtype
        dim, nq = self.dim, self.n_q
        # q-points in lexicographic order, x fastest
        idx = np.array([[(q // nq ** d) % nq for d in range(dim)]
                        for q in range(nq ** dim)])
        self.q_ref = self.shape.q_points[idx]                       # (nqp, dim)
        self.q_w = np.prod(self.shape.q_weights[idx], axis=1)       # (nqp,)
        J = jacobians_q1(mesh.cell_vertices, self.q_ref)            # (c, q, d, e)
        self.inv_jacobian = np.linalg.inv(J)                        # [c,q,e,d] = dxi_e/dx_d
        self.JxW = np.linalg.det(J) * self.q_w[None, :]
        self.q_points = map_q1(mesh.cell_vertices, self.q_ref)      # (c, q, dim)
        self.grad_coef = self._coef(grad_coefficient)
        self.mass_coef = self._coef(mass_coefficient)
        self.constrained = (np.zeros(0, dtype=np.int64) if constrained_dofs is None
                            else np.asarray(constrained_dofs, dtype=np.int64))
        self._cmask = np.zeros(mesh.n_dofs, dtype=bool)
        self._cmask[self.constrained] = True

    def _coef(self, c):
        if c is None:
            return None
        nc, nqp = self.mesh.n_cells, self.n_q ** self.dim
        if callable(c):
            return np.asarray(c(self.q_points.reshape(-1, self.dim))).reshape(nc, nqp)
        c = np.asarray(c, dtype=np.float64)
        if c.ndim == 0:
            return np.full((nc, nqp), float(c))
        return c.reshape(nc, nqp)

    # ------------------------------------------------------------------ cell kernel
    def _cell_apply(self, u_cells):
        """u_cells: (c, npc) lexicographic -> (c, npc) local results."""
        dim, n, nq = self.dim, self.n, self.n_q
        dt = self.dtype
        S = self.shape.shape_values.astype(dt)
        D = self.shape.shape_gradients_collocation.astype(dt)
        nc = u_cells.shape[0]
        u = u_cells.astype(dt).reshape((nc,) + (n,) * dim)   # axes: [c, z, y, x]
        ax = [dim - d for d in range(dim)]                   # axis of direction d
        # values at q-points (basis change to the collocation space)
        uq = u
        for d in range(dim):
            uq = _apply_1d(S, uq, ax[d])
        # reference gradients: D[i, q] = l_i'(x_q)
        g = np.stack([_apply_1d(D, uq, ax[d]) for d in range(dim)], axis=-1)
        g = g.reshape(nc, nq ** dim, dim)
        invJ = self.inv_jacobian.astype(dt)
        JxW = self.JxW.astype(dt)
        # get_gradient: grad[d1] = sum_d2 inv_jac[d2][d1] g[d2]
        grad = np.einsum("cqed,cqe->cqd", invJ, g)
        if self.grad_coef is not None:
            grad = grad * self.grad_coef.astype(dt)[:, :, None]
        # submit_gradient: g'[d1] = (sum_d2 inv_jac[d1][d2] grad[d2]) * JxW
        gs = np.einsum("cqed,cqd->cqe", invJ, grad) * JxW[:, :, None]
        gs = gs.reshape((nc,) + (nq,) * dim + (dim,))
        vq = np.zeros_like(uq)
        if self.mass_coef is not None:
            vq = (uq.reshape(nc, -1) * self.mass_coef.astype(dt) * JxW).reshape(uq.shape)
        for d in range(dim):
            vq = vq + _apply_1d(D.T, gs[..., d], ax[d])
        out = vq
        for d in range(dim):
            out = _apply_1d(S.T, out, ax[d])
        return out.reshape(nc, n ** dim)

    def cell_loop(self, src, read_constrained_as_zero=False):
        src = np.asarray(src)
        l2g = self.mesh.l2g
        u = src[l2g]
        if read_constrained_as_zero:
            u = np.where(self._cmask[l2g], 0.0, u)
        r = self._cell_apply(u)
        if read_constrained_as_zero:
            r = np.where(self._cmask[l2g], 0.0, r)
        dst = np.zeros(self.mesh.n_dofs, dtype=self.dtype)
        np.add.at(dst, l2g.ravel(), r.ravel())
        return dst

    def vmult(self, src):
        """Portable::MatrixFree operator: dst = 0; cell_loop; copy_constrained_values."""
        dst = self.cell_loop(src)
        dst[self.constrained] = np.asarray(src, dtype=self.dtype)[self.constrained]
        return dst

    def vmult_cpu_matrixfree(self, src):
        """CPU MatrixFree + MatrixFreeOperators::Base::vmult semantics."""
        dst = self.cell_loop(src, read_constrained_as_zero=True)
        dst[self.constrained] += np.asarray(src, dtype=self.dtype)[self.constrained]
        return dst

    # ------------------------------------------------------------------ diagonal
    def compute_diagonal(self):
        """MatrixFreeTools::compute_diagonal (matrix_free/tools.h:1392-1569): apply the
        cell operator to every local unit vector, scatter-add the diagonal entry, then
        set constrained entries to 1."""
        npc = self.n ** self.dim
        nc = self.mesh.n_cells
        diag_local = np.zeros((nc, npc), dtype=self.dtype)
        for i in range(npc):
            e = np.zeros((nc, npc), dtype=self.dtype)
            e[:, i] = 1.0
            diag_local[:, i] = self._cell_apply(e)[:, i]
        diag = np.zeros(self.mesh.n_dofs, dtype=self.dtype)
        np.add.at(diag, self.mesh.l2g.ravel(), diag_local.ravel())
        diag[self.constrained] = 1.0
        return diag

    # ------------------------------------------------------------------ assembled matrix
    def assemble_sparse(self):
        """Classical assembly with full tensor-product basis tabulation -- an
        independent code path used the way tests/matrix_free_kokkos/
        matrix_vector_device_common.h:110-180 uses SparseMatrix.  Constrained rows and
        columns are eliminated (homogeneous) with a unit diagonal."""
        import scipy.sparse as sp
        dim, n, nq = self.dim, self.n, self.n_q
        S, G = self.shape.shape_values, self.shape.shape_gradients
        npc, nqp = n ** dim, nq ** dim
        li = np.array([[(i // n ** d) % n for d in range(dim)] for i in range(npc)])
        qi = np.array([[(q // nq ** d) % nq for d in range(dim)] for q in range(nqp)])
        phi = np.ones((nqp, npc))
        dphi = np.ones((nqp, npc, dim))
        for d in range(dim):
            v = S[li[:, d]][:, qi[:, d]].T          # (nqp, npc)
            gv = G[li[:, d]][:, qi[:, d]].T
            phi *= v
            for e in range(dim):
                dphi[:, :, e] *= gv if e == d else v
        rows, cols, vals = [], [], []
        l2g = self.mesh.l2g
        for c in range(self.mesh.n_cells):
            dp = np.einsum("qed,qie->qid", self.inv_jacobian[c], dphi)
            w = self.JxW[c] * (self.grad_coef[c] if self.grad_coef is not None else 1.0)
            A = np.einsum("qid,qjd,q->ij", dp, dp, w)
            if self.mass_coef is not None:
                A += np.einsum("qi,qj,q->ij", phi, phi, self.JxW[c] * self.mass_coef[c])
            idx = l2g[c]
            rows.append(np.repeat(idx, npc))
            cols.append(np.tile(idx, npc))
            vals.append(A.ravel())
        rows, cols, vals = map(np.concatenate, (rows, cols, vals))
        keep = ~(self._cmask[rows] | self._cmask[cols])
        N = self.mesh.n_dofs
        A = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(N, N)).tocsr()
        A = A + sp.diags(self._cmask.astype(np.float64))
        return A
