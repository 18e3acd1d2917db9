"""SolverCG, PreconditionChebyshev, DiagonalMatrix (Jacobi) -- oracle restatement.

Follows (paths relative to the deal.II tree):
  SolverCG::solve                      include/deal.II/lac/solver_cg.h:1391-1470
  IterationWorker::startup/do_iteration  solver_cg.h:625-655, 703-763 (generic branch)
  Lanczos tridiagonal from CG          solver_cg.h:1440-1452 (compute_eigs_and_cond)
  SolverControl::check                 source/lac/solver_control.cc (success: res <= tol;
                                       failure: step >= max_steps)
  IterationNumberControl::check        source/lac/solver_control.cc (success at max_steps)
  PreconditionChebyshev                include/deal.II/lac/precondition.h:
       set_initial_guess :2378-2408, estimate_eigenvalues :2465-2572 and :3928-4020,
       apply_internal :4029-4121, vector_updates :3154-3190
  DiagonalMatrix::vmult                include/deal.II/lac/diagonal_matrix.h:435

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np


class DiagonalMatrix:
    """dst = diag * src (for Jacobi, ``diag`` holds the inverse diagonal)."""

    def __init__(self, diagonal):
        self.diagonal = np.asarray(diagonal)

    def vmult(self, src):
        return self.diagonal * src


class PreconditionIdentity:
    def vmult(self, src):
        return src.copy()


class NoConvergence(RuntimeError):
    pass


def solver_cg(A_vmult, b, preconditioner=None, x0=None, tol=1e-12, max_steps=1000,
              iteration_number_control=False, track_eigenvalues=False):
    """Returns dict(x, iterations, residual, history[, eigenvalues]).

    ``tol`` is the absolute tolerance on ||r||_2 (SolverControl semantics);
    with ``iteration_number_control`` reaching ``max_steps`` counts as success."""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    identity = preconditioner is None or isinstance(preconditioner, PreconditionIdentity)
    # startup(): residual, short-circuit for zero start vector
    if np.any(x != 0):
        r = b - A_vmult(x)
    else:
        r = b.copy()
    res = float(np.sqrt(r @ r))
    history = [res]
    diag, offdiag = [], []
    eig_beta_alpha = 0.0

    def state(step, value):
        if value <= tol or (iteration_number_control and step >= max_steps):
            return "success"
        if step >= max_steps or np.isnan(value):
            return "failure"
        return "iterate"

    st = state(0, res)
    it = 0
    p = None
    rpr = 0.0
    alpha = beta = prev_alpha = 0.0
    while st == "iterate":
        it += 1
        prev_rpr = rpr
        if not identity:
            v = preconditioner.vmult(r)
            rpr = float(r @ v)
            direction = v
        else:
            rpr = res * res
            direction = r
        if it > 1:
            beta = rpr / prev_rpr
            p = beta * p + direction
        else:
            p = direction.copy()
        v = A_vmult(p)
        pAp = float(p @ v)
        prev_alpha = alpha
        alpha = rpr / pAp
        x = x + alpha * p
        r = r - alpha * v
        res = float(np.sqrt(abs(r @ r)))
        history.append(res)
        if it > 1 and track_eigenvalues:
            diag.append(1.0 / prev_alpha + eig_beta_alpha)
            eig_beta_alpha = beta / prev_alpha
            offdiag.append(np.sqrt(beta) / prev_alpha)
        st = state(it, res)
    out = dict(x=x, iterations=it, residual=res, history=history, state=st)
    if track_eigenvalues:
        out["eigenvalues"] = lanczos_eigenvalues(diag, offdiag)
    if st != "success":
        raise NoConvergence(f"CG did not converge: it={it} res={res}")
    return out


def lanczos_eigenvalues(diagonal, offdiagonal):
    """Eigenvalues of the CG-generated tridiagonal matrix
    (solver_cg.h:502-520 compute_eigs_and_cond: T(i,i) = diagonal[i],
    T(i,i+1) = offdiagonal[i]; eigenvalues by LAPACK stev, here numpy eigvalsh)."""
    n = len(diagonal)
    if n == 0:
        return np.zeros(0)
    T = np.diag(np.asarray(diagonal, dtype=float))
    for i in range(n - 1):
        T[i, i + 1] = T[i + 1, i] = offdiagonal[i]
    return np.sort(np.abs(np.linalg.eigvalsh(T)))


class PreconditionChebyshev:
    def __init__(self, A_vmult, inverse_diagonal, degree=1, smoothing_range=0.0,
                 eig_cg_n_iterations=8, eig_cg_residual=1e-2, max_eigenvalue=1.0,
                 constrained_dofs=None, safety_factor=1.2, first_owned_index=0):
        self.A = A_vmult
        self.P = DiagonalMatrix(inverse_diagonal)
        self.degree = degree
        self.smoothing_range = smoothing_range
        self.eig_cg_n_iterations = eig_cg_n_iterations
        self.max_eigenvalue = max_eigenvalue
        self.constrained = constrained_dofs
        self.safety_factor = safety_factor
        self.first_owned_index = first_owned_index
        self.initialized = False
        self.info = {}

    def estimate_eigenvalues(self, n):
        if self.eig_cg_n_iterations > 0:
            t = ((np.arange(n) + self.first_owned_index) % 11).astype(np.float64)
            t -= t.sum() / n                       # vector.add(-mean_value)
            if self.constrained is not None:
                t[self.constrained] = 0.0          # constraints.set_zero
            out = solver_cg(self.A, t, self.P, tol=1e-10,
                            max_steps=self.eig_cg_n_iterations,
                            iteration_number_control=True, track_eigenvalues=True)
            ev = out["eigenvalues"]
            if len(ev) == 0:
                lmin = lmax = 1.0
            else:
                lmin, lmax = ev[0], self.safety_factor * ev[-1]
            self.info["cg_iterations"] = out["iterations"]
        else:
            lmax = self.max_eigenvalue
            lmin = lmax / self.smoothing_range
        alpha = (lmax / self.smoothing_range if self.smoothing_range > 1.0
                 else min(0.9 * lmax, lmin))
        self.delta = (lmax - alpha) * 0.5
        self.theta = (lmax + alpha) * 0.5
        self.info.update(min_eigenvalue=lmin, max_eigenvalue=lmax)
        self.initialized = True

    def vmult(self, rhs):
        if not self.initialized:
            self.estimate_eigenvalues(len(rhs))
        # iteration_index 0: solution = P * (factor2 * rhs)
        sol = self.P.vmult((1.0 / self.theta) * rhs)
        sol_old = np.zeros_like(sol)
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return sol
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for k in range(self.degree - 1):
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            t = self.P.vmult(rhs - self.A(sol))
            if k == 0:        # iteration_index 1
                new = (1.0 + f1) * sol + f2 * t
            else:
                new = (1.0 + f1) * sol - f1 * sol_old + f2 * t
            sol_old, sol = sol, new
        return sol
