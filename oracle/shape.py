"""1D shape data of FE_Q(p) with QGauss(n_q) -- oracle restatement.

Follows ``include/deal.II/matrix_free/shape_info.templates.h:861-985``
(UnivariateShapeData::evaluate_shape_functions / evaluate_collocation_space):
row = dof ``i`` (lexicographic), column = quadrature point ``q``, stored
row-major ``[i * n_q + q]``.  FE_Q support points are the Gauss-Lobatto points
on [0,1] (``source/fe/fe_q.cc``: ``QGaussLobatto<1>(degree+1)``), quadrature is
Gauss-Legendre on [0,1] (``source/base/quadrature_lib.cc`` QGauss).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
from numpy.polynomial import legendre as _leg


def gauss_points_weights(n):
    """QGauss<1>(n) on [0,1] (source/base/quadrature_lib.cc:38-110)."""
    x, w = _leg.leggauss(n)
    # symmetrise to kill the last-bit asymmetry of the eigenvalue solver
    x = 0.5 * (x - x[::-1])
    w = 0.5 * (w + w[::-1])
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto_points(n):
    """QGaussLobatto<1>(n) nodes on [0,1] (source/base/quadrature_lib.cc:120-200):
    end points plus the roots of P'_{n-1}."""
    if n == 1:
        return np.array([0.5])
    if n == 2:
        return np.array([0.0, 1.0])
    c = np.zeros(n)
    c[n - 1] = 1.0
    dc = _leg.legder(c)
    r = np.sort(_leg.legroots(dc).real)
    # two Newton steps on P'_{n-1} for full double accuracy
    d2c = _leg.legder(dc)
    for _ in range(2):
        r = r - _leg.legval(r, dc) / _leg.legval(r, d2c)
    r = 0.5 * (r - r[::-1])
    x = np.concatenate(([-1.0], r, [1.0]))
    return 0.5 * (x + 1.0)


def lagrange_values_and_derivatives(nodes, x):
    """values[i, q] = l_i(x_q), derivs[i, q] = l_i'(x_q) for the Lagrange basis on
    ``nodes`` (Polynomials::generate_complete_Lagrange_basis,
    include/deal.II/base/polynomial.h)."""
    nodes = np.asarray(nodes, dtype=np.longdouble)
    x = np.asarray(x, dtype=np.longdouble)
    n = len(nodes)
    vals = np.zeros((n, len(x)), dtype=np.longdouble)
    ders = np.zeros((n, len(x)), dtype=np.longdouble)
    for i in range(n):
        denom = np.longdouble(1.0)
        for j in range(n):
            if j != i:
                denom *= nodes[i] - nodes[j]
        for q, xq in enumerate(x):
            v = np.longdouble(1.0)
            for j in range(n):
                if j != i:
                    v *= xq - nodes[j]
            d = np.longdouble(0.0)
            for k in range(n):
                if k == i:
                    continue
                t = np.longdouble(1.0)
                for j in range(n):
                    if j != i and j != k:
                        t *= xq - nodes[j]
                d += t
            vals[i, q] = v / denom
            ders[i, q] = d / denom
    return vals.astype(np.float64), ders.astype(np.float64)


class ShapeInfo:
    """The subset of internal::MatrixFreeFunctions::ShapeInfo the hot path uses.

    Attributes (all float64, row-major [i, q]):
      shape_values, shape_gradients              shape_info.templates.h:895-898
      shape_gradients_collocation                shape_info.templates.h:973-984
      subface_interpolation_matrix               shape_info.templates.h:~1100
          (= subface_interpolation_matrices[0], the weights Portable::MatrixFree
          ships as ``constraint_weights``, portable_matrix_free.templates.h:1306-1322)
      q_points, q_weights                        QGauss<1>(n_q)
      support_points                             FE_Q 1D node positions
    """

    def __init__(self, degree, n_q_points_1d=None):
        self.degree = degree
        self.n = degree + 1
        self.n_q = n_q_points_1d if n_q_points_1d is not None else degree + 1
        self.q_points, self.q_weights = gauss_points_weights(self.n_q)
        self.support_points = gauss_lobatto_points(self.n)
        self.shape_values, self.shape_gradients = lagrange_values_and_derivatives(
            self.support_points, self.q_points)
        _, self.shape_gradients_collocation = lagrange_values_and_derivatives(
            self.q_points, self.q_points)
        # exact (skew-)symmetry, as check_and_set_shapes_symmetric relies on
        # (shape_info.templates.h:1102-1150)
        self.shape_values = 0.5 * (self.shape_values + self.shape_values[::-1, ::-1])
        self.shape_gradients = 0.5 * (self.shape_gradients - self.shape_gradients[::-1, ::-1])
        self.shape_gradients_collocation = 0.5 * (
            self.shape_gradients_collocation - self.shape_gradients_collocation[::-1, ::-1])
        # interpolation from the coarse (parent) line onto the first child half:
        # W[i, j] = l_j(0.5 * x_i)   (values of parent basis j at child node i)
        # shape_info.templates.h: subface_interpolation_matrices[0][i*n + j]
        v0, _ = lagrange_values_and_derivatives(self.support_points,
                                                0.5 * self.support_points)
        self.subface_interpolation_matrix = np.ascontiguousarray(v0.T)


def dealii_testing_rand(count):
    """Testing::rand() of tests/tests.h:232-275: the glibc TYPE_3 additive-feedback
    generator re-implemented there so that test output is platform independent.
    Returns ``count`` successive values (ints in [0, 2^31))."""
    r = [0] * 32
    r[0] = 1
    for i in range(1, 31):
        r[i] = (16807 * r[i - 1]) % 2147483647
        if r[i] < 0:
            r[i] += 2147483647
    k = 31
    for i in range(31, 34):
        r[k % 32] = r[(k + 32 - 31) % 32]
        k = (k + 1) % 32
    for i in range(34, 344):
        r[k % 32] = (r[(k + 32 - 31) % 32] + r[(k + 32 - 3) % 32]) & 0xFFFFFFFF
        k = (k + 1) % 32
    out = []
    for _ in range(count):
        r[k % 32] = (r[(k + 32 - 31) % 32] + r[(k + 32 - 3) % 32]) & 0xFFFFFFFF
        ret = r[k % 32]
        k = (k + 1) % 32
        out.append(ret >> 1)
    return out
