// 1D shape data of FE_Q(p) x QGauss(p+1), built on the host when the caller does not hand
// over ShapeInfo arrays.  Same definitions as the reference
// (matrix_free/shape_info.templates.h:861-985): Lagrange basis on the Gauss-Lobatto points
// evaluated at the Gauss points of [0,1]; collocation derivative = derivative of the
// Lagrange basis on the Gauss points themselves; subface matrix = parent basis at the
// nodes of the first child half.
#include <cmath>

#include "internal.h"

namespace b200mf {
namespace {

// Legendre polynomial P_k and derivative at x in [-1,1]
void legendre(int k, long double x, long double &P, long double &dP) {
  long double p0 = 1.0L, p1 = x;
  if (k == 0) { P = 1.0L; dP = 0.0L; return; }
  for (int j = 2; j <= k; ++j) {
    long double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
    p0 = p1; p1 = p2;
  }
  P = p1;
  dP = k * (x * p1 - p0) / (x * x - 1.0L);
}

std::vector<long double> gauss_points(int n, std::vector<long double> &w) {
  std::vector<long double> x(n);
  w.resize(n);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < n; ++i) {
    long double z = -std::cos(pi * (i + 0.75L) / (n + 0.5L));
    for (int it = 0; it < 100; ++it) {
      long double P, dP;
      legendre(n, z, P, dP);
      long double dz = P / dP;
      z -= dz;
      if (std::fabs((double)dz) < 1e-19) break;
    }
    long double P, dP;
    legendre(n, z, P, dP);
    x[i] = 0.5L * (z + 1.0L);
    w[i] = 1.0L / ((1.0L - z * z) * dP * dP); // = 0.5 * 2/((1-z^2) P'^2)
  }
  for (int i = 0; i < n / 2; ++i) { // exact symmetry
    long double a = 0.5L * (x[i] + (1.0L - x[n - 1 - i]));
    x[i] = a; x[n - 1 - i] = 1.0L - a;
    long double b = 0.5L * (w[i] + w[n - 1 - i]);
    w[i] = w[n - 1 - i] = b;
  }
  if (n % 2) x[n / 2] = 0.5L;
  return x;
}

// Gauss-Lobatto nodes: +-1 and the roots of P'_{n-1}
std::vector<long double> gauss_lobatto_points(int n) {
  std::vector<long double> x(n);
  if (n == 1) { x[0] = 0.5L; return x; }
  x[0] = 0.0L; x[n - 1] = 1.0L;
  const int k = n - 1;
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 1; i < n - 1; ++i) {
    long double z = -std::cos(pi * i / k);
    for (int it = 0; it < 100; ++it) {
      long double P, dP;
      legendre(k, z, P, dP);
      // (1-z^2) P'' = 2 z P' - k(k+1) P
      long double d2P = (2.0L * z * dP - k * (k + 1) * P) / (1.0L - z * z);
      long double dz = dP / d2P;
      z -= dz;
      if (std::fabs((double)dz) < 1e-19) break;
    }
    x[i] = 0.5L * (z + 1.0L);
  }
  for (int i = 0; i < n / 2; ++i) {
    long double a = 0.5L * (x[i] + (1.0L - x[n - 1 - i]));
    x[i] = a; x[n - 1 - i] = 1.0L - a;
  }
  if (n % 2) x[n / 2] = 0.5L;
  return x;
}

void lagrange(const std::vector<long double> &nodes, long double x, int i, long double &v,
              long double &d) {
  const int n = (int)nodes.size();
  long double denom = 1.0L;
  for (int j = 0; j < n; ++j)
    if (j != i) denom *= nodes[i] - nodes[j];
  v = 1.0L;
  for (int j = 0; j < n; ++j)
    if (j != i) v *= x - nodes[j];
  d = 0.0L;
  for (int k = 0; k < n; ++k) {
    if (k == i) continue;
    long double t = 1.0L;
    for (int j = 0; j < n; ++j)
      if (j != i && j != k) t *= x - nodes[j];
    d += t;
  }
  v /= denom;
  d /= denom;
}

} // namespace

void build_fe_q_support_points(int degree, std::vector<double> &points) {
  const std::vector<long double> xn = gauss_lobatto_points(degree + 1);
  points.assign(xn.begin(), xn.end());
}

// 1D prolongation of FE_Q(degree) from a cell to its two children: P[X * n + i] = l_i(x_X), X = 0..2p the
// nodes of both children in the parent's unit cell (child 0: X <= p at gl[X] / 2, child 1 at
// 1/2 + gl[X - p] / 2); the tensor product of three of them is the embedding matrix deal.II's
// MGTransferMatrixFree applies cell by cell (multigrid/mg_transfer_internal.cc: setup_element_info,
// "prolongation_matrix_1d")
void build_prolongation_1d(int degree, std::vector<double> &P) {
  const int n = degree + 1, M = 2 * degree + 1;
  const std::vector<long double> xn = gauss_lobatto_points(n);
  P.assign((size_t)M * n, 0.0);
  for (int X = 0; X < M; ++X) {
    const long double x = X <= degree ? 0.5L * xn[X] : 0.5L + 0.5L * xn[X - degree];
    for (int i = 0; i < n; ++i) {
      long double v, d;
      lagrange(xn, x, i, v, d);
      P[(size_t)X * n + i] = (double)v;
    }
  }
}

void build_overint_shape_data(int degree, int Q, std::vector<double> &S, std::vector<double> &Dq,
                              std::vector<double> &weights, std::vector<double> &points) {
  const int n = degree + 1;
  std::vector<long double> w;
  const std::vector<long double> xq = gauss_points(Q, w), xn = gauss_lobatto_points(n);
  S.assign((size_t)n * Q, 0.0);
  Dq.assign((size_t)Q * Q, 0.0);
  weights.resize(Q);
  points.resize(Q);
  for (int q = 0; q < Q; ++q) { weights[q] = (double)w[q]; points[q] = (double)xq[q]; }
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < Q; ++q) {
      long double v, d;
      lagrange(xn, xq[q], i, v, d);
      S[(size_t)i * Q + q] = (double)v;
    }
  for (int a = 0; a < Q; ++a)
    for (int b = 0; b < Q; ++b) {
      long double v, d;
      lagrange(xq, xq[b], a, v, d);
      Dq[(size_t)a * Q + b] = (double)d;
    }
}

void build_fe_q_shape_data(int degree, std::vector<double> &shape_values,
                           std::vector<double> &shape_grad_colloc, std::vector<double> &q_weights,
                           std::vector<double> &q_points, std::vector<double> &subface) {
  const int n = degree + 1;
  std::vector<long double> w;
  std::vector<long double> xq = gauss_points(n, w);
  std::vector<long double> xn = gauss_lobatto_points(n);
  shape_values.assign(n * n, 0.0);
  shape_grad_colloc.assign(n * n, 0.0);
  subface.assign(n * n, 0.0);
  q_weights.resize(n);
  q_points.resize(n);
  for (int q = 0; q < n; ++q) { q_weights[q] = (double)w[q]; q_points[q] = (double)xq[q]; }
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q) {
      long double v, d;
      lagrange(xn, xq[q], i, v, d);
      shape_values[i * n + q] = (double)v;
      lagrange(xq, xq[q], i, v, d);
      shape_grad_colloc[i * n + q] = (double)d;
      // subface[i][j] = l_j(x_i / 2): row = child node i, column = parent basis j
      lagrange(xn, 0.5L * xn[i], q, v, d);
      subface[i * n + q] = (double)v;
    }
  // enforce exact (skew-)symmetry, cf. shape_info.templates.h:1102-1150
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q) {
      int i2 = n - 1 - i, q2 = n - 1 - q;
      if (i * n + q < i2 * n + q2) {
        double a = 0.5 * (shape_values[i * n + q] + shape_values[i2 * n + q2]);
        shape_values[i * n + q] = shape_values[i2 * n + q2] = a;
        double b = 0.5 * (shape_grad_colloc[i * n + q] - shape_grad_colloc[i2 * n + q2]);
        shape_grad_colloc[i * n + q] = b;
        shape_grad_colloc[i2 * n + q2] = -b;
      }
    }
  if (n % 2) shape_grad_colloc[(n / 2) * n + n / 2] = 0.0;
}

namespace {
template <typename Number, int n>
void pack_eo(const double *M /* [i*n+q] */, bool transpose, EoMatrix<Number, n> &out) {
  constexpr int h = n / 2, hq = (n + 1) / 2;
  auto at = [&](int i, int q) { return transpose ? M[q * n + i] : M[i * n + q]; };
  for (int i = 0; i < h; ++i) {
    for (int q = 0; q < hq; ++q) out.E[i * hq + q] = Number(0.5 * (at(i, q) + at(i, n - 1 - q)));
    for (int q = 0; q < h; ++q) out.O[i * h + q] = Number(0.5 * (at(i, q) - at(i, n - 1 - q)));
  }
  for (int q = 0; q < hq; ++q) out.mid[q] = (n % 2) ? Number(at(h, q)) : Number(0);
}
} // namespace

template <typename Number, int n>
void fill_shape_data(const Setup &s, ShapeData<Number, n> &out) {
  pack_eo<Number, n>(s.shape_values.data(), false, out.S);
  pack_eo<Number, n>(s.shape_values.data(), true, out.St);
  pack_eo<Number, n>(s.shape_grad_colloc.data(), false, out.D);
  pack_eo<Number, n>(s.shape_grad_colloc.data(), true, out.Dt);
  // DtW: out[i] = sum_q (w[q] D[i][q]) in[q]; symmetric weights keep the skew structure
  std::vector<double> dw(n * n);
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q) dw[i * n + q] = s.q_weights[q] * s.shape_grad_colloc[i * n + q];
  pack_eo<Number, n>(dw.data(), true, out.DtW);
  for (int q = 0; q < n; ++q) out.w[q] = Number(s.q_weights[q]);
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) out.w2[a * n + b] = Number(s.q_weights[a] * s.q_weights[b]);
}

// 1D mass and stiffness matrices of the reference cell in the nodal basis,
//   M[i][j] = sum_q w_q phi_i(x_q) phi_j(x_q),  K[i][j] = sum_q w_q phi_i'(x_q) phi_j'(x_q),
// phi_i'(x_q) = sum_r S[i][r] D[r][q] (the collocation derivative is exact for degree p), scaled
// by the metric of the mesh's single Cartesian cell shape and the operator's constants:
//   A_cell = g (m0 K(x)M(x)M + m1 M(x)K(x)M + m2 M(x)M(x)K) + c det M(x)M(x)M
template <typename Number, int n>
void fill_brick_matrices(const Setup &s, const b200mf_operator &op, BrickMatrices<Number, n> &out,
                         uint32_t geom) {
  const double *G0 = (s.h_geom_table.size() >= 4 * ((size_t)geom + 1)) ? s.h_geom_table.data() + 4 * (size_t)geom : s.geom0;
  const double *S = s.shape_values.data(), *D = s.shape_grad_colloc.data(), *w = s.q_weights.data();
  double G[n * n], M[n * n], K[n * n];
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q) {
      double g = 0.0;
      for (int r = 0; r < n; ++r) g += S[i * n + r] * D[r * n + q];
      G[i * n + q] = g;
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double m = 0.0, k = 0.0;
      for (int q = 0; q < n; ++q) {
        m += w[q] * S[i * n + q] * S[j * n + q];
        k += w[q] * G[i * n + q] * G[j * n + q];
      }
      M[i * n + j] = m;
      K[i * n + j] = k;
    }
  // exact symmetry and centro-symmetry (round-off only)
  auto symmetrise = [&](double *A) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        const int i2 = n - 1 - i, j2 = n - 1 - j;
        const double a = 0.25 * (A[i * n + j] + A[j * n + i] + A[i2 * n + j2] + A[j2 * n + i2]);
        A[i * n + j] = A[j * n + i] = A[i2 * n + j2] = A[j2 * n + i2] = a;
      }
  };
  symmetrise(M);
  symmetrise(K);
  const double g = op.grad_constant;
  const bool has_mass = op.mass_coefficient != nullptr || op.mass_constant != 0.0;
  const double cm = has_mass ? op.mass_constant * G0[3] : 0.0;
  double Kx[n * n], Ky[n * n], Kz[n * n];
  for (int i = 0; i < n * n; ++i) {
    Kx[i] = g * G0[0] * K[i];
    Ky[i] = g * G0[1] * K[i];
    Kz[i] = g * G0[2] * K[i] + cm * M[i];
  }
  pack_eo<Number, n>(M, false, out.M);
  pack_eo<Number, n>(Kx, false, out.Kx);
  pack_eo<Number, n>(Ky, false, out.Ky);
  pack_eo<Number, n>(Kz, false, out.Kz);
}

#define INST(N)                                                                     \
  template void fill_brick_matrices<double, N>(const Setup &, const b200mf_operator &, BrickMatrices<double, N> &, uint32_t); \
  template void fill_brick_matrices<float, N>(const Setup &, const b200mf_operator &, BrickMatrices<float, N> &, uint32_t);   \
  template void fill_shape_data<double, N>(const Setup &, ShapeData<double, N> &);  \
  template void fill_shape_data<float, N>(const Setup &, ShapeData<float, N> &);
INST(2) INST(3) INST(4) INST(5) INST(6) INST(7) INST(8) INST(9)
#undef INST

} // namespace b200mf
