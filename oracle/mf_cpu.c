/* mf_cpu.c -- CPU restatement (plain C, GCC vector extensions) of deal.II's vectorised CPU
 * MatrixFree Laplace operator:  dst = sum_cells P^T (grad phi_i, grad phi_j) P src.
 *
 * TEST INFRASTRUCTURE ONLY.  Part of oracle/: used by tests/ as a second, independent
 * checker next to the numpy oracle (oracle/mf_oracle.py) and by bench.py's `cpu_baseline` /
 * `--impl reference` legs as the thing timed on the host cores.  The product
 * (dealii_b200 + libb200mf.so) never links, loads or calls it.
 *
 * Parity pinning: deal.II needs its cmake build, generated headers and a 200 MB library, so
 * it is treated as unbuildable here (DESIGN.md "Oracle"); this file is checked against the
 * numpy oracle, which in turn is pinned on the golden vectors the reference ships
 * (tests/test_oracle_golden.py).
 *
 * What it follows (paths relative to the deal.II tree):
 *   operator           tests/performance/timing_matrix_free_kokkos.cc:56-111 (LaplaceOperator
 *                      on CPU MatrixFree: read_dof_values, evaluate(gradients),
 *                      submit_gradient(get_gradient), integrate, distribute_local_to_global)
 *   loop               matrix_free/matrix_free.h:5090 (cell_loop, zero_dst_vector = true)
 *   kernels            see mf_cpu_kernel.inc
 *   geometry storage   matrix_free/mapping_info_storage.h:50-73,194-305: one inverse Jacobian
 *                      and determinant per Cartesian cell, per-q-point data on general cells
 *   threading          the reference image has neither MPI nor TBB, cell_loop is serial per
 *                      process (source/matrix_free/task_info.cc:362); callers run one
 *                      independent instance per thread, like one MPI rank per core with zero
 *                      communication cost.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LANES 8
typedef double vd __attribute__((vector_size(8 * LANES), aligned(64)));

#define MAXN 9
typedef struct {
  double E[(MAXN / 2) * ((MAXN + 1) / 2)], O[(MAXN / 2) * (MAXN / 2)], mid[(MAXN + 1) / 2];
} eo_matrix;

typedef struct mfcpu {
  int dim, n;
  uint64_t n_cells, n_dofs;
  uint32_t *l2g;
  int cartesian;
  double *inv_jac_diag, *det; /* cartesian: [cells][dim], [cells]          */
  double *inv_jac, *jxw;      /* general:   [cells][nq][dim][dim], [cells][nq] */
  double *qw;                 /* tensor-product quadrature weights [nq]     */
  eo_matrix S, St, D, Dt;
} mfcpu;

/* even-odd packing of out[q] = sum_i M[i*n+q] in[i]
 * (shape_info.templates.h:1153-1180 convert_to_eo) */
static void pack_eo(const double *M, int n, eo_matrix *out) {
  const int h = n / 2, hq = (n + 1) / 2;
  for (int i = 0; i < h; ++i) {
    for (int q = 0; q < hq; ++q) out->E[i * hq + q] = 0.5 * (M[i * n + q] + M[i * n + n - 1 - q]);
    for (int q = 0; q < h; ++q) out->O[i * h + q] = 0.5 * (M[i * n + q] - M[i * n + n - 1 - q]);
  }
  for (int q = 0; q < hq; ++q) out->mid[q] = (n % 2) ? M[h * n + q] : 0.0;
}

static void q1_jacobian(int dim, const double *v, const double *xi, double J[3][3]) {
  for (int d = 0; d < 3; ++d)
    for (int e = 0; e < 3; ++e) J[d][e] = 0.0;
  for (int k = 0; k < (1 << dim); ++k) {
    double dN[3];
    for (int e = 0; e < dim; ++e) dN[e] = 1.0;
    for (int d = 0; d < dim; ++d) {
      const int b = (k >> d) & 1;
      const double f = b ? xi[d] : 1.0 - xi[d], df = b ? 1.0 : -1.0;
      for (int e = 0; e < dim; ++e) dN[e] *= (e == d) ? df : f;
    }
    for (int d = 0; d < dim; ++d)
      for (int e = 0; e < dim; ++e) J[d][e] += dN[e] * v[k * dim + d];
  }
}

static double invert(int dim, double J[3][3], double Ji[3][3]) {
  if (dim == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det;
    Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
    return det;
  }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  Ji[0][0] = c00 / det; Ji[1][0] = c01 / det; Ji[2][0] = c02 / det;
  Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  return det;
}

/* shape_values / shape_gradients_collocation: row-major [i*n+q]
 * (shape_info.templates.h:895-898, 973-984); q_points / q_weights: QGauss<1>(n) on [0,1].
 * The arrays l2g ([cells][n^dim], lexicographic) and cell_vertices ([cells][2^dim][dim]) are
 * copied. */
mfcpu *mfcpu_create(int dim, int degree, uint64_t n_cells, uint64_t n_dofs, const uint32_t *l2g,
                    const double *cell_vertices, const double *shape_values,
                    const double *shape_gradients_collocation, const double *q_points,
                    const double *q_weights) {
  const int n = degree + 1;
  if (dim < 2 || dim > 3 || n < 2 || n > MAXN) return NULL;
  mfcpu *s = (mfcpu *)calloc(1, sizeof(mfcpu));
  s->dim = dim; s->n = n; s->n_cells = n_cells; s->n_dofs = n_dofs;
  const int npc = dim == 2 ? n * n : n * n * n, nv = 1 << dim;
  s->l2g = (uint32_t *)malloc(sizeof(uint32_t) * n_cells * npc);
  memcpy(s->l2g, l2g, sizeof(uint32_t) * n_cells * npc);
  double T[MAXN * MAXN];
  pack_eo(shape_values, n, &s->S);
  pack_eo(shape_gradients_collocation, n, &s->D);
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q) T[q * n + i] = shape_values[i * n + q];
  pack_eo(T, n, &s->St);
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q) T[q * n + i] = shape_gradients_collocation[i * n + q];
  pack_eo(T, n, &s->Dt);
  s->qw = (double *)malloc(sizeof(double) * npc);
  for (int q = 0; q < npc; ++q) {
    double w = q_weights[q % n] * q_weights[(q / n) % n];
    if (dim == 3) w *= q_weights[q / (n * n)];
    s->qw[q] = w;
  }
  /* cell type detection (mapping_info.templates.h:428-573): all cells Cartesian? */
  int cart = 1;
  for (uint64_t c = 0; c < n_cells && cart; ++c) {
    const double *v = cell_vertices + c * nv * dim;
    double scale = 0.0;
    for (int e = 0; e < dim; ++e) scale = fmax(scale, fabs(v[(1 << e) * dim + e] - v[e]));
    for (int k = 0; k < nv && cart; ++k)
      for (int x = 0; x < dim; ++x) {
        const double pred = v[x] + (((k >> x) & 1) ? v[(1 << x) * dim + x] - v[x] : 0.0);
        if (fabs(pred - v[k * dim + x]) > 1e-12 * scale) { cart = 0; break; }
      }
  }
  s->cartesian = cart;
  if (cart) {
    s->inv_jac_diag = (double *)malloc(sizeof(double) * n_cells * dim);
    s->det = (double *)malloc(sizeof(double) * n_cells);
    for (uint64_t c = 0; c < n_cells; ++c) {
      const double *v = cell_vertices + c * nv * dim;
      double det = 1.0;
      for (int d = 0; d < dim; ++d) {
        const double h = v[(1 << d) * dim + d] - v[d];
        s->inv_jac_diag[c * dim + d] = 1.0 / h;
        det *= h;
      }
      s->det[c] = det;
    }
  } else {
    s->inv_jac = (double *)malloc(sizeof(double) * n_cells * npc * dim * dim);
    s->jxw = (double *)malloc(sizeof(double) * n_cells * npc);
    for (uint64_t c = 0; c < n_cells; ++c)
      for (int q = 0; q < npc; ++q) {
        const double xi[3] = {q_points[q % n], q_points[(q / n) % n],
                              dim == 3 ? q_points[q / (n * n)] : 0.0};
        double J[3][3], Ji[3][3];
        q1_jacobian(dim, cell_vertices + c * nv * dim, xi, J);
        const double det = invert(dim, J, Ji);
        for (int a = 0; a < dim; ++a)
          for (int b = 0; b < dim; ++b) s->inv_jac[((c * npc + q) * dim + a) * dim + b] = Ji[a][b];
        s->jxw[c * npc + q] = det * s->qw[q];
      }
  }
  return s;
}

void mfcpu_destroy(mfcpu *s) {
  if (!s) return;
  free(s->l2g); free(s->inv_jac_diag); free(s->det); free(s->inv_jac); free(s->jxw); free(s->qw);
  free(s);
}

#define DIM 2
#define N 2
#include "mf_cpu_kernel.inc"
#undef N
#define N 3
#include "mf_cpu_kernel.inc"
#undef N
#define N 4
#include "mf_cpu_kernel.inc"
#undef N
#define N 5
#include "mf_cpu_kernel.inc"
#undef N
#define N 6
#include "mf_cpu_kernel.inc"
#undef N
#define N 7
#include "mf_cpu_kernel.inc"
#undef N
#define N 8
#include "mf_cpu_kernel.inc"
#undef N
#define N 9
#include "mf_cpu_kernel.inc"
#undef N
#undef DIM
#define DIM 3
#define N 2
#include "mf_cpu_kernel.inc"
#undef N
#define N 3
#include "mf_cpu_kernel.inc"
#undef N
#define N 4
#include "mf_cpu_kernel.inc"
#undef N
#define N 5
#include "mf_cpu_kernel.inc"
#undef N
#define N 6
#include "mf_cpu_kernel.inc"
#undef N
#define N 7
#include "mf_cpu_kernel.inc"
#undef N
#define N 8
#include "mf_cpu_kernel.inc"
#undef N
#define N 9
#include "mf_cpu_kernel.inc"
#undef N
#undef DIM

typedef void (*range_fn)(const mfcpu *, const double *, double *, uint64_t, uint64_t);
static const range_fn k_table[2][MAXN + 1] = {
    {0, 0, cell_range_2_2, cell_range_2_3, cell_range_2_4, cell_range_2_5, cell_range_2_6,
     cell_range_2_7, cell_range_2_8, cell_range_2_9},
    {0, 0, cell_range_3_2, cell_range_3_3, cell_range_3_4, cell_range_3_5, cell_range_3_6,
     cell_range_3_7, cell_range_3_8, cell_range_3_9}};

/* dst = A src (dst zeroed first, as cell_loop(..., zero_dst_vector = true)); serial. */
void mfcpu_vmult(const mfcpu *s, const double *src, double *dst) {
  memset(dst, 0, sizeof(double) * s->n_dofs);
  k_table[s->dim - 2][s->n](s, src, dst, 0, s->n_cells);
}

/* `repeat` back-to-back applications (the timed loop of the baseline; keeps the Python
 * caller out of the measurement). */
void mfcpu_vmult_repeat(const mfcpu *s, const double *src, double *dst, int repeat) {
  for (int r = 0; r < repeat; ++r) mfcpu_vmult(s, src, dst);
}

int mfcpu_is_cartesian(const mfcpu *s) { return s->cartesian; }
