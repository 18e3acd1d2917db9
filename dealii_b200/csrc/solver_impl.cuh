// SolverCG + DiagonalMatrix (Jacobi) + PreconditionChebyshev on the device.
//
// Algebra follows the reference exactly (iteration counts must agree):
//   SolverCG::solve / IterationWorker::do_iteration   lac/solver_cg.h:1391-1470, 703-763
//   PreconditionChebyshev                             lac/precondition.h:2378-2408 (initial
//       guess), 2465-2572 + 3928-4020 (eigenvalue estimate), 4029-4121 (polynomial),
//       3154-3190 / 3358-3430 (vector updates)
// but the pass structure is the merged one of lac/solver_cg.h:862-1137: with Jacobi (or no)
// preconditioning an iteration is   [vmult kernel, p.Ap fused into its scatter epilogue]
// -> [post kernel: r -= alpha v, r.r and r.D^-1 r] -> [pre kernel: x += alpha p,
// p = beta p + D^-1 r], all scalars staying on the device; the host reads back one
// residual norm per iteration for SolverControl.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "vector_ops.cuh"

namespace b200mf {

// scalar slots on the device: slot(k) = scratch + 8*(k%3): [0] p.Ap  [1] r.r  [2] r.z
__device__ __forceinline__ double *slot(double *scratch, int k) { return scratch + 8 * (k % 3); }

template <typename Number>
__global__ void cg_init_kernel(Number *r, Number *p, const Number *b, const Number *Ax,
                               const Number *d, uint64_t n, double *scratch) {
  double rr = 0.0, rz = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Number ri = b[i];
    if (Ax) ri -= Ax[i];
    const Number zi = d ? d[i] * ri : ri;
    r[i] = ri;
    p[i] = zi;
    rr += double(ri) * double(ri);
    rz += double(ri) * double(zi);
  }
  rr = block_sum(rr);
  if (threadIdx.x == 0) atomicAdd(slot(scratch, 1) + 1, rr);
  rz = block_sum(rz);
  if (threadIdx.x == 0) atomicAdd(slot(scratch, 1) + 2, rz);
}

// after v = A p of iteration `it`:  alpha = rz/pAp;  r -= alpha v;  rr, rz of the new residual
template <typename Number>
__global__ void cg_post_kernel(Number *r, const Number *v, const Number *d, uint64_t n,
                               double *scratch, int it) {
  const double *cur = slot(scratch, it);
  double *nxt = slot(scratch, it + 1);
  const Number alpha = Number(cur[2] / cur[0]);
  double rr = 0.0, rz = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Number ri = r[i] - alpha * v[i];
    r[i] = ri;
    const Number zi = d ? d[i] * ri : ri;
    rr += double(ri) * double(ri);
    rz += double(ri) * double(zi);
  }
  rr = block_sum(rr);
  if (threadIdx.x == 0) atomicAdd(nxt + 1, rr);
  rz = block_sum(rz);
  if (threadIdx.x == 0) atomicAdd(nxt + 2, rz);
}

// end of iteration `it`: x += alpha p; p = beta p + D^-1 r  (beta = rz_new / rz_old)
template <typename Number>
__global__ void cg_pre_kernel(Number *x, Number *p, const Number *r, const Number *d, uint64_t n,
                              double *scratch, int it) {
  const double *cur = slot(scratch, it);
  const double *nxt = slot(scratch, it + 1);
  const Number alpha = Number(cur[2] / cur[0]);
  const Number beta = Number(nxt[2] / cur[2]);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Number pi = p[i];
    x[i] += alpha * pi;
    const Number zi = d ? d[i] * r[i] : r[i];
    p[i] = beta * pi + zi;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double *z = slot(scratch, it + 2);
    z[0] = z[1] = z[2] = 0.0;
  }
}

template <typename Number>
__global__ void cg_final_kernel(Number *x, const Number *p, uint64_t n, const double *scratch,
                                int it) {
  const double *cur = slot(const_cast<double *>(scratch), it);
  const Number alpha = Number(cur[2] / cur[0]);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    x[i] += alpha * p[i];
}

// ---- generic (host-scalar) pieces used with the Chebyshev preconditioner
template <typename Number>
__global__ void axpy_dot_kernel(Number *y, Number a, const Number *x, uint64_t n, double *out) {
  // y += a x; *out += y.y   (Vector::add_and_dot, lac/vector_operations_internal.h:2590)
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Number yi = y[i] + a * x[i];
    y[i] = yi;
    acc += double(yi) * double(yi);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
template <typename Number>
__global__ void dot2_kernel(const Number *x, const Number *y, uint64_t n, double *out) {
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    acc += double(x[i]) * double(y[i]);
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
template <typename Number>
__global__ void sadd2_kernel(Number *y, Number s, Number a, const Number *x, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = s * y[i] + a * x[i];
}

// Chebyshev: first step  sol = f2 * D^-1 rhs
template <typename Number>
__global__ void cheb_first_kernel(Number *sol, const Number *rhs, const Number *d, Number f2,
                                  uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    sol[i] = f2 * d[i] * rhs[i];
}
// later steps: sol_old <- (1+f1) sol - f1 sol_old + f2 D^-1 (rhs - t)   (then swap)
// (lac/precondition.h:3391-3425 "dealii::ChebyshevItk")
template <typename Number>
__global__ void cheb_update_kernel(Number *sol_old, const Number *sol, const Number *rhs,
                                   const Number *t, const Number *d, Number f1, Number f2,
                                   int use_old, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Number v = (Number(1) + f1) * sol[i] + f2 * d[i] * (rhs[i] - t[i]);
    if (use_old) v -= f1 * sol_old[i];
    sol_old[i] = v;
  }
}
template <typename Number>
__global__ void initial_guess_kernel(Number *v, uint64_t first, uint64_t n, double *sum) {
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Number x = Number((i + first) % 11);
    v[i] = x;
    acc += double(x);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(sum, acc);
}
template <typename Number>
__global__ void shift_kernel(Number *v, Number a, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    v[i] += a;
}

// eigenvalues of a small symmetric tridiagonal matrix (the reference calls LAPACK stev,
// source/lac/tridiagonal_matrix.cc:223-241); cyclic Jacobi rotations on the dense form.
static std::vector<double> tridiagonal_eigenvalues(const std::vector<double> &diag,
                                                   const std::vector<double> &off) {
  const int n = (int)diag.size();
  std::vector<double> A(n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    A[i * n + i] = diag[i];
    if (i + 1 < n) A[i * n + i + 1] = A[(i + 1) * n + i] = off[i];
  }
  for (int sweep = 0; sweep < 100; ++sweep) {
    double offn = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) offn += A[i * n + j] * A[i * n + j];
    if (offn < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
      }
  }
  std::vector<double> ev(n);
  for (int i = 0; i < n; ++i) ev[i] = std::fabs(A[i * n + i]);
  std::sort(ev.begin(), ev.end());
  return ev;
}

struct CgOptions {
  double tol;
  int max_it;
  bool iteration_number_control; // reaching max_it counts as success
  bool track_eigenvalues;
  int check_every = 1; // read the residual back only every k-th iteration (not with track_eigenvalues)
};
struct CgOutcome {
  int iterations = 0;
  double residual = 0.0, initial_residual = 0.0;
  bool success = false;
  std::vector<double> eigenvalues;
  uint64_t vmults = 0;
};

static int ensure_work(Setup &s, int count) {
  const uint64_t elems = s.n_owned + s.n_ghost;
  for (int i = 0; i < count; ++i)
    if (!s.d_work[i])
      B200MF_CUDA_CHECK(cudaMalloc(&s.d_work[i], std::max<uint64_t>(elems, 1) * number_size(s.number)));
  return B200MF_OK;
}

// Fused CG with diagonal (or no) preconditioner; d == nullptr => PreconditionIdentity.
template <typename Number>
static int cg_fused(Setup &s, const b200mf_operator &op, Number *x, const Number *b,
                    const Number *d, const CgOptions &opt, CgOutcome &out, cudaStream_t st,
                    const b200mf_partitioner *part = nullptr) {
  int rc = ensure_work(s, 3);
  if (rc != B200MF_OK) return rc;
  Number *r = (Number *)s.d_work[0], *p = (Number *)s.d_work[1], *v = (Number *)s.d_work[2];
  const uint64_t n = s.n_owned;
  const unsigned grid = vec_grid(n);
  double *sc = s.d_scratch;
  double *h = s.h_pinned;
  B200MF_CUDA_CHECK(cudaMemsetAsync(sc, 0, 64 * sizeof(double), st));
  // ghost part of work vectors must be defined (zero) for the gather
  if (s.n_ghost) B200MF_CUDA_CHECK(cudaMemsetAsync(p + n, 0, s.n_ghost * sizeof(Number), st));

  // startup(): r = b - A x unless x == 0 (solver_cg.h:640-652)
  dot2_kernel<Number><<<grid, kVecThreads, 0, st>>>(x, x, n, sc + 32);
  count_launch();
  if ((rc = level_allreduce(part, sc + 32, 1, st)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(h, sc + 32, sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  const bool x_zero = (h[0] == 0.0);
  if (!x_zero) {
    rc = level_vmult(s, part, op, v, x, st, nullptr);
    if (rc != B200MF_OK) return rc;
    out.vmults++;
  }
  cg_init_kernel<Number><<<grid, kVecThreads, 0, st>>>(r, p, b, x_zero ? nullptr : v, d, n, sc);
  count_launch();
  if ((rc = level_allreduce(part, sc + 8 + 1, 2, st)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(h, sc + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  double res = std::sqrt(h[1]);
  out.initial_residual = out.residual = res;
  auto state = [&](int step, double value) {
    if (value <= opt.tol || (opt.iteration_number_control && step >= opt.max_it)) return 1;
    if (step >= opt.max_it || std::isnan(value)) return -1;
    return 0;
  };
  int stt = state(0, res);
  int it = 0;
  std::vector<double> diag, off;
  double eig_beta_alpha = 0.0, alpha = 0.0;
  while (stt == 0) {
    ++it;
    rc = level_vmult(s, part, op, v, p, st, sc + 8 * (it % 3));
    if (rc != B200MF_OK) return rc;
    out.vmults++;
    if ((rc = level_allreduce(part, sc + 8 * (it % 3), 1, st)) != B200MF_OK) return rc;
    cg_post_kernel<Number><<<grid, kVecThreads, 0, st>>>(r, v, d, n, sc, it);
    count_launch();
    if ((rc = level_allreduce(part, sc + 8 * ((it + 1) % 3) + 1, 2, st)) != B200MF_OK) return rc;
    const int every = opt.track_eigenvalues ? 1 : std::max(opt.check_every, 1);
    if (it % every != 0 && it < opt.max_it) {
      // no look at the residual this iteration: keep the device busy
      cg_pre_kernel<Number><<<grid, kVecThreads, 0, st>>>(x, p, r, d, n, sc, it);
      count_launch();
      continue;
    }
    B200MF_CUDA_CHECK(cudaMemcpyAsync(h, sc, 24 * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
    const double *cur = h + 8 * (it % 3), *nxt = h + 8 * ((it + 1) % 3);
    res = std::sqrt(std::fabs(nxt[1]));
    alpha = cur[2] / cur[0];
    stt = state(it, res);
    // Lanczos coefficients (solver_cg.h:1440-1452): recorded at the *next* iteration in the
    // reference (it > 1 uses previous_alpha and the current beta); beta_{it+1} = nxt[2]/cur[2]
    if (opt.track_eigenvalues) {
      // at reference iteration it+1 (if it happens): diag.push(1/alpha_it + eig_beta_alpha),
      // eig_beta_alpha = beta_{it+1}/alpha_it, off.push(sqrt(beta_{it+1})/alpha_it)
      if (stt == 0) {
        const double beta_next = nxt[2] / cur[2];
        diag.push_back(1.0 / alpha + eig_beta_alpha);
        eig_beta_alpha = beta_next / alpha;
        off.push_back(std::sqrt(beta_next) / alpha);
      }
    }
    if (stt != 0) {
      cg_final_kernel<Number><<<grid, kVecThreads, 0, st>>>(x, p, n, sc, it);
      count_launch();
      break;
    }
    cg_pre_kernel<Number><<<grid, kVecThreads, 0, st>>>(x, p, r, d, n, sc, it);
    count_launch();
  }
  B200MF_CUDA_CHECK(cudaGetLastError());
  out.iterations = it;
  out.residual = res;
  out.success = (stt == 1);
  if (opt.track_eigenvalues) out.eigenvalues = tridiagonal_eigenvalues(diag, off);
  return B200MF_OK;
}

template <typename Number>
struct Chebyshev {
  Setup &s;
  const b200mf_operator &op;
  const Number *d;
  int degree;
  double theta = 1.0, delta = 1.0;
  uint64_t vmults = 0;
  const b200mf_partitioner *part = nullptr; // the level is partitioned over ranks
  // z = P(rhs); uses work vectors 3 (sol_old) and 4 (t)
  int apply(Number *z, const Number *rhs, cudaStream_t st) {
    const uint64_t n = s.n_owned;
    const unsigned grid = vec_grid(n);
    Number *sol = z, *sol_old = (Number *)s.d_work[3], *t = (Number *)s.d_work[4];
    cheb_first_kernel<Number><<<grid, kVecThreads, 0, st>>>(sol, rhs, d, Number(1.0 / theta), n);
    count_launch();
    if (degree < 2 || std::fabs(delta) < 1e-40) return B200MF_OK;
    double rhok = delta / theta, sigma = theta / delta;
    for (int k = 0; k < degree - 1; ++k) {
      const double rhokp = 1.0 / (2.0 * sigma - rhok);
      const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / delta;
      rhok = rhokp;
      int rc = level_vmult(s, part, op, t, sol, st, nullptr);
      if (rc != B200MF_OK) return rc;
      vmults++;
      cheb_update_kernel<Number><<<grid, kVecThreads, 0, st>>>(sol_old, sol, rhs, t, d, Number(f1),
                                                               Number(f2), k > 0, n);
      count_launch();
      std::swap(sol, sol_old);
    }
    if (sol != z)
      B200MF_CUDA_CHECK(cudaMemcpyAsync(z, sol, n * sizeof(Number), cudaMemcpyDeviceToDevice, st));
    return B200MF_OK;
  }
  // x <- x + P(rhs - A x): PreconditionChebyshev::step, the polynomial started from a nonzero guess
  // (apply_internal with zero_out_dst = false, precondition.h:4029-4121): one more operator
  // application than apply()
  int step(Number *x, const Number *rhs, cudaStream_t st) {
    const uint64_t n = s.n_owned;
    const unsigned grid = vec_grid(n);
    Number *sol = x, *sol_old = (Number *)s.d_work[3], *t = (Number *)s.d_work[4];
    double rhok = delta / theta, sigma = theta / delta;
    const int n_updates = (degree < 2 || std::fabs(delta) < 1e-40) ? 1 : degree;
    for (int k = 0; k < n_updates; ++k) {
      double f1 = 0.0, f2 = 1.0 / theta;
      if (k > 0) {
        const double rhokp = 1.0 / (2.0 * sigma - rhok);
        f1 = rhokp * rhok;
        f2 = 2.0 * rhokp / delta;
        rhok = rhokp;
      }
      int rc = level_vmult(s, part, op, t, sol, st, nullptr);
      if (rc != B200MF_OK) return rc;
      vmults++;
      cheb_update_kernel<Number><<<grid, kVecThreads, 0, st>>>(sol_old, sol, rhs, t, d, Number(f1),
                                                               Number(f2), k > 0, n);
      count_launch();
      std::swap(sol, sol_old);
    }
    if (sol != x)
      B200MF_CUDA_CHECK(cudaMemcpyAsync(x, sol, n * sizeof(Number), cudaMemcpyDeviceToDevice, st));
    return B200MF_OK;
  }
};

// internal::estimate_eigenvalues of PreconditionChebyshev (precondition.h:2465-2572): the Lanczos
// eigenvalues of `eig_cg_n_iterations` Jacobi-CG iterations on the vector of set_initial_guess
// (entries (i + first) % 11 minus their mean, constrained entries zero).  Uses work vectors 0..4.
template <typename Number>
static int estimate_eigenvalues(Setup &s, const b200mf_operator &op, const Number *d, int eig_cg_n_iterations,
                                uint64_t first_owned_global_index, double safety_factor, cudaStream_t st,
                                double &lmin, double &lmax, uint64_t &vmults, int *cg_iterations,
                                bool zero_constrained = true, const b200mf_partitioner *part = nullptr) {
  lmin = lmax = 1.0;
  if (cg_iterations) *cg_iterations = 0;
  if (eig_cg_n_iterations <= 0) return B200MF_OK;
  int rc = ensure_work(s, 5);
  if (rc != B200MF_OK) return rc;
  const uint64_t n = s.n_owned;
  const unsigned grid = vec_grid(n);
  Number *t1 = (Number *)s.d_work[4], *sol = (Number *)s.d_work[3];
  // [56] sum of the entries, [57] number of entries (both summed over the ranks)
  const double n_local = double(n);
  B200MF_CUDA_CHECK(cudaMemsetAsync(s.d_scratch + 56, 0, sizeof(double), st));
  B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_scratch + 57, &n_local, sizeof(double), cudaMemcpyHostToDevice, st));
  B200MF_CUDA_CHECK(cudaMemsetAsync(sol, 0, (s.n_owned + s.n_ghost) * sizeof(Number), st));
  B200MF_CUDA_CHECK(cudaMemsetAsync(t1, 0, (s.n_owned + s.n_ghost) * sizeof(Number), st));
  initial_guess_kernel<Number><<<grid, kVecThreads, 0, st>>>(t1, first_owned_global_index, n, s.d_scratch + 56);
  count_launch();
  if ((rc = level_allreduce(part, s.d_scratch + 56, 2, st)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(s.h_pinned, s.d_scratch + 56, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  const double mean = s.h_pinned[0] / s.h_pinned[1];
  shift_kernel<Number><<<grid, kVecThreads, 0, st>>>(t1, Number(-mean), n);
  count_launch();
  // constraints.set_zero(temp_vector1): AdditionalData::constraints -- empty in step-37's smoothers, where
  // the operator itself is the identity on constrained rows
  if (zero_constrained && (rc = set_constrained_impl(s, t1, 0.0, st)) != B200MF_OK) return rc;
  // temp_vector1.all_zero(): every dof constrained (a one-cell level) -> both estimates stay 1
  B200MF_CUDA_CHECK(cudaMemsetAsync(s.d_scratch + 57, 0, sizeof(double), st));
  dot2_kernel<Number><<<grid, kVecThreads, 0, st>>>(t1, t1, n, s.d_scratch + 57);
  count_launch();
  if ((rc = level_allreduce(part, s.d_scratch + 57, 1, st)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(s.h_pinned, s.d_scratch + 57, sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  if (s.h_pinned[0] == 0.0) return B200MF_OK;
  CgOptions eopt{1e-10, eig_cg_n_iterations, true, true};
  CgOutcome eout;
  // the Lanczos CG must not clobber t1 (its rhs) nor sol: it uses work 0..2 only
  if ((rc = cg_fused<Number>(s, op, sol, t1, d, eopt, eout, st, part)) != B200MF_OK) return rc;
  vmults += eout.vmults;
  if (cg_iterations) *cg_iterations = eout.iterations;
  if (!eout.eigenvalues.empty()) {
    lmin = eout.eigenvalues.front();
    lmax = safety_factor * eout.eigenvalues.back();
  }
  return B200MF_OK;
}

// Textbook PCG with a general preconditioner (do_iteration, solver_cg.h:703-763), scalars
// through the host; used with the Chebyshev polynomial whose own cost dominates.
template <typename Number, typename Preconditioner>
static int cg_generic(Setup &s, const b200mf_operator &op, Number *x, const Number *b,
                      Preconditioner &prec, const CgOptions &opt, CgOutcome &out,
                      cudaStream_t st, const b200mf_partitioner *part = nullptr) {
  int rc = ensure_work(s, 6);
  if (rc != B200MF_OK) return rc;
  Number *r = (Number *)s.d_work[0], *p = (Number *)s.d_work[1], *v = (Number *)s.d_work[2];
  Number *z = (Number *)s.d_work[5];
  const uint64_t n = s.n_owned;
  const unsigned grid = vec_grid(n);
  double *sc = s.d_scratch, *h = s.h_pinned;
  if (s.n_ghost) {
    B200MF_CUDA_CHECK(cudaMemsetAsync(p + n, 0, s.n_ghost * sizeof(Number), st));
    B200MF_CUDA_CHECK(cudaMemsetAsync(z + n, 0, s.n_ghost * sizeof(Number), st));
    B200MF_CUDA_CHECK(cudaMemsetAsync((Number *)s.d_work[3] + n, 0, s.n_ghost * sizeof(Number), st));
  }
  auto dot = [&](const Number *a, const Number *c, double &result) -> int {
    B200MF_CUDA_CHECK(cudaMemsetAsync(sc + 40, 0, sizeof(double), st));
    dot2_kernel<Number><<<grid, kVecThreads, 0, st>>>(a, c, n, sc + 40);
    count_launch();
    if (int rca = level_allreduce(part, sc + 40, 1, st)) return rca;
    B200MF_CUDA_CHECK(cudaMemcpyAsync(h, sc + 40, sizeof(double), cudaMemcpyDeviceToHost, st));
    B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
    result = h[0];
    return B200MF_OK;
  };
  double xx;
  if ((rc = dot(x, x, xx)) != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaMemcpyAsync(r, b, n * sizeof(Number), cudaMemcpyDeviceToDevice, st));
  if (xx != 0.0) {
    if ((rc = level_vmult(s, part, op, v, x, st, nullptr)) != B200MF_OK) return rc;
    out.vmults++;
    sadd2_kernel<Number><<<grid, kVecThreads, 0, st>>>(r, Number(1), Number(-1), v, n);
    count_launch();
  }
  double rr;
  if ((rc = dot(r, r, rr)) != B200MF_OK) return rc;
  double res = std::sqrt(rr);
  out.initial_residual = res;
  auto state = [&](int step, double value) {
    if (value <= opt.tol || (opt.iteration_number_control && step >= opt.max_it)) return 1;
    if (step >= opt.max_it || std::isnan(value)) return -1;
    return 0;
  };
  int stt = state(0, res), it = 0;
  double rpr = 0.0;
  while (stt == 0) {
    ++it;
    const double prev = rpr;
    if ((rc = prec.apply(z, r, st)) != B200MF_OK) return rc;
    if ((rc = dot(r, z, rpr)) != B200MF_OK) return rc;
    if (it > 1) {
      sadd2_kernel<Number><<<grid, kVecThreads, 0, st>>>(p, Number(rpr / prev), Number(1), z, n);
      count_launch();
    } else {
      B200MF_CUDA_CHECK(cudaMemcpyAsync(p, z, n * sizeof(Number), cudaMemcpyDeviceToDevice, st));
    }
    B200MF_CUDA_CHECK(cudaMemsetAsync(sc + 48, 0, 2 * sizeof(double), st));
    if ((rc = level_vmult(s, part, op, v, p, st, sc + 48)) != B200MF_OK) return rc;
    out.vmults++;
    if ((rc = level_allreduce(part, sc + 48, 1, st)) != B200MF_OK) return rc;
    B200MF_CUDA_CHECK(cudaMemcpyAsync(h, sc + 48, sizeof(double), cudaMemcpyDeviceToHost, st));
    B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
    const double alpha = rpr / h[0];
    sadd2_kernel<Number><<<grid, kVecThreads, 0, st>>>(x, Number(1), Number(alpha), p, n);
    axpy_dot_kernel<Number><<<grid, kVecThreads, 0, st>>>(r, Number(-alpha), v, n, sc + 49);
    count_launch(2);
    if ((rc = level_allreduce(part, sc + 49, 1, st)) != B200MF_OK) return rc;
    B200MF_CUDA_CHECK(cudaMemcpyAsync(h, sc + 49, sizeof(double), cudaMemcpyDeviceToHost, st));
    B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
    res = std::sqrt(std::fabs(h[0]));
    stt = state(it, res);
  }
  out.iterations = it;
  out.residual = res;
  out.success = (stt == 1);
  out.vmults += prec.vmults;
  return B200MF_OK;
}

} // namespace b200mf
