// C entry points of SolverCG with Jacobi / Chebyshev preconditioning; the kernels and the two CG
// drivers live in solver_impl.cuh (shared with multigrid.cu).
#include "solver_impl.cuh"

namespace b200mf {

template <typename Number>
static int solve_impl(Setup &s, const b200mf_operator &op, const b200mf_solver_desc &sd, void *xv,
                      const void *bv, b200mf_solver_result *result, cudaStream_t st) {
  Number *x = (Number *)xv;
  const Number *b = (const Number *)bv;
  const Number *d = (const Number *)sd.inverse_diagonal;
  CgOptions opt{sd.tolerance, sd.max_iterations, false, false, sd.check_every};
  CgOutcome out;
  int rc;
  double ev_min = 0.0, ev_max = 0.0;
  if (sd.preconditioner == B200MF_PRECOND_CHEBYSHEV) {
    B200MF_REQUIRE(d != nullptr, "Chebyshev needs the inverse diagonal");
    B200MF_REQUIRE(sd.chebyshev_degree >= 1, "Chebyshev degree must be positive");
    if ((rc = ensure_work(s, 6)) != B200MF_OK) return rc;
    double lmax = 1.0, lmin = 1.0;
    uint64_t extra_vmults = 0;
    if ((rc = estimate_eigenvalues<Number>(s, op, d, sd.eig_cg_n_iterations, sd.first_owned_global_index,
                                           sd.safety_factor > 0 ? sd.safety_factor : 1.2, st, lmin, lmax,
                                           extra_vmults, nullptr,
                                           /*zero_constrained=*/sd.eig_keep_constrained_entries == 0)) != B200MF_OK)
      return rc;
    if (sd.eig_cg_n_iterations <= 0) {
      // no estimate: AdditionalData::max_eigenvalue and the smoothing range define the interval
      lmax = sd.max_eigenvalue > 0.0 ? sd.max_eigenvalue : 1.0;
      lmin = sd.smoothing_range > 0.0 ? lmax / sd.smoothing_range : lmax;
    }
    const double alpha = sd.smoothing_range > 1.0 ? lmax / sd.smoothing_range
                                                  : std::min(0.9 * lmax, lmin);
    Chebyshev<Number> prec{s, op, d, sd.chebyshev_degree};
    prec.delta = (lmax - alpha) * 0.5;
    prec.theta = (lmax + alpha) * 0.5;
    ev_min = lmin; ev_max = lmax;
    rc = cg_generic<Number>(s, op, x, b, prec, opt, out, st);
    out.vmults += extra_vmults;
  } else {
    if (sd.preconditioner == B200MF_PRECOND_JACOBI)
      B200MF_REQUIRE(d != nullptr, "Jacobi needs the inverse diagonal");
    rc = cg_fused<Number>(s, op, x, b, sd.preconditioner == B200MF_PRECOND_JACOBI ? d : nullptr,
                          opt, out, st);
  }
  if (rc != B200MF_OK) return rc;
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  if (result) {
    result->iterations = out.iterations;
    result->residual = out.residual;
    result->initial_residual = out.initial_residual;
    result->chebyshev_max_eigenvalue = ev_max;
    result->chebyshev_min_eigenvalue = ev_min;
    result->operator_applications = out.vmults;
  }
  if (!out.success) {
    set_error("CG did not converge: %d iterations, residual %g (SolverControl::NoConvergence)",
              out.iterations, out.residual);
    return B200MF_ERR_NOCONVERGENCE;
  }
  return B200MF_OK;
}

} // namespace b200mf

using namespace b200mf;

extern "C" {

int b200mf_cg_solve(const b200mf_setup *h, const b200mf_operator *op,
                    const b200mf_solver_desc *solver, void *x, const void *b,
                    b200mf_solver_result *result, void *stream) {
  B200MF_REQUIRE(h && op && solver && x && b, "null argument");
  Setup &s = const_cast<Setup &>(h->impl);
  if (s.number == B200MF_F64)
    return solve_impl<double>(s, *op, *solver, x, b, result, (cudaStream_t)stream);
  return solve_impl<float>(s, *op, *solver, x, b, result, (cudaStream_t)stream);
}

/* ---- building blocks of the fused CG for callers that interleave their own communication
 * (multi-GPU: one all-reduce of the scalar slots between the kernels).  `scratch` is a
 * caller-owned, zero-initialised device array of 24 doubles: slot(k) = scratch + 8*(k%3)
 * holds [p.Ap, r.r, r.z] of iteration k. */
#define B200MF_DISPATCH_NUMBER(number, CALL)                      \
  do {                                                            \
    if ((number) == B200MF_F64) { using T = double; CALL; }       \
    else if ((number) == B200MF_F32) { using T = float; CALL; }   \
    else { set_error("bad number type"); return B200MF_ERR_INVALID; } \
  } while (0)

int b200mf_cg_init(int number, void *r, void *p, const void *b, const void *Ax, const void *d,
                   uint64_t n, double *scratch, void *stream) {
  B200MF_REQUIRE(r && p && b && scratch, "null argument");
  B200MF_DISPATCH_NUMBER(number, (cg_init_kernel<T><<<vec_grid(n), kVecThreads, 0, (cudaStream_t)stream>>>(
      (T *)r, (T *)p, (const T *)b, (const T *)Ax, (const T *)d, n, scratch)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}
int b200mf_cg_post(int number, void *r, const void *v, const void *d, uint64_t n, double *scratch,
                   int it, void *stream) {
  B200MF_REQUIRE(r && v && scratch, "null argument");
  B200MF_DISPATCH_NUMBER(number, (cg_post_kernel<T><<<vec_grid(n), kVecThreads, 0, (cudaStream_t)stream>>>(
      (T *)r, (const T *)v, (const T *)d, n, scratch, it)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}
int b200mf_cg_pre(int number, void *x, void *p, const void *r, const void *d, uint64_t n,
                  double *scratch, int it, void *stream) {
  B200MF_REQUIRE(x && p && r && scratch, "null argument");
  B200MF_DISPATCH_NUMBER(number, (cg_pre_kernel<T><<<vec_grid(n), kVecThreads, 0, (cudaStream_t)stream>>>(
      (T *)x, (T *)p, (const T *)r, (const T *)d, n, scratch, it)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}
int b200mf_cg_final(int number, void *x, const void *p, uint64_t n, const double *scratch, int it,
                    void *stream) {
  B200MF_REQUIRE(x && p && scratch, "null argument");
  B200MF_DISPATCH_NUMBER(number, (cg_final_kernel<T><<<vec_grid(n), kVecThreads, 0, (cudaStream_t)stream>>>(
      (T *)x, (const T *)p, n, scratch, it)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_cg_solve_host(const b200mf_setup *h, const b200mf_operator *op,
                         const b200mf_solver_desc *solver, void *x_host, const void *b_host,
                         b200mf_solver_result *result) {
  B200MF_REQUIRE(h && op && solver && x_host && b_host, "null argument");
  Setup &s = const_cast<Setup &>(h->impl);
  const size_t elems = s.n_owned + s.n_ghost, ns = number_size(s.number);
  for (int i = 0; i < 2; ++i)
    if (!s.d_stage[i]) B200MF_CUDA_CHECK(cudaMalloc(&s.d_stage[i], std::max<size_t>(elems, 1) * ns));
  B200MF_CUDA_CHECK(cudaMemsetAsync(s.d_stage[0], 0, elems * ns, 0));
  B200MF_CUDA_CHECK(cudaMemsetAsync(s.d_stage[1], 0, elems * ns, 0));
  B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_stage[0], x_host, s.n_owned * ns, cudaMemcpyHostToDevice, 0));
  B200MF_CUDA_CHECK(cudaMemcpyAsync(s.d_stage[1], b_host, s.n_owned * ns, cudaMemcpyHostToDevice, 0));
  int rc = b200mf_cg_solve(h, op, solver, s.d_stage[0], s.d_stage[1], result, nullptr);
  cudaMemcpyAsync(x_host, s.d_stage[0], s.n_owned * ns, cudaMemcpyDeviceToHost, 0);
  cudaStreamSynchronize(0);
  return rc;
}

} // extern "C"
