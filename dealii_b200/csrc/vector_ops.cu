// BLAS-1 entry points (LinearAlgebra::distributed::Vector<Number, MemorySpace::Default>,
// lac/vector_operations_internal.h:2140-2660).  Grid-stride kernels, 148-SM sized grids.
#include <cmath>

#include "vector_ops.cuh"

namespace b200mf {

template <typename Number>
__global__ void set_kernel(Number *x, Number v, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    x[i] = v;
}
template <typename Number>
__global__ void sadd_kernel(Number *y, Number s, Number a, const Number *x, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = s * y[i] + a * x[i];
}
template <typename Number>
__global__ void scale_by_kernel(Number *y, const Number *d, const Number *x, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = d[i] * x[i];
}
template <typename Number>
__global__ void dot_kernel(const Number *x, const Number *y, uint64_t n, double *out) {
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    acc += double(x[i]) * double(y[i]);
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// y = a x  (Vector::equ, Vectorization_equ_au :2223); w = a x + b v folded in as equ(a, x, b, v)
template <typename Number>
__global__ void equ_kernel(Number *y, Number a, const Number *x, Number b, const Number *v, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = v ? a * x[i] + b * v[i] : a * x[i];
}
// y = s y + a x + b w  (Vector::sadd(s, a, V, b, W), Vectorization_sadd_xavbw :2312)
template <typename Number>
__global__ void sadd_xavbw_kernel(Number *y, Number s, Number a, const Number *x, Number b, const Number *w,
                                  uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = s * y[i] + a * x[i] + b * w[i];
}
// y *= a  (Vector::operator*=, Vectorization_multiply_factor :2188);  y = y .* d (Vector::scale :2338)
template <typename Number>
__global__ void scale_kernel(Number *y, Number a, const Number *d, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = d ? y[i] * d[i] : y[i] * a;
}
// y += a x; *out += y . w   (Vector::add_and_dot, AddAndDot :2590)
template <typename Number>
__global__ void add_and_dot_kernel(Number *y, Number a, const Number *x, const Number *w, uint64_t n, double *out) {
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Number yi = y[i] + a * x[i];
    y[i] = yi;
    acc += double(yi) * double(w[i]);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
// *out += sum |x_i|  (l1_norm, Norm1 :2517) / max |x_i| (linfty_norm)
template <typename Number>
__global__ void norm1_kernel(const Number *x, uint64_t n, double *out) {
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    acc += fabs(double(x[i]));
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
template <typename Number>
__global__ void norm_inf_kernel(const Number *x, uint64_t n, unsigned long long *out) {
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    acc = fmax(acc, fabs(double(x[i])));
  // non-negative doubles order like their bit patterns
  atomicMax(out, (unsigned long long)__double_as_longlong(acc));
}

// ghost exchange pack / unpack (base/partitioner.templates.h:119-137 and :604-671)
template <typename Number>
__global__ void pack_kernel(Number *buf, const Number *vec, const uint32_t *idx, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    buf[i] = vec[idx[i]];
}
template <typename Number>
__global__ void unpack_add_kernel(Number *vec, const Number *buf, const uint32_t *idx, uint64_t n) {
  // an owned dof can be imported from several neighbours: atomics
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(vec + idx[i], buf[i]);
}

// runs a reduction kernel into a zeroed device double and returns it through a host pointer
template <typename Launch>
int reduce_to_host(double *result_host, cudaStream_t st, Launch launch) {
  double *d_out = nullptr;
  B200MF_CUDA_CHECK(cudaMalloc((void **)&d_out, sizeof(double)));
  B200MF_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(double), st));
  launch(d_out);
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  B200MF_CUDA_CHECK(cudaMemcpyAsync(result_host, d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
  B200MF_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(d_out);
  return B200MF_OK;
}

template <typename Number>
int dot_impl(const void *x, const void *y, uint64_t n, double *result_host, cudaStream_t st) {
  return reduce_to_host(result_host, st, [&](double *d_out) {
    dot_kernel<Number><<<vec_grid(n), kVecThreads, 0, st>>>((const Number *)x, (const Number *)y, n, d_out);
  });
}

} // namespace b200mf

using namespace b200mf;

#define DISPATCH(number, call64, call32)                 \
  do {                                                   \
    if ((number) == B200MF_F64) { call64; }              \
    else if ((number) == B200MF_F32) { call32; }         \
    else { set_error("bad number type"); return B200MF_ERR_INVALID; } \
  } while (0)

extern "C" {

int b200mf_vec_set(int number, void *x, double value, uint64_t n, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number, (set_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)x, value, n)),
           (set_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)x, (float)value, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_sadd(int number, void *y, double s, double a, const void *x, uint64_t n,
                    void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (sadd_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)y, s, a, (const double *)x, n)),
           (sadd_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)y, (float)s, (float)a, (const float *)x, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_axpy(int number, void *y, double a, const void *x, uint64_t n, void *stream) {
  return b200mf_vec_sadd(number, y, 1.0, a, x, n, stream);
}

int b200mf_vec_scale_by(int number, void *y, const void *d, const void *x, uint64_t n,
                        void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (scale_by_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)y, (const double *)d, (const double *)x, n)),
           (scale_by_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)y, (const float *)d, (const float *)x, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_dot(int number, const void *x, const void *y, uint64_t n, double *result_host,
                   void *stream) {
  B200MF_REQUIRE(result_host, "null result pointer");
  if (number == B200MF_F64) return dot_impl<double>(x, y, n, result_host, (cudaStream_t)stream);
  if (number == B200MF_F32) return dot_impl<float>(x, y, n, result_host, (cudaStream_t)stream);
  set_error("bad number type");
  return B200MF_ERR_INVALID;
}

int b200mf_vec_dot_device(int number, const void *x, const void *y, uint64_t n, double *result_device,
                          void *stream) {
  B200MF_REQUIRE(result_device, "null result pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (dot_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((const double *)x, (const double *)y, n, result_device)),
           (dot_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((const float *)x, (const float *)y, n, result_device)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_norm_sqr(int number, const void *x, uint64_t n, double *result_host, void *stream) {
  return b200mf_vec_dot(number, x, x, n, result_host, stream);
}

int b200mf_vec_norm_2(int number, const void *x, uint64_t n, double *result_host, void *stream) {
  int rc = b200mf_vec_dot(number, x, x, n, result_host, stream);
  if (rc == B200MF_OK) *result_host = std::sqrt(*result_host);
  return rc;
}

int b200mf_vec_norm_1(int number, const void *x, uint64_t n, double *result_host, void *stream) {
  B200MF_REQUIRE(result_host, "null result pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (number == B200MF_F64)
    return reduce_to_host(result_host, st, [&](double *o) { norm1_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((const double *)x, n, o); });
  if (number == B200MF_F32)
    return reduce_to_host(result_host, st, [&](double *o) { norm1_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((const float *)x, n, o); });
  set_error("bad number type");
  return B200MF_ERR_INVALID;
}

int b200mf_vec_norm_linfty(int number, const void *x, uint64_t n, double *result_host, void *stream) {
  B200MF_REQUIRE(result_host, "null result pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (number != B200MF_F64 && number != B200MF_F32) { set_error("bad number type"); return B200MF_ERR_INVALID; }
  return reduce_to_host(result_host, st, [&](double *o) {
    unsigned long long *u = reinterpret_cast<unsigned long long *>(o);
    if (number == B200MF_F64) norm_inf_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((const double *)x, n, u);
    else                      norm_inf_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((const float *)x, n, u);
  });
}

int b200mf_vec_equ(int number, void *y, double a, const void *x, double b, const void *v, uint64_t n,
                   void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (equ_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)y, a, (const double *)x, b, (const double *)v, n)),
           (equ_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)y, (float)a, (const float *)x, (float)b, (const float *)v, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_sadd_xavbw(int number, void *y, double s, double a, const void *x, double b, const void *w,
                          uint64_t n, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (sadd_xavbw_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)y, s, a, (const double *)x, b, (const double *)w, n)),
           (sadd_xavbw_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)y, (float)s, (float)a, (const float *)x, (float)b, (const float *)w, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_scale(int number, void *y, double a, const void *d, uint64_t n, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (scale_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)y, a, (const double *)d, n)),
           (scale_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)y, (float)a, (const float *)d, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_vec_add_and_dot(int number, void *y, double a, const void *x, const void *w, uint64_t n,
                           double *result_host, void *stream) {
  B200MF_REQUIRE(result_host, "null result pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (number == B200MF_F64)
    return reduce_to_host(result_host, st, [&](double *o) {
      add_and_dot_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)y, a, (const double *)x, (const double *)w, n, o); });
  if (number == B200MF_F32)
    return reduce_to_host(result_host, st, [&](double *o) {
      add_and_dot_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)y, (float)a, (const float *)x, (const float *)w, n, o); });
  set_error("bad number type");
  return B200MF_ERR_INVALID;
}

int b200mf_ghost_pack(int number, void *buf, const void *vec, const uint32_t *idx, uint64_t n,
                      void *stream) {
  if (n == 0) return B200MF_OK;
  B200MF_REQUIRE(buf && vec && idx, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (pack_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)buf, (const double *)vec, idx, n)),
           (pack_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)buf, (const float *)vec, idx, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

int b200mf_ghost_unpack_add(int number, void *vec, const void *buf, const uint32_t *idx, uint64_t n,
                            void *stream) {
  if (n == 0) return B200MF_OK;
  B200MF_REQUIRE(buf && vec && idx, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(number,
           (unpack_add_kernel<double><<<vec_grid(n), kVecThreads, 0, st>>>((double *)vec, (const double *)buf, idx, n)),
           (unpack_add_kernel<float><<<vec_grid(n), kVecThreads, 0, st>>>((float *)vec, (const float *)buf, idx, n)));
  count_launch();
  B200MF_CUDA_CHECK(cudaGetLastError());
  return B200MF_OK;
}

} // extern "C"
