// FP64 / FP32 FMA pipe peak microbenchmark (roofline denominator for the compute-bound
// affine high-degree cases; SURVEY.md section 7 "measure it with a FP64 FMA microbenchmark").
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T *out, T a, T b, int iters) {
  T acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = T(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T>
double run(const char *name) {
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  constexpr int ILP = 8;
  T *out;
  cudaMalloc(&out, sizeof(T) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fma_kernel<T, ILP><<<blocks, threads>>>(out, T(1.0000001), T(1e-9), 100);
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, T(1.0000001), T(1e-9), iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * blocks * threads * (double)iters * ILP / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  printf("{\"pipe\": \"%s\", \"tflops\": %.2f}\n", name, best);
  cudaFree(out);
  return best;
}
int main() {
  run<double>("fp64_fma");
  run<float>("fp32_fma");
  return 0;
}
