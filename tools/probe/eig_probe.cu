// eig_probe.cu — symmetric 300x300 eigensolvers for the CMA-ES refit (cma_es.py:195-198 computes svd(C) once per iteration):
// cusolverDnSsyevd vs cusolverDnSsyevj from scratch and warm-started (C' = B^T C B is nearly diagonal when B are the
// previous iteration's eigenvectors and C moved by one rank-mu update).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probe/eig_probe.cu -o tools/probe/eig_probe -lcusolver -lcublas
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <functional>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cublas_v2.h>


#include <functional>
static float timeit(cudaStream_t st, int reps, const std::function<void()>& f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b, st); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
  const int N = 300;
  cusolverDnHandle_t h; cusolverDnCreate(&h);
  cublasHandle_t cb; cublasCreate(&cb);
  cudaStream_t st; cudaStreamCreate(&st); cusolverDnSetStream(h, st); cublasSetStream(cb, st);
  // C = I + small symmetric positive update (like CMA-ES after a few iterations), C2 = C + another small update
  std::vector<float> C(N * N), C2(N * N);
  srand(1);
  std::vector<float> Y(50 * N);
  auto rnd = [] { return (rand() / float(RAND_MAX) - 0.5f) * 2.0f; };
  for (int it = 0; it < 2; ++it) {
    for (auto& y : Y) y = rnd();
    std::vector<float>& M = it ? C2 : C;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) {
      float s = 0; for (int k = 0; k < 50; ++k) s += Y[k * N + i] * Y[k * N + j];
      M[i * N + j] = (it ? C[i * N + j] * 0.9f : (i == j ? 1.0f : 0.0f)) + 0.1f * s / 50;
    }
  }
  float *dC, *dC2, *dA, *dW, *dB, *dT, *dwork; int* info;
  cudaMalloc(&dC, N * N * 4); cudaMalloc(&dC2, N * N * 4); cudaMalloc(&dA, N * N * 4); cudaMalloc(&dB, N * N * 4); cudaMalloc(&dT, N * N * 4);
  cudaMalloc(&dW, N * 4); cudaMalloc(&info, 4);
  cudaMemcpy(dC, C.data(), N * N * 4, cudaMemcpyHostToDevice); cudaMemcpy(dC2, C2.data(), N * N * 4, cudaMemcpyHostToDevice);
  int lw1 = 0, lw2 = 0;
  cusolverDnSsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, dA, N, dW, &lw1);
  syevjInfo_t ji; cusolverDnCreateSyevjInfo(&ji);
  cusolverDnXsyevjSetTolerance(ji, 1e-7); cusolverDnXsyevjSetMaxSweeps(ji, 20); cusolverDnXsyevjSetSortEig(ji, 1);
  cusolverDnSsyevj_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, dA, N, dW, &lw2, ji);
  cudaMalloc(&dwork, (lw1 > lw2 ? lw1 : lw2) * 4);
  printf("workspace floats: syevd %d syevj %d\n", lw1, lw2);
  float t = timeit(st, 10, [&] { cudaMemcpyAsync(dA, dC2, N * N * 4, cudaMemcpyDeviceToDevice, st);
                                 cusolverDnSsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, dA, N, dW, dwork, lw1, info); });
  printf("syevd  from scratch: %.3f ms\n", t);
  for (double tol : {1e-7, 1e-5}) {
    cusolverDnXsyevjSetTolerance(ji, tol);
    t = timeit(st, 10, [&] { cudaMemcpyAsync(dA, dC2, N * N * 4, cudaMemcpyDeviceToDevice, st);
                             cusolverDnSsyevj(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, dA, N, dW, dwork, lw2, info, ji); });
    int sweeps = 0; double res = 0; cusolverDnXsyevjGetSweeps(h, ji, &sweeps); cusolverDnXsyevjGetResidual(h, ji, &res);
    printf("syevj  from scratch tol %.0e: %.3f ms (%d sweeps, residual %.2e)\n", tol, t, sweeps, res);
  }
  // warm start: B = eigenvectors of C (previous iteration); C' = B^T C2 B; syevj(C'); B2 = B V
  cusolverDnXsyevjSetTolerance(ji, 1e-7);
  cudaMemcpy(dB, dC, N * N * 4, cudaMemcpyDeviceToDevice);
  cusolverDnSsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, dB, N, dW, dwork, lw1, info);
  const float one = 1.0f, zero = 0.0f;
  for (double tol : {1e-7, 1e-5}) {
    cusolverDnXsyevjSetTolerance(ji, tol);
    t = timeit(st, 10, [&] {
      cublasSgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &one, dC2, N, dB, N, &zero, dT, N);      // T = C2 B
      cublasSgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, N, N, N, &one, dB, N, dT, N, &zero, dA, N);       // A = B^T T
      cusolverDnSsyevj(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, dA, N, dW, dwork, lw2, info, ji);
      cublasSgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &one, dB, N, dA, N, &zero, dT, N);       // B2 = B V
    });
    int sweeps = 0; double res = 0; cusolverDnXsyevjGetSweeps(h, ji, &sweeps); cusolverDnXsyevjGetResidual(h, ji, &res);
    printf("syevj  warm start   tol %.0e: %.3f ms incl. 3 SGEMMs (%d sweeps, residual %.2e)\n", tol, t, sweeps, res);
  }
  return 0;
}
