// cmaes.cu — CMA-ES (optimizers/cma_es.py:7-227 of the reference) behind the begin / iter_local /
// iter_merge / finish protocol of optimizers.cu.
//
//   iter_local : BD = B diag(d);  z ~ N(0,I) (Philox, keyed on the GLOBAL row);
//                x = clip(m + sigma * (z BD))   -- one fused tiled SGEMM, z is never stored (cma_es.py:139-149);
//                penalty; rollout; reward summed over agents (:157-158); local top-E -> partial message
//   iter_merge : exact global top-E of the gathered candidates (reward desc, global row asc) = the first E
//                rows of the reference's full argsort (:159); the weights are zero beyond E (:62-68), so
//                every reduction over the population collapses to E rows (the reference materialises
//                [P,N,N] here: 18 GB at P=50 000, N=300).  Mean / evolution paths / step size (:161-177),
//                rank-mu + rank-one covariance update and symmetrisation (:180-190), eigendecomposition
//                (the SVD of a symmetric PSD matrix, :195-198) with cuSOLVER syevd, loaded lazily.
// sigma is a per-coordinate VECTOR as in the reference (:97); D is kept as its diagonal.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include "common.cuh"
#include "device_fns.cuh"
#include "refit.cuh"
#include "opt_state.cuh"

namespace bbmpc {
namespace {

// ---- constants (cma_consts slots)
enum { K_MUEFF = 0, K_CSIGMA, K_DSIGMA, K_CC, K_C1, K_CMU, K_EN, K_N };

// ---- lazily bound cuSOLVER (only CMA-ES needs it; libbbmpc.so itself does not link it)
// The dlopen'ed function table is process-wide (bound once); the cuSOLVER handle, its workspace and devInfo belong to the
// optimizer handle (bbmpc_opt::eig_*): one per device / stream user, freed with the handle.
struct Solver {
  void* lib = nullptr;
  bool ok = false; std::string err;
  int (*create)(void**) = nullptr;
  int (*destroy)(void*) = nullptr;
  int (*set_stream)(void*, cudaStream_t) = nullptr;
  int (*buffer_size)(void*, int, int, int, const float*, int, const float*, int*) = nullptr;
  int (*syevd)(void*, int, int, int, float*, int, float*, float*, int, int*) = nullptr;
};
Solver g_solver;
std::once_flag g_solver_once;

int solver_init(bbmpc_ctx* ctx) {
  Solver& s = g_solver;
  std::call_once(g_solver_once, [&s] {
    const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so.12"};
    for (const char* n : names) { s.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (s.lib) break; }
    if (!s.lib) { s.err = std::string("CMA-ES needs libcusolver (dlopen failed: ") + dlerror() + ")"; return; }
    s.create = reinterpret_cast<decltype(s.create)>(dlsym(s.lib, "cusolverDnCreate"));
    s.destroy = reinterpret_cast<decltype(s.destroy)>(dlsym(s.lib, "cusolverDnDestroy"));
    s.set_stream = reinterpret_cast<decltype(s.set_stream)>(dlsym(s.lib, "cusolverDnSetStream"));
    s.buffer_size = reinterpret_cast<decltype(s.buffer_size)>(dlsym(s.lib, "cusolverDnSsyevd_bufferSize"));
    s.syevd = reinterpret_cast<decltype(s.syevd)>(dlsym(s.lib, "cusolverDnSsyevd"));
    if (!s.create || !s.destroy || !s.set_stream || !s.buffer_size || !s.syevd) { s.err = "libcusolver lacks cusolverDnSsyevd"; return; }
    s.ok = true;
  });
  if (!s.ok) return fail(ctx, BBMPC_ECUDA, "%s", s.err.c_str());
  return BBMPC_OK;
}

// ---- kernels
__global__ void bd_kernel(const float* __restrict__ B, const float* __restrict__ d, float* __restrict__ BD, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * N) BD[i] = __fmul_rn(B[i], d[i % N]);   // B @ diag(d): column j scaled by d[j]
}

// x[p, n] = clip(m[n] + sigma[n] * sum_k z[p,k] BD[k,n]);  excess_sq[p,n] = (unclipped - clipped)^2.
// 64x64 output tile, K step 16, 256 threads x (4x4).  z is drawn in the loader: element (row, k) is word
// k&3 of Philox block k>>2 of the GLOBAL row (same counter layout as every other sampler).
constexpr int CT = 64, CK = 16;
__global__ void __launch_bounds__(256) cmaes_sample_kernel(const float* __restrict__ BD, const float* __restrict__ m,
                                                           const float* __restrict__ sigma, const float* __restrict__ lb,
                                                           const float* __restrict__ ub, float* __restrict__ x,
                                                           float* __restrict__ excess_sq, float* __restrict__ z_trace,
                                                           const float* __restrict__ z_inject, int P_local, int p0, int N, int dU, uint64_t seed,
                                                           const uint32_t* act_ctr, uint32_t iter) {
  __shared__ float zs[CK][CT + 1];
  __shared__ float bs[CK][CT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * CT, col0 = blockIdx.y * CT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < N; k0 += CK) {
    {  // z tile: thread -> (row = tid / 4, 4 consecutive k)
      const int r = tid >> 2, kq = (tid & 3) * 4, row = row0 + r;
      Philox4 w{0u, 0u, 0u, 0u};
      if (row < P_local && k0 + kq < N) w = draw_block(seed, *act_ctr, STREAM_SAMPLES, iter, static_cast<uint32_t>(p0 + row), (k0 + kq) >> 2);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        const float zv = (row < P_local && k < N) ? (z_inject ? z_inject[static_cast<size_t>(p0 + row) * N + k] : std_normal(ws[j])) : 0.0f;
        zs[kq + j][r] = zv;
        if (z_trace && blockIdx.y == 0 && row < P_local && k < N) z_trace[static_cast<size_t>(row) * N + k] = zv;
      }
    }
#pragma unroll
    for (int i = tid; i < CK * CT; i += 256) {
      const int k = i / CT, c = i % CT;
      bs[k][c] = (k0 + k < N && col0 + c < N) ? BD[static_cast<size_t>(k0 + k) * N + col0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = zs[k][ty * 4 + i]; b[i] = bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    if (row >= P_local) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = col0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = __fadd_rn(m[n], __fmul_rn(sigma[n], acc[i][j]));
      const int u = n % dU;
      const float f = fminf(fmaxf(v, lb[u]), ub[u]);
      const float d = __fsub_rn(v, f);
      x[static_cast<size_t>(row) * N + n] = f;
      excess_sq[static_cast<size_t>(row) * N + n] = __fmul_rn(d, d);
    }
  }
}

// rewards[p] = sum_a returns[p, a]   (cma_es.py:158; penalty already subtracted per (p, a))
__global__ void reward_sum_kernel(const float* __restrict__ returns, float* __restrict__ out, int P_local, int A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P_local) return;
  float s = 0.0f;
  for (int a = 0; a < A; ++a) s = __fadd_rn(s, returns[static_cast<size_t>(p) * A + a]);
  out[p] = s;
}

struct UpdArgs {
  const float* partials; int world; int64_t partial_stride; int E, N;
  const float* w;       // [E] recombination weights
  float *m, *sigma, *C, *B, *d, *ps, *pc;
  float* yu;            // [E, N] scratch: x_diff / sigma of the elites, in rank order
  float* tmp;           // [2N] scratch
  float c_sigma, d_sigma, cc, c1, c_mu, mu_eff, h_sigma, exp_norm;
};

// One CTA: rank the world*E candidates, then mean / paths / step size (cma_es.py:161-177).
__global__ void __launch_bounds__(SEL_THREADS) cmaes_paths_kernel(const UpdArgs a) {
  __shared__ uint32_t keys[SEL_MAX_K];
  __shared__ int gp[SEL_MAX_K];
  __shared__ int src[SEL_MAX_K];
  __shared__ float red[33];
  const int tid = threadIdx.x, E = a.E, N = a.N, rec = 2 + N, n = a.world * E;
  for (int i = tid; i < n; i += SEL_THREADS) {
    const float* r = a.partials + (i / E) * a.partial_stride + static_cast<size_t>(i % E) * rec;
    keys[i] = f2key(r[0]); gp[i] = __float_as_int(r[1]);
  }
  __syncthreads();
  for (int i = tid; i < n; i += SEL_THREADS) {
    int pos = 0;
    for (int j = 0; j < n; ++j) pos += before(keys[j], gp[j], keys[i], gp[i]) ? 1 : 0;
    if (pos < SEL_MAX_K) src[pos] = i;
  }
  __syncthreads();
  // yu[i, :] = (x_i - m) / sigma ; x_mean = sum_i w_i (x_i - m)
  for (int c = tid; c < N; c += SEL_THREADS) {
    const float mc = a.m[c], sc = a.sigma[c];
    float xm = 0.0f;
    for (int i = 0; i < E; ++i) {
      const int s = src[i];
      const float xv = a.partials[(s / E) * a.partial_stride + static_cast<size_t>(s % E) * rec + 2 + c];
      const float diff = __fsub_rn(xv, mc);
      a.yu[static_cast<size_t>(i) * N + c] = __fdiv_rn(diff, sc);
      xm = __fadd_rn(xm, __fmul_rn(diff, a.w[i]));
    }
    a.m[c] = __fadd_rn(mc, xm);
    a.tmp[c] = __fdiv_rn(xm, sc);          // y_mean
  }
  __syncthreads();
  // C^{-1/2} y_mean = B D^{-1} B^T y_mean
  for (int j = tid; j < N; j += SEL_THREADS) {
    float s = 0.0f;
    for (int i = 0; i < N; ++i) s = fmaf(a.B[static_cast<size_t>(i) * N + j], a.tmp[i], s);
    a.tmp[N + j] = __fmul_rn(s, __fdiv_rn(1.0f, a.d[j]));
  }
  __syncthreads();
  float sq = 0.0f;
  const float k_ps = sqrtf(a.c_sigma * (2.0f - a.c_sigma) * a.mu_eff);
  const float k_pc = a.h_sigma * sqrtf(a.cc * (2.0f - a.cc) * a.mu_eff);
  for (int i = tid; i < N; i += SEL_THREADS) {
    float s = 0.0f;
    for (int j = 0; j < N; ++j) s = fmaf(a.B[static_cast<size_t>(i) * N + j], a.tmp[N + j], s);
    const float ps = __fadd_rn(__fmul_rn(1.0f - a.c_sigma, a.ps[i]), __fmul_rn(k_ps, s));
    a.ps[i] = ps;
    sq = fmaf(ps, ps, sq);
    a.pc[i] = __fadd_rn(__fmul_rn(1.0f - a.cc, a.pc[i]), __fmul_rn(k_pc, a.tmp[i]));
  }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  if (tid < 32) {
    float v = red[tid];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tid == 0) red[32] = expf((a.c_sigma / a.d_sigma) * (sqrtf(v) / a.exp_norm - 1.0f));
  }
  __syncthreads();
  const float g = red[32];
  for (int i = tid; i < N; i += SEL_THREADS) a.sigma[i] = __fmul_rn(a.sigma[i], g);
}

// C' = (1 - c1 - c_mu) C + c1 pc pc^T + c_mu sum_i w_i yu_i yu_i^T, computed on the upper triangle and
// mirrored (cma_es.py:183-190).
__global__ void cmaes_cov_kernel(const UpdArgs a) {
  const int r = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = a.N;
  if (r >= N || c >= N || r > c) return;
  float ys = 0.0f;
  for (int i = 0; i < a.E; ++i) ys = fmaf(__fmul_rn(a.yu[static_cast<size_t>(i) * N + r], a.yu[static_cast<size_t>(i) * N + c]), a.w[i], ys);
  const float v = __fadd_rn(__fadd_rn(__fmul_rn(1.0f - a.c1 - a.c_mu, a.C[static_cast<size_t>(r) * N + c]),
                                      __fmul_rn(a.c1, __fmul_rn(a.pc[r], a.pc[c]))),
                            __fmul_rn(a.c_mu, ys));
  a.C[static_cast<size_t>(r) * N + c] = v;
  a.C[static_cast<size_t>(c) * N + r] = v;
}

// syevd output (eigenvalues ascending in wv, eigenvectors in the COLUMNS of a column-major matrix, i.e.
// eigenvector j is row j of the row-major view) -> B[i][j] = component i of the j-th LARGEST eigenvector,
// d[j] = sqrt(max(eigenvalue_j, 0))   (tf.linalg.svd returns singular values in descending order)
__global__ void eig_finish_kernel(const float* __restrict__ V, const float* __restrict__ wv, float* __restrict__ B,
                                  float* __restrict__ d, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  const int r = i / N, j = i % N, jj = N - 1 - j;
  B[i] = V[static_cast<size_t>(jj) * N + r];
  if (r == 0) d[j] = sqrtf(fmaxf(wv[jj], 0.0f));
}

__global__ void set_identity_kernel(float* M, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * N) M[i] = (i / N == i % N) ? 1.0f : 0.0f;
}
__global__ void cmaes_init_vec_kernel(float* m, float* sigma, const float* lb, const float* ub, int N, int dU) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float l = lb[i % dU], u = ub[i % dU];
  m[i] = __fdiv_rn(__fadd_rn(l, u), 2.0f);
  const float dd = __fsub_rn(l, u);
  sigma[i] = sqrtf(__fdiv_rn(__fmul_rn(dd, dd), 16.0f));
}

template <typename T>
int dalloc(bbmpc_opt* o, T** p, size_t n) {
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
  if (e != cudaSuccess) return fail(o->ctx, BBMPC_ENOMEM, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
  o->owned.push_back(*p);
  return BBMPC_OK;
}
inline int grid_for(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

}  // namespace

int cmaes_create(bbmpc_opt* o) {
  bbmpc_ctx* ctx = o->ctx;
  const bbmpc_opt_config& c = o->cfg;
  const int N = o->AHU, E = c.num_elite, P = c.population_size;
  if (E < 1 || E > P || E > SEL_MAX_K) return fail(ctx, BBMPC_EINVAL, "num_elite=%d must be in [1, min(P, %d)]", E, SEL_MAX_K);
  if (N > 2048) return fail(ctx, BBMPC_EINVAL, "CMA-ES solution size %d > 2048 unsupported", N);
  if (int rc = solver_init(ctx)) return rc;
  // constants, evaluated in fp32 in the reference's order (cma_es.py:62-126)
  std::vector<float> w(E);
  float wsum = 0.0f;
  for (int i = 0; i < E; ++i) { w[i] = logf(static_cast<float>(E) + 0.5f) - logf(static_cast<float>(i + 1)); wsum += w[i]; }
  float s1 = 0.0f, s2 = 0.0f;
  for (int i = 0; i < E; ++i) { w[i] = w[i] / wsum; s1 += w[i]; s2 += w[i] * w[i]; }
  const float nf = static_cast<float>(N);
  const float mu_eff = s1 * s1 / s2;
  const float c_sigma = (mu_eff + 2.0f) / (nf + mu_eff + 5.0f);
  const float d_sigma = 1.0f + 2.0f * fmaxf(0.0f, sqrtf((mu_eff - 1.0f) / (nf + 1.0f)) - 1.0f) + c_sigma;
  const float cc = (4.0f + mu_eff / nf) / (nf + 4.0f + 2.0f * mu_eff / nf);
  const float c1 = c.alpha_cov / ((nf + 1.3f) * (nf + 1.3f) + mu_eff);
  const float c_mu2 = c.alpha_cov * (mu_eff - 2.0f + 1.0f / mu_eff) / ((nf + 2.0f) * (nf + 2.0f) + c.alpha_cov * mu_eff / 2.0f);
  const float c_mu = fminf(1.0f - c1, c_mu2);
  const float en = sqrtf(nf * (1.0f - 1.0f / (4.0f * nf) + 1.0f / (21.0f * nf * nf)));
  o->cma_consts[K_MUEFF] = mu_eff; o->cma_consts[K_CSIGMA] = c_sigma; o->cma_consts[K_DSIGMA] = d_sigma;
  o->cma_consts[K_CC] = cc; o->cma_consts[K_C1] = c1; o->cma_consts[K_CMU] = c_mu; o->cma_consts[K_EN] = en; o->cma_consts[K_N] = N;
  int rc = BBMPC_OK;
  auto A_ = [&](int r) { if (rc == BBMPC_OK) rc = r; };
  const size_t NN = static_cast<size_t>(N) * N;
  A_(dalloc(o, &o->d_sigma, N)); A_(dalloc(o, &o->d_C, NN)); A_(dalloc(o, &o->d_B, NN)); A_(dalloc(o, &o->d_D, N));
  A_(dalloc(o, &o->d_ps, N)); A_(dalloc(o, &o->d_pc, N)); A_(dalloc(o, &o->d_BD, NN)); A_(dalloc(o, &o->d_cma_w, E));
  A_(dalloc(o, &o->d_z, NN + 3 * static_cast<size_t>(N) + static_cast<size_t>(E) * N));   // eig matrix | eigenvalues | tmp[2N] | yu[E,N]
  if (rc != BBMPC_OK) return rc;
  o->d_m = o->d_mean;   // the mean IS the loop variable / solution [A,H,dU] flattened (cma_es.py:95,211)
  cudaMemcpy(o->d_cma_w, w.data(), E * sizeof(float), cudaMemcpyHostToDevice);
  set_identity_kernel<<<grid_for(NN, 256), 256>>>(o->d_C, N);
  set_identity_kernel<<<grid_for(NN, 256), 256>>>(o->d_B, N);
  {
    std::vector<float> ones(N, 1.0f);
    cudaMemcpy(o->d_D, ones.data(), N * sizeof(float), cudaMemcpyHostToDevice);
  }
  cudaMemset(o->d_ps, 0, N * sizeof(float)); cudaMemset(o->d_pc, 0, N * sizeof(float));
  cmaes_init_vec_kernel<<<grid_for(N, 256), 256>>>(o->d_m, o->d_sigma, o->d_lb, o->d_ub, N, c.dU);
  ctx->launches += 3;
  // eigensolver workspace
  Solver& s = g_solver;
  if (s.create(&o->eig_handle) != 0) { o->eig_handle = nullptr; return fail(ctx, BBMPC_ECUDA, "cusolverDnCreate failed"); }
  int lwork = 0;
  if (s.buffer_size(o->eig_handle, 1 /*CUSOLVER_EIG_MODE_VECTOR*/, 0 /*CUBLAS_FILL_MODE_LOWER*/, N, o->d_z, N, o->d_z + NN, &lwork) != 0)
    return fail(ctx, BBMPC_ECUDA, "cusolverDnSsyevd_bufferSize failed");
  if (cudaMalloc(&o->eig_work, static_cast<size_t>(lwork) * sizeof(float)) != cudaSuccess) return fail(ctx, BBMPC_ENOMEM, "eigensolver workspace");
  o->eig_lwork = lwork;
  if (cudaMalloc(&o->eig_info, sizeof(int)) != cudaSuccess) return fail(ctx, BBMPC_ENOMEM, "eigensolver info");
  return BBMPC_OK;
}

void cmaes_destroy(bbmpc_opt* o) {
  if (o->eig_handle) g_solver.destroy(o->eig_handle);
  cudaFree(o->eig_work); cudaFree(o->eig_info);
  o->eig_handle = nullptr; o->eig_work = nullptr; o->eig_info = nullptr; o->eig_lwork = 0;
}

int cmaes_set_shard(bbmpc_opt* o) {
  // excess^2 scratch [P_local*A, HU] and the per-population reward sums live in d_work / d_penalty-sized buffers
  const size_t rows = static_cast<size_t>(o->P_local) * o->cfg.num_agents;
  int rc = dalloc(o, &o->d_work, rows * o->HU + static_cast<size_t>(o->P_local));
  return rc;
}

int cmaes_reset(bbmpc_opt* o, cudaStream_t st) {   // cma_es.py:215-227: m and sigma only
  cmaes_init_vec_kernel<<<grid_for(o->AHU, 256), 256, 0, st>>>(o->d_m, o->d_sigma, o->d_lb, o->d_ub, o->AHU, o->cfg.dU);
  BB_LAUNCH_CHECK(o->ctx);
  return BBMPC_OK;
}

int cmaes_iter_local(bbmpc_opt* o, int iter, float* partial, cudaStream_t st) {
  bbmpc_ctx* ctx = o->ctx;
  const bbmpc_opt_config& c = o->cfg;
  const int N = o->AHU, A = c.num_agents, H = c.planning_horizon, E = c.num_elite;
  const size_t NN = static_cast<size_t>(N) * N;
  if (o->P_local > 0) {
    bd_kernel<<<grid_for(NN, 256), 256, 0, st>>>(o->d_B, o->d_D, o->d_BD, N); BB_LAUNCH_CHECK(ctx);
    const int64_t per_iter = static_cast<int64_t>(o->P_local) * N;
    float* z_trace = (o->trace && (static_cast<int64_t>(iter) + 1) * per_iter <= o->trace_floats) ? o->trace + iter * per_iter : nullptr;
    dim3 grid(grid_for(o->P_local, CT), grid_for(N, CT));
    const float* z_inject = nullptr;
    if (o->inject) {   // one block of [P, N] standard normals per iteration since bbmpc_opt_set_draw_injection
      const int64_t blk = static_cast<int64_t>(c.population_size) * N;
      if ((o->inject_iter + 1) * blk > o->inject_floats) return fail(ctx, BBMPC_EINVAL, "draw injection buffer exhausted");
      z_inject = o->inject + o->inject_iter * blk;
      ++o->inject_iter;
    }
    cmaes_sample_kernel<<<grid, 256, 0, st>>>(o->d_BD, o->d_m, o->d_sigma, o->d_lb, o->d_ub, o->d_samples, o->d_work, z_trace,
                                              z_inject, o->P_local, o->p0, N, c.dU, ctx->seed, o->d_act_ctr, static_cast<uint32_t>(iter));
    BB_LAUNCH_CHECK(ctx);
    const int64_t rows = static_cast<int64_t>(o->P_local) * A;
    launch_penalty(o->d_work, o->d_penalty, rows, o->HU, st); BB_LAUNCH_CHECK(ctx);
    if (int rc = rollout_dispatch(ctx, o->d_state, o->d_samples, o->d_returns, o->d_penalty, static_cast<int>(rows), A, H, st)) return rc;
    float* rew = o->d_work + static_cast<size_t>(rows) * o->HU;
    reward_sum_kernel<<<grid_for(o->P_local, 256), 256, 0, st>>>(o->d_returns, rew, o->P_local, A); BB_LAUNCH_CHECK(ctx);
    launch_topk_partial(rew, o->d_samples, partial, o->P_local, o->p0, 1, N, E, st);
  } else {
    launch_topk_partial(nullptr, o->d_samples, partial, 0, o->p0, 1, N, E, st);
  }
  BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

int cmaes_iter_merge(bbmpc_opt* o, int /*iter*/, const float* partials, int world, cudaStream_t st) {
  bbmpc_ctx* ctx = o->ctx;
  const bbmpc_opt_config& c = o->cfg;
  const int N = o->AHU, E = c.num_elite;
  const size_t NN = static_cast<size_t>(N) * N;
  if (world * E > SEL_MAX_K) return fail(ctx, BBMPC_EINVAL, "world*num_elite exceeds %d", SEL_MAX_K);
  float* eigm = o->d_z; float* eigw = o->d_z + NN; float* tmp = eigw + N; float* yu = tmp + 2 * N;
  UpdArgs a{partials, world, static_cast<int64_t>(E) * (2 + N), E, N, o->d_cma_w, o->d_m, o->d_sigma, o->d_C, o->d_B, o->d_D,
            o->d_ps, o->d_pc, yu, tmp,
            static_cast<float>(o->cma_consts[K_CSIGMA]), static_cast<float>(o->cma_consts[K_DSIGMA]), static_cast<float>(o->cma_consts[K_CC]),
            static_cast<float>(o->cma_consts[K_C1]), static_cast<float>(o->cma_consts[K_CMU]), static_cast<float>(o->cma_consts[K_MUEFF]),
            c.h_sigma, static_cast<float>(o->cma_consts[K_EN])};
  cmaes_paths_kernel<<<1, SEL_THREADS, 0, st>>>(a); BB_LAUNCH_CHECK(ctx);
  dim3 blk(16, 16), grd(grid_for(N, 16), grid_for(N, 16));
  cmaes_cov_kernel<<<grd, blk, 0, st>>>(a); BB_LAUNCH_CHECK(ctx);
  BB_CUDA(ctx, cudaMemcpyAsync(eigm, o->d_C, NN * sizeof(float), cudaMemcpyDeviceToDevice, st));
  Solver& s = g_solver;
  if (s.set_stream(o->eig_handle, st) != 0) return fail(ctx, BBMPC_ECUDA, "cusolverDnSetStream failed");
  if (s.syevd(o->eig_handle, 1, 0, N, eigm, N, eigw, o->eig_work, o->eig_lwork, o->eig_info) != 0) return fail(ctx, BBMPC_ECUDA, "cusolverDnSsyevd failed");
  ctx->launches++;
  if (getenv("BBMPC_DEBUG")) {   // BBMPC_DEBUG=1: devInfo of the eigensolve (costs a synchronisation per iteration)
    int info = 0;
    BB_CUDA(ctx, cudaMemcpyAsync(&info, o->eig_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    BB_CUDA(ctx, cudaStreamSynchronize(st));
    if (info != 0) return fail(ctx, BBMPC_ECUDA, "cusolverDnSsyevd: devInfo = %d (eigensolve of the CMA-ES covariance did not converge)", info);
  }
  eig_finish_kernel<<<grid_for(NN, 256), 256, 0, st>>>(eigm, eigw, o->d_B, o->d_D, N); BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

}  // namespace bbmpc
