// optimizers.cu — sampling + refit kernels and host orchestration of the six reference
// optimizers (optimizers/{cem,pi2,random_search,pso,spsa,cma_es}.py) behind bbmpc_opt_*.
//
// Every optimizer iteration is "local" (sample this rank's slice of the population, roll it out,
// reduce it to a small partial message) followed by "merge" (combine the partial messages of all
// ranks, identical arithmetic on every rank, update the search distribution).  With one GPU the
// two halves run back to back; with several, the host all-gathers the partials in between.
// Draws are Philox streams keyed on (seed, act-call, stream, iteration, GLOBAL row): results do
// not depend on how the population is sharded.
#include <cmath>
#include <cstring>
#include <limits>
#include "common.cuh"
#include "device_fns.cuh"
#include "refit.cuh"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight tools, no-ops otherwise
#include "opt_state.cuh"

using namespace bbmpc;

namespace {

int opt_fail(bbmpc_opt* o, int code, const char* msg) { return fail(o->ctx, code, "%s", msg); }

template <typename T>
int dalloc(bbmpc_opt* o, T** p, size_t n) {
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
  if (e != cudaSuccess) return fail(o->ctx, BBMPC_ENOMEM, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
  o->owned.push_back(*p);
  return BBMPC_OK;
}

struct NvtxRange {   // sample / rollout / refit phases of an act()
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ============================================================================ kernels
// ---- fills / small vector ops
__global__ void fill_midpoint_kernel(float* out, const float* lb, const float* ub, int n, int dU) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fdiv_rn(__fadd_rn(lb[i % dU], ub[i % dU]), 2.0f);   // (lb + ub) / 2
}
__global__ void fill_var0_kernel(float* out, const float* lb, const float* ub, int n, int dU) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float d = __fsub_rn(lb[i % dU], ub[i % dU]); out[i] = __fdiv_rn(__fmul_rn(d, d), 16.0f); }
}
__global__ void fill_kernel(float* out, float v, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
// out[a,h,:] = in[a,min(h+1,H-1),:]   (pi2.py:92-93, spsa.py:114-115)
__global__ void act_ctr_bump_kernel(uint32_t* ctr) { *ctr += 1u; }   // last kernel of an act(): next call draws from fresh Philox counters
__global__ void shift_left_kernel(const float* in, float* out, int A, int H, int dU) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A * H * dU) return;
  const int u = i % dU, h = (i / dU) % H, a = i / (dU * H);
  const int hs = h + 1 < H ? h + 1 : H - 1;
  out[i] = in[(a * H + hs) * dU + u];
}
// action[a,:] = sol[a,0,:]  (+ exploration noise and clip, optimizer_base.py:82-90)
__global__ void first_action_kernel(const float* sol, float* action, const float* lb, const float* ub, int A,
                                    int H, int dU, int sol_stride_a, int add_noise, uint64_t seed, const uint32_t* act_ctr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A * dU) return;
  const int a = i / dU, u = i % dU;
  float v = sol[a * sol_stride_a + u];
  if (add_noise) {
    const Philox4 r = draw_block(seed, *act_ctr, STREAM_EXPLORE, 0, a, u);
    const float d = __fsub_rn(lb[u], ub[u]);
    const float var = __fmul_rn(__fdiv_rn(__fmul_rn(d, d), 16.0f), 0.05f);
    const float mean = __fdiv_rn(__fadd_rn(ub[u], lb[u]), 2.0f);
    const float noise = __fadd_rn(__fmul_rn(std_truncnorm(r.x), sqrtf(var)), mean);
    v = fminf(fmaxf(__fadd_rn(v, noise), lb[u]), ub[u]);
  }
  action[i] = v;
}

// ---- samplers.  One thread per 4 consecutive elements of a (population row, agent) sequence.
// Global Philox row = p_global * A + a.
struct SampleArgs {
  float* samples; float* penalty; const float* mean; const float* var; const float* lb; const float* ub;
  int P_local, p0, A, H, dU; uint64_t seed;
  const uint32_t* act_ctr;   // act() calls completed so far, on the device: a captured graph replays with a fresh Philox counter
  uint32_t iter;
  float ck;   // SPSA perturbation size
  float* raw_trace;  // optional: un-clipped draws of this iteration (PI2), for oracle injection
  const float* inject;  // optional: this iteration's standard variates [P, A, H*dU] by GLOBAL row, instead of Philox
};

// cem.py:81-94: cvar = min(((mean-lb)/2)^2, ((ub-mean)/2)^2, var); x = mean + sqrt(cvar)*tn
__global__ void cem_sample_kernel(const SampleArgs s) {
  const int HU = s.H * s.dU, nb = (HU + 3) / 4;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(s.P_local) * s.A * nb) return;
  const int blk = gid % nb; const int64_t pa = gid / nb;
  const int a = pa % s.A; const int p = pa / s.A;
  const Philox4 r = draw_block(s.seed, *s.act_ctr, STREAM_SAMPLES, s.iter, (s.p0 + p) * s.A + a, blk);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  for (int j = 0; j < 4; ++j) {
    const int e = 4 * blk + j;
    if (e >= HU) break;
    const int u = e % s.dU;
    const float m = s.mean[a * HU + e];
    const float lo = __fdiv_rn(__fsub_rn(m, s.lb[u]), 2.0f), hi = __fdiv_rn(__fsub_rn(s.ub[u], m), 2.0f);
    const float cvar = fminf(fminf(__fmul_rn(lo, lo), __fmul_rn(hi, hi)), s.var[a * HU + e]);
    const float z = s.inject ? s.inject[(static_cast<int64_t>(s.p0 + p) * s.A + a) * HU + e] : std_truncnorm(w[j]);
    s.samples[pa * HU + e] = __fadd_rn(__fmul_rn(z, sqrtf(cvar)), m);
  }
}
// pi2.py:65-76: x = mean + sqrt(var)*tn; clip; penalty[p,a] = ||x - clip(x)||^2 (summed by a
// second kernel in a fixed order).  Writes the CLIPPED sample and the raw excess^2 into `penalty`
// scratch laid out like samples?  No: excess is accumulated per (p,a) by penalty_kernel below.
__global__ void pi2_sample_kernel(const SampleArgs s, float* raw_excess_sq) {
  const int HU = s.H * s.dU, nb = (HU + 3) / 4;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(s.P_local) * s.A * nb) return;
  const int blk = gid % nb; const int64_t pa = gid / nb;
  const int a = pa % s.A; const int p = pa / s.A;
  const Philox4 r = draw_block(s.seed, *s.act_ctr, STREAM_SAMPLES, s.iter, (s.p0 + p) * s.A + a, blk);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  for (int j = 0; j < 4; ++j) {
    const int e = 4 * blk + j;
    if (e >= HU) break;
    const int u = e % s.dU;
    const float z = s.inject ? s.inject[(static_cast<int64_t>(s.p0 + p) * s.A + a) * HU + e] : std_truncnorm(w[j]);
    const float x = __fadd_rn(__fmul_rn(z, sqrtf(s.var[a * HU + e])), s.mean[a * HU + e]);
    const float xf = fminf(fmaxf(x, s.lb[u]), s.ub[u]);
    const float d = __fsub_rn(x, xf);
    s.samples[pa * HU + e] = xf;
    raw_excess_sq[pa * HU + e] = __fmul_rn(d, d);
    if (s.raw_trace) s.raw_trace[pa * HU + e] = x;
  }
}
// random_search.py:40-41: x = lb + (ub - lb) * U[0,1)
__global__ void rs_sample_kernel(const SampleArgs s) {
  const int HU = s.H * s.dU, nb = (HU + 3) / 4;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(s.P_local) * s.A * nb) return;
  const int blk = gid % nb; const int64_t pa = gid / nb;
  const int a = pa % s.A; const int p = pa / s.A;
  const Philox4 r = draw_block(s.seed, *s.act_ctr, STREAM_SAMPLES, s.iter, (s.p0 + p) * s.A + a, blk);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  for (int j = 0; j < 4; ++j) {
    const int e = 4 * blk + j;
    if (e >= HU) break;
    const int u = e % s.dU;
    const float u01 = s.inject ? s.inject[(static_cast<int64_t>(s.p0 + p) * s.A + a) * HU + e] : u01_halfopen(w[j]);
    s.samples[pa * HU + e] = __fadd_rn(__fmul_rn(u01, __fsub_rn(s.ub[u], s.lb[u])), s.lb[u]);
  }
}
// spsa.py:73-91: delta = +-1; theta+- = sol +- ck*delta; clip; excess^2 (plus rows first, then minus)
__global__ void spsa_sample_kernel(const SampleArgs s, float* raw_excess_sq) {
  const int HU = s.H * s.dU, nb = (HU + 3) / 4;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t half = static_cast<int64_t>(s.P_local) * s.A;
  if (gid >= half * nb) return;
  const int blk = gid % nb; const int64_t pa = gid / nb;
  const int a = pa % s.A; const int p = pa / s.A;
  const Philox4 r = draw_block(s.seed, *s.act_ctr, STREAM_SAMPLES, s.iter, (s.p0 + p) * s.A + a, blk);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  for (int j = 0; j < 4; ++j) {
    const int e = 4 * blk + j;
    if (e >= HU) break;
    const int u = e % s.dU;
    const float delta = s.inject ? s.inject[(static_cast<int64_t>(s.p0 + p) * s.A + a) * HU + e] : ((w[j] >> 31) ? 1.0f : -1.0f);
    const float step = __fmul_rn(s.ck, delta), m = s.mean[a * HU + e];
    const float xp = __fadd_rn(m, step), xm = __fsub_rn(m, step);
    const float xpf = fminf(fmaxf(xp, s.lb[u]), s.ub[u]), xmf = fminf(fmaxf(xm, s.lb[u]), s.ub[u]);
    s.samples[pa * HU + e] = xpf;
    s.samples[(half + pa) * HU + e] = xmf;
    const float dp = __fsub_rn(xp, xpf), dm = __fsub_rn(xm, xmf);
    raw_excess_sq[pa * HU + e] = __fmul_rn(dp, dp);
    raw_excess_sq[(half + pa) * HU + e] = __fmul_rn(dm, dm);
  }
}
// penalty[row] = (sqrt(sum_e excess_sq[row,e]))^2 — tf.norm(...,axis=2)**2 [TF]; one warp per row,
// fixed summation order.
__global__ void penalty_kernel(const float* excess_sq, float* penalty, int64_t rows, int HU) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.0f;
  for (int e = lane; e < HU; e += 32) acc += excess_sq[row * HU + e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) { const float n = sqrtf(acc); penalty[row] = __fmul_rn(n, n); }
}

// ---- CEM local top-E -> partial message.  One CTA per (agent, population slice): gridDim.y slices of the local rows,
// each with its own message (slice y at partial + y * slice_stride); the merge kernel treats slices like ranks.
// partial layout per agent: E records of (reward, global_p bits, seq[HU]).
__global__ void __launch_bounds__(SEL_THREADS) topk_partial_kernel(const float* returns, const float* samples,
                                                                   float* partial, int P_all, int p0_all, int A,
                                                                   int HU, int E, int use_cache, int64_t slice_stride) {
  const int n_slices = static_cast<int>(gridDim.y), y = static_cast<int>(blockIdx.y);
  const int pb = static_cast<int>(static_cast<int64_t>(P_all) * y / n_slices), pe = static_cast<int>(static_cast<int64_t>(P_all) * (y + 1) / n_slices);
  const int P_local = pe - pb, p0 = p0_all + pb;
  returns += static_cast<size_t>(pb) * A; samples += static_cast<size_t>(pb) * A * HU; partial += y * slice_stride;
  __shared__ int hist[256];
  __shared__ int misc[40];
  __shared__ uint32_t keys[SEL_MAX_K];
  __shared__ int idx[SEL_MAX_K];
  extern __shared__ uint32_t topk_kcache[];   // P_local words when the launch provides them, else none
  const int a = blockIdx.x;
  const SelectScratch sc{hist, misc, keys, idx};
  block_topk(returns + a, P_local, A, E, sc, use_cache ? topk_kcache : nullptr);
  const int rec = 2 + HU;
  float* out = partial + static_cast<size_t>(a) * E * rec;
  for (int i = threadIdx.x; i < E * rec; i += SEL_THREADS) {
    const int r = i / rec, c = i % rec;
    const int p = idx[r];
    const bool ok = p != 0x7FFFFFFF;
    float v;
    if (c == 0) v = ok ? key2f(keys[r]) : -INFINITY;
    else if (c == 1) v = __int_as_float(ok ? p0 + p : 0x7FFFFFFF);
    else v = ok ? samples[(static_cast<size_t>(p) * A + a) * HU + (c - 2)] : 0.0f;
    out[i] = v;
  }
}

// ---- CEM top-E + refit in ONE kernel for the unsharded act() (no message, no second launch): the elite rows are staged
// in shared memory by the whole CTA, then mean / ddof-0 variance / alpha blend with exactly the arithmetic (and the
// summation order: elite order = reward descending, row ascending) of cem_refit_kernel, so that sharded and
// unsharded runs stay bit-identical.  elite_cache: E*HU floats of dynamic shared memory behind the key cache, or 0.
__global__ void __launch_bounds__(SEL_THREADS) cem_topk_refit_kernel(const float* returns, const float* samples, int P_local, int A, int HU,
                                                                     int E, int use_cache, int key_words, int elite_cache, float alpha,
                                                                     float* mean, float* var) {
  __shared__ int hist[256];
  __shared__ int misc[40];
  __shared__ uint32_t keys[SEL_MAX_K];
  __shared__ int idx[SEL_MAX_K];
  extern __shared__ uint32_t topk_kcache[];
  const int a = blockIdx.x;
  const SelectScratch sc{hist, misc, keys, idx};
  block_topk(returns + a, P_local, A, E, sc, use_cache ? topk_kcache : nullptr);
  __syncthreads();
  float* elite = reinterpret_cast<float*>(topk_kcache + key_words);
  if (elite_cache)
    for (int i = threadIdx.x; i < E * HU; i += SEL_THREADS) {
      const int k = i / HU, e = i - k * HU, p = idx[k];
      elite[i] = (p != 0x7FFFFFFF) ? samples[(static_cast<size_t>(p) * A + a) * HU + e] : 0.0f;
    }
  __syncthreads();
  for (int e = threadIdx.x; e < HU; e += SEL_THREADS) {
    auto at = [&](int k) -> float {
      if (elite_cache) return elite[k * HU + e];
      const int p = idx[k];
      return (p != 0x7FFFFFFF) ? samples[(static_cast<size_t>(p) * A + a) * HU + e] : 0.0f;
    };
    float sum = 0.0f;
    for (int k = 0; k < E; ++k) sum += at(k);
    const float nm = __fdiv_rn(sum, static_cast<float>(E));
    float sq = 0.0f;
    for (int k = 0; k < E; ++k) { const float d = __fsub_rn(at(k), nm); sq += __fmul_rn(d, d); }
    const float nv = __fdiv_rn(sq, static_cast<float>(E));
    const float one_m = __fsub_rn(1.0f, alpha);
    mean[a * HU + e] = __fadd_rn(__fmul_rn(alpha, mean[a * HU + e]), __fmul_rn(one_m, nm));
    var[a * HU + e] = __fadd_rn(__fmul_rn(alpha, var[a * HU + e]), __fmul_rn(one_m, nv));
  }
}

// ---- CEM merge + refit (cem.py:98-125).  One CTA per agent; candidates = world x E records.
__global__ void __launch_bounds__(SEL_THREADS) cem_refit_kernel(const float* partials, int world, int A, int HU,
                                                                int E, float alpha, float* mean, float* var,
                                                                int64_t partial_stride, int elite_cache) {
  __shared__ uint32_t keys[SEL_MAX_K];
  __shared__ int gp[SEL_MAX_K];
  __shared__ int src[SEL_MAX_K];
  const int a = blockIdx.x, rec = 2 + HU, n = world * E;
  int n2 = 1; while (n2 < n) n2 <<= 1;
  // sort candidates by (reward desc, global p asc); `src` carries the candidate slot along.
  for (int i = threadIdx.x; i < n2; i += SEL_THREADS) {
    if (i < n) {
      const float* r = partials + (i / E) * partial_stride + (static_cast<size_t>(a) * E + (i % E)) * rec;
      keys[i] = f2key(r[0]); gp[i] = __float_as_int(r[1]);
    } else { keys[i] = 0u; gp[i] = 0x7FFFFFFF; }
  }
  __syncthreads();
  if (world > 1) {
    // rank the candidates: position = number of candidates strictly before it (n <= 1024, O(n^2/threads))
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
      int pos = 0;
      for (int j = 0; j < n; ++j) pos += before(keys[j], gp[j], keys[i], gp[i]) ? 1 : 0;
      src[pos] = i;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) src[i] = i;  // local list is already sorted
  }
  __syncthreads();
  // elites = first E in order; mean, ddof-0 variance in elite order; alpha blend.  The E elite sequences are staged in
  // shared memory by the whole CTA first (elite_cache: E*HU floats of dynamic shared memory, else read in place): HU
  // threads each walking E rows twice was a chain of 2E dependent loads.
  extern __shared__ float refit_elite[];
  if (elite_cache) {
    for (int t = threadIdx.x; t < E * HU; t += SEL_THREADS) {
      const int k = t / HU, e = t - k * HU, i = src[k];
      refit_elite[t] = partials[(i / E) * partial_stride + (static_cast<size_t>(a) * E + (i % E)) * rec + 2 + e];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < HU; e += SEL_THREADS) {
    auto at = [&](int k) -> float {
      if (elite_cache) return refit_elite[k * HU + e];
      const int i = src[k];
      return partials[(i / E) * partial_stride + (static_cast<size_t>(a) * E + (i % E)) * rec + 2 + e];
    };
    float sum = 0.0f;
    for (int k = 0; k < E; ++k) sum += at(k);
    const float nm = __fdiv_rn(sum, static_cast<float>(E));
    float sq = 0.0f;
    for (int k = 0; k < E; ++k) {
      const float d = __fsub_rn(at(k), nm);
      sq += __fmul_rn(d, d);
    }
    const float nv = __fdiv_rn(sq, static_cast<float>(E));
    const float one_m = __fsub_rn(1.0f, alpha);
    mean[a * HU + e] = __fadd_rn(__fmul_rn(alpha, mean[a * HU + e]), __fmul_rn(one_m, nm));
    var[a * HU + e] = __fadd_rn(__fmul_rn(alpha, var[a * HU + e]), __fmul_rn(one_m, nv));
  }
}

// ---- PI2 local partial (pi2.py:79-87): per agent (max_r, sum e, sum e*x[HU]), e = exp((r - max_r)/lambda)
__global__ void __launch_bounds__(1024) pi2_partial_kernel(const float* returns, const float* samples, float* partial,
                                                           int P_local, int A, int HU, float lamda) {
  __shared__ float red[32];
  __shared__ float s_max, s_sum;
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (int p = tid; p < P_local; p += 1024) mx = fmaxf(mx, returns[p * A + a]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    mx = red[lane];
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_max = mx;
  }
  __syncthreads();
  mx = s_max;
  const float inv_l = __fdiv_rn(1.0f, lamda);
  float sm = 0.0f;
  for (int p = tid; p < P_local; p += 1024) sm += expf(-inv_l * ((-returns[p * A + a]) - (-mx)));
  for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
  __syncthreads();
  if (lane == 0) red[warp] = sm;
  __syncthreads();
  if (warp == 0) {
    sm = red[lane];
    for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
    if (lane == 0) s_sum = sm;
  }
  __syncthreads();
  float* out = partial + static_cast<size_t>(a) * (2 + HU);
  if (tid == 0) { out[0] = mx; out[1] = s_sum; }
}
// weighted sums sum_p e_p x[p, a, :]: block (a, b) takes a contiguous slice of the population, thread = element
// (coalesced sample reads, the slice's weights staged in shared memory), slices are added in slice order by
// pi2_wsum_reduce_kernel: deterministic, and no longer one CTA striding 3.6 MB with 32-byte-sector reads.
constexpr int PI2_SLICES = 64, PI2_ROWS = 128;
__global__ void __launch_bounds__(256) pi2_wsum_kernel(const float* __restrict__ returns, const float* __restrict__ samples,
                                                       const float* __restrict__ partial, float* __restrict__ scratch,
                                                       int P_local, int A, int HU, float lamda) {
  __shared__ float wgt[PI2_ROWS];
  const int a = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float mx = partial[static_cast<size_t>(a) * (2 + HU)];
  const float inv_l = __fdiv_rn(1.0f, lamda);
  const int per = (P_local + PI2_SLICES - 1) / PI2_SLICES;
  const int p_begin = b * per, p_end = (p_begin + per < P_local) ? p_begin + per : P_local;
  for (int e0 = 0; e0 < HU; e0 += 256) {
    const int e = e0 + tid;
    float acc = 0.0f;
    for (int p0 = p_begin; p0 < p_end; p0 += PI2_ROWS) {
      const int n = (p_end - p0 < PI2_ROWS) ? p_end - p0 : PI2_ROWS;
      __syncthreads();
      if (tid < n) wgt[tid] = expf(-inv_l * ((-returns[static_cast<size_t>(p0 + tid) * A + a]) - (-mx)));
      __syncthreads();
      if (e < HU)
        for (int i = 0; i < n; ++i) acc = fmaf(wgt[i], samples[(static_cast<size_t>(p0 + i) * A + a) * HU + e], acc);
    }
    if (e < HU) scratch[(static_cast<size_t>(a) * PI2_SLICES + b) * HU + e] = acc;
  }
}
__global__ void pi2_wsum_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ partial, int A, int HU) {
  const int a = blockIdx.x;
  for (int e = threadIdx.x; e < HU; e += blockDim.x) {
    float acc = 0.0f;
    for (int b = 0; b < PI2_SLICES; ++b) acc = __fadd_rn(acc, scratch[(static_cast<size_t>(a) * PI2_SLICES + b) * HU + e]);
    partial[static_cast<size_t>(a) * (2 + HU) + 2 + e] = acc;
  }
}
// PI2 merge: log-sum-exp combination of the ranks' partials -> new mean (pi2.py:80-87)
__global__ void pi2_merge_kernel(const float* partials, int world, int A, int HU, float lamda, float* mean,
                                 int64_t partial_stride) {
  const int a = blockIdx.x;
  const int rec = 2 + HU;
  float gmax = -INFINITY;
  for (int g = 0; g < world; ++g) gmax = fmaxf(gmax, partials[g * partial_stride + a * rec]);
  const float inv_l = __fdiv_rn(1.0f, lamda);
  float eta = 0.0f;
  for (int g = 0; g < world; ++g) {
    const float* r = partials + g * partial_stride + a * rec;
    eta += r[1] * expf(-inv_l * (gmax - r[0]));
  }
  for (int e = threadIdx.x; e < HU; e += blockDim.x) {
    float acc = 0.0f;
    for (int g = 0; g < world; ++g) {
      const float* r = partials + g * partial_stride + a * rec;
      acc += r[2 + e] * expf(-inv_l * (gmax - r[0]));
    }
    mean[a * HU + e] = __fmul_rn(__fdiv_rn(1.0f, eta), acc);
  }
}

// ---- RandomSearch / PSO: local argmax over the population (first index wins) -> (value, global p, seq[HU])
__global__ void __launch_bounds__(1024) argmax_partial_kernel(const float* vals, const float* seqs, float* partial,
                                                              int P_local, int p0, int A, int HU) {
  __shared__ uint32_t s_key[32];
  __shared__ int s_idx[32];
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t bk = 0; int bi = 0x7FFFFFFF;
  for (int p = tid; p < P_local; p += 1024) {
    const uint32_t k = f2key(vals[p * A + a]);
    if (before(k, p, bk, bi)) { bk = k; bi = p; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (before(ok, oi, bk, bi)) { bk = ok; bi = oi; }
  }
  if (lane == 0) { s_key[warp] = bk; s_idx[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bk = s_key[lane]; bi = s_idx[lane];
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (before(ok, oi, bk, bi)) { bk = ok; bi = oi; }
    }
    if (lane == 0) { s_key[0] = bk; s_idx[0] = bi; }
  }
  __syncthreads();
  bk = s_key[0]; bi = s_idx[0];
  float* out = partial + static_cast<size_t>(a) * (2 + HU);
  const bool ok = bi != 0x7FFFFFFF;
  if (tid == 0) { out[0] = ok ? key2f(bk) : -INFINITY; out[1] = __int_as_float(ok ? p0 + bi : 0x7FFFFFFF); }
  for (int e = tid; e < HU; e += 1024) out[2 + e] = ok ? seqs[(static_cast<size_t>(bi) * A + a) * HU + e] : 0.0f;
}
// merge: best record over ranks (value desc, global p asc) -> best_seq[a,:], best_val[a]
__global__ void argmax_merge_kernel(const float* partials, int world, int HU, float* best_seq, float* best_val,
                                    int64_t partial_stride) {
  const int a = blockIdx.x, rec = 2 + HU;
  int bg = 0; uint32_t bk = 0; int bi = 0x7FFFFFFF;
  for (int g = 0; g < world; ++g) {
    const float* r = partials + g * partial_stride + a * rec;
    const uint32_t k = f2key(r[0]); const int i = __float_as_int(r[1]);
    if (g == 0 || before(k, i, bk, bi)) { bk = k; bi = i; bg = g; }
  }
  const float* r = partials + bg * partial_stride + a * rec;
  for (int e = threadIdx.x; e < HU; e += blockDim.x) best_seq[a * HU + e] = r[2 + e];
  if (threadIdx.x == 0 && best_val) best_val[a] = r[0];
}

// ---- SPSA partial: sum_p (r+ - r-) / (2 ck delta) per (a,e) (spsa.py:98-103); delta regenerated
__global__ void spsa_partial_kernel(const float* returns, float* partial, int P_local, int p0, int A, int HU,
                                    float ck, uint64_t seed, const uint32_t* act_ctr, uint32_t iter, const float* inject) {
  // one warp per (a, e): lanes stride the population, fixed-order tree reduction
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= A * HU) return;
  const int a = w / HU, e = w % HU;
  const int64_t half = static_cast<int64_t>(P_local) * A;
  float acc = 0.0f;
  for (int p = lane; p < P_local; p += 32) {
    const Philox4 r = draw_block(seed, *act_ctr, STREAM_SAMPLES, iter, (p0 + p) * A + a, e >> 2);
    const uint32_t word = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
    const float delta = inject ? inject[(static_cast<int64_t>(p0 + p) * A + a) * HU + e] : ((word >> 31) ? 1.0f : -1.0f);
    const float diff = __fsub_rn(returns[p * A + a], returns[half + p * A + a]);
    acc += __fdiv_rn(diff, __fmul_rn(__fmul_rn(2.0f, ck), delta));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) partial[w] = acc;
}
// SPSA merge (spsa.py:101-107): ghat = sum / P; sol = clip(sol + ak * ghat)
__global__ void spsa_merge_kernel(const float* partials, int world, int n, int P, float ak, float* sol,
                                  const float* lb, const float* ub, int dU, int64_t partial_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int g = 0; g < world; ++g) s += partials[g * partial_stride + i];
  const float ghat = __fdiv_rn(s, static_cast<float>(P));
  const float v = __fadd_rn(sol[i], __fmul_rn(ak, ghat));
  sol[i] = fminf(fmaxf(v, lb[i % dU]), ub[i % dU]);
}

// ---- PSO (pso.py:79-111)
// clip + excess^2 in place, before evaluation
__global__ void pso_clip_kernel(float* x, float* excess_sq, const float* lb, const float* ub, int64_t n, int dU) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i], f = fminf(fmaxf(v, lb[i % dU]), ub[i % dU]);
  const float d = __fsub_rn(v, f);
  x[i] = f; excess_sq[i] = __fmul_rn(d, d);
}
// personal best update (pso.py:87-94): where pbest_r < reward
__global__ void pso_pbest_kernel(const float* x, const float* rewards, float* pbx, float* pbr, int64_t rows, int HU) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * HU) return;
  const int64_t row = i / HU;
  if (pbr[row] < rewards[row]) pbx[i] = x[i];
}
__global__ void pso_pbest_r_kernel(const float* rewards, float* pbr, int64_t rows) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < rows && pbr[i] < rewards[i]) pbr[i] = rewards[i];
}
// velocity / position update (pso.py:107-111); r1, r2: ONE N(0,1) scalar each per iteration
__global__ void pso_move_kernel(float* x, float* v, const float* pbx, const float* gbx, int64_t n, int AHU, float w,
                                float c1, float c2, uint64_t seed, const uint32_t* act_ctr, uint32_t iter, float* r_record) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Philox4 r = draw_block(seed, *act_ctr, STREAM_PSO_R, iter, 0, 0);
  const float r1 = std_normal(r.x), r2 = std_normal(r.y);
  if (i == 0 && r_record) { r_record[2 * iter] = r1; r_record[2 * iter + 1] = r2; }   // inspection: get_tensor("pso_r")
  const float xi = x[i];
  const float nv = __fadd_rn(__fadd_rn(__fmul_rn(v[i], w), __fmul_rn(__fmul_rn(__fsub_rn(pbx[i], xi), c1), r1)),
                             __fmul_rn(__fmul_rn(__fsub_rn(gbx[i % AHU], xi), c2), r2));
  v[i] = nv;
  x[i] = __fadd_rn(xi, nv);
}
// swarm (re-)seeding: mode 0 = reset() uniform positions (pso.py:147-159); mode 1 = the tail of
// _optimize: x ~ truncnorm(shift(gbest), sqrt(cvar(UNSHIFTED gbest))) (pso.py:116-138)
__global__ void pso_seed_kernel(float* x, float* v, float* pbx, const float* gbx, const float* var0, const float* lb,
                                const float* ub, int P_local, int p0, int A, int H, int dU, float v0frac, int mode,
                                uint64_t seed, const uint32_t* act_ctr) {
  const int HU = H * dU;
  const uint32_t act_call = *act_ctr;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(P_local) * A * HU) return;
  const int e = i % HU; const int64_t pa = i / HU; const int a = pa % A; const int p = pa / A;
  const int u = e % dU, h = e / dU;
  const uint32_t row = (p0 + p) * A + a;
  const Philox4 rp = draw_block(seed, act_call, STREAM_PSO_POS, mode, row, e >> 2);
  const Philox4 rv = draw_block(seed, act_call, STREAM_PSO_VEL, mode, row, e >> 2);
  const uint32_t wp = (e & 3) == 0 ? rp.x : (e & 3) == 1 ? rp.y : (e & 3) == 2 ? rp.z : rp.w;
  const uint32_t wv = (e & 3) == 0 ? rv.x : (e & 3) == 1 ? rv.y : (e & 3) == 2 ? rv.z : rv.w;
  float pos;
  if (mode == 0) {
    pos = __fadd_rn(__fmul_rn(u01_halfopen(wp), __fsub_rn(ub[u], lb[u])), lb[u]);
  } else {
    const float g = gbx[a * HU + e];
    const int hs = h + 1 < H ? h + 1 : H - 1;
    const float gs = gbx[a * HU + hs * dU + u];
    const float lo = __fdiv_rn(__fsub_rn(g, lb[u]), 2.0f), hi = __fdiv_rn(__fsub_rn(ub[u], g), 2.0f);
    const float cvar = fminf(fminf(__fmul_rn(lo, lo), __fmul_rn(hi, hi)), var0[a * HU + e]);
    pos = __fadd_rn(__fmul_rn(std_truncnorm(wp), sqrtf(cvar)), gs);
  }
  const float v0 = __fmul_rn(v0frac, __fsub_rn(ub[u], lb[u]));
  const float vel = __fadd_rn(__fmul_rn(u01_halfopen(wv), __fsub_rn(v0, -v0)), -v0);
  x[i] = pos; pbx[i] = pos; v[i] = vel;
}

// ---- peer-memory exchange (NVLink P2P).  Publish: the partial message is already in this rank's exchange
// buffer; make it visible system-wide, then store the sequence number.  Gather: block g waits until rank g
// has published `seq`, then copies its message into the local gather buffer with plain P2P loads.
__global__ void p2p_publish_kernel(uint32_t* flag, uint32_t seq) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(seq) : "memory");
}
// `src_stride` (a multiple of 4 floats) separates the two parity copies of a message inside a rank's exchange buffer, so
// the 16-byte loads are aligned for any message length; the gathered copies stay dense (n_floats apart), hence the
// vector path only when n_floats itself is a multiple of 4.
__global__ void __launch_bounds__(1024) p2p_gather_kernel(float* const* peers, float* gathered, int n_floats, int src_stride,
                                                         size_t flag_off_floats, uint32_t seq, int parity) {
  const int g = blockIdx.x;
  const float* src_base = peers[g];
  if (threadIdx.x == 0) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(src_base + flag_off_floats);
    uint32_t seen = 0, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
      if (seen < seq && ++spins > (1u << 26)) asm volatile("trap;");   // a peer died: abort instead of hanging
    } while (seen < seq);
  }
  __syncthreads();
  const float* msg = src_base + static_cast<size_t>(parity) * src_stride;
  const int n_vec = (n_floats & 3) == 0 ? n_floats / 4 : 0;
  const float4* src = reinterpret_cast<const float4*>(msg);
  float4* dst = reinterpret_cast<float4*>(gathered + static_cast<size_t>(g) * n_floats);
  // four peer loads in flight per thread before the first store: a load over NVLink is a ~2 us round trip, and a
  // load -> store -> load chain per thread made the gather of a 36 KB message take nine of them
  constexpr int U = 4;
  for (int i0 = threadIdx.x; i0 < n_vec; i0 += U * blockDim.x) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n_vec)
        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n_vec) dst[i] = v[u];
    }
  }
  for (int i = 4 * n_vec + threadIdx.x; i < n_floats; i += blockDim.x) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(msg + i) : "memory");
    gathered[static_cast<size_t>(g) * n_floats + i] = v;
  }
}

inline int grid_for(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

}  // namespace

namespace bbmpc {
void launch_penalty(const float* excess_sq, float* penalty, int64_t rows, int HU, cudaStream_t st) {
  penalty_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(excess_sq, penalty, rows, HU);
}
void launch_topk_partial(const float* returns, const float* samples, float* partial, int P_local, int p0, int A, int HU, int E,
                         cudaStream_t st, int n_slices, int64_t slice_stride) {
  // opt in to > 48 KB of dynamic shared memory (per device and function: set on every launch, it is a host-side table write)
  cudaFuncSetAttribute(topk_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int per = (P_local + n_slices - 1) / n_slices + 1;
  const int use_cache = (P_local > 0 && per <= 16384) ? 1 : 0;   // 64 KB of dynamic shared memory at most
  topk_partial_kernel<<<dim3(A, n_slices), SEL_THREADS, use_cache ? per * sizeof(uint32_t) : 0, st>>>(returns, samples, partial, P_local, p0, A, HU, E,
                                                                                                    use_cache, slice_stride);
}
void launch_cem_topk_refit(const float* returns, const float* samples, int P_local, int A, int HU, int E, float alpha, float* mean,
                           float* var, cudaStream_t st) {
  cudaFuncSetAttribute(cem_topk_refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int use_cache = (P_local > 0 && P_local + 1 <= 16384) ? 1 : 0;
  const int key_words = use_cache ? P_local + 1 : 0;
  const size_t elite_bytes = static_cast<size_t>(E) * HU * sizeof(float);
  const int elite_cache = (key_words * sizeof(uint32_t) + elite_bytes <= 160 * 1024) ? 1 : 0;
  cem_topk_refit_kernel<<<A, SEL_THREADS, key_words * sizeof(uint32_t) + (elite_cache ? elite_bytes : 0), st>>>(
      returns, samples, P_local, A, HU, E, use_cache, key_words, elite_cache, alpha, mean, var);
}
}  // namespace bbmpc

// ============================================================================ host orchestration
static int partial_floats(const bbmpc_opt* o) {
  const int A = o->cfg.num_agents, HU = o->HU;
  switch (o->cfg.kind) {
    case BBMPC_OPT_CEM: return A * o->cfg.num_elite * (2 + HU);
    case BBMPC_OPT_PI2: return A * (2 + HU);
    case BBMPC_OPT_RANDOM_SEARCH: return A * (2 + HU);
    case BBMPC_OPT_PSO: return A * (2 + HU);
    case BBMPC_OPT_SPSA: return A * HU;
    case BBMPC_OPT_CMAES: return o->cfg.num_elite * (2 + A * HU);
    default: return 0;
  }
}

static int set_shard(bbmpc_opt* o, int rank, int world);
// population slices of the unsharded CEM top-E: candidates of all slices (slices x E) must fit the merge kernel's sort
// (measured at P = 10 000: 8 slices take the selection from 34 to 29 us but the 400-candidate merge from 17 to 27 us: the
// kernels are bound by their ~25 block-wide phases, not by the scan, so slices only pay for much larger populations)
static int cem_slice_count(const bbmpc_opt* o) {
  int s = o->P_local / 32768;
  const int cap = SEL_MAX_K / (o->cfg.num_elite > 0 ? o->cfg.num_elite : 1);
  if (s > 8) s = 8;
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}
static void opt_graph_drop(bbmpc_opt* o) {
  if (o->graph_exec) cudaGraphExecDestroy(o->graph_exec);
  o->graph_exec = nullptr; o->graph_warm = 0; o->graph_noise = -1;
}

extern "C" {

int bbmpc_opt_create(bbmpc_ctx* ctx, const bbmpc_opt_config* cfg, bbmpc_opt** out) {
  if (!ctx || !cfg || !out) return BBMPC_EINVAL;
  *out = nullptr;
  const int P = cfg->population_size, A = cfg->num_agents, H = cfg->planning_horizon, dU = cfg->dU, dS = cfg->dS;
  if (cfg->kind < BBMPC_OPT_CEM || cfg->kind > BBMPC_OPT_CMAES) return fail(ctx, BBMPC_EINVAL, "unknown optimizer kind %d", cfg->kind);
  if (P < 1 || A < 1 || H < 1 || dU < 1 || dS < 1 || dU > MAX_DU || dS > MAX_DS)
    return fail(ctx, BBMPC_EINVAL, "bad optimizer shape P=%d A=%d H=%d dS=%d dU=%d", P, A, H, dS, dU);
  if (!cfg->lb_host || !cfg->ub_host) return fail(ctx, BBMPC_EINVAL, "action bounds missing");
  if (cfg->kind != BBMPC_OPT_RANDOM_SEARCH && cfg->max_iterations < 0) return fail(ctx, BBMPC_EINVAL, "max_iterations < 0");
  if (cfg->kind == BBMPC_OPT_CEM && (cfg->num_elite < 1 || cfg->num_elite > P || cfg->num_elite > SEL_MAX_K))
    return fail(ctx, BBMPC_EINVAL, "num_elite=%d must be in [1, min(P, %d)]", cfg->num_elite, SEL_MAX_K);
  if (ctx->model.set && (ctx->model.dS != dS || ctx->model.dU != dU))
    return fail(ctx, BBMPC_EINVAL, "optimizer dS=%d dU=%d differ from the model's dS=%d dU=%d", dS, dU, ctx->model.dS, ctx->model.dU);
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  bbmpc_opt* o = new bbmpc_opt();
  o->ctx = ctx; o->cfg = *cfg; o->cfg.lb_host = nullptr; o->cfg.ub_host = nullptr;
  o->HU = H * dU; o->AHU = A * H * dU;
  int rc = BBMPC_OK;
  auto A_ = [&](int r) { if (rc == BBMPC_OK) rc = r; };
  A_(dalloc(o, &o->d_lb, dU)); A_(dalloc(o, &o->d_ub, dU));
  A_(dalloc(o, &o->d_state, A * dS)); A_(dalloc(o, &o->d_mean, o->AHU)); A_(dalloc(o, &o->d_var, o->AHU));
  A_(dalloc(o, &o->d_prev, o->AHU)); A_(dalloc(o, &o->d_var0, o->AHU));
  A_(dalloc(o, &o->d_action, A * dU)); A_(dalloc(o, &o->d_next, A * dS)); A_(dalloc(o, &o->d_reward, A));
  A_(dalloc(o, &o->d_act_ctr, 1));
  if (cfg->kind == BBMPC_OPT_PSO) { A_(dalloc(o, &o->d_gbx, o->AHU)); A_(dalloc(o, &o->d_gbr, A)); A_(dalloc(o, &o->d_sol, A * dU)); A_(dalloc(o, &o->d_pso_r, 128)); }
  if (rc == BBMPC_OK && cudaMallocHost(reinterpret_cast<void**>(&o->h_pinned), (A * dS * 2 + A * dU + A) * sizeof(float)) != cudaSuccess)
    rc = fail(ctx, BBMPC_ENOMEM, "cudaMallocHost failed");
  if (rc != BBMPC_OK) { bbmpc_opt_destroy(o); return rc; }
  cudaMemset(o->d_act_ctr, 0, sizeof(uint32_t));
  cudaMemcpy(o->d_lb, cfg->lb_host, dU * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(o->d_ub, cfg->ub_host, dU * sizeof(float), cudaMemcpyHostToDevice);
  fill_midpoint_kernel<<<grid_for(o->AHU, 256), 256>>>(o->d_prev, o->d_lb, o->d_ub, o->AHU, dU);
  fill_var0_kernel<<<grid_for(o->AHU, 256), 256>>>(o->d_var0, o->d_lb, o->d_ub, o->AHU, dU);
  ctx->launches += 2;
  if (cfg->kind == BBMPC_OPT_PSO) {  // every tf.Variable starts at zero (pso.py:50-68)
    cudaMemset(o->d_gbx, 0, o->AHU * sizeof(float)); cudaMemset(o->d_gbr, 0, A * sizeof(float));
    cudaMemset(o->d_sol, 0, A * dU * sizeof(float));
  }
  if ((rc = set_shard(o, 0, 1)) != BBMPC_OK) { bbmpc_opt_destroy(o); return rc; }
  if (cfg->kind == BBMPC_OPT_CMAES && (rc = cmaes_create(o)) != BBMPC_OK) { bbmpc_opt_destroy(o); return rc; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { rc = fail(ctx, BBMPC_ECUDA, "optimizer init: %s", cudaGetErrorString(e)); bbmpc_opt_destroy(o); return rc; }
  *out = o;
  return BBMPC_OK;
}

void bbmpc_opt_destroy(bbmpc_opt* o) {
  if (!o) return;
  cudaSetDevice(o->ctx->device);
  cudaDeviceSynchronize();
  if (o->graph_exec) cudaGraphExecDestroy(o->graph_exec);
  if (o->graph_stream) cudaStreamDestroy(o->graph_stream);
  if (o->cfg.kind == BBMPC_OPT_CMAES) cmaes_destroy(o);
  for (void* p : o->p2p_opened) cudaIpcCloseMemHandle(p);
  for (void* p : o->owned) cudaFree(p);
  if (o->h_pinned) cudaFreeHost(o->h_pinned);
  delete o;
}

}  // extern "C"

// (Re)allocates the population-sized buffers for this rank's slice.
static int set_shard(bbmpc_opt* o, int rank, int world) {
  bbmpc_ctx* ctx = o->ctx;
  opt_graph_drop(o);   // the population-sized buffers are reallocated below
  const int P = o->cfg.population_size, A = o->cfg.num_agents;
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, BBMPC_EINVAL, "bad shard %d/%d", rank, world);
  if (world > 1 && o->cfg.kind == BBMPC_OPT_CEM && world * o->cfg.num_elite > SEL_MAX_K)
    return fail(ctx, BBMPC_EINVAL, "world*num_elite = %d exceeds %d", world * o->cfg.num_elite, SEL_MAX_K);
  BB_CUDA(ctx, cudaDeviceSynchronize());
  o->rank = rank; o->world = world;
  o->p0 = static_cast<int>(static_cast<int64_t>(P) * rank / world);
  o->P_local = static_cast<int>(static_cast<int64_t>(P) * (rank + 1) / world) - o->p0;
  o->n_eval = o->cfg.kind == BBMPC_OPT_SPSA ? 2 * o->P_local : o->P_local;
  const size_t rows = static_cast<size_t>(o->n_eval) * A;
  int rc = BBMPC_OK;
  auto A_ = [&](int r) { if (rc == BBMPC_OK) rc = r; };
  // (old population buffers stay in `owned` until destroy; resharding is a setup-time operation)
  A_(dalloc(o, &o->d_samples, rows * o->HU)); A_(dalloc(o, &o->d_returns, rows)); A_(dalloc(o, &o->d_penalty, rows));
  A_(dalloc(o, &o->d_partial, static_cast<size_t>(partial_floats(o)) * (o->cfg.kind == BBMPC_OPT_CEM ? 8 : 1)));
  if (o->cfg.kind == BBMPC_OPT_PI2 || o->cfg.kind == BBMPC_OPT_SPSA || o->cfg.kind == BBMPC_OPT_PSO)
    A_(dalloc(o, &o->d_work, rows * o->HU));
  if (o->cfg.kind == BBMPC_OPT_CMAES) A_(cmaes_set_shard(o));
  if (o->cfg.kind == BBMPC_OPT_PI2 && !o->d_pi2_scratch) A_(dalloc(o, &o->d_pi2_scratch, static_cast<size_t>(A) * PI2_SLICES * o->HU));
  if (o->cfg.kind == BBMPC_OPT_PSO) {
    A_(dalloc(o, &o->d_v, rows * o->HU)); A_(dalloc(o, &o->d_pbx, rows * o->HU)); A_(dalloc(o, &o->d_pbr, rows));
    if (rc == BBMPC_OK) {
      cudaMemset(o->d_samples, 0, rows * o->HU * sizeof(float)); cudaMemset(o->d_v, 0, rows * o->HU * sizeof(float));
      cudaMemset(o->d_pbx, 0, rows * o->HU * sizeof(float)); cudaMemset(o->d_pbr, 0, rows * sizeof(float));
    }
  }
  return rc;
}

extern "C" {

int bbmpc_opt_set_shard(bbmpc_opt* o, int rank, int world) {
  if (!o) return BBMPC_EINVAL;
  BB_CUDA(o->ctx, cudaSetDevice(o->ctx->device));
  return set_shard(o, rank, world);
}

int bbmpc_opt_num_iterations(const bbmpc_opt* o) {
  if (!o) return BBMPC_EINVAL;
  return o->cfg.kind == BBMPC_OPT_RANDOM_SEARCH ? 1 : o->cfg.max_iterations;
}
int bbmpc_opt_partial_floats(const bbmpc_opt* o) { return o ? partial_floats(o) : BBMPC_EINVAL; }

int bbmpc_opt_reset(bbmpc_opt* o, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int dU = o->cfg.dU;
  switch (o->cfg.kind) {
    case BBMPC_OPT_CEM: case BBMPC_OPT_PI2: case BBMPC_OPT_SPSA:  // mean/params back to the midpoint only
      fill_midpoint_kernel<<<grid_for(o->AHU, 256), 256, 0, st>>>(o->d_prev, o->d_lb, o->d_ub, o->AHU, dU);
      BB_LAUNCH_CHECK(ctx);
      break;
    case BBMPC_OPT_PSO: {
      const int64_t n = static_cast<int64_t>(o->P_local) * o->AHU;
      pso_seed_kernel<<<grid_for(n, 256), 256, 0, st>>>(o->d_samples, o->d_v, o->d_pbx, o->d_gbx, o->d_var0, o->d_lb, o->d_ub,
                                                        o->P_local, o->p0, o->cfg.num_agents, o->cfg.planning_horizon, dU,
                                                        o->cfg.initial_velocity_fraction, 0, ctx->seed, o->d_act_ctr);
      BB_LAUNCH_CHECK(ctx);
      const int64_t rows = static_cast<int64_t>(o->P_local) * o->cfg.num_agents;
      fill_kernel<<<grid_for(rows, 256), 256, 0, st>>>(o->d_pbr, -INFINITY, rows); BB_LAUNCH_CHECK(ctx);
      fill_kernel<<<1, 256, 0, st>>>(o->d_gbr, -INFINITY, o->cfg.num_agents); BB_LAUNCH_CHECK(ctx);
      o->act_call++;  // a reset consumes its own Philox sub-stream
      break;
    }
    case BBMPC_OPT_CMAES:
      return cmaes_reset(o, st);
    default: break;  // RandomSearch: nothing
  }
  return BBMPC_OK;
}

int bbmpc_opt_begin(bbmpc_opt* o, const float* state, int time_step, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!state) return opt_fail(o, BBMPC_EINVAL, "state is NULL");
  if (!ctx->model.set) return opt_fail(o, BBMPC_ESTATE, "optimizer called before a dynamics model was set");
  if (!ctx->reward_id) return opt_fail(o, BBMPC_ESTATE, "optimizer called before a reward function was set");
  if (ctx->model.dS != o->cfg.dS || ctx->model.dU != o->cfg.dU) return opt_fail(o, BBMPC_EINVAL, "optimizer/model dS,dU mismatch");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  o->time_step = time_step;  // accepted and ignored, as in deterministic.py:26
  if (state != o->d_state)
    BB_CUDA(ctx, cudaMemcpyAsync(o->d_state, state, o->cfg.num_agents * o->cfg.dS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  switch (o->cfg.kind) {
    case BBMPC_OPT_CEM:  // loop_vars = [_previous_solution, _solution_variance] (cem.py:129-132)
      BB_CUDA(ctx, cudaMemcpyAsync(o->d_mean, o->d_prev, o->AHU * sizeof(float), cudaMemcpyDeviceToDevice, st));
      BB_CUDA(ctx, cudaMemcpyAsync(o->d_var, o->d_var0, o->AHU * sizeof(float), cudaMemcpyDeviceToDevice, st));
      break;
    case BBMPC_OPT_PI2: case BBMPC_OPT_SPSA:
      BB_CUDA(ctx, cudaMemcpyAsync(o->d_mean, o->d_prev, o->AHU * sizeof(float), cudaMemcpyDeviceToDevice, st));
      break;
    default: break;
  }
  o->began = true;
  return BBMPC_OK;
}

int bbmpc_opt_iter_local(bbmpc_opt* o, int iter, float* partial_out, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!o->began) return opt_fail(o, BBMPC_ESTATE, "iter_local before begin");
  NvtxRange nvtx_range("bbmpc:sample+rollout+local_refit");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bbmpc_opt_config& c = o->cfg;
  const int A = c.num_agents, H = c.planning_horizon, dU = c.dU, HU = o->HU;
  float* partial = partial_out ? partial_out : o->d_partial;
  if (c.kind == BBMPC_OPT_CMAES) return cmaes_iter_local(o, iter, partial, st);
  SampleArgs s{o->d_samples, o->d_penalty, o->d_mean, o->d_var, o->d_lb, o->d_ub, o->P_local, o->p0, A, H, dU,
               ctx->seed, o->d_act_ctr, static_cast<uint32_t>(iter), 0.0f, nullptr, nullptr};
  {  // injected standard variates: one block of [P, A, H*dU] per iteration since bbmpc_opt_set_draw_injection
    const int64_t blk = static_cast<int64_t>(c.population_size) * A * HU;
    if (o->inject && c.kind != BBMPC_OPT_PSO) {
      if ((o->inject_iter + 1) * blk > o->inject_floats) return opt_fail(o, BBMPC_EINVAL, "draw injection buffer exhausted");
      s.inject = o->inject + o->inject_iter * blk;
      ++o->inject_iter;
    }
  }
  const int64_t rows = static_cast<int64_t>(o->n_eval) * A;
  const int64_t nthr = static_cast<int64_t>(o->P_local) * A * ((HU + 3) / 4);
  const float* penalty = nullptr;
  const int64_t per_iter = rows * HU;
  float* trace_slot = (o->trace && (static_cast<int64_t>(iter) + 1) * per_iter <= o->trace_floats) ? o->trace + iter * per_iter : nullptr;
  if (o->P_local > 0) {
    switch (c.kind) {
      case BBMPC_OPT_CEM:
        cem_sample_kernel<<<grid_for(nthr, 256), 256, 0, st>>>(s); BB_LAUNCH_CHECK(ctx);
        break;
      case BBMPC_OPT_PI2:
        s.var = o->d_var0;
        s.raw_trace = trace_slot;
        pi2_sample_kernel<<<grid_for(nthr, 256), 256, 0, st>>>(s, o->d_work); BB_LAUNCH_CHECK(ctx);
        penalty_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(o->d_work, o->d_penalty, rows, HU); BB_LAUNCH_CHECK(ctx);
        penalty = o->d_penalty;
        break;
      case BBMPC_OPT_RANDOM_SEARCH:
        rs_sample_kernel<<<grid_for(nthr, 256), 256, 0, st>>>(s); BB_LAUNCH_CHECK(ctx);
        break;
      case BBMPC_OPT_SPSA:
        s.ck = c.noise_parameter / powf(static_cast<float>(iter) + 1.0f, c.gamma);
        spsa_sample_kernel<<<grid_for(nthr, 256), 256, 0, st>>>(s, o->d_work); BB_LAUNCH_CHECK(ctx);
        penalty_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(o->d_work, o->d_penalty, rows, HU); BB_LAUNCH_CHECK(ctx);
        penalty = o->d_penalty;
        break;
      case BBMPC_OPT_PSO:
        pso_clip_kernel<<<grid_for(rows * HU, 256), 256, 0, st>>>(o->d_samples, o->d_work, o->d_lb, o->d_ub, rows * HU, dU); BB_LAUNCH_CHECK(ctx);
        penalty_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(o->d_work, o->d_penalty, rows, HU); BB_LAUNCH_CHECK(ctx);
        penalty = o->d_penalty;
        break;
      default: return opt_fail(o, BBMPC_EINVAL, "unsupported optimizer kind");
    }
    if (trace_slot && c.kind != BBMPC_OPT_PI2)
      BB_CUDA(ctx, cudaMemcpyAsync(trace_slot, o->d_samples, per_iter * sizeof(float), cudaMemcpyDeviceToDevice, st));
    {
      NvtxRange nvtx_rollout("bbmpc:rollout");
      if (int rc = rollout_dispatch(ctx, o->d_state, o->d_samples, o->d_returns, penalty, static_cast<int>(rows), A, H, st)) return rc;
    }
  }
  switch (c.kind) {
    case BBMPC_OPT_CEM:
      // unsharded, internal message buffer: the local rows are cut into slices, one CTA each (a single CTA scanning
      // 10 000 returns took 34 us per iteration); bbmpc_opt_iter_merge ranks the slices' candidates like ranks'
      o->cem_slices = (!partial_out && o->world == 1) ? cem_slice_count(o) : 1;
      if (o->cem_fuse && !partial_out && o->world == 1 && o->cem_slices == 1) {   // inside bbmpc_opt_call: top-E and refit in one kernel
        launch_cem_topk_refit(o->d_returns, o->d_samples, o->P_local, A, HU, c.num_elite, c.alpha, o->d_mean, o->d_var, st);
        o->cem_fused_done = true;
        break;
      }
      launch_topk_partial(o->d_returns, o->d_samples, partial, o->P_local, o->p0, A, HU, c.num_elite, st, o->cem_slices, partial_floats(o));
      break;
    case BBMPC_OPT_PI2:
      pi2_partial_kernel<<<A, 1024, 0, st>>>(o->d_returns, o->d_samples, partial, o->P_local, A, HU, c.lamda); BB_LAUNCH_CHECK(ctx);
      pi2_wsum_kernel<<<dim3(A, PI2_SLICES), 256, 0, st>>>(o->d_returns, o->d_samples, partial, o->d_pi2_scratch, o->P_local, A, HU, c.lamda); BB_LAUNCH_CHECK(ctx);
      pi2_wsum_reduce_kernel<<<A, 256, 0, st>>>(o->d_pi2_scratch, partial, A, HU);
      break;
    case BBMPC_OPT_RANDOM_SEARCH:
      argmax_partial_kernel<<<A, 1024, 0, st>>>(o->d_returns, o->d_samples, partial, o->P_local, o->p0, A, HU);
      break;
    case BBMPC_OPT_SPSA:
      spsa_partial_kernel<<<grid_for(static_cast<int64_t>(A) * HU * 32, 256), 256, 0, st>>>(
          o->d_returns, partial, o->P_local, o->p0, A, HU, c.noise_parameter / powf(static_cast<float>(iter) + 1.0f, c.gamma),
          ctx->seed, o->d_act_ctr, static_cast<uint32_t>(iter), s.inject);
      break;
    case BBMPC_OPT_PSO: {
      pso_pbest_kernel<<<grid_for(rows * HU, 256), 256, 0, st>>>(o->d_samples, o->d_returns, o->d_pbx, o->d_pbr, rows, HU); BB_LAUNCH_CHECK(ctx);
      pso_pbest_r_kernel<<<grid_for(rows, 256), 256, 0, st>>>(o->d_returns, o->d_pbr, rows); BB_LAUNCH_CHECK(ctx);
      argmax_partial_kernel<<<A, 1024, 0, st>>>(o->d_pbr, o->d_pbx, partial, o->P_local, o->p0, A, HU);
      break;
    }
    default: break;
  }
  BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

int bbmpc_opt_iter_merge(bbmpc_opt* o, int iter, const float* partials, int world, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!o->began) return opt_fail(o, BBMPC_ESTATE, "iter_merge before begin");
  NvtxRange nvtx_range("bbmpc:merge+refit");
  if (world < 1) return opt_fail(o, BBMPC_EINVAL, "world < 1");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bbmpc_opt_config& c = o->cfg;
  const int A = c.num_agents, HU = o->HU;
  const float* in = partials ? partials : o->d_partial;
  const int64_t stride = partial_floats(o);
  if (c.kind == BBMPC_OPT_CMAES) return cmaes_iter_merge(o, iter, in, world, st);
  switch (c.kind) {
    case BBMPC_OPT_CEM:
      if (world * c.num_elite > SEL_MAX_K) return opt_fail(o, BBMPC_EINVAL, "world*num_elite exceeds 1024");
      if (o->cem_fused_done) { o->cem_fused_done = false; return BBMPC_OK; }   // refit already done by cem_topk_refit_kernel
      {
        const size_t eb = static_cast<size_t>(c.num_elite) * HU * sizeof(float);
        const int elite_cache = eb <= 160 * 1024 ? 1 : 0;
        if (elite_cache) cudaFuncSetAttribute(cem_refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cem_refit_kernel<<<A, SEL_THREADS, elite_cache ? eb : 0, st>>>(in, (!partials && world == 1) ? o->cem_slices : world, A, HU, c.num_elite,
                                                                     c.alpha, o->d_mean, o->d_var, stride, elite_cache);
      }
      break;
    case BBMPC_OPT_PI2:
      pi2_merge_kernel<<<A, 256, 0, st>>>(in, world, A, HU, c.lamda, o->d_mean, stride);
      break;
    case BBMPC_OPT_RANDOM_SEARCH:
      argmax_merge_kernel<<<A, 256, 0, st>>>(in, world, HU, o->d_mean, nullptr, stride);
      break;
    case BBMPC_OPT_SPSA: {
      const float big_a = static_cast<float>(c.max_iterations) / 10.0f;
      const float ak = c.a_par / powf(static_cast<float>(iter) + 1.0f + big_a, c.alpha);
      spsa_merge_kernel<<<grid_for(o->AHU, 256), 256, 0, st>>>(in, world, o->AHU, c.population_size, ak, o->d_mean, o->d_lb, o->d_ub, c.dU, stride);
      break;
    }
    case BBMPC_OPT_PSO: {
      argmax_merge_kernel<<<A, 256, 0, st>>>(in, world, HU, o->d_gbx, o->d_gbr, stride); BB_LAUNCH_CHECK(ctx);
      const int64_t n = static_cast<int64_t>(o->P_local) * o->AHU;
      if (n > 0)
        pso_move_kernel<<<grid_for(n, 256), 256, 0, st>>>(o->d_samples, o->d_v, o->d_pbx, o->d_gbx, n, o->AHU, c.w, c.c1, c.c2,
                                                          ctx->seed, o->d_act_ctr, static_cast<uint32_t>(iter),
                                                          iter < 64 ? o->d_pso_r : nullptr);
      else return BBMPC_OK;
      break;
    }
    default: return opt_fail(o, BBMPC_EINVAL, "unsupported optimizer kind");
  }
  BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

int bbmpc_opt_finish(bbmpc_opt* o, int add_noise, float* action, float* next_state, float* reward, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!o->began) return opt_fail(o, BBMPC_ESTATE, "finish before begin");
  NvtxRange nvtx_range("bbmpc:finish(first action, predict, reward)");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bbmpc_opt_config& c = o->cfg;
  const int A = c.num_agents, H = c.planning_horizon, dU = c.dU, dS = c.dS, HU = o->HU;
  const float* sol = o->d_mean; int sol_stride = HU;
  switch (c.kind) {
    case BBMPC_OPT_PI2: case BBMPC_OPT_SPSA:  // warm start: shift left, repeat last (pi2.py:92-93, spsa.py:114-115)
      shift_left_kernel<<<grid_for(o->AHU, 256), 256, 0, st>>>(o->d_mean, o->d_prev, A, H, dU); BB_LAUNCH_CHECK(ctx);
      break;
    case BBMPC_OPT_PSO: {  // _solution = gbest[:,0,:]; then re-seed the swarm (pso.py:114-138)
      sol = o->d_gbx;
      const int64_t n = static_cast<int64_t>(o->P_local) * o->AHU;
      if (n > 0) {
        pso_seed_kernel<<<grid_for(n, 256), 256, 0, st>>>(o->d_samples, o->d_v, o->d_pbx, o->d_gbx, o->d_var0, o->d_lb, o->d_ub,
                                                          o->P_local, o->p0, A, H, dU, c.initial_velocity_fraction, 1, ctx->seed, o->d_act_ctr);
        BB_LAUNCH_CHECK(ctx);
        const int64_t rows = static_cast<int64_t>(o->P_local) * A;
        fill_kernel<<<grid_for(rows, 256), 256, 0, st>>>(o->d_pbr, -INFINITY, rows); BB_LAUNCH_CHECK(ctx);
      }
      fill_kernel<<<1, 256, 0, st>>>(o->d_gbr, -INFINITY, A); BB_LAUNCH_CHECK(ctx);
      break;
    }
    default: break;  // CEM: no warm start (cem.py:133-134); RandomSearch: stateless
  }
  float* act_out = action ? action : o->d_action;
  float* next_out = next_state ? next_state : o->d_next;
  float* rew_out = reward ? reward : o->d_reward;
  first_action_kernel<<<grid_for(A * dU, 128), 128, 0, st>>>(sol, act_out, o->d_lb, o->d_ub, A, H, dU, sol_stride, add_noise,
                                                            ctx->seed, o->d_act_ctr);
  BB_LAUNCH_CHECK(ctx);
  // predict_next_state + evaluate_next_reward on the A executed actions (optimizer_base.py:91-94)
  StepIO io{o->d_state, act_out, nullptr, next_out, rew_out, nullptr, A, 3};
  if (int rc = launch_step_simt(ctx, io, st)) return rc;
  act_ctr_bump_kernel<<<1, 1, 0, st>>>(o->d_act_ctr); BB_LAUNCH_CHECK(ctx);
  o->act_call++;
  o->began = false;
  (void)dS;
  return BBMPC_OK;
}

// The kernels of one act() on the handle's own buffers (state in d_state; results in d_action / d_next / d_reward).
static int opt_call_body(bbmpc_opt* o, int time_step, int add_noise, void* stream) {
  if (int rc = bbmpc_opt_begin(o, o->d_state, time_step, stream)) return rc;
  const int n = bbmpc_opt_num_iterations(o);
  for (int it = 0; it < n; ++it) {
    if (o->p2p_on) {
      // publish this rank's message in its exchange buffer, pull the peers' over NVLink, merge locally
      cudaStream_t st = static_cast<cudaStream_t>(stream);
      const int nf = partial_floats(o);
      const uint32_t seq = ++o->p2p_seq;
      const int parity = static_cast<int>(seq & 1u);
      const int stride4 = (nf + 3) & ~3;     // parity copies 16-byte aligned for any message length
      if (int rc = bbmpc_opt_iter_local(o, it, o->p2p_buf + static_cast<size_t>(parity) * stride4, stream)) return rc;
      const size_t flag_off = 2 * static_cast<size_t>(stride4);
      p2p_publish_kernel<<<1, 1, 0, st>>>(reinterpret_cast<uint32_t*>(o->p2p_buf + flag_off), seq); BB_LAUNCH_CHECK(o->ctx);
      p2p_gather_kernel<<<o->world, 1024, 0, st>>>(o->d_peer, o->d_gather, nf, stride4, flag_off, seq, parity); BB_LAUNCH_CHECK(o->ctx);
      if (int rc = bbmpc_opt_iter_merge(o, it, o->d_gather, o->world, stream)) return rc;
      continue;
    }
    o->cem_fuse = !getenv("BBMPC_NO_CEM_FUSE");
    int rc = bbmpc_opt_iter_local(o, it, nullptr, stream);
    o->cem_fuse = false;
    if (rc) return rc;
    if ((rc = bbmpc_opt_iter_merge(o, it, nullptr, 1, stream))) return rc;
  }
  return bbmpc_opt_finish(o, add_noise, o->d_action, o->d_next, o->d_reward, stream);
}


// optimizers/optimizer_base.py:55-56: the reference runs one tf.function graph per act().  Here the ~25 kernels of an act()
// are captured once (after two eager calls that size every scratch buffer) into a CUDA graph over the handle's own buffers
// and replayed: one launch per act(), no launch gaps between the small kernels.  The Philox act-call counter lives on the
// device (bumped by the last kernel), so a replay draws fresh samples.  Not captured: sharded handles (their sequence
// numbers are launch arguments), calls with a sample trace / draw injection / rollout profiling, BBMPC_NO_GRAPH=1.
int bbmpc_opt_call(bbmpc_opt* o, const float* state, int time_step, int add_noise, float* action, float* next_state,
                   float* reward, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (o->world != 1 && !o->p2p_on)
    return opt_fail(o, BBMPC_ESTATE, "bbmpc_opt_call on a sharded optimizer without a peer-memory exchange: use begin/iter_local/iter_merge/finish");
  if (!state) return opt_fail(o, BBMPC_EINVAL, "state is NULL");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int A = o->cfg.num_agents, dS = o->cfg.dS, dU = o->cfg.dU;
  if (state != o->d_state)
    BB_CUDA(ctx, cudaMemcpyAsync(o->d_state, state, A * dS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const bool no_graph = getenv("BBMPC_NO_GRAPH") != nullptr;
  const bool graphable = !no_graph && o->world == 1 && !o->trace && !o->inject && !ctx->prof_on && ctx->model.set && ctx->reward_id &&
                         o->cfg.kind != BBMPC_OPT_CMAES;   // (cuSOLVER's syevd is not capturable)
  int rc = BBMPC_OK;
  if (!graphable) {
    rc = opt_call_body(o, time_step, add_noise, stream);
  } else {
    if (o->graph_exec && (o->graph_epoch != ctx->epoch || o->graph_noise != add_noise)) opt_graph_drop(o);
    if (o->graph_exec) {
      BB_CUDA(ctx, cudaGraphLaunch(o->graph_exec, st));
      o->act_call++;
      ctx->launches += o->graph_launches;
    } else if (o->graph_warm < 2) {
      ++o->graph_warm;                      // eager: scratch buffers are (re)allocated on the first calls
      rc = opt_call_body(o, time_step, add_noise, stream);
    } else {
      // Capture on a stream of our own (the caller's may be the legacy default stream, which cannot be captured); the
      // instantiated graph is launched into the caller's stream.
      const uint64_t before = ctx->launches;
      const uint32_t act_before = o->act_call;
      cudaGraph_t graph = nullptr;
      bool ok = true;
      if (!o->graph_stream) ok = cudaStreamCreateWithFlags(&o->graph_stream, cudaStreamNonBlocking) == cudaSuccess;
      ok = ok && cudaStreamBeginCapture(o->graph_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        rc = opt_call_body(o, time_step, add_noise, o->graph_stream);
        const cudaError_t ce = cudaStreamEndCapture(o->graph_stream, &graph);
        ok = rc == BBMPC_OK && ce == cudaSuccess && graph != nullptr;
      }
      cudaGraphExec_t exec = nullptr;
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      o->act_call = act_before;               // the captured body advanced the host mirrors without running
      o->began = false;
      if (ok) {
        o->graph_exec = exec; o->graph_epoch = ctx->epoch; o->graph_noise = add_noise;
        o->graph_launches = ctx->launches - before;
        ctx->launches = before;
        BB_CUDA(ctx, cudaGraphLaunch(o->graph_exec, st));
        o->act_call++;
        ctx->launches += o->graph_launches;
        rc = BBMPC_OK;
      } else {
        cudaGetLastError();
        ctx->launches = before;
        o->graph_warm = -1000000;             // this configuration cannot be captured: stay eager
        rc = opt_call_body(o, time_step, add_noise, stream);
      }
    }
  }
  if (rc != BBMPC_OK) return rc;
  if (action && action != o->d_action) BB_CUDA(ctx, cudaMemcpyAsync(action, o->d_action, A * dU * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (next_state && next_state != o->d_next) BB_CUDA(ctx, cudaMemcpyAsync(next_state, o->d_next, A * dS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (reward && reward != o->d_reward) BB_CUDA(ctx, cudaMemcpyAsync(reward, o->d_reward, A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return BBMPC_OK;
}

int bbmpc_opt_call_host(bbmpc_opt* o, const float* state_host, int time_step, int add_noise, float* action_host,
                        float* next_state_host, float* reward_host, void* stream) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!state_host || !action_host) return opt_fail(o, BBMPC_EINVAL, "NULL host buffer");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int A = o->cfg.num_agents, dS = o->cfg.dS, dU = o->cfg.dU;
  float* h_state = o->h_pinned; float* h_action = h_state + A * dS; float* h_next = h_action + A * dU; float* h_rew = h_next + A * dS;
  std::memcpy(h_state, state_host, A * dS * sizeof(float));
  BB_CUDA(ctx, cudaMemcpyAsync(o->d_state, h_state, A * dS * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = bbmpc_opt_call(o, o->d_state, time_step, add_noise, o->d_action, o->d_next, o->d_reward, stream)) return rc;
  BB_CUDA(ctx, cudaMemcpyAsync(h_action, o->d_action, A * dU * sizeof(float), cudaMemcpyDeviceToHost, st));
  BB_CUDA(ctx, cudaMemcpyAsync(h_next, o->d_next, A * dS * sizeof(float), cudaMemcpyDeviceToHost, st));
  BB_CUDA(ctx, cudaMemcpyAsync(h_rew, o->d_reward, A * sizeof(float), cudaMemcpyDeviceToHost, st));
  BB_CUDA(ctx, cudaStreamSynchronize(st));
  std::memcpy(action_host, h_action, A * dU * sizeof(float));
  if (next_state_host) std::memcpy(next_state_host, h_next, A * dS * sizeof(float));
  if (reward_host) std::memcpy(reward_host, h_rew, A * sizeof(float));
  return BBMPC_OK;
}

int bbmpc_opt_p2p_export(bbmpc_opt* o, void* handle_out_host, void** ptr_out_host) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!o->p2p_buf) {
    const size_t nf = (static_cast<size_t>(partial_floats(o)) + 3) & ~static_cast<size_t>(3);
    o->p2p_bytes = (2 * nf + 16) * sizeof(float);
    BB_CUDA(ctx, cudaMalloc(&o->p2p_buf, o->p2p_bytes));     // own allocation: an IPC handle exports the whole allocation
    o->owned.push_back(o->p2p_buf);
    BB_CUDA(ctx, cudaMemset(o->p2p_buf, 0, o->p2p_bytes));
  }
  if (handle_out_host) {
    cudaIpcMemHandle_t h;
    BB_CUDA(ctx, cudaIpcGetMemHandle(&h, o->p2p_buf));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle_out_host, &h, sizeof(h));
  }
  if (ptr_out_host) *ptr_out_host = o->p2p_buf;
  return BBMPC_OK;
}

int bbmpc_opt_p2p_connect(bbmpc_opt* o, const void* handles_host, void* const* ptrs_host) {
  if (!o) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  if (!o->p2p_buf) return opt_fail(o, BBMPC_ESTATE, "p2p_connect before p2p_export");
  if (!handles_host && !ptrs_host) return opt_fail(o, BBMPC_EINVAL, "p2p_connect needs IPC handles or device pointers");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<float*> peers(o->world, nullptr);
  for (int g = 0; g < o->world; ++g) {
    if (g == o->rank) { peers[g] = o->p2p_buf; continue; }
    if (ptrs_host && ptrs_host[g]) { peers[g] = static_cast<float*>(ptrs_host[g]); continue; }
    if (!handles_host) return opt_fail(o, BBMPC_EINVAL, "p2p_connect: no handle or pointer for a peer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles_host) + 64 * g, sizeof(h));
    void* mapped = nullptr;
    BB_CUDA(ctx, cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
    o->p2p_opened.push_back(mapped);
    peers[g] = static_cast<float*>(mapped);
  }
  if (!o->d_peer) {
    int rc = dalloc(o, &o->d_peer, o->world);
    if (rc == BBMPC_OK) rc = dalloc(o, &o->d_gather, static_cast<size_t>(o->world) * partial_floats(o));
    if (rc != BBMPC_OK) return rc;
  }
  BB_CUDA(ctx, cudaMemcpy(o->d_peer, peers.data(), o->world * sizeof(float*), cudaMemcpyHostToDevice));
  o->p2p_on = true;
  return BBMPC_OK;
}

int64_t bbmpc_opt_get_tensor(bbmpc_opt* o, const char* name, float* out, int64_t n_floats, void* stream) {
  if (!o || !name) return BBMPC_EINVAL;
  bbmpc_ctx* ctx = o->ctx;
  const int A = o->cfg.num_agents;
  const int64_t rows = static_cast<int64_t>(o->n_eval) * A, pop = static_cast<int64_t>(o->P_local) * A;
  const float* src = nullptr; int64_t n = 0;
  const std::string s(name);
  if (s == "mean" || s == "solution") { src = o->d_mean; n = o->AHU; }
  else if (s == "variance") { src = o->d_var; n = o->AHU; }
  else if (s == "previous_solution" || s == "current_parameters") { src = o->d_prev; n = o->AHU; }
  else if (s == "samples" || s == "x") { src = o->d_samples; n = rows * o->HU; }
  else if (s == "returns") { src = o->d_returns; n = rows; }
  else if (s == "partial") { src = o->d_partial; n = partial_floats(o); }
  else if (s == "v") { src = o->d_v; n = pop * o->HU; }
  else if (s == "pbest_x") { src = o->d_pbx; n = pop * o->HU; }
  else if (s == "pbest_r") { src = o->d_pbr; n = pop; }
  else if (s == "gbest_x") { src = o->d_gbx; n = o->AHU; }
  else if (s == "gbest_r") { src = o->d_gbr; n = A; }
  else if (s == "pso_r") { src = o->d_pso_r; n = o->d_pso_r ? 128 : 0; }   // (r1, r2) of the iterations of the last act() call
  if (o->cfg.kind == BBMPC_OPT_CMAES) {   // tf.Variables of cma_es.py:95-117 (D as its diagonal)
    const int64_t N = o->AHU;
    if (s == "m") { src = o->d_mean; n = N; }
    else if (s == "sigma") { src = o->d_sigma; n = N; }
    else if (s == "C") { src = o->d_C; n = N * N; }
    else if (s == "B") { src = o->d_B; n = N * N; }
    else if (s == "D") { src = o->d_D; n = N; }
    else if (s == "p_sigma") { src = o->d_ps; n = N; }
    else if (s == "p_C") { src = o->d_pc; n = N; }
  }
  if (!src) return fail(ctx, BBMPC_EINVAL, "unknown or unavailable tensor '%s'", name);
  if (out && n_floats > 0) {
    const int64_t m = n_floats < n ? n_floats : n;
    cudaError_t e = cudaMemcpyAsync(out, src, m * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(ctx, BBMPC_ECUDA, "get_tensor copy: %s", cudaGetErrorString(e));
  }
  return n;
}

int bbmpc_opt_set_draw_injection(bbmpc_opt* o, const float* std_draws, int64_t n_floats) {
  if (!o) return BBMPC_EINVAL;
  if (o->cfg.kind == BBMPC_OPT_PSO) return opt_fail(o, BBMPC_EINVAL, "draw injection is not available for PSO");
  o->inject = std_draws; o->inject_floats = std_draws ? n_floats : 0; o->inject_iter = 0;
  return BBMPC_OK;
}

int bbmpc_opt_set_sample_trace(bbmpc_opt* o, float* trace, int64_t n_floats) {
  if (!o) return BBMPC_EINVAL;
  o->trace = trace; o->trace_floats = trace ? n_floats : 0;
  return BBMPC_OK;
}

}  // extern "C"
