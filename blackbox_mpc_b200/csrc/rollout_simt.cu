// rollout_simt.cu — fp32 CUDA-core path of the trajectory evaluator.
//
// Parity-grade restatement of trajectory_evaluators/deterministic.py:48-77 fused into one kernel:
// process_input -> dynamics_function -> process_output -> reward, H times, state on chip.
// Serves (a) BBMPC_PREC_FP32, (b) models the tensor-core path cannot hold (wide layers),
// (c) analytical dynamics (pendulum), (d) the A-row tail of OptimizerBase.__call__
// (optimizers/optimizer_base.py:91-94) and the stand-alone predict/reward/forward entry points.
//
// Layout: a CTA owns TR=32 trajectories (lane = trajectory), its 4 warps split the output
// features of each layer in blocks of 16.  Activations live in shared memory as [feature][row]
// (conflict-free for lane = row); weights are read through the read-only path with warp-uniform
// 128-bit loads from the zero-padded fp32 image W32[K][ldw].
#include "common.cuh"
#include "device_fns.cuh"

namespace bbmpc {

constexpr int TR = 32;   // trajectories per CTA
constexpr int NG = 4;    // warps = feature groups
constexpr int FT = 16;   // features per thread per pass

struct SimtParams {
  MlpDev mlp;
  NormDev norm;
  int dyn_id, reward_id, dS, dU;
  // rollout
  const float* states; const float* actions; float* returns; const float* penalty;
  int rows, A, H;
  float* traj;   // user reward: visited states [rows][H][dS] (else nullptr)
  // single step
  StepIO io;
};

struct SimtSmem {
  float *X, *bufA, *bufB, *S, *S2, *Y, *Act;
};

__device__ inline SimtSmem carve(float* base, int dS, int dU, int maxw) {
  SimtSmem s;
  s.X = base;                base += (dS + dU) * TR;
  s.bufA = base;             base += maxw * TR;
  s.bufB = base;             base += maxw * TR;
  s.S = base;                base += dS * TR;
  s.S2 = base;               base += dS * TR;
  s.Y = base;                base += dS * TR;
  s.Act = base;
  return s;
}
static size_t simt_smem_bytes(int dS, int dU, int maxw) {
  return static_cast<size_t>((dS + dU) + 2 * maxw + 3 * dS + dU) * TR * sizeof(float);
}

// One Dense layer for the CTA's 32 rows: out[f][lane] = act(sum_k in[k][lane] * W[k][f] + b[f]).
// When `accum` is set the (linear or activated) result is added to out instead (ensemble sum).
__device__ inline void dense_layer(const float* __restrict__ in, float* __restrict__ out,
                                   const float* __restrict__ W, const float* __restrict__ b, int K, int N,
                                   int ldw, int act, bool accum, int lane, int g) {
  for (int f0 = g * FT; f0 < ldw; f0 += NG * FT) {
    float acc[FT];
#pragma unroll
    for (int j = 0; j < FT; ++j) acc[j] = 0.0f;
    const float* wp = W + f0;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      const float a = in[k * TR + lane];
      const float4* w4 = reinterpret_cast<const float4*>(wp + static_cast<size_t>(k) * ldw);
#pragma unroll
      for (int q = 0; q < FT / 4; ++q) {
        const float4 w = __ldg(w4 + q);
        acc[4 * q + 0] = fmaf(a, w.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(a, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(a, w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(a, w.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < FT; ++j) {
      const int f = f0 + j;
      if (f < N) {
        const float v = act_exact(__fadd_rn(acc[j], __ldg(b + f)), act);
        float* o = out + f * TR + lane;
        *o = accum ? __fadd_rn(*o, v) : v;
      }
    }
  }
}

// Raw dynamics_function(x): X[dS+dU][TR] -> Y[dS][TR]  (ensemble: member-order sum, then / n).
template <int DYN>
__device__ inline void dynamics_raw(const SimtParams& p, const SimtSmem& sm, int lane, int g) {
  if (DYN == BBMPC_DYN_MLP) {
    const MlpDev& m = p.mlp;
    for (int mm = 0; mm < m.n_members; ++mm) {
      const float* in = sm.X;
      for (int l = 0; l < m.n_layers; ++l) {
        const LayerDev& L = m.layer[l];
        const bool last = (l == m.n_layers - 1);
        float* out = last ? sm.Y : ((l & 1) ? sm.bufB : sm.bufA);
        const float* W = m.w32 + L.w_off + mm * m.w_member_stride;
        const float* b = m.w32 + L.b_off + mm * m.w_member_stride;
        dense_layer(in, out, W, b, L.K, L.N, L.ldw, L.act, last && mm > 0, lane, g);
        __syncthreads();
        in = out;
      }
    }
    if (m.n_members > 1) {
      const float n = static_cast<float>(m.n_members);
      for (int i = g; i < p.dS; i += NG) sm.Y[i * TR + lane] = __fdiv_rn(sm.Y[i * TR + lane], n);
      __syncthreads();
    }
  } else {  // pendulum: X = [cos, sin, thdot, u]
    if (g == 0) {
      float s[3], a[1], dev[3];
      s[0] = sm.X[0 * TR + lane]; s[1] = sm.X[1 * TR + lane]; s[2] = sm.X[2 * TR + lane];
      a[0] = sm.X[3 * TR + lane];
      pendulum_deviation(s, a, dev);
      sm.Y[0 * TR + lane] = dev[0]; sm.Y[1 * TR + lane] = dev[1]; sm.Y[2 * TR + lane] = dev[2];
    }
    __syncthreads();
  }
}

// process_input (system_dynamics_handler.py:97-126): S, Act -> X.
__device__ inline void build_input(const SimtParams& p, const SimtSmem& sm, int lane, int g, bool norm_on) {
  for (int k = g; k < p.dS + p.dU; k += NG) {
    float v;
    if (k < p.dS) {
      v = sm.S[k * TR + lane];
      if (norm_on) v = __fdiv_rn(__fsub_rn(v, p.norm.mean_s[k]), p.norm.den_s[k]);
    } else {
      v = sm.Act[(k - p.dS) * TR + lane];
      if (norm_on) v = __fdiv_rn(__fsub_rn(v, p.norm.mean_a[k - p.dS]), p.norm.den_a[k - p.dS]);
    }
    sm.X[k * TR + lane] = v;
  }
}
// process_output (system_dynamics_handler.py:128-161) + utils/transforms.py:34: Y, S -> S2.
__device__ inline void build_output(const SimtParams& p, const SimtSmem& sm, int lane, int g, bool norm_on) {
  for (int i = g; i < p.dS; i += NG) {
    float d = sm.Y[i * TR + lane];
    if (norm_on) d = __fadd_rn(p.norm.mean_t[i], __fmul_rn(d, p.norm.den_t[i]));
    sm.S2[i * TR + lane] = __fadd_rn(d, sm.S[i * TR + lane]);
  }
}

__device__ inline float row_reward(const SimtParams& p, const SimtSmem& sm, int lane) {
  float s[MAX_DS], s2[MAX_DS], a[MAX_DU];
  for (int i = 0; i < p.dS; ++i) { s[i] = sm.S[i * TR + lane]; s2[i] = sm.S2[i * TR + lane]; }
  for (int i = 0; i < p.dU; ++i) a[i] = sm.Act[i * TR + lane];
  return reward_dispatch(p.reward_id, s, a, s2, p.dS, p.dU);
}

template <int DYN>
__global__ void __launch_bounds__(TR* NG) rollout_simt_kernel(const SimtParams p) {
  extern __shared__ float smem_f[];
  const SimtSmem sm = carve(smem_f, p.dS, p.dU, DYN == BBMPC_DYN_MLP ? p.mlp.max_width : 4);
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int row = blockIdx.x * TR + lane;
  const bool valid = row < p.rows;
  const int arow = valid ? row : 0;
  const bool norm_on = (DYN == BBMPC_DYN_MLP) && p.norm.enabled;

  for (int i = g; i < p.dS; i += NG) sm.S[i * TR + lane] = p.states[(arow % p.A) * p.dS + i];
  float ret = 0.0f;
  for (int t = 0; t < p.H; ++t) {
    for (int i = g; i < p.dU; i += NG)
      sm.Act[i * TR + lane] = p.actions[(static_cast<size_t>(arow) * p.H + t) * p.dU + i];
    __syncthreads();
    build_input(p, sm, lane, g, norm_on);
    __syncthreads();
    dynamics_raw<DYN>(p, sm, lane, g);
    build_output(p, sm, lane, g, norm_on);
    __syncthreads();
    if (p.traj && valid)
      for (int i = g; i < p.dS; i += NG) p.traj[(static_cast<size_t>(row) * p.H + t) * p.dS + i] = sm.S2[i * TR + lane];
    if (g == 0) ret = __fadd_rn(ret, row_reward(p, sm, lane));
    __syncthreads();
    for (int i = g; i < p.dS; i += NG) sm.S[i * TR + lane] = sm.S2[i * TR + lane];
  }
  if (g == 0 && valid) {
    float r = isnan(ret) ? -1e6f : ret;  // deterministic.py:75-77 (NaN only, not +-inf)
    if (p.penalty) r = __fsub_rn(r, p.penalty[row]);
    p.returns[row] = r;
  }
}

template <int DYN>
__global__ void __launch_bounds__(TR* NG) step_simt_kernel(const SimtParams p) {
  extern __shared__ float smem_f[];
  const SimtSmem sm = carve(smem_f, p.dS, p.dU, DYN == BBMPC_DYN_MLP ? p.mlp.max_width : 4);
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int row = blockIdx.x * TR + lane;
  const bool valid = row < p.io.B;
  const int arow = valid ? row : 0;
  const bool norm_on = (DYN == BBMPC_DYN_MLP) && p.norm.enabled;
  const int mode = p.io.mode;

  if (mode & 4) {  // raw dynamics_function(x)
    for (int k = g; k < p.dS + p.dU; k += NG) sm.X[k * TR + lane] = p.io.s[static_cast<size_t>(arow) * (p.dS + p.dU) + k];
    __syncthreads();
    dynamics_raw<DYN>(p, sm, lane, g);
    if (valid) for (int i = g; i < p.dS; i += NG) p.io.raw_out[static_cast<size_t>(row) * p.dS + i] = sm.Y[i * TR + lane];
    return;
  }
  for (int i = g; i < p.dS; i += NG) sm.S[i * TR + lane] = p.io.s[static_cast<size_t>(arow) * p.dS + i];
  for (int i = g; i < p.dU; i += NG) sm.Act[i * TR + lane] = p.io.a[static_cast<size_t>(arow) * p.dU + i];
  __syncthreads();
  if (mode & 1) {
    build_input(p, sm, lane, g, norm_on);
    __syncthreads();
    dynamics_raw<DYN>(p, sm, lane, g);
    build_output(p, sm, lane, g, norm_on);
    __syncthreads();
    if (valid && p.io.s2_out)
      for (int i = g; i < p.dS; i += NG) p.io.s2_out[static_cast<size_t>(row) * p.dS + i] = sm.S2[i * TR + lane];
  } else {
    for (int i = g; i < p.dS; i += NG) sm.S2[i * TR + lane] = p.io.s2_in[static_cast<size_t>(arow) * p.dS + i];
    __syncthreads();
  }
  if ((mode & 2) && g == 0 && valid) p.io.reward_out[row] = row_reward(p, sm, lane);
}

// Few-row variant of step_simt_kernel for the optimizer tail (optimizers/optimizer_base.py:91-94:
// A rows) and small predict/forward calls.  One CTA per (row, ensemble member); a layer's K axis is split
// over the warps and its features over the lanes (fp32 FMA, partial sums added in warp order).  The
// last CTA to finish a row (arrival counter) sums the members in member order and applies
// process_output + reward.
constexpr int STEP_THREADS = 1024;   // 32 warps: a 200-row layer is 7 weight rows per warp, all loads of a warp in flight at once
constexpr int STEP_MAX_ROWS = 64;
__global__ void __launch_bounds__(STEP_THREADS) step_mlp_kernel(const SimtParams p, float* __restrict__ scratch,
                                                               unsigned* __restrict__ counters) {
  extern __shared__ float smem_f[];
  __shared__ bool is_last;
  const MlpDev& m = p.mlp;
  const int b = blockIdx.x, mm = blockIdx.y, tid = threadIdx.x;
  const int dS = p.dS, dU = p.dU, mode = p.io.mode;
  const bool norm_on = p.norm.enabled;
  float* in = smem_f;
  float* out = smem_f + m.max_width;
  for (int k = tid; k < dS + dU; k += STEP_THREADS) {
    float v;
    if (mode & 4) {
      v = p.io.s[static_cast<size_t>(b) * (dS + dU) + k];
    } else if (k < dS) {
      v = p.io.s[static_cast<size_t>(b) * dS + k];
      if (norm_on) v = __fdiv_rn(__fsub_rn(v, p.norm.mean_s[k]), p.norm.den_s[k]);
    } else {
      v = p.io.a[static_cast<size_t>(b) * dU + (k - dS)];
      if (norm_on) v = __fdiv_rn(__fsub_rn(v, p.norm.mean_a[k - dS]), p.norm.den_a[k - dS]);
    }
    in[k] = v;
  }
  __syncthreads();
  // K is split over the warps (contiguous ranges), a lane owns output features lane, lane+32, ...: every
  // weight is read exactly once, coalesced, and all loads of a warp are independent (the former one-thread-
  // per-feature loop was a chain of ~K/8 dependent DRAM round trips per layer: 286 us for the C4 ensemble).
  // The warps' partial sums are added in warp order, then bias and activation.
  float* part = smem_f + 2 * m.max_width;   // [STEP_WARPS][max_width]
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int STEP_WARPS = STEP_THREADS / 32;
  for (int l = 0; l < m.n_layers; ++l) {
    const LayerDev& L = m.layer[l];
    const bool last = (l == m.n_layers - 1);
    const float* __restrict__ W = m.w32 + L.w_off + mm * m.w_member_stride;
    const float* __restrict__ bias = m.w32 + L.b_off + mm * m.w_member_stride;
    const int kc = (L.K + STEP_WARPS - 1) / STEP_WARPS;
    const int k0 = warp * kc, k1 = (k0 + kc < L.K) ? k0 + kc : L.K;
    for (int f0 = 0; f0 < L.N; f0 += 256) {          // 8 features per lane per sweep
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const float x = in[k];
        const float* __restrict__ wr = W + static_cast<size_t>(k) * L.ldw + f0 + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f0 + lane + 32 * j < L.N) acc[j] = fmaf(x, __ldg(wr + 32 * j), acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f0 + lane + 32 * j < L.N) part[warp * m.max_width + f0 + lane + 32 * j] = acc[j];
    }
    __syncthreads();
    for (int f = tid; f < L.N; f += STEP_THREADS) {
      float acc = part[f];
#pragma unroll
      for (int w2 = 1; w2 < STEP_WARPS; ++w2) acc = __fadd_rn(acc, part[w2 * m.max_width + f]);
      const float v = act_exact(__fadd_rn(acc, __ldg(bias + f)), L.act);
      if (last) scratch[(static_cast<size_t>(b) * m.n_members + mm) * dS + f] = v;
      else out[f] = v;
    }
    __syncthreads();
    float* t = in; in = out; out = t;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(&counters[b], 1u) == static_cast<unsigned>(m.n_members - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid == 0) counters[b] = 0;  // self-resetting for the next launch
  float* s2s = smem_f;            // [dS]
  for (int i = tid; i < dS; i += STEP_THREADS) {
    float y = __ldcg(scratch + (static_cast<size_t>(b) * m.n_members) * dS + i);
    for (int k = 1; k < m.n_members; ++k) y = __fadd_rn(y, __ldcg(scratch + (static_cast<size_t>(b) * m.n_members + k) * dS + i));
    if (m.n_members > 1) y = __fdiv_rn(y, static_cast<float>(m.n_members));
    if (mode & 4) {
      p.io.raw_out[static_cast<size_t>(b) * dS + i] = y;
    } else {
      const float d = norm_on ? __fadd_rn(p.norm.mean_t[i], __fmul_rn(y, p.norm.den_t[i])) : y;
      const float s2 = __fadd_rn(d, p.io.s[static_cast<size_t>(b) * dS + i]);
      s2s[i] = s2;
      if (p.io.s2_out) p.io.s2_out[static_cast<size_t>(b) * dS + i] = s2;
    }
  }
  __syncthreads();
  if ((mode & 2) && tid == 0) {
    float s[MAX_DS], s2[MAX_DS], a[MAX_DU];
    for (int i = 0; i < dS; ++i) { s[i] = p.io.s[static_cast<size_t>(b) * dS + i]; s2[i] = s2s[i]; }
    for (int i = 0; i < dU; ++i) a[i] = p.io.a[static_cast<size_t>(b) * dU + i];
    p.io.reward_out[b] = reward_dispatch(p.reward_id, s, a, s2, dS, dU);
  }
}

static int fill_params(bbmpc_ctx* ctx, SimtParams& p) {
  const ModelHost& m = ctx->model;
  p.mlp = m.mlp; p.norm = m.norm; p.dyn_id = m.dyn_id; p.reward_id = ctx->reward_id; p.dS = m.dS; p.dU = m.dU;
  if (m.dyn_id == BBMPC_DYN_MLP && m.mlp.max_width > 800)
    return fail(ctx, BBMPC_EINVAL, "fp32 path supports layer widths up to 800 (got %d)", m.mlp.max_width);
  return BBMPC_OK;
}

template <typename K>
static int set_smem(bbmpc_ctx* ctx, K kernel, size_t bytes) {
  BB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return BBMPC_OK;
}

int launch_rollout_simt(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                        const float* penalty, int rows, int A, int H, int, cudaStream_t st) {
  SimtParams p{};
  if (int rc = fill_params(ctx, p)) return rc;
  p.states = states; p.actions = actions; p.returns = returns; p.penalty = penalty;
  p.rows = rows; p.A = A; p.H = H; p.traj = ctx->traj_cur;
  const int grid = (rows + TR - 1) / TR;
  if (p.dyn_id == BBMPC_DYN_MLP) {
    const size_t sb = simt_smem_bytes(p.dS, p.dU, p.mlp.max_width);
    if (int rc = set_smem(ctx, rollout_simt_kernel<BBMPC_DYN_MLP>, sb)) return rc;
    rollout_simt_kernel<BBMPC_DYN_MLP><<<grid, TR * NG, sb, st>>>(p);
  } else {
    const size_t sb = simt_smem_bytes(p.dS, p.dU, 4);
    rollout_simt_kernel<BBMPC_DYN_PENDULUM><<<grid, TR * NG, sb, st>>>(p);
  }
  BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

int launch_step_simt(bbmpc_ctx* ctx, const StepIO& io_in, cudaStream_t st) {
  StepIO io = io_in;
  if (ctx->reward_id == BBMPC_REWARD_USER && (io.mode & 2)) {
    // user reward (user_reward.cu): predict with the built-in path, then the JIT-compiled row kernel on (s, a, s')
    const float* s2 = io.s2_in;
    if (io.mode & 1) {
      if (!io.s2_out) return fail(ctx, BBMPC_EINVAL, "user reward after predict needs a next-state buffer");
      StepIO pred = io; pred.mode = 1; pred.reward_out = nullptr;
      if (int rc = launch_step_simt(ctx, pred, st)) return rc;
      s2 = io.s2_out;
    }
    return launch_user_reward_rows(ctx, io.s, io.a, s2, io.reward_out, io.B, st);
  }
  SimtParams p{};
  if (int rc = fill_params(ctx, p)) return rc;
  p.io = io;
  const int grid = (io.B + TR - 1) / TR;
  if (p.dyn_id == BBMPC_DYN_MLP && ctx->model.set && io.B <= STEP_MAX_ROWS && (io.mode & 5)) {
    if (!ctx->step_scratch) {
      BB_CUDA(ctx, cudaMalloc(&ctx->step_scratch, sizeof(float) * STEP_MAX_ROWS * MAX_MEMBERS * MAX_DS));
      BB_CUDA(ctx, cudaMalloc(&ctx->step_counters, sizeof(unsigned) * STEP_MAX_ROWS));
      BB_CUDA(ctx, cudaMemset(ctx->step_counters, 0, sizeof(unsigned) * STEP_MAX_ROWS));
    }
    const size_t sb = sizeof(float) * (2 + STEP_THREADS / 32) * static_cast<size_t>(p.mlp.max_width > MAX_DS ? p.mlp.max_width : MAX_DS);
    if (int rc = set_smem(ctx, step_mlp_kernel, sb)) return rc;
    step_mlp_kernel<<<dim3(io.B, p.mlp.n_members), STEP_THREADS, sb, st>>>(p, ctx->step_scratch, ctx->step_counters);
    BB_LAUNCH_CHECK(ctx);
    return BBMPC_OK;
  }
  if (p.dyn_id == BBMPC_DYN_MLP && ctx->model.set) {
    const size_t sb = simt_smem_bytes(p.dS, p.dU, p.mlp.max_width);
    if (int rc = set_smem(ctx, step_simt_kernel<BBMPC_DYN_MLP>, sb)) return rc;
    step_simt_kernel<BBMPC_DYN_MLP><<<grid, TR * NG, sb, st>>>(p);
  } else {
    const size_t sb = simt_smem_bytes(p.dS, p.dU, 4);
    step_simt_kernel<BBMPC_DYN_PENDULUM><<<grid, TR * NG, sb, st>>>(p);
  }
  BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

}  // namespace bbmpc
