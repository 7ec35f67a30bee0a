// rollout_tc.cu — tensor-core (tcgen05 / TMEM) path of the trajectory evaluator.
//
// One persistent CTA owns a tile of 128 trajectories for the whole horizon
// (trajectory_evaluators/deterministic.py:48-77 fused: process_input -> Dense chain ->
// process_output -> reward, H times).  Per Dense layer the CTA issues
//      D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, bf16) * W[K x N] (SMEM, bf16, K-major)
// as tcgen05.mma M=128 / N=Npad / K=16 instructions.  fp32-grade accuracy comes from splitting
// both operands into bf16 hi + lo and issuing three MMAs per K-chunk (hi*hi, lo*hi, hi*lo;
// the dropped lo*lo term is O(2^-18)); BBMPC_PREC_BF16 issues only hi*hi.
//
//   warp 0      producer: streams weight chunks L2 -> SMEM ring with cp.async.bulk + mbarrier tx
//   warp 1      MMA issuer (one elected lane): tcgen05.mma A-from-TMEM, tcgen05.commit
//   warps 2..9  epilogue: thread <-> (trajectory, column half).  tcgen05.ld accumulator ->
//               activation -> bf16 hi/lo split -> tcgen05.st as the next layer's A operand.
//               Trajectory state, action and return live in registers for all H steps.
//
// Activations therefore never touch shared or global memory; HBM traffic is the action read
// (H*dU floats per trajectory) and the 4-byte return.  Bias is folded into the GEMM: the A
// operand carries three ones-columns that meet three bf16 bias rows of the weight image.
// The ensemble mean is folded too: every member's output layer accumulates into the same TMEM
// tile (accumulate flag), the epilogue divides by n_members.
#include "common.cuh"
#include "device_fns.cuh"
#include "tc05.cuh"

namespace bbmpc {
using namespace tc05;

constexpr int TC_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int TC_MAX_STAGES = 24;
constexpr int TILE_ROWS = 128;

struct TcParams {
  MlpDev mlp;
  NormDev norm;
  int reward_id, dS, dU;
  const float* states; const float* actions; float* returns; const float* penalty;
  int rows, A, H, n_tiles, passes;
  int stage_bytes, n_stages;
  int col_dmain, col_dout, col_x, col_a;  // TMEM column map (hi at col, lo at col + *_half)
  int x_half, a_half;
};

struct TcSmemLayout {
  uint32_t stages, table, bars, tmem_slot, stats, total;
};
__host__ __device__ inline TcSmemLayout tc_layout(int stage_bytes, int n_stages, int chunks_per_step) {
  TcSmemLayout L;
  uint32_t off = 0;
  L.stages = off; off += static_cast<uint32_t>(stage_bytes) * n_stages;
  L.table = off;  off += static_cast<uint32_t>(chunks_per_step) * 8;
  off = (off + 7u) & ~7u;
  L.bars = off;   off += (2 * TC_MAX_STAGES + 2) * 8;
  L.tmem_slot = off; off += 16;
  L.stats = off;  off += (4 * MAX_DS + 2 * MAX_DU) * 4;
  L.total = off;
  return L;
}

template <int DS_T, int DU_T>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const TcSmemLayout lay = tc_layout(p.stage_bytes, p.n_stages, p.mlp.chunks_per_step);
  const uint32_t smem_base = smem_u32(smem);
  const uint2* table = reinterpret_cast<const uint2*>(smem + lay.table);
  const uint32_t bar_full = smem_base + lay.bars;
  const uint32_t bar_empty = bar_full + TC_MAX_STAGES * 8;
  const uint32_t bar_a = bar_empty + TC_MAX_STAGES * 8;   // epilogue -> MMA: A operand written
  const uint32_t bar_d = bar_a + 8;                       // MMA -> epilogue: accumulator complete
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + lay.tmem_slot);
  float* st_mean_s = reinterpret_cast<float*>(smem + lay.stats);
  float* st_den_s = st_mean_s + MAX_DS;
  float* st_mean_t = st_den_s + MAX_DS;
  float* st_den_t = st_mean_t + MAX_DS;
  float* st_mean_a = st_den_t + MAX_DS;
  float* st_den_a = st_mean_a + MAX_DU;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const MlpDev& M = p.mlp;
  const int nL = M.n_layers;
  const bool norm_on = p.norm.enabled != 0;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < M.chunks_per_step; i += TC_THREADS)
    reinterpret_cast<uint2*>(smem + lay.table)[i] = M.chunk_table[i];
  for (int i = tid; i < MAX_DS; i += TC_THREADS) {
    const bool in = norm_on && i < p.dS;
    st_mean_s[i] = in ? p.norm.mean_s[i] : 0.0f;
    st_den_s[i] = in ? p.norm.den_s[i] : 1.0f;
    st_mean_t[i] = in ? p.norm.mean_t[i] : 0.0f;
    st_den_t[i] = in ? p.norm.den_t[i] : 1.0f;
  }
  for (int i = tid; i < MAX_DU; i += TC_THREADS) {
    const bool in = norm_on && i < p.dU;
    st_mean_a[i] = in ? p.norm.mean_a[i] : 0.0f;
    st_den_a[i] = in ? p.norm.den_a[i] : 1.0f;
  }
  if (tid == 0) {
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_a, EPI_THREADS);
    mbar_init(bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_base + lay.tmem_slot, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = (p.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ============================================================ producer
    if (lane == 0) {
      const long long total = static_cast<long long>(my_tiles) * p.H * M.chunks_per_step;
      int stage = 0, ci = 0;
      uint32_t phase = 0;
      for (long long i = 0; i < total; ++i) {
        const uint2 e = table[ci];
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * stage, e.y);
        bulk_g2s(smem_base + lay.stages + stage * p.stage_bytes, M.wimg + e.x, e.y, bar_full + 8 * stage);
        if (++ci == M.chunks_per_step) ci = 0;
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, pa = 0;
      for (int tile = 0; tile < my_tiles; ++tile) {
        for (int t = 0; t < p.H; ++t) {
          for (int mm = 0; mm < M.n_members; ++mm) {
            for (int l = 0; l < nL; ++l) {
              const LayerDev& L = M.layer[l];
              const bool last = (l == nL - 1);
              if (l > 0 || mm == 0) { mbar_wait(bar_a, pa); pa ^= 1; }
              fence_after_sync();
              const uint32_t a_hi = tmem_base + (l == 0 ? p.col_x : p.col_a);
              const uint32_t a_lo = a_hi + (l == 0 ? p.x_half : p.a_half);
              const uint32_t d = tmem_base + (last ? p.col_dout : p.col_dmain);
              const uint32_t idesc = idesc_bf16_f32(TILE_ROWS, L.Npad);
              const uint32_t kstep = static_cast<uint32_t>(L.Npad) * 16;  // bytes between 8-wide K slabs
              const int nchunks = L.Kpad >> 4;
              for (int c = 0; c < nchunks; ++c) {
                mbar_wait(bar_full + 8 * stage, phase);
                fence_after_sync();
                const uint32_t sa = smem_base + lay.stages + stage * p.stage_bytes;
                const uint64_t bhi = smem_desc_kmajor_noswz(sa, kstep, 128);
                const uint32_t acc0 = (c > 0 || (last && mm > 0)) ? 1u : 0u;
                mma_ts(d, a_hi + 8 * c, bhi, idesc, acc0);
                if (p.passes == 3) {
                  const uint64_t blo = smem_desc_kmajor_noswz(sa + 2 * kstep, kstep, 128);
                  mma_ts(d, a_lo + 8 * c, bhi, idesc, 1u);
                  mma_ts(d, a_hi + 8 * c, blo, idesc, 1u);
                }
                mma_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
                if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
              }
              if (!last || mm == M.n_members - 1) mma_commit(bar_d);
            }
          }
        }
      }
    }
  } else {
    // ============================================================ epilogue warps
    const int e = warp - 2;                 // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = e >> 2;                // which interleaved half of the 16-column chunks
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    uint32_t pd = 0;
    const float inv_pad = 0.0f;
    (void)inv_pad;
    const int kp0_chunks = M.layer[0].Kpad >> 4;
    const float n_members_f = static_cast<float>(M.n_members);

    for (int tile = 0; tile < my_tiles; ++tile) {
      const int row = (static_cast<int>(blockIdx.x) + tile * static_cast<int>(gridDim.x)) * TILE_ROWS + row_in_tile;
      const bool valid = row < p.rows;
      const int arow = valid ? row : 0;
      float s[DS_T];
#pragma unroll
      for (int i = 0; i < DS_T; ++i) s[i] = (i < p.dS) ? p.states[(arow % p.A) * p.dS + i] : 0.0f;
      float ret = 0.0f;
      const float* arow_ptr = p.actions + static_cast<size_t>(arow) * p.H * p.dU;
      float a_next[DU_T];
#pragma unroll
      for (int i = 0; i < DU_T; ++i) a_next[i] = (i < p.dU && p.H > 0) ? arow_ptr[i] : 0.0f;

      for (int t = 0; t < p.H; ++t) {
        float a[DU_T];
#pragma unroll
        for (int i = 0; i < DU_T; ++i) a[i] = a_next[i];
        if (t + 1 < p.H) {
#pragma unroll
          for (int i = 0; i < DU_T; ++i) if (i < p.dU) a_next[i] = arow_ptr[(t + 1) * p.dU + i];
        }
        // ---- process_input: X = [norm(a) (DU_T slots) | norm(s) | 1 1 1 | 0...] -> TMEM
        constexpr int KP0_T = (DU_T + DS_T + BIAS_COLS + 15) / 16;
#pragma unroll
        for (int c = 0; c < KP0_T; ++c) {
          if (c < kp0_chunks && (c & 1) == half) {
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = 16 * c + j;
              float v;
              if (k < DU_T) {
                v = (k < p.dU) ? __fdiv_rn(__fsub_rn(a[k < DU_T ? k : 0], st_mean_a[k < DU_T ? k : 0]), st_den_a[k < DU_T ? k : 0]) : 0.0f;
              } else {
                const int i = k - DU_T;
                const float one_or_zero = (i >= p.dS && i < p.dS + BIAS_COLS) ? 1.0f : 0.0f;
                if (i < DS_T)
                  v = (i < p.dS) ? __fdiv_rn(__fsub_rn(s[i < DS_T ? i : 0], st_mean_s[i < DS_T ? i : 0]), st_den_s[i < DS_T ? i : 0]) : one_or_zero;
                else
                  v = one_or_zero;
              }
              x[j] = v;
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split_bf16x2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
            tmem_st8(tmem_base + lane_base + p.col_x + 8 * c, hi);
            if (p.passes == 3) tmem_st8(tmem_base + lane_base + p.col_x + p.x_half + 8 * c, lo);
          }
        }
        wait_st();
        fence_before_sync();
        mbar_arrive(bar_a);

        // ---- hidden layers of every member: D_main -> activation -> A
        for (int mm = 0; mm < M.n_members; ++mm) {
          for (int l = 0; l + 1 < nL; ++l) {
            const LayerDev& L = M.layer[l];
            const int n_a_chunks = M.layer[l + 1].Kpad >> 4;
            mbar_wait(bar_d, pd); pd ^= 1;
            fence_after_sync();
            for (int c = half; c < n_a_chunks; c += 2) {
              uint32_t r[16];
              if (16 * c < L.Npad) {
                tmem_ld16(tmem_base + lane_base + p.col_dmain + 16 * c, r);
                wait_ld();
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
              }
              float v[16];
              if (16 * c + 16 <= L.N) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = act_fast(__uint_as_float(r[j]), L.act);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int f = 16 * c + j;
                  v[j] = f < L.N ? act_fast(__uint_as_float(r[j]), L.act)
                                 : (f < L.N + BIAS_COLS ? 1.0f : 0.0f);
                }
              }
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
              tmem_st8(tmem_base + lane_base + p.col_a + 8 * c, hi);
              if (p.passes == 3) tmem_st8(tmem_base + lane_base + p.col_a + p.a_half + 8 * c, lo);
            }
            wait_st();
            fence_before_sync();
            mbar_arrive(bar_a);
          }
        }

        // ---- output layer (sum over members) -> process_output -> reward
        mbar_wait(bar_d, pd); pd ^= 1;
        fence_after_sync();
        float s2[DS_T];
        {
          const LayerDev& Lo = M.layer[nL - 1];
#pragma unroll
          for (int c = 0; c < DS_T / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem_base + lane_base + p.col_dout + 16 * c, r);
            wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int i = 16 * c + j;
              float y = act_fast(__uint_as_float(r[j]), M.n_members == 1 ? Lo.act : BBMPC_ACT_NONE);
              if (M.n_members > 1) y = __fdiv_rn(y, n_members_f);
              const float dlt = norm_on ? __fadd_rn(st_mean_t[i], __fmul_rn(y, st_den_t[i])) : y;
              s2[i] = (i < p.dS) ? __fadd_rn(dlt, s[i]) : 0.0f;
            }
          }
        }
        // all of this thread's TMEM reads of D_out are complete (wait_ld) before the next arrive.
        float r_t = 0.0f;
        if (p.reward_id == BBMPC_REWARD_HALFCHEETAH) {
          if constexpr (DS_T >= 18) {
            if (s[5] >= 0.2f) r_t += -10.0f;
            if (s[6] >= 0.0f) r_t += -10.0f;
            if (s[7] >= 0.0f) r_t += -10.0f;
            r_t = __fadd_rn(r_t, __fdiv_rn(__fsub_rn(s2[17], s[17]), 0.01f));
            float ss = 0.0f;
#pragma unroll
            for (int i = 0; i < DU_T; ++i) if (i < p.dU) ss = __fadd_rn(ss, __fmul_rn(a[i], a[i]));
            r_t = __fsub_rn(r_t, __fmul_rn(0.0f, ss));
          }
        } else if (p.reward_id == BBMPC_REWARD_PENDULUM || p.reward_id == BBMPC_REWARD_PENDULUM_GYM) {
          const float pi = 3.14159274101257324f, two_pi = 6.28318548202514648f;
          const float ang = __fsub_rn(floormod_f(__fadd_rn(atan2f(s[1], s[0]), pi), two_pi), pi);
          float ss = 0.0f;
          if (p.reward_id == BBMPC_REWARD_PENDULUM) {  // `actions` parameter receives next_state
#pragma unroll
            for (int i = 0; i < DS_T; ++i) if (i < p.dS) ss = __fadd_rn(ss, __fmul_rn(s2[i], s2[i]));
          } else {
#pragma unroll
            for (int i = 0; i < DU_T; ++i) if (i < p.dU) ss = __fadd_rn(ss, __fmul_rn(a[i], a[i]));
          }
          const float sc = __fadd_rn(__fmul_rn(ang, ang), __fmul_rn(0.1f, __fmul_rn(s[2], s[2])));
          r_t = __fsub_rn(-sc, __fmul_rn(0.001f, ss));
        }
        ret = __fadd_rn(ret, r_t);
#pragma unroll
        for (int i = 0; i < DS_T; ++i) s[i] = s2[i];
      }
      if (half == 0 && valid) {
        float r = isnan(ret) ? -1e6f : ret;  // deterministic.py:75-77
        if (p.penalty) r = __fsub_rn(r, p.penalty[row]);
        p.returns[row] = r;
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------- host launcher
template <int DS_T, int DU_T>
static int launch_t(bbmpc_ctx* ctx, const TcParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
  BB_CUDA(ctx, cudaFuncSetAttribute(rollout_tc_kernel<DS_T, DU_T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem_bytes)));
  rollout_tc_kernel<DS_T, DU_T><<<grid, TC_THREADS, smem_bytes, st>>>(p);
  BB_LAUNCH_CHECK(ctx);
  return BBMPC_OK;
}

int tc_du_slots(int dU) { return dU <= 8 ? 8 : 16; }

int launch_rollout_tc(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                      const float* penalty, int rows, int A, int H, int passes, cudaStream_t st) {
  const ModelHost& m = ctx->model;
  TcParams p{};
  p.mlp = m.mlp; p.norm = m.norm; p.reward_id = ctx->reward_id; p.dS = m.dS; p.dU = m.dU;
  p.states = states; p.actions = actions; p.returns = returns; p.penalty = penalty;
  p.rows = rows; p.A = A; p.H = H; p.passes = passes;
  p.n_tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  const int nL = m.mlp.n_layers;
  int d_main = 0, a_cols = 0, stage = 0;
  for (int l = 0; l < nL; ++l) {
    const LayerDev& L = m.mlp.layer[l];
    if (l < nL - 1 && L.Npad > d_main) d_main = L.Npad;
    if (l > 0 && L.Kpad > a_cols) a_cols = L.Kpad;
    if (L.chunk_bytes > stage) stage = L.chunk_bytes;
  }
  p.col_dmain = 0;
  p.col_dout = d_main;
  p.col_x = p.col_dout + m.mlp.layer[nL - 1].Npad;
  p.x_half = m.mlp.layer[0].Kpad / 2;
  p.col_a = p.col_x + m.mlp.layer[0].Kpad;
  p.a_half = a_cols / 2;
  p.stage_bytes = stage;
  const size_t budget = 227 * 1024;
  int n_stages = TC_MAX_STAGES;
  while (n_stages > 2 && tc_layout(stage, n_stages, m.mlp.chunks_per_step).total > budget) --n_stages;
  if (tc_layout(stage, n_stages, m.mlp.chunks_per_step).total > budget)
    return fail(ctx, BBMPC_EINVAL, "tensor-core path: shared memory budget exceeded");
  p.n_stages = n_stages;
  const size_t smem_bytes = tc_layout(stage, n_stages, m.mlp.chunks_per_step).total;
  const int grid = p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count;
  if (m.dS <= 16 && m.dU <= 8) return launch_t<16, 8>(ctx, p, grid, smem_bytes, st);
  if (m.dU <= 8) return launch_t<32, 8>(ctx, p, grid, smem_bytes, st);
  return launch_t<32, 16>(ctx, p, grid, smem_bytes, st);
}

}  // namespace bbmpc
