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
// Layer pipeline.  The accumulator of layer l is converted IN PLACE, 16 columns at a time, into
// the A operand of layer l+1 (columns [16c,16c+8) = hi, [16c+8,16c+16) = lo of K-chunk c), and
// every converted chunk is published on its own mbarrier.  The MMA issuer consumes chunks as
// they appear and accumulates layer l+1 into the other of two TMEM buffers, so the tensor pipe
// works on layer l+1 while the epilogue warps are still converting layer l.  For an ensemble the
// first layer of member m+1 (whose input X is already in TMEM) is issued ahead of the output
// layer of member m, which keeps the epilogue fed across the member boundary.
//
// Activations never touch shared or global memory; HBM traffic is the action read
// (H*dU floats per trajectory) and the 4-byte return.  Bias is folded into the GEMM: the A
// operand carries three ones-columns that meet three bf16 bias rows of the weight image.
// The ensemble mean is folded too: every member's output layer accumulates into the same TMEM
// tile (accumulate flag), the epilogue divides by n_members.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "device_fns.cuh"
#include "tc05.cuh"

namespace bbmpc {
using namespace tc05;

constexpr int TC_THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr int TC_MAX_STAGES = 24;
constexpr int TC_MAX_CHUNKS = 16;   // K chunks of one layer (Kpad <= 256)
constexpr int TILE_ROWS = 128;

struct TcParams {
  MlpDev mlp;
  NormDev norm;
  int reward_id, dS, dU;
  const float* states; const float* actions; float* returns; const float* penalty;
  int rows, A, H, n_tiles, passes;
  int stage_bytes, n_stages;
  int col_buf0, col_buf1, col_x, col_dout;  // TMEM column map
  uint32_t* dbg;  // host-mapped watchdog record (BBMPC_DEBUG=1), else nullptr
};

struct TcSmemLayout {
  uint32_t stages, table, jobs, bars, tmem_slot, stats, total;
};
constexpr int TC_NUM_BARS = 2 * TC_MAX_STAGES + 2 * TC_MAX_CHUNKS + 4;
__host__ __device__ inline TcSmemLayout tc_layout(int stage_bytes, int n_stages, int chunks_per_step, int jobs_per_step) {
  TcSmemLayout L;
  uint32_t off = 0;
  L.stages = off; off += static_cast<uint32_t>(stage_bytes) * n_stages;
  L.table = off;  off += static_cast<uint32_t>(chunks_per_step) * 8;
  off = (off + 15u) & ~15u;
  L.jobs = off;   off += static_cast<uint32_t>(jobs_per_step) * sizeof(TcJob);
  L.bars = off;   off += TC_NUM_BARS * 8;
  L.tmem_slot = off; off += 16;
  L.stats = off;  off += (4 * MAX_DS + 2 * MAX_DU) * 4;
  L.total = off;
  return L;
}

// One hidden-layer epilogue of one thread: accumulator chunks c = half, half+2, ... of its TMEM
// lane -> activation -> bf16 hi/lo -> written back over the same 16 columns -> chunk published.
template <int ACT>
__device__ __forceinline__ void epi_hidden(uint32_t taddr, int Npad, int N, int n_a_chunks, int half, int passes,
                                           uint32_t bar_achunk, int lane) {
  for (int c = half; c < n_a_chunks; c += 2) {
    float v[16];
    if (16 * c < Npad) {
      uint32_t r[16];
      tmem_ld16(taddr + 16 * c, r);
      wait_ld();
      if (16 * c + 16 <= N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = act_fast_t<ACT>(__uint_as_float(r[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int f = 16 * c + j;
          v[j] = f < N ? act_fast_t<ACT>(__uint_as_float(r[j])) : (f < N + BIAS_COLS ? 1.0f : 0.0f);
        }
      }
    } else {  // pure padding chunk of the next layer's K axis: ones-columns / zeros
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int f = 16 * c + j;
        v[j] = (f >= N && f < N + BIAS_COLS) ? 1.0f : 0.0f;
      }
    }
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    tmem_st8(taddr + 16 * c, hi);
    if (passes == 3) tmem_st8(taddr + 16 * c + 8, lo);
    wait_st();
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_achunk + 8 * c);
  }
}

template <int DS_T, int DU_T>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const TcSmemLayout lay = tc_layout(p.stage_bytes, p.n_stages, p.mlp.chunks_per_step, p.mlp.jobs_per_step);
  const uint32_t smem_base = smem_u32(smem);
  const uint2* table = reinterpret_cast<const uint2*>(smem + lay.table);
  const uint32_t bar_full = smem_base + lay.bars;
  const uint32_t bar_empty = bar_full + TC_MAX_STAGES * 8;
  // epilogue -> MMA: K-chunk c of the A operand written.  Two sets, alternating per hidden round: with
  // the early first-layer issue the epilogue may run ONE round ahead of the MMA issuer's consumption,
  // and an mbarrier must never complete two phases ahead of its waiter (parity aliasing).
  const uint32_t bar_achunk = bar_empty + TC_MAX_STAGES * 8;
  const uint32_t bar_x = bar_achunk + 2 * TC_MAX_CHUNKS * 8;  // epilogue -> MMA: layer-0 input of this step written
  const uint32_t bar_d0 = bar_x + 8;                          // MMA -> epilogue: first-layer accumulator complete
  const uint32_t bar_d = bar_d0 + 8;                          // MMA -> epilogue: accumulator of a layer l >= 1 complete
  const uint32_t bar_dout = bar_d + 8;                        // MMA -> epilogue: output accumulator complete
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
  const int nM = M.n_members;
  const bool norm_on = p.norm.enabled != 0;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < M.chunks_per_step; i += TC_THREADS)
    reinterpret_cast<uint2*>(smem + lay.table)[i] = M.chunk_table[i];
  for (int i = tid; i < M.jobs_per_step * static_cast<int>(sizeof(TcJob) / 16); i += TC_THREADS)
    reinterpret_cast<uint4*>(smem + lay.jobs)[i] = reinterpret_cast<const uint4*>(M.jobs)[i];
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
    for (int c = 0; c < 2 * TC_MAX_CHUNKS; ++c) mbar_init(bar_achunk + 8 * c, EPI_WARPS / 2);
    mbar_init(bar_x, EPI_WARPS);
    mbar_init(bar_d0, 1);
    mbar_init(bar_d, 1);
    mbar_init(bar_dout, 1);
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
        mbar_wait(bar_empty + 8 * stage, phase ^ 1, p.dbg, 0x6000000u | static_cast<uint32_t>(i));
        mbar_arrive_expect_tx(bar_full + 8 * stage, e.y);
        bulk_g2s(smem_base + lay.stages + stage * p.stage_bytes, M.wimg + e.x, e.y, bar_full + 8 * stage);
        if (++ci == M.chunks_per_step) ci = 0;
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    // The whole warp runs the (warp-uniform) control flow and the barrier waits; one elected lane
    // issues the tcgen05 instructions.  Per K-chunk the issue path is a handful of integer ops:
    // everything layer-specific comes pre-digested from the job table.
    const TcJob* jobs = reinterpret_cast<const TcJob*>(smem + lay.jobs);
    const int n_jobs = M.jobs_per_step;
    const bool three = (p.passes == 3);
    uint32_t stage = 0, phase = 0, px = 0, cph0 = 0, cph1 = 0, rj = 0;  // rj: hidden rounds consumed so far
    const uint32_t stages16 = (smem_base + lay.stages) >> 4, stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4;
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1, no swizzle
    for (int tile = 0; tile < my_tiles; ++tile) {
      for (int t = 0; t < p.H; ++t) {
        for (int j = 0; j < n_jobs; ++j) {
          const TcJob job = jobs[j];
          const uint32_t d = tmem_base + job.d_col, a = tmem_base + job.a_col;
          const bool from_epi = (job.flags & TCJ_FROM_EPI) != 0;
          if (job.flags & TCJ_WAIT_X) { mbar_wait(bar_x, px, p.dbg, 0x1000000u); px ^= 1; }
          const uint32_t set = rj & 1u;
          const uint32_t cph = set ? cph1 : cph0;
          const uint32_t bar_c = bar_achunk + 8 * set * TC_MAX_CHUNKS;
          uint32_t acc = (job.flags & TCJ_ACC_FIRST) ? 1u : 0u;
          for (uint32_t c = 0; c < job.nchunks; ++c) {
            if (from_epi) mbar_wait(bar_c + 8 * c, (cph >> c) & 1u, p.dbg, 0x2000000u | (j << 8) | c);
            mbar_wait(bar_full + 8 * stage, phase, p.dbg, 0x3000000u | (j << 8) | c);
            fence_after_sync();
            if (elect_one()) {
              const uint32_t lo = job.desc_lo_base | ((stages16 + stage * stage16) & 0x3FFFu);
              const uint64_t bhi = (static_cast<uint64_t>(DESC_HI) << 32) | lo;
              mma_ts(d, a + 16 * c, bhi, job.idesc, acc);
              if (three) {
                mma_ts(d, a + 16 * c + 8, bhi, job.idesc, 1u);
                mma_ts(d, a + 16 * c, bhi + job.lo_off16, job.idesc, 1u);
              }
              mma_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
            }
            __syncwarp();
            acc = 1u;
            if (++stage == static_cast<uint32_t>(p.n_stages)) { stage = 0; phase ^= 1; }
          }
          if (from_epi) {
            const uint32_t used = (1u << job.nchunks) - 1u;
            if (set) cph1 ^= used; else cph0 ^= used;
            ++rj;
          }
          const uint32_t commit = job.flags & TCJ_COMMIT_MASK;
          if (commit && elect_one())
            mma_commit(commit == TCJ_COMMIT_D0 ? bar_d0 : (commit == TCJ_COMMIT_D ? bar_d : bar_dout));
          __syncwarp();
        }
      }
    }
  } else {
    // ============================================================ epilogue warps
    const int e = warp - 2;                 // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = e >> 2;                // which interleaved half of the 16-column chunks
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tm = tmem_base + lane_base;
    const int row_in_tile = q * 32 + lane;
    uint32_t pd0 = 0, pd = 0, pdout = 0, rj = 0;  // rj: hidden rounds produced so far (chunk-barrier set = rj & 1)
    const int kp0_chunks = M.layer[0].Kpad >> 4;
    const float n_members_f = static_cast<float>(nM);

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
            tmem_st8(tm + p.col_x + 16 * c, hi);
            if (p.passes == 3) tmem_st8(tm + p.col_x + 16 * c + 8, lo);
          }
        }
        wait_st();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x);

        // ---- hidden layers of every member: accumulator -> activation -> next A operand, in place
        int idx = 0;
        for (int mm = 0; mm < nM; ++mm) {
          for (int l = 0; l + 1 < nL; ++l, ++idx) {
            const LayerDev& L = M.layer[l];
            const int n_a_chunks = M.layer[l + 1].Kpad >> 4;
            const uint32_t taddr = tm + ((idx & 1) ? p.col_buf1 : p.col_buf0);
            if (l == 0) { mbar_wait(bar_d0, pd0, p.dbg, 0x4000000u | (mm << 16)); pd0 ^= 1; }
            else { mbar_wait(bar_d, pd, p.dbg, 0x4000000u | (mm << 16) | (l << 8)); pd ^= 1; }
            fence_after_sync();
            const uint32_t bar_set = bar_achunk + 8 * ((rj & 1u) * TC_MAX_CHUNKS);
            ++rj;
            switch (L.act) {
              case BBMPC_ACT_TANH: epi_hidden<BBMPC_ACT_TANH>(taddr, L.Npad, L.N, n_a_chunks, half, p.passes, bar_set, lane); break;
              case BBMPC_ACT_RELU: epi_hidden<BBMPC_ACT_RELU>(taddr, L.Npad, L.N, n_a_chunks, half, p.passes, bar_set, lane); break;
              case BBMPC_ACT_SIGMOID: epi_hidden<BBMPC_ACT_SIGMOID>(taddr, L.Npad, L.N, n_a_chunks, half, p.passes, bar_set, lane); break;
              default: epi_hidden<BBMPC_ACT_NONE>(taddr, L.Npad, L.N, n_a_chunks, half, p.passes, bar_set, lane); break;
            }
          }
        }

        // ---- output layer (sum over members) -> process_output -> reward
        mbar_wait(bar_dout, pdout, p.dbg, 0x5000000u); pdout ^= 1;
        fence_after_sync();
        float s2[DS_T];
        {
          const LayerDev& Lo = M.layer[nL - 1];
#pragma unroll
          for (int c = 0; c < DS_T / 8; ++c) {
            uint32_t r[8];
            tmem_ld8(tm + p.col_dout + 8 * c, r);
            wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int i = 8 * c + j;
              float y = act_fast(__uint_as_float(r[j]), nM == 1 ? Lo.act : BBMPC_ACT_NONE);
              if (nM > 1) y = __fdiv_rn(y, n_members_f);
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
  if (p.dbg) {  // BBMPC_DEBUG=1: synchronise and dump the watchdog record of a starved pipeline
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      uint32_t* h = static_cast<uint32_t*>(ctx->dbg_host);
      for (int b = 0; b < 2; ++b)
        for (int w = 0; w < 10; ++w) {
          const uint32_t* r = h + 128 * b + 8 * w;
          if (r[0]) fprintf(stderr, "[bbmpc watchdog] blk%%2=%d warp=%d tid=%u bar=+%u parity=%u tag=%08x\n", b, w, r[0] & 0xFFFF,
                            r[1], r[2], r[3]);
        }
      return fail(ctx, BBMPC_ECUDA, "rollout_tc_kernel: %s", cudaGetErrorString(e));
    }
  }
  return BBMPC_OK;
}

int tc_du_slots(int dU) { return dU <= 8 ? 8 : 16; }
uint32_t tc_idesc(int Npad) { return idesc_bf16_f32(TILE_ROWS, static_cast<uint32_t>(Npad)); }

// TMEM column map: two ping-pong buffers (accumulator of layer l / in-place A operand of layer
// l+1) in the two 256-column halves, the layer-0 input X behind buffer 0 and the output
// accumulator behind buffer 1.
bool tc_column_map(const MlpDev& m, int* buf_w, int* col_x, int* col_dout) {
  const int nL = m.n_layers;
  int w = 0;
  for (int l = 0; l + 1 < nL; ++l) {
    if (m.layer[l].Npad > w) w = m.layer[l].Npad;
    if (m.layer[l + 1].Kpad > w) w = m.layer[l + 1].Kpad;
  }
  const int x_w = m.layer[0].Kpad, out_w = (m.layer[nL - 1].Npad + 15) / 16 * 16;
  if (buf_w) *buf_w = w;
  if (col_x) *col_x = w;
  if (col_dout) *col_dout = 256 + w;
  return w + x_w <= 256 && w + (out_w > 32 ? out_w : 32) <= 256;
}

int launch_rollout_tc(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                      const float* penalty, int rows, int A, int H, int passes, cudaStream_t st) {
  const ModelHost& m = ctx->model;
  TcParams p{};
  p.mlp = m.mlp; p.norm = m.norm; p.reward_id = ctx->reward_id; p.dS = m.dS; p.dU = m.dU;
  p.states = states; p.actions = actions; p.returns = returns; p.penalty = penalty;
  p.rows = rows; p.A = A; p.H = H; p.passes = passes;
  p.n_tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  const int nL = m.mlp.n_layers;
  int stage = 0;
  for (int l = 0; l < nL; ++l)
    if (m.mlp.layer[l].chunk_bytes > stage) stage = m.mlp.layer[l].chunk_bytes;
  int buf_w = 0;
  if (!tc_column_map(m.mlp, &buf_w, &p.col_x, &p.col_dout))
    return fail(ctx, BBMPC_EINVAL, "tensor-core path: TMEM column budget exceeded");
  p.col_buf0 = 0;
  p.col_buf1 = 256;
  if (getenv("BBMPC_DEBUG")) {
    if (!ctx->dbg_host) {
      BB_CUDA(ctx, cudaHostAlloc(&ctx->dbg_host, 1024, cudaHostAllocMapped));
      memset(ctx->dbg_host, 0, 1024);
    }
    BB_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&p.dbg), ctx->dbg_host, 0));
  }
  p.stage_bytes = stage;
  const size_t budget = 227 * 1024;
  int n_stages = TC_MAX_STAGES;
  const int cps = m.mlp.chunks_per_step, jps = m.mlp.jobs_per_step;
  while (n_stages > 2 && tc_layout(stage, n_stages, cps, jps).total > budget) --n_stages;
  if (tc_layout(stage, n_stages, cps, jps).total > budget)
    return fail(ctx, BBMPC_EINVAL, "tensor-core path: shared memory budget exceeded");
  p.n_stages = n_stages;
  const size_t smem_bytes = tc_layout(stage, n_stages, cps, jps).total;
  const int grid = p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count;
  if (m.dS <= 8 && m.dU <= 8) return launch_t<8, 8>(ctx, p, grid, smem_bytes, st);
  if (m.dS <= 24 && m.dU <= 8) return launch_t<24, 8>(ctx, p, grid, smem_bytes, st);
  if (m.dU <= 8) return launch_t<32, 8>(ctx, p, grid, smem_bytes, st);
  return launch_t<32, 16>(ctx, p, grid, smem_bytes, st);
}

}  // namespace bbmpc
