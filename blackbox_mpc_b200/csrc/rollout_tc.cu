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
//   warp 0       producer: streams the weight image L2 -> SMEM ring in groups of K-chunks
//                (cp.async.bulk + mbarrier tx bytes; one full/empty handshake per group)
//   warp 1       MMA issuer (one elected lane): tcgen05.mma A-from-TMEM, tcgen05.commit
//   warp 2       owns the TMEM allocation; warp 3 idle (warpgroup padding)
//   warps 4..7   state warps, one per TMEM lane quarter: trajectory state, action and return of 32 rows in
//                registers for all H steps; build the layer-0 operand, consume the output accumulator
//   warps 8..    conversion warps, EPI_SUB per lane quarter: tcgen05.ld accumulator chunk -> activation ->
//                bf16 hi/lo split -> tcgen05.st as the next layer's A operand
//
// Layer pipeline.  The accumulator of layer l is converted IN PLACE, 16 columns at a time, into
// the A operand of layer l+1 (columns [16c,16c+8) = hi, [16c+8,16c+16) = lo of K-chunk c); converted
// chunks are published per UNIT (pair of chunks) on the unit's mbarrier.  The MMA issuer consumes units as
// they appear (probing the next unit's barrier before it issues the current unit's MMAs) and accumulates
// layer l+1 into the other of two TMEM buffers, so the tensor pipe works on layer l+1 while layer l is
// still being converted.  One CTA per tile (single models): the first layer of member m+1 is issued ahead
// of the output layer of member m and all members' output layers accumulate into one TMEM tile.
// Member-parallel mode (ensembles): the n_members CTAs of a group share a tile, one member each, and
// exchange their raw outputs through L2 once per horizon step (see the state-warp code).
//
// Activations never touch shared or global memory; HBM traffic is the action read
// (H*dU floats per trajectory) and the 4-byte return.  Bias is folded into the GEMM: the A
// operand carries three ones-columns that meet three bf16 bias rows of the weight image; hidden tanh
// layers carry 2 log2(e) in their weights so the activation starts at ex2.
// DESIGN.md section 4.1 has the measurements behind each of these choices.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "device_fns.cuh"
#include "tc05.cuh"
#include "tc_epi.cuh"

namespace bbmpc {
using namespace tc05;

// 2 conversion warps per quarter => 512 threads => 128 registers per thread.  Measured (tools/debug/ab.sh,
// r1c, C4 rollout): 2 warps 1.40 ms, 3 warps (96 registers) 1.44 ms, 4 warps (80 registers, state warps
// spill) 1.53 ms: the conversion is bound by the MUFU pipe of its SM sub-partition, not by warp count.
#ifndef BBMPC_EPI_SUB
#define BBMPC_EPI_SUB 2
#endif
constexpr int EPI_SUB = BBMPC_EPI_SUB;                      // conversion warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * EPI_SUB;
constexpr int TC_WARPS = 8 + EPI_WARPS;         // warpgroup 0: producer, MMA issuer, TMEM owner, idle; warpgroup 1: state warps
constexpr int TC_THREADS = 32 * TC_WARPS;
// (512 threads -> 128 registers per thread; ptxas does not raise the cap after setmaxnreg.inc, so every
// role is written to fit 128 registers and no setmaxnreg is used.)
constexpr int TC_MAX_STAGES = 24;
constexpr int TC_MAX_CHUNKS = 16;   // K chunks of one layer (Kpad <= 256)
constexpr int TILE_ROWS = 128;
constexpr int BBMPC_ERETRY = -100;   // internal: relaunch with the member-parallel mode disabled
constexpr int MAX_DU_T = 16;
#ifndef EPI_STAGGER
#define EPI_STAGGER 150   // clocks between the first chunks of the conversion warps of a quarter
#endif

struct TcParams {
  MlpDev mlp;
  NormDev norm;
  int reward_id, dS, dU;
  const float* states; const float* actions; float* returns; const float* penalty;
  int rows, A, H, n_tiles, passes;
  int stage_bytes, n_stages;
  int col_buf0, col_buf1, col_x, col_dout;  // TMEM column map
  // issue tables of this launch (full ensemble per CTA, or one member per CTA)
  const TcJob* jobs; int n_jobs; const uint2* table; int n_table;
  // member-parallel mode: the n_members CTAs of a group share a tile; each contracts ONE member and the
  // members' raw outputs are exchanged through `xchg` (L2) once per horizon step.  0 = one CTA per tile.
  int group_mode; float* xchg; unsigned* flags;
  int xs_bytes;   // > 0: the group's outputs are staged in shared memory by bulk copies (else read from L2 directly)
  uint32_t* dbg;  // host-mapped watchdog record (BBMPC_DEBUG=1), else nullptr
  uint32_t* trace; // BBMPC_TC_TRACE: per-warp (tag, clock) records of CTA 0, step 1
  float* traj;    // user reward: visited states [rows][H][dS] (else nullptr)
  int xflags;     // BBMPC_TC_X timing experiments (results are garbage): 1 = no weight loads, 2 = identity activations
};

struct TcSmemLayout {
  uint32_t stages, xs, acts, table, jobs, bars, tmem_slot, stats, conv, total;
};
constexpr int TC_NUM_BARS = 2 * TC_MAX_STAGES + 2 * TC_MAX_CHUNKS + 7;
__host__ __device__ inline TcSmemLayout tc_layout(int stage_bytes, int n_stages, int groups_per_step, int jobs_per_step, int xs_bytes) {
  TcSmemLayout L;
  uint32_t off = 0;
  L.stages = off; off += static_cast<uint32_t>(stage_bytes) * n_stages;
  L.xs = off;     off += static_cast<uint32_t>(xs_bytes);   // member-parallel mode: staging of the group's outputs
  L.acts = off;   off += 2u * TILE_ROWS * MAX_DU_T * 4u;    // next step's actions, prefetched with cp.async (double buffer)
  L.table = off;  off += static_cast<uint32_t>(groups_per_step) * 8;
  off = (off + 15u) & ~15u;
  L.jobs = off;   off += static_cast<uint32_t>(jobs_per_step) * sizeof(TcJob);
  L.bars = off;   off += TC_NUM_BARS * 8;
  off = (off + 15u) & ~15u;
  L.tmem_slot = off; off += 16;
  L.stats = off;  off += (4 * MAX_DS + 2 * MAX_DU + 3 * 64) * 4;   // + layer-0 operand tables (mean / 1/den / add per K position)
  L.conv = off;   off += MAX_LAYERS * (32 + 256);   // per hidden layer: 8 ints + mask/add vectors of its 2 trailing chunks
  L.total = off;
  return L;
}

template <bool ON>
struct Tracer {
  uint32_t* buf; uint32_t n; bool on;
  __device__ __forceinline__ void rec(uint32_t tag) {
    if (ON) { if (on && n < 500) { buf[2 * n] = tag; buf[2 * n + 1] = static_cast<uint32_t>(clock64()); ++n; } }
  }
  __device__ __forceinline__ void arm(bool v) { if (ON) on = v; }
};

// Converts one 16-column accumulator chunk (already in registers) into the bf16 hi/lo A-operand chunk
// of the next layer and stores it in place.  Full chunks (all 16 columns are real features) take the
// lean path; the one chunk per layer that holds the ones-columns / padding takes convert_edge, kept
// out of line so that its column tests are not if-converted into the hot path.
__device__ __forceinline__ void store_split(uint32_t taddr, int c, const float (&v)[16], int passes) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_bf16x2_veltkamp(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tmem_st8(taddr + 16 * c, hi);
  if (passes == 3) tmem_st8(taddr + 16 * c + 8, lo);
}
template <int ACT>
__device__ __forceinline__ void convert_full(uint32_t taddr, int c, const uint32_t (&r)[16], int passes) {
  if constexpr (ACT == BBMPC_ACT_TANH && PACKED_TANH) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      tanh_split_quad(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]),
                      hi[2 * j], lo[2 * j], hi[2 * j + 1], lo[2 * j + 1]);
    if (passes == 3) tmem_st16(taddr + 16 * c, hi, lo);
    else tmem_st8(taddr + 16 * c, hi);
  } else {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    act16<ACT>(v);
    store_split(taddr, c, v, passes);
  }
}
// The (at most two) trailing chunks of a layer hold real features, the three ones-columns that meet the
// bias rows of the next layer's weights, and zero padding: v = act(D) * mask + add with per-column
// mask / add vectors prepared in shared memory (no per-column tests in the instruction stream).
template <int ACT>
__device__ __forceinline__ void convert_tail(uint32_t taddr, int c, bool has_data, int n_real, const float* mk_ad, int passes) {
  uint32_t r[16], hi[8], lo[8];
  if (has_data) { tmem_ld16(taddr + 16 * c, r); wait_ld(); }
  tail16<ACT>(r, has_data, n_real, mk_ad, hi, lo);   // tc_epi.cuh (shared with rollout_pipe.cu: identical arithmetic)
  if (passes == 3) tmem_st16(taddr + 16 * c, hi, lo);
  else tmem_st8(taddr + 16 * c, hi);
}

// One hidden-layer epilogue of one warp: 16-column chunks c = sub, sub+EPI_SUB, ... of its TMEM
// lanes: accumulator -> activation -> bf16 hi/lo -> written back over the same columns -> one
// arrival on the mbarrier of the chunk's UNIT (pair of chunks, the MMA issuer's wait granularity).
template <int ACT, bool TR>
__device__ __forceinline__ void epi_hidden(uint32_t taddr, int Npad, int N, int n_a_chunks, int cb, int ce, int sub,
                                           int passes, uint32_t bar_unit, int lane, const float* tail_tab, Tracer<TR>& tr) {
  const int n_full = N >> 4;   // chunks whose 16 columns are all real features
  if (EPI_STAGGER > 0 && sub > 0) {
    // The EPI_SUB warps of a quarter share one MUFU pipe; started together they finish their chunks in
    // bursts and the MMA issuer idles, then has a whole round of chunks left when the epilogue is done.
    // A small start offset per warp makes the chunks complete at a steady rate instead.
    const long long t0 = clock64();
    while (clock64() - t0 < static_cast<long long>(sub) * EPI_STAGGER) {}
  }
  // (processing whole units per iteration saves ~150 cycles of per-iteration overhead per chunk but coarsens the
  // pipeline towards the MMA issuer: measured slower, 1.45 vs 1.39 ms)
  for (int c = cb + ((sub - cb % EPI_SUB + EPI_SUB) % EPI_SUB); c < ce; c += EPI_SUB) {   // chunks of [cb, ce) with c % EPI_SUB == sub
    tr.rec(0x100u | c);
    if (c < n_full) {
      uint32_t r[16];
      tmem_ld16(taddr + 16 * c, r);
      wait_ld();
      convert_full<ACT>(taddr, c, r, passes);
    } else {
      convert_tail<ACT>(taddr, c, 16 * c < Npad, N - 16 * c, tail_tab + 32 * (c - n_full), passes);
    }
    tr.rec(0x300u | c);
    wait_st();
    fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar_unit + 8 * (c >> 1));
      if ((c ^ 1) >= n_a_chunks) mbar_arrive(bar_unit + 8 * (c >> 1));   // lone last chunk stands for its missing partner
    }
    tr.rec(0x500u | c);
  }
}

template <int DS_T, int DU_T, bool TR, int ACT_T>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(const TcParams p_in) {
  const TcParams& p = p_in;
  // watchdog records only exist in the debug/trace build (keeps the production kernel's code small)
  volatile uint32_t* const dbgp = TR ? p_in.dbg : nullptr;
  extern __shared__ __align__(128) uint8_t smem[];
  const TcSmemLayout lay = tc_layout(p.stage_bytes, p.n_stages, p.n_table, p.n_jobs, p.xs_bytes);
  const uint32_t smem_base = smem_u32(smem);
  const uint2* table = reinterpret_cast<const uint2*>(smem + lay.table);
  const uint32_t bar_full = smem_base + lay.bars;
  const uint32_t bar_empty = bar_full + TC_MAX_STAGES * 8;
  // epilogue -> MMA: K-chunk c of the A operand written.  Two sets, alternating per hidden round: with
  // the early first-layer issue the epilogue may run ONE round ahead of the MMA issuer's consumption,
  // and an mbarrier must never complete two phases ahead of its waiter (parity aliasing).
  const uint32_t bar_achunk = bar_empty + TC_MAX_STAGES * 8;
  const uint32_t bar_x = bar_achunk + 2 * TC_MAX_CHUNKS * 8;  // epilogue -> MMA: layer-0 input of this step written
  // MMA -> conversion warps: accumulator column half h of a first layer (bar_d0 + 8h) / of a later hidden
  // layer (bar_d + 8h) complete.  One barrier per (kind, half): each is at most one phase ahead of its waiters.
  const uint32_t bar_d0 = bar_x + 8;
  const uint32_t bar_d = bar_d0 + 16;
  const uint32_t bar_dout = bar_d + 16;                       // MMA -> state warps: output accumulator complete
  const uint32_t bar_xs = bar_dout + 8;                       // bulk copies of the group's outputs landed (member-parallel mode)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + lay.tmem_slot);
  float* st_mean_s = reinterpret_cast<float*>(smem + lay.stats);
  float* st_rden_s = st_mean_s + MAX_DS;   // 1 / (std_s + 1e-7)
  float* st_mean_t = st_rden_s + MAX_DS;
  float* st_den_t = st_mean_t + MAX_DS;
  float* st_mean_a = st_den_t + MAX_DS;
  float* st_rden_a = st_mean_a + MAX_DU;
  float* xt_mean = st_rden_a + MAX_DU;     // per K position of the layer-0 operand: x = (src - mean) * rden + add
  float* xt_rden = xt_mean + 64;
  float* xt_add = xt_rden + 64;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const MlpDev& M = p.mlp;
  const int nL = M.n_layers;
  const int nM = M.n_members;
  const bool norm_on = p.norm.enabled != 0;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < p.n_table; i += TC_THREADS)
    reinterpret_cast<uint2*>(smem + lay.table)[i] = p.table[i];
  for (int i = tid; i < p.n_jobs * static_cast<int>(sizeof(TcJob) / 16); i += TC_THREADS)
    reinterpret_cast<uint4*>(smem + lay.jobs)[i] = reinterpret_cast<const uint4*>(p.jobs)[i];
  for (int i = tid; i < MAX_DS; i += TC_THREADS) {
    const bool in = norm_on && i < p.dS;
    st_mean_s[i] = in ? p.norm.mean_s[i] : 0.0f;
    st_rden_s[i] = in ? __frcp_rn(p.norm.den_s[i]) : 1.0f;
    st_mean_t[i] = in ? p.norm.mean_t[i] : 0.0f;
    st_den_t[i] = in ? p.norm.den_t[i] : (i < p.dS ? 1.0f : 0.0f);
  }
  for (int i = tid; i < MAX_DU; i += TC_THREADS) {
    const bool in = norm_on && i < p.dU;
    st_mean_a[i] = in ? p.norm.mean_a[i] : 0.0f;
    st_rden_a[i] = in ? __frcp_rn(p.norm.den_a[i]) : 1.0f;
  }
  for (int k = tid; k < 64; k += TC_THREADS) {   // operand layout: [actions (DU_T slots) | state | 1 1 1 | 0 ...]
    float mean = 0.0f, rden = 0.0f, add = 0.0f;
    if (k < DU_T) {
      if (k < p.dU) { mean = norm_on ? p.norm.mean_a[k] : 0.0f; rden = norm_on ? __frcp_rn(p.norm.den_a[k]) : 1.0f; }
    } else {
      const int i = k - DU_T;
      if (i < p.dS) { mean = norm_on ? p.norm.mean_s[i] : 0.0f; rden = norm_on ? __frcp_rn(p.norm.den_s[i]) : 1.0f; }
      else if (i < p.dS + BIAS_COLS) add = 1.0f;
    }
    xt_mean[k] = mean; xt_rden[k] = rden; xt_add[k] = add;
  }
  if (tid == 0) {
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int c = 0; c < 2 * TC_MAX_CHUNKS; ++c) mbar_init(bar_achunk + 8 * c, 8);   // 2 chunks x 4 lane quarters per unit
    mbar_init(bar_x, 4);
    mbar_init(bar_d0, 1); mbar_init(bar_d0 + 8, 1);
    mbar_init(bar_d, 1); mbar_init(bar_d + 8, 1);
    mbar_init(bar_dout, 1);
    mbar_init(bar_xs, 1);
    fence_mbar_init();
    // what the conversion warps need to know about hidden layer l, in shared memory (LDS instead of
    // dynamically indexed kernel-parameter loads in their per-layer prologue)
    int* cv = reinterpret_cast<int*>(smem + lay.conv);
    for (int l = 0; l + 1 < nL; ++l) {
      cv[8 * l + 0] = M.layer[l].Npad; cv[8 * l + 1] = M.layer[l].N; cv[8 * l + 2] = M.layer[l].act;
      cv[8 * l + 3] = M.layer[l + 1].Kpad >> 4;                  // A-operand chunks of the next layer
      cv[8 * l + 4] = M.layer[l].nsplit >> 4;                    // first chunk of column half 1 (0: not split)
      float* tt = reinterpret_cast<float*>(smem + lay.conv + MAX_LAYERS * 32) + 64 * l;
      const int N = M.layer[l].N, f0 = (N >> 4) << 4;
      for (int q2 = 0; q2 < 2; ++q2)
        for (int j = 0; j < 16; ++j) {
          const int f = f0 + 16 * q2 + j;
          tt[32 * q2 + j] = f < N ? 1.0f : 0.0f;                                   // mask
          tt[32 * q2 + 16 + j] = (f >= N && f < N + BIAS_COLS) ? 1.0f : 0.0f;      // add
        }
    }
  }
  if (warp == 2) {
    tmem_alloc(smem_base + lay.tmem_slot, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // tile ownership: CTA b owns tiles b, b+grid, ... ; in member-parallel mode the nM consecutive CTAs of
  // group g = b / nM own tiles g, g+G, ... together and CTA b contracts member b % nM.
  const int n_owners = p.group_mode ? static_cast<int>(gridDim.x) / nM : static_cast<int>(gridDim.x);
  const int owner = p.group_mode ? static_cast<int>(blockIdx.x) / nM : static_cast<int>(blockIdx.x);
  const int my_member = p.group_mode ? static_cast<int>(blockIdx.x) % nM : 0;
  const int my_tiles = (p.n_tiles - owner + n_owners - 1) / n_owners;
  const int nM_here = p.group_mode ? 1 : nM;   // members contracted by this CTA

  if (warp < 4) {
  if (warp == 0) {
    // ============================================================ producer
    // One bulk copy per chunk GROUP (up to ~52 KB of consecutive K-chunks of one layer).
    if (lane == 0) {
      const long long total = static_cast<long long>(my_tiles) * p.H * p.n_table;
      const uint8_t* wimg = M.wimg + static_cast<size_t>(my_member) * M.img_member_stride;
      int stage = 0, ci = 0;
      uint32_t phase = 0;
      for (long long i = 0; i < total; ++i) {
        const uint2 e = table[ci];
        mbar_wait(bar_empty + 8 * stage, phase ^ 1, dbgp, 0x6000000u | static_cast<uint32_t>(i));
        if (p.xflags & 1) { mbar_arrive(bar_full + 8 * stage); }
        else {
          mbar_arrive_expect_tx(bar_full + 8 * stage, e.y);
          bulk_g2s(smem_base + lay.stages + stage * p.stage_bytes, wimg + e.x, e.y, bar_full + 8 * stage);
        }
        if (++ci == p.n_table) ci = 0;
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    // The whole warp runs the (warp-uniform) control flow; one elected lane issues the tcgen05
    // instructions.  A successful mbarrier wait costs ~150 cycles of latency in this serial
    // instruction stream (measured, tools/probe/issue_cost.cu), a K-chunk is only 312 cycles of
    // tensor work: so weights arrive in multi-chunk groups (one full/empty handshake and one
    // commit per group), and the state of the NEXT chunk's barrier is probed (non-blocking
    // test_wait) before the current chunk's MMAs are issued, which hides the probe latency.
    const TcJob* jobs = reinterpret_cast<const TcJob*>(smem + lay.jobs);
    const int n_jobs = p.n_jobs;
    const bool three = (p.passes == 3);
    uint32_t stage = 0, phase = 0, px = 0, cph0 = 0, cph1 = 0, rj = 0;  // rj: hidden rounds consumed so far
    const uint32_t stages16 = (smem_base + lay.stages) >> 4, stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4;
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1, no swizzle
    Tracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};
    for (int tile = 0; tile < my_tiles; ++tile) {
      for (int t = 0; t < p.H; ++t) {
        tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && tile == 0 && t == 1);
        for (int j = 0; j < n_jobs; ++j) {
          const TcJob job = jobs[j];
          const uint32_t d = tmem_base + job.d_col, a = tmem_base + job.a_col;
          const bool from_epi = (job.flags & TCJ_FROM_EPI) != 0;
          if (job.flags & TCJ_WAIT_X) { mbar_wait(bar_x, px, dbgp, 0x1000000u); px ^= 1; }
          const uint32_t set = rj & 1u;
          const uint32_t cph = set ? cph1 : cph0;
          const uint32_t bar_c = bar_achunk + 8 * set * TC_MAX_CHUNKS;
          uint32_t acc = (job.flags & TCJ_ACC_FIRST) ? 1u : 0u;
          uint32_t u = 0, pre_ok = 0;
          const uint32_t n_units = (job.nchunks + 1u) >> 1;
          for (uint32_t g = 0; g < job.ngroups; ++g) {
            const uint32_t n = (job.gsz >> (4 * g)) & 15u;
            mbar_wait(bar_full + 8 * stage, phase, dbgp, 0x3000000u | (j << 8) | g);
            const uint32_t sbase = (stages16 + stage * stage16);
            for (uint32_t k = 0; k < n; ++k, ++u) {
              if (from_epi) {
                if (!pre_ok) {
                  if (p.xflags & 8) mbar_wait(bar_c + 8 * u, (cph >> u) & 1u, dbgp, 0x2000000u | (j << 8) | u);
                  else mbar_wait_poll(bar_c + 8 * u, (cph >> u) & 1u);
                }
                pre_ok = (u + 1 < n_units) ? mbar_test_wait(bar_c + 8 * (u + 1), (cph >> (u + 1)) & 1u) : 0u;
              }
              fence_after_sync();
              tr.rec(0x1000u | (j << 4) | u);
              const bool two = 2 * u + 1 < job.nchunks;
              if (elect_one()) {
                const uint32_t lo = job.desc_lo_base | ((sbase + 2 * k * job.chunk16) & 0x3FFFu);
                const uint64_t b0 = (static_cast<uint64_t>(DESC_HI) << 32) | lo;
                const uint32_t a0 = a + 32 * u;
                mma_ts(d, a0, b0, job.idesc, acc);
                if (three) {
                  mma_ts(d, a0 + 8, b0, job.idesc, 1u);
                  mma_ts(d, a0, b0 + job.lo_off16, job.idesc, 1u);
                }
                if (two) {
                  const uint64_t b1 = b0 + job.chunk16;
                  mma_ts(d, a0 + 16, b1, job.idesc, 1u);
                  if (three) {
                    mma_ts(d, a0 + 24, b1, job.idesc, 1u);
                    mma_ts(d, a0 + 16, b1 + job.lo_off16, job.idesc, 1u);
                  }
                }
              }
              __syncwarp();
              acc = 1u;
            }
            if (elect_one()) mma_commit(bar_empty + 8 * stage);  // frees the ring stage when these MMAs retire
            __syncwarp();
            if (++stage == static_cast<uint32_t>(p.n_stages)) { stage = 0; phase ^= 1; }
          }
          if (job.flags & TCJ_ROUND_END) {
            const uint32_t used = (1u << n_units) - 1u;
            if (set) cph1 ^= used; else cph0 ^= used;
            ++rj;
          }
          const uint32_t commit = (job.flags & TCJ_COMMIT_MASK) >> 4;   // 1..4: bar_d0 + 8*(commit-1) (bar_d follows bar_d0)
          if (commit && elect_one()) mma_commit(commit == 5u ? bar_dout : bar_d0 + 8 * (commit - 1u));
          __syncwarp();
        }
      }
    }
  }
  } else if (warp < 8) {
    // ============================================================ state warps (one per TMEM lane quarter)
    // Own the trajectory state, action and return of their 32 rows in registers for all H steps:
    // build the layer-0 input (process_input), consume the output accumulator (process_output,
    // reward, NaN guard).  They sleep on bar_dout while the hidden layers run.
    const int q = warp & 3;
    const uint32_t tm = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int row_in_tile = q * 32 + lane;
    uint32_t pdout = 0;
    const int kp0_chunks = M.layer[0].Kpad >> 4;
    const float inv_members = __frcp_rn(static_cast<float>(nM));
    Tracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};

    for (int tile = 0; tile < my_tiles; ++tile) {
      const int row = (owner + tile * n_owners) * TILE_ROWS + row_in_tile;
      const bool valid = row < p.rows;
      const int arow = valid ? row : 0;
      float s[DS_T];
      const float* srow_ptr = p.states + static_cast<size_t>(arow % p.A) * p.dS;
#pragma unroll
      for (int i = 0; i < DS_T; ++i) s[i] = (i < p.dS) ? srow_ptr[i] : 0.0f;
      float ret = 0.0f;
      const float* arow_ptr = p.actions + static_cast<size_t>(arow) * p.H * p.dU;
      // actions reach the thread through shared memory: step t+1's dU floats are fetched with cp.async while
      // step t runs (a register prefetch gets sunk to its use by the compiler and exposes the L2 latency).
      float* act_s = reinterpret_cast<float*>(smem + lay.acts) + row_in_tile;   // [buf][i][row]
      auto fetch_actions = [&](int t_next) {
        const uint32_t dst = smem_u32(act_s + (t_next & 1) * (MAX_DU_T * TILE_ROWS));
        const float* src = arow_ptr + static_cast<size_t>(t_next) * p.dU;
#pragma unroll
        for (int i = 0; i < DU_T; ++i)
          if (i < p.dU) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + i * TILE_ROWS * 4), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      if (p.H > 0) fetch_actions(0);

      for (int t = 0; t < p.H; ++t) {
        tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && tile == 0 && t == 1);
        tr.rec(0x10u);
        float a[DU_T];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int i = 0; i < DU_T; ++i) a[i] = (i < p.dU) ? act_s[(t & 1) * (MAX_DU_T * TILE_ROWS) + i * TILE_ROWS] : 0.0f;
        if (t + 1 < p.H) fetch_actions(t + 1);
        tr.rec(0x12u);
        // ---- process_input: X = [norm(a) (DU_T slots) | norm(s) | 1 1 1 | 0...] -> TMEM
        constexpr int KP0_T = (DU_T + DS_T + BIAS_COLS + 15) / 16;
#pragma unroll
        for (int c = 0; c < KP0_T; ++c) {
          if (c < kp0_chunks) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 mn = *reinterpret_cast<const float4*>(xt_mean + 16 * c + j4);
              const float4 rd = *reinterpret_cast<const float4*>(xt_rden + 16 * c + j4);
              const float4 ad = *reinterpret_cast<const float4*>(xt_add + 16 * c + j4);
              float src[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = 16 * c + j4 + j;        // compile-time position -> register of a[] / s[]
                src[j] = (k < DU_T) ? a[k < DU_T ? k : 0] : ((k - DU_T < DS_T) ? s[(k - DU_T < DS_T && k >= DU_T) ? k - DU_T : 0] : 0.0f);
              }
              const float x0 = fmaf(__fsub_rn(src[0], mn.x), rd.x, ad.x), x1 = fmaf(__fsub_rn(src[1], mn.y), rd.y, ad.y);
              const float x2 = fmaf(__fsub_rn(src[2], mn.z), rd.z, ad.z), x3 = fmaf(__fsub_rn(src[3], mn.w), rd.w, ad.w);
              split_bf16x2_packed(pk2(x0, x1), hi[j4 / 2], lo[j4 / 2]);
              split_bf16x2_packed(pk2(x2, x3), hi[j4 / 2 + 1], lo[j4 / 2 + 1]);
            }
            tmem_st8(tm + p.col_x + 16 * c, hi);
            if (p.passes == 3) tmem_st8(tm + p.col_x + 16 * c + 8, lo);
          }
        }
        tr.rec(0x13u);
        wait_st();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x);
        tr.rec(0x11u);

        // ---- output layer (sum over members) -> process_output -> reward
        mbar_wait(bar_dout, pdout, dbgp, 0x5000000u); pdout ^= 1;
        fence_after_sync();
        tr.rec(0x30u);
        float s2[DS_T];
        {
          const LayerDev& Lo = M.layer[nL - 1];
          uint32_t r[DS_T / 8][8];
#pragma unroll
          for (int c = 0; c < DS_T / 8; ++c) tmem_ld8(tm + p.col_dout + 8 * c, r[c]);
          wait_ld();
          tr.rec(0x32u);
          if (p.group_mode) {
            // ---- member-parallel exchange: publish this member's raw output, wait for the whole group,
            // sum the members in member order (identical on every CTA of the group).
            const long long gstep = static_cast<long long>(tile) * p.H + t;
            const int par = static_cast<int>(gstep & 1);
            float* base = p.xchg + (static_cast<size_t>(owner) * 2 + par) * nM * (DS_T * TILE_ROWS);
            float4* mine = reinterpret_cast<float4*>(base + (static_cast<size_t>(my_member) * TILE_ROWS + row_in_tile) * DS_T);
#pragma unroll
            for (int c = 0; c < DS_T / 4; ++c)
              __stcg(mine + c, make_float4(__uint_as_float(r[c / 2][4 * (c & 1)]), __uint_as_float(r[c / 2][4 * (c & 1) + 1]),
                                           __uint_as_float(r[c / 2][4 * (c & 1) + 2]), __uint_as_float(r[c / 2][4 * (c & 1) + 3])));
            tr.rec(0x33u);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            tr.rec(0x34u);
            const uint32_t xs_member_bytes = TILE_ROWS * DS_T * 4;
            if (warp == 4 && lane == 0) {
              // release this CTA's rows (cumulative over the CTA barrier above), then wait for the group
              asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(p.flags + owner) : "memory");
              const unsigned target = static_cast<unsigned>(nM) * static_cast<unsigned>(gstep + 1);
              unsigned seen = 0, spins = 0;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + owner) : "memory");
                if (++spins > (1u << 26)) asm volatile("trap;");
              } while (seen < target);
              tr.rec(0x35u);
              if (p.xs_bytes) {   // L2 -> shared memory, one bulk copy per member
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_arrive_expect_tx(bar_xs, xs_member_bytes * nM);
                for (int m2 = 0; m2 < nM; ++m2)
                  bulk_g2s(smem_base + lay.xs + m2 * xs_member_bytes, base + static_cast<size_t>(m2) * (TILE_ROWS * DS_T), xs_member_bytes, bar_xs);
              }
            }
            float acc[DS_T];
#pragma unroll
            for (int i = 0; i < DS_T; ++i) acc[i] = 0.0f;
            if (p.xs_bytes) {
              mbar_wait(bar_xs, static_cast<uint32_t>(gstep & 1), dbgp, 0x7000000u);
              tr.rec(0x36u);
              const float4* xs = reinterpret_cast<const float4*>(smem + lay.xs) + row_in_tile * (DS_T / 4);
              // member order, all loads of a column group in flight together (group mode has nM <= 8)
#pragma unroll
              for (int c = 0; c < DS_T / 4; ++c) {
                float4 v[8];
#pragma unroll
                for (int m2 = 0; m2 < 8; ++m2)
                  if (m2 < nM) v[m2] = xs[m2 * (TILE_ROWS * DS_T / 4) + c];
#pragma unroll
                for (int m2 = 0; m2 < 8; ++m2)
                  if (m2 < nM) {
                    acc[4 * c] = __fadd_rn(acc[4 * c], v[m2].x); acc[4 * c + 1] = __fadd_rn(acc[4 * c + 1], v[m2].y);
                    acc[4 * c + 2] = __fadd_rn(acc[4 * c + 2], v[m2].z); acc[4 * c + 3] = __fadd_rn(acc[4 * c + 3], v[m2].w);
                  }
              }
              // every state thread is done with the staging area before the next step's copies may land:
              // guaranteed by the bar.sync at the top of the next exchange (the copies are issued after it).
            } else {
              asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
              for (int m2 = 0; m2 < nM; ++m2) {
                const float4* src = reinterpret_cast<const float4*>(base + (static_cast<size_t>(m2) * TILE_ROWS + row_in_tile) * DS_T);
                float4 v[DS_T / 4];
#pragma unroll
                for (int c = 0; c < DS_T / 4; ++c) v[c] = __ldcg(src + c);
#pragma unroll
                for (int c = 0; c < DS_T / 4; ++c) {
                  acc[4 * c] = __fadd_rn(acc[4 * c], v[c].x); acc[4 * c + 1] = __fadd_rn(acc[4 * c + 1], v[c].y);
                  acc[4 * c + 2] = __fadd_rn(acc[4 * c + 2], v[c].z); acc[4 * c + 3] = __fadd_rn(acc[4 * c + 3], v[c].w);
                }
              }
            }
            tr.rec(0x37u);
#pragma unroll
            for (int i = 0; i < DS_T; ++i) r[i / 8][i % 8] = __float_as_uint(acc[i]);
          }
          if (nM == 1 && Lo.act != BBMPC_ACT_NONE) {   // non-linear output layer (single model only): rare, kept out of line
#pragma unroll
            for (int c = 0; c < DS_T / 8; ++c)
#pragma unroll
              for (int j = 0; j < 8; ++j) r[c][j] = __float_as_uint(act_fast(__uint_as_float(r[c][j]), Lo.act));
          }
#pragma unroll
          for (int i = 0; i < DS_T; ++i) {
            float y = __uint_as_float(r[i / 8][i % 8]);
            if (nM > 1) y = __fmul_rn(y, inv_members);
            s2[i] = __fadd_rn(fmaf(y, st_den_t[i], st_mean_t[i]), s[i]);   // tables hold (0, 1) without normalisation, (0, 0) beyond dS
          }
        }
        // all of this thread's TMEM reads of D_out are complete (wait_ld) before the next arrive.
        if (p.traj && valid && my_member == 0) {   // user reward: dump the visited state (user_reward.cu)
          float* tp = p.traj + (static_cast<size_t>(row) * p.H + t) * p.dS;
#pragma unroll
          for (int i = 0; i < DS_T; ++i) if (i < p.dS) tp[i] = s2[i];
        }
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
        tr.rec(0x31u);
      }
      if (valid && my_member == 0) {
        float r = isnan(ret) ? -1e6f : ret;  // deterministic.py:75-77
        if (p.penalty) r = __fsub_rn(r, p.penalty[row]);
        p.returns[row] = r;
      }
    }
  } else {
    // ============================================================ conversion warps
    // EPI_SUB warps per TMEM lane quarter turn hidden-layer accumulators into the next layer's A operand.
    const int e = warp - 8;
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch (hardware: warp id % 4)
    const int sub = e >> 2;                 // 0 .. EPI_SUB-1
    const uint32_t tm = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t pd = 0, rj = 0;  // pd: phase bits of the four accumulator barriers; rj: hidden rounds produced so far
    const int* cv = reinterpret_cast<const int*>(smem + lay.conv);
    Tracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};
    const long long n_steps = static_cast<long long>(my_tiles) * p.H;
    for (long long st = 0; st < n_steps; ++st) {
      tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && st == 1);
      int idx = 0;
      for (int mm = 0; mm < nM_here; ++mm) {
        for (int l = 0; l + 1 < nL; ++l, ++idx) {
          const int Npad = cv[8 * l], N = cv[8 * l + 1], act = (p.xflags & 2) ? BBMPC_ACT_NONE : cv[8 * l + 2];
          const int n_a_chunks = cv[8 * l + 3], csplit = cv[8 * l + 4];
          const float* tail_tab = reinterpret_cast<const float*>(smem + lay.conv + MAX_LAYERS * 32) + 64 * l;
          const uint32_t taddr = tm + ((idx & 1) ? p.col_buf1 : p.col_buf0);
          const uint32_t bar_set = bar_achunk + 8 * ((rj & 1u) * TC_MAX_CHUNKS);
          ++rj;
          for (int h = 0; h < (csplit ? 2 : 1); ++h) {
            const uint32_t kind = (l == 0 ? 0u : 2u) + h;
            mbar_wait(bar_d0 + 8 * kind, (pd >> kind) & 1u, dbgp, 0x4000000u | (mm << 16) | (l << 8) | h);
            pd ^= 1u << kind;
            fence_after_sync();
            tr.rec(0x20u | (mm << 12) | (l << 8) | (h << 6));
            const int cb = h ? csplit : 0, ce = (csplit && !h) ? csplit : n_a_chunks;
            if (ACT_T >= 0) {
              epi_hidden<(ACT_T >= 0 ? ACT_T : 0), TR>(taddr, Npad, N, n_a_chunks, cb, ce, sub, p.passes, bar_set, lane, tail_tab, tr);
            } else {
              switch (act) {
                case BBMPC_ACT_TANH: epi_hidden<BBMPC_ACT_TANH, TR>(taddr, Npad, N, n_a_chunks, cb, ce, sub, p.passes, bar_set, lane, tail_tab, tr); break;
                case BBMPC_ACT_RELU: epi_hidden<BBMPC_ACT_RELU, TR>(taddr, Npad, N, n_a_chunks, cb, ce, sub, p.passes, bar_set, lane, tail_tab, tr); break;
                case BBMPC_ACT_SIGMOID: epi_hidden<BBMPC_ACT_SIGMOID, TR>(taddr, Npad, N, n_a_chunks, cb, ce, sub, p.passes, bar_set, lane, tail_tab, tr); break;
                default: epi_hidden<BBMPC_ACT_NONE, TR>(taddr, Npad, N, n_a_chunks, cb, ce, sub, p.passes, bar_set, lane, tail_tab, tr); break;
              }
            }
          }
        }
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
  // all hidden layers tanh (every reference tutorial): specialised kernel without the per-layer switch
  bool all_tanh = !(p.xflags & 2);
  for (int l = 0; l + 1 < p.mlp.n_layers; ++l) all_tanh = all_tanh && p.mlp.layer[l].act == BBMPC_ACT_TANH;
  auto kern = (p.trace || p.dbg) ? (all_tanh ? rollout_tc_kernel<DS_T, DU_T, true, BBMPC_ACT_TANH> : rollout_tc_kernel<DS_T, DU_T, true, -1>)
                                 : (all_tanh ? rollout_tc_kernel<DS_T, DU_T, false, BBMPC_ACT_TANH> : rollout_tc_kernel<DS_T, DU_T, false, -1>);
  BB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  if (p.group_mode) {
    // the CTAs of a group wait for each other every horizon step: all of them must be resident
    void* args[] = {const_cast<TcParams*>(&p)};
    const cudaError_t ce = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(TC_THREADS), args, smem_bytes, st);
    if (ce == cudaErrorCooperativeLaunchTooLarge) {   // fewer SMs than expected (MPS / green context): one CTA per tile instead
      cudaGetLastError();
      ctx->tc_no_groups = true;
      return BBMPC_ERETRY;
    }
    BB_CUDA(ctx, ce);
  } else {
    kern<<<grid, TC_THREADS, smem_bytes, st>>>(p);
  }
  BB_LAUNCH_CHECK(ctx);
  if (p.trace) {
    cudaStreamSynchronize(st);
    static std::vector<uint32_t> h(32 * 1024);
    cudaMemcpy(h.data(), p.trace, h.size() * 4, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("BBMPC_TC_TRACE"), "w")) {
      for (int w = 0; w < TC_WARPS; ++w)
        for (int i = 0; i < 500; ++i)
          if (h[w * 1024 + 2 * i]) fprintf(f, "%d %x %u\n", w, h[w * 1024 + 2 * i], h[w * 1024 + 2 * i + 1]);
      fclose(f);
    }
    cudaMemset(p.trace, 0, h.size() * 4);
  }
  if (p.dbg) {  // BBMPC_DEBUG=1: synchronise and dump the watchdog record of a starved pipeline
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      uint32_t* h = static_cast<uint32_t*>(ctx->dbg_host);
      for (int b = 0; b < 2; ++b)
        for (int w = 0; w < TC_WARPS; ++w) {
          const uint32_t* r = h + 256 * b + 8 * w;
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

static int launch_rollout_tc_once(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                                  const float* penalty, int rows, int A, int H, int passes, cudaStream_t st);
int launch_rollout_tc(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                      const float* penalty, int rows, int A, int H, int passes, cudaStream_t st) {
  int rc = launch_rollout_tc_once(ctx, states, actions, returns, penalty, rows, A, H, passes, st);
  if (rc == BBMPC_ERETRY) rc = launch_rollout_tc_once(ctx, states, actions, returns, penalty, rows, A, H, passes, st);
  return rc;
}
static int launch_rollout_tc_once(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                                  const float* penalty, int rows, int A, int H, int passes, cudaStream_t st) {
  const ModelHost& m = ctx->model;
  TcParams p{};
  p.mlp = m.mlp; p.norm = m.norm; p.reward_id = ctx->reward_id; p.dS = m.dS; p.dU = m.dU;
  p.states = states; p.actions = actions; p.returns = returns; p.penalty = penalty;
  p.rows = rows; p.A = A; p.H = H; p.passes = passes; p.traj = ctx->traj_cur;
  p.n_tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  const int stage = m.mlp.stage_bytes;
  int buf_w = 0;
  if (!tc_column_map(m.mlp, &buf_w, &p.col_x, &p.col_dout))
    return fail(ctx, BBMPC_EINVAL, "tensor-core path: TMEM column budget exceeded");
  p.col_buf0 = 0;
  p.col_buf1 = 256;
  if (const char* x = getenv("BBMPC_TC_X")) p.xflags = atoi(x);
  if (getenv("BBMPC_TC_TRACE")) {
    static uint32_t* tbuf = nullptr;
    if (!tbuf) { BB_CUDA(ctx, cudaMalloc(&tbuf, 32 * 1024 * 4)); BB_CUDA(ctx, cudaMemset(tbuf, 0, 32 * 1024 * 4)); }
    p.trace = tbuf;
  }
  if (getenv("BBMPC_DEBUG")) {
    if (!ctx->dbg_host) {
      BB_CUDA(ctx, cudaHostAlloc(&ctx->dbg_host, 4096, cudaHostAllocMapped));
      memset(ctx->dbg_host, 0, 4096);
    }
    BB_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&p.dbg), ctx->dbg_host, 0));
  }
  p.stage_bytes = stage;
  const size_t budget = 227 * 1024;
  // Ensembles run member-parallel: the n_members CTAs of a group share a tile.  The mode depends on the
  // model only (never on P), so results are bit-identical for any population size / sharding.
  const int nM = m.mlp.n_members;
  const bool group = nM > 1 && nM <= ctx->sm_count && !ctx->tc_no_groups && !getenv("BBMPC_NO_GROUPS");
  p.group_mode = group ? 1 : 0;
  p.jobs = group ? m.mlp.solo_jobs : m.mlp.jobs;          p.n_jobs = group ? m.mlp.solo_jobs_per_step : m.mlp.jobs_per_step;
  p.table = group ? m.mlp.solo_table : m.mlp.chunk_table; p.n_table = group ? m.mlp.solo_groups_per_step : m.mlp.chunks_per_step;
  {
    const int ds_t = m.dU > 8 ? 32 : (m.dS <= 8 ? 8 : (m.dS <= 24 ? 24 : 32));   // DS_T of the kernel instance launched below
    const int xs = nM * TILE_ROWS * ds_t * 4;
    p.xs_bytes = (group && xs <= 64 * 1024 && !getenv("BBMPC_NO_XS")) ? xs : 0;
  }
  int n_stages = TC_MAX_STAGES;
  const int cps = p.n_table, jps = p.n_jobs;
  while (n_stages > 2 && tc_layout(stage, n_stages, cps, jps, p.xs_bytes).total > budget) --n_stages;
  if (tc_layout(stage, n_stages, cps, jps, p.xs_bytes).total > budget)
    return fail(ctx, BBMPC_EINVAL, "tensor-core path: shared memory budget exceeded");
  p.n_stages = n_stages;
  const size_t smem_bytes = tc_layout(stage, n_stages, cps, jps, p.xs_bytes).total;
  int grid = p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count;
  if (group) {
    int n_groups = ctx->sm_count / nM;
    if (n_groups > p.n_tiles) n_groups = p.n_tiles;
    grid = n_groups * nM;
    const size_t need = static_cast<size_t>(n_groups) * 2 * nM * TILE_ROWS * 32;
    if (ctx->tc_xchg_floats < need) {
      BB_CUDA(ctx, cudaStreamSynchronize(st));
      cudaFree(ctx->tc_xchg);
      BB_CUDA(ctx, cudaMalloc(&ctx->tc_xchg, need * sizeof(float)));
      ctx->tc_xchg_floats = need;
    }
    if (ctx->tc_flags_n < n_groups) {
      BB_CUDA(ctx, cudaStreamSynchronize(st));
      cudaFree(ctx->tc_flags);
      BB_CUDA(ctx, cudaMalloc(&ctx->tc_flags, n_groups * sizeof(unsigned)));
      ctx->tc_flags_n = n_groups;
    }
    BB_CUDA(ctx, cudaMemsetAsync(ctx->tc_flags, 0, n_groups * sizeof(unsigned), st));
    p.xchg = ctx->tc_xchg; p.flags = ctx->tc_flags;
  }
  if (m.dS <= 8 && m.dU <= 8) return launch_t<8, 8>(ctx, p, grid, smem_bytes, st);
  if (m.dS <= 24 && m.dU <= 8) return launch_t<24, 8>(ctx, p, grid, smem_bytes, st);
  if (m.dU <= 8) return launch_t<32, 8>(ctx, p, grid, smem_bytes, st);
  return launch_t<32, 16>(ctx, p, grid, smem_bytes, st);
}

}  // namespace bbmpc
