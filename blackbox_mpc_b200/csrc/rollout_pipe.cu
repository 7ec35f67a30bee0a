// rollout_pipe.cu — pipelined tensor-core (tcgen05 / TMEM) trajectory evaluator: two member-tiles in
// flight per CTA (trajectory_evaluators/deterministic.py:48-77 fused, as rollout_tc.cu).
//
// Why.  One tile of 128 trajectories is a serial chain per horizon step (Dense -> tanh -> Dense -> ... ->
// process_output -> reward -> process_input): in rollout_tc.cu the tensor pipe idles while the conversion
// warps work and during the whole step boundary (34 % active, profiles/r1e_*).  Here the unit of work is
// a JOB = (member-tile, horizon step): a 128-row tile x one ensemble member x one step.  A CTA owns up
// to three member-tiles and runs their jobs round-robin, two at a time: while the conversion warps turn
// the accumulator of job a into the next layer's operand, the tensor pipe contracts a layer of job b, and
// the step boundary of a third member-tile (output exchange, state update, reward) runs on its own warps.
//
// What makes two jobs fit.  The accumulator of a layer stays in TMEM (two D buffers, one per job in
// flight), but the converted activations go to SHARED memory as the bf16 hi/lo K-major A operand of an
// SS-mode tcgen05.mma, in a ring of UNITS (2 K-chunks = 16 KB): the tensor pipe frees a unit
// (tcgen05.commit) as soon as its MMAs retire, and the conversion of the other job refills it, so the
// ring holds little more than one layer (8 units) instead of one layer per job.
//
//   warp 0       weights producer: L2 -> SMEM ring (cp.async.bulk + mbarrier tx bytes), in stage order
//   warp 1       MMA issuer (one lane): layer 0 A-from-TMEM (X region), layers >= 1 A-from-SMEM ring
//   warp 2       TMEM allocation; warp 3 idle
//   warps 4..7   publishers (ensembles only): output accumulator -> exchange buffer in L2, release flag
//   warps 8..11  integrators: per member-tile state / return in registers; sum of the members' outputs,
//                process_output, reward, process_input of the next step -> X region in TMEM
//   warps 12..19 conversion warps, 2 per TMEM lane quarter: tcgen05.ld -> activation -> hi/lo split ->
//                st.shared into the A ring -> fence.proxy.async -> arrive
//
// Every role walks the same static STAGE ORDER (StageSeq): the stages (job, layer) of the two TMEM slots
// alternate, so production and consumption of ring units are FIFO and no role ever waits on a stage that
// is behind it in the order (tests/test_pipe_schedule_cpu.py simulates the barrier protocol).
// A CTA with one member-tile (few-tile launches) alternates the D buffers per LAYER instead and overlaps
// the conversion of layer l with the contraction of layer l+1 chunk by chunk, as rollout_tc.cu does.
//
// Ensembles: member-tile id = tile * n_members + member, CTA = id % grid, so the members of a tile sit on
// different CTAs (cooperative launch) and exchange their raw outputs through L2 once per step; all CTAs
// visit tiles in increasing order, which makes the cross-CTA waits acyclic.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "device_fns.cuh"
#include "pipe_sched.h"
#include "tc05.cuh"
#include "tc_epi.cuh"

namespace bbmpc {
using namespace tc05;

constexpr int PIPE_WARPS = 20;
constexpr int PIPE_THREADS = 32 * PIPE_WARPS;
constexpr int PIPE_MAX_MT = 3;          // member-tiles per CTA and round
constexpr int PIPE_MAX_WSTAGES = 8;
constexpr int PIPE_MAX_UNITS = 16;      // A-ring positions
constexpr int PIPE_ROWS = 128;
// register budgets per warpgroup role.  setmaxnreg moves registers inside the CTA's LAUNCH allocation (640 threads x 96), not the
// SM's file: the budgets must sum to 5 warpgroups x 96 = 480 (64 + 32 + 144 + 2 x 120), or the last setmaxnreg.inc never returns.
constexpr int PIPE_REGS_CTRL = 64, PIPE_REGS_PUB = 32, PIPE_REGS_INT = 144, PIPE_REGS_CONV = 120;

struct PipeParams {
  MlpDev mlp;
  NormDev norm;
  int reward_id, dS, dU;
  const float* states; const float* actions; float* returns; const float* penalty;
  int rows, A, H, n_tiles, passes;
  int stage_bytes, n_wstages;        // weights ring
  int a_units, a_chunk_bytes;        // A ring: positions (units = 2 K-chunks) and bytes per chunk
  int col_d0, col_d1, col_x, x_w;    // TMEM column map
  int max_mt;                        // member-tiles per CTA and round (TMEM X regions available)
  int tiles_per_round, xchg_tiles;
  const TcJob* jobs; const uint2* table; int n_table;
  float* xchg; unsigned* flags;
  float* traj;                       // user reward: visited states [rows][H][dS] (else nullptr)
  float* park;                       // integrators' parked states: [grid][PIPE_MAX_MT][DS_T/4 + 1][128] float4
  uint32_t* trace; int xflags;
  int dfull_polls;                   // conversion warps: non-blocking probes of an accumulator barrier before they sleep on it
  int stagger;                       // cycles by which the odd-chunk conversion warps start late (de-phases compute / store)
  uint32_t* dbg;                     // host-mapped watchdog record (BBMPC_DEBUG=1), else nullptr
};

struct PipeSmem { uint32_t wring, aring, table, jobs, bars, tmem_slot, stats, conv, total; };
constexpr int PIPE_NUM_BARS = 2 * PIPE_MAX_WSTAGES + 2 * PIPE_MAX_UNITS + 8 + PIPE_MAX_MT;
__host__ __device__ inline PipeSmem pipe_layout(int stage_bytes, int n_wstages, int a_units, int a_chunk_bytes, int n_table, int n_layers) {
  PipeSmem L;
  uint32_t off = 0;
  L.wring = off;  off += static_cast<uint32_t>(stage_bytes) * n_wstages;
  off = (off + 127u) & ~127u;
  L.aring = off;  off += static_cast<uint32_t>(a_units) * 2u * a_chunk_bytes;
  L.table = off;  off += static_cast<uint32_t>(n_table) * 8;
  off = (off + 15u) & ~15u;
  L.jobs = off;   off += static_cast<uint32_t>(n_layers) * sizeof(TcJob);
  L.bars = off;   off += PIPE_NUM_BARS * 8;
  off = (off + 15u) & ~15u;
  L.tmem_slot = off; off += 16;
  L.stats = off;  off += (4 * MAX_DS + 2 * MAX_DU + 3 * 64) * 4;
  L.conv = off;   off += MAX_LAYERS * (32 + 256);
  L.total = off;
  return L;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ float ldg_f32(const float* p) {   // volatile: stays where it is written (ahead of the waits)
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
#ifndef PIPE_NO_SETMAXNREG
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
#else   // debugging build: every role at the launch-time register count (spills)
template <int N> __device__ __forceinline__ void setmaxnreg_inc() {}
template <int N> __device__ __forceinline__ void setmaxnreg_dec() {}
#endif

template <bool ON>
struct PTracer {
  uint32_t* buf; uint32_t n; bool on;
  __device__ __forceinline__ void rec(uint32_t tag) {
    if (ON) { if (on && n < 500) { buf[2 * n] = tag; buf[2 * n + 1] = static_cast<uint32_t>(clock64()); ++n; } }
  }
  __device__ __forceinline__ void arm(bool v) { if (ON) on = v; }
};

// One 16-column accumulator chunk -> bf16 hi/lo words of the next layer's A operand (8 packed pairs each).
template <int ACT>
__device__ __forceinline__ void pconv_full(const uint32_t (&r)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  if constexpr (ACT == BBMPC_ACT_TANH && PACKED_TANH) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      tanh_split_quad(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]),
                      hi[2 * j], lo[2 * j], hi[2 * j + 1], lo[2 * j + 1]);
  } else {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    act16<ACT>(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16x2_veltkamp(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  }
}
// One hidden-layer conversion of one warp: chunks c = sub, sub+2, ... of its 32 TMEM lanes.  Measured (ncu source
// page, profiles/r2*): with two conversion warps per scheduler the loop is bound by the dependent-issue latency of
// one chunk's tanh + hi/lo split (about 120 instructions, 300 cycles of fixed stalls), not by a pipe.  So a warp
// converts TWO chunks per iteration (the compiler interleaves their eight quads: twice the instruction-level
// parallelism), ring and barrier addresses advance by constants, and EVERY lane arrives on a unit's barrier after its
// own stores and proxy fence (no warp-level sync or lane election on the way).
template <int ACT, bool TR>
__device__ __forceinline__ void pipe_convert(uint32_t taddr, int Npad, int N, int n_a_chunks, int sub, int passes,
                                            uint32_t aring, uint32_t chunk_bytes, uint32_t a_units, uint32_t pu, uint32_t wrap,
                                            uint32_t bar_afull, uint32_t bar_afree, uint32_t bar_drained, int row, int lane,
                                            const float* tail_tab, int stagger, uint32_t polls, volatile uint32_t* dbgp, PTracer<TR>& tr) {
  const int n_full = N >> 4;        // chunks whose 16 columns are all real features
  const int n_data = Npad >> 4;     // chunks that carry accumulator data at all
  if (sub && stagger > 0) {
    // The two warps of a quarter share one scheduler and one MUFU pipe: started together they compute together and
    // store together; half a step apart one converts while the other waits for its stores / ring units.
    const long long t0 = clock64();
    while (clock64() - t0 < stagger) {}
  }
  const uint32_t unit_bytes = 2 * chunk_bytes;
  uint32_t slot = aring + pu * unit_bytes + static_cast<uint32_t>(sub) * chunk_bytes + static_cast<uint32_t>(row) * 16;
  uint32_t bf = bar_afull + 8 * pu;
  const uint32_t bf_end = bar_afull + 8 * a_units, free_off = bar_afree - bar_afull;
  auto store = [&](const uint32_t (&hi)[8], const uint32_t (&lo)[8], uint32_t at) {
    st_shared_v4(at, hi[0], hi[1], hi[2], hi[3]);
    st_shared_v4(at + PIPE_ROWS * 16, hi[4], hi[5], hi[6], hi[7]);
    if (passes == 3) {
      st_shared_v4(at + 2 * PIPE_ROWS * 16, lo[0], lo[1], lo[2], lo[3]);
      st_shared_v4(at + 3 * PIPE_ROWS * 16, lo[4], lo[5], lo[6], lo[7]);
    }
  };
  auto advance = [&]() {
    slot += unit_bytes; bf += 8;
    if (bf == bf_end) { slot -= a_units * unit_bytes; bf = bar_afull; ++wrap; }
  };
  auto drained = [&]() {   // this warp's last read of the accumulator is complete: the D buffer may be overwritten
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_drained);
  };
  int c = sub;
  // ---- pairs of full chunks
  for (; c + 2 < n_full; c += 4) {
    tr.rec(0x100u | c);
    uint32_t ra[16], rb[16];
    tmem_ld16(taddr + 16 * c, ra);
    tmem_ld16(taddr + 16 * (c + 2), rb);
    // probe the two ring units now (non-blocking); the answers are needed only after the conversion
    uint32_t ok0 = 1, ok1 = 1;
    {
      uint32_t bf1 = bf + 8, wrap1 = wrap;
      if (bf1 == bf_end) { bf1 = bar_afull; ++wrap1; }
      if (wrap > 0) ok0 = mbar_test_wait(bf + free_off, (wrap - 1) & 1u);
      if (wrap1 > 0) ok1 = mbar_test_wait(bf1 + free_off, (wrap1 - 1) & 1u);
    }
    wait_ld();
    if (c + 4 >= n_data) drained();
    uint32_t hia[8], loa[8], hib[8], lob[8];
    pconv_full<ACT>(ra, hia, loa);
    pconv_full<ACT>(rb, hib, lob);
    tr.rec(0x200u | c);
    if (!ok0) mbar_wait_poll_then_sleep(bf + free_off, (wrap - 1) & 1u, polls, dbgp, 0x8000000u | (c << 8));
    const uint32_t slot0 = slot, bf0 = bf;
    store(hia, loa, slot0);
    advance();
    if (!ok1) mbar_wait_poll_then_sleep(bf + free_off, (wrap - 1) & 1u, polls, dbgp, 0x8100000u | (c << 8));
    store(hib, lob, slot);
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
    mbar_arrive(bf0);   // (per-unit arrivals in every mode: the barrier phases must stay in step with the ring across rounds)
    if (((c + 2) ^ 1) < n_a_chunks) mbar_arrive(bf); else mbar_arrive_n(bf, 2);
    advance();
    tr.rec(0x300u | c);
  }
  // ---- remaining chunks one at a time (at most one full chunk, then the trailing chunks with ones-columns / padding)
  for (; c < n_a_chunks; c += 2) {
    tr.rec(0x100u | c);
    uint32_t r[16];
    if (c < n_data) { tmem_ld16(taddr + 16 * c, r); wait_ld(); if (c + 2 >= n_data) drained(); }
    uint32_t hi[8], lo[8];
    if (c < n_full) pconv_full<ACT>(r, hi, lo);
    else tail16<ACT>(r, c < n_data, N - 16 * c, tail_tab + 32 * (c - n_full), hi, lo);
    tr.rec(0x200u | c);
    if (wrap > 0) mbar_wait_poll_then_sleep(bf + free_off, (wrap - 1) & 1u, polls, dbgp, 0x8200000u | (c << 8));
    store(hi, lo, slot);
    fence_proxy_async_smem();
    if ((c ^ 1) < n_a_chunks) mbar_arrive(bf); else mbar_arrive_n(bf, 2);   // a lone last chunk stands for its missing partner
    advance();
    tr.rec(0x300u | c);
  }
  if (sub >= n_data) drained();   // no chunk of this warp carried accumulator data
}

template <int DS_T, int DU_T, bool TR, int ACT_T>
__global__ void __launch_bounds__(PIPE_THREADS, 1) rollout_pipe_kernel(const __grid_constant__ PipeParams p) {
  volatile uint32_t* const dbgp = TR ? p.dbg : nullptr;   // watchdog records only in the debug / trace build
  extern __shared__ __align__(128) uint8_t smem[];
  const MlpDev& M = p.mlp;
  const int nL = M.n_layers, nM = M.n_members;
  const PipeSmem lay = pipe_layout(p.stage_bytes, p.n_wstages, p.a_units, p.a_chunk_bytes, p.n_table, nL);
  const uint32_t smem_base = smem_u32(smem);
  const uint2* table = reinterpret_cast<const uint2*>(smem + lay.table);
  const TcJob* jobs = reinterpret_cast<const TcJob*>(smem + lay.jobs);
  const uint32_t bar_wfull = smem_base + lay.bars;
  const uint32_t bar_wempty = bar_wfull + PIPE_MAX_WSTAGES * 8;
  const uint32_t bar_afull = bar_wempty + PIPE_MAX_WSTAGES * 8;    // conversion -> MMA: ring unit written (256 arrivals: every lane of the 8 warps)
  const uint32_t bar_afree = bar_afull + PIPE_MAX_UNITS * 8;       // MMA -> conversion: ring unit consumed (commit)
  const uint32_t bar_dfull = bar_afree + PIPE_MAX_UNITS * 8;       // MMA -> conversion: hidden accumulator in D[b] complete
  const uint32_t bar_dout = bar_dfull + 16;                        // MMA -> publishers / integrators: output accumulator in D[b]
  const uint32_t bar_drained = bar_dout + 16;                      // readers -> MMA: D[b] has been read (8 arrivals)
  const uint32_t bar_adone = bar_drained + 16;                     // conversion -> MMA (two jobs in flight): the whole A operand of slot b is in the ring (8 arrivals)
  const uint32_t bar_xfull = bar_adone + 16;                       // integrators -> MMA: X region i written (4 arrivals)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + lay.tmem_slot);
  float* st_mean_t = reinterpret_cast<float*>(smem + lay.stats);
  float* st_den_t = st_mean_t + MAX_DS;
  float* xt_mean = st_den_t + MAX_DS;      // per K position of the layer-0 operand: x = (src - mean) * rden + add
  float* xt_rden = xt_mean + 64;
  float* xt_add = xt_rden + 64;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool norm_on = p.norm.enabled != 0;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < p.n_table; i += PIPE_THREADS) reinterpret_cast<uint2*>(smem + lay.table)[i] = p.table[i];
  for (int i = tid; i < nL * static_cast<int>(sizeof(TcJob) / 16); i += PIPE_THREADS)
    reinterpret_cast<uint4*>(smem + lay.jobs)[i] = reinterpret_cast<const uint4*>(p.jobs)[i];
  for (int i = tid; i < MAX_DS; i += PIPE_THREADS) {
    const bool in = norm_on && i < p.dS;
    st_mean_t[i] = in ? p.norm.mean_t[i] : 0.0f;
    st_den_t[i] = in ? p.norm.den_t[i] : (i < p.dS ? 1.0f : 0.0f);
  }
  for (int k = tid; k < 64; k += PIPE_THREADS) {   // operand layout: [actions (DU_T slots) | state | 1 1 1 | 0 ...]
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
    for (int s = 0; s < p.n_wstages; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
    for (int u = 0; u < p.a_units; ++u) { mbar_init(bar_afull + 8 * u, 256); mbar_init(bar_afree + 8 * u, 1); }   // 2 chunks x 4 quarters x 32 lanes
    for (int b = 0; b < 2; ++b) { mbar_init(bar_dfull + 8 * b, 1); mbar_init(bar_dout + 8 * b, 1); mbar_init(bar_drained + 8 * b, 8); }
    for (int b = 0; b < 2; ++b) mbar_init(bar_adone + 8 * b, 8);
    for (int i = 0; i < PIPE_MAX_MT; ++i) mbar_init(bar_xfull + 8 * i, 4);
    fence_mbar_init();
    int* cv = reinterpret_cast<int*>(smem + lay.conv);
    for (int l = 0; l + 1 < nL; ++l) {
      cv[8 * l + 0] = M.layer[l].Npad; cv[8 * l + 1] = M.layer[l].N; cv[8 * l + 2] = M.layer[l].act;
      cv[8 * l + 3] = M.layer[l + 1].Kpad >> 4;                  // A-operand chunks of the next layer
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
  const int G = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  const int n_rounds = (p.n_tiles + p.tiles_per_round - 1) / p.tiles_per_round;

  // member-tiles of this CTA in round r: local ids cta, cta + G, ... (id = local tile * nM + member)
  auto round_count = [&](int r) -> int {
    const int t0 = r * p.tiles_per_round;
    const int nt = (p.n_tiles - t0 < p.tiles_per_round) ? p.n_tiles - t0 : p.tiles_per_round;
    const int ids = nt * nM - cta;
    int n = ids <= 0 ? 0 : (ids + G - 1) / G;
    return n < p.max_mt ? n : p.max_mt;
  };
  auto mt_tile = [&](int r, int i) -> int { return r * p.tiles_per_round + (cta + i * G) / nM; };
  auto mt_member = [&](int i) -> int { return (cta + i * G) % nM; };

  if (warp < 4) {
    setmaxnreg_dec<PIPE_REGS_CTRL>();
    if (warp == 0 && lane == 0) {
      // ============================================================ weights producer
      int first_group[MAX_LAYERS];
      { int g = 0; for (int l = 0; l < nL; ++l) { first_group[l] = g; g += static_cast<int>(jobs[l].ngroups); } }
      uint32_t stage = 0, phase = 0;
      for (int r = 0; r < n_rounds; ++r) {
        const int n_mt = round_count(r);
        if (n_mt == 0) continue;
        StageSeq seq; seq.init(n_mt * p.H, nL, n_mt);
        int j, l, b;
        while (seq.next(j, l, b)) {
          const uint8_t* wimg = M.wimg + static_cast<size_t>(mt_member(j % n_mt)) * M.img_member_stride;
          const int g0 = first_group[l], g1 = g0 + static_cast<int>(jobs[l].ngroups);
          for (int g = g0; g < g1; ++g) {
            const uint2 e = table[g];
            mbar_wait_sleep(bar_wempty + 8 * stage, phase ^ 1, dbgp, 0x6000000u);
            mbar_arrive_expect_tx(bar_wfull + 8 * stage, e.y);
            bulk_g2s(smem_base + lay.wring + stage * p.stage_bytes, wimg + e.x, e.y, bar_wfull + 8 * stage);
            if (++stage == static_cast<uint32_t>(p.n_wstages)) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ============================================================ MMA issuer
      // The whole warp runs the (warp-uniform) control flow and one elected lane issues the tcgen05 instructions of a
      // unit back to back.  (A loop executed by a single lane makes ptxas wrap EVERY tcgen05.mma in an elect /
      // broadcast / vote sequence of ~60 stall cycles that also waits for the previous MMA to read its operands:
      // measured ~200 cycles per MMA instead of ~50, profiles/r2*.)
      const bool three = (p.passes == 3);
      uint32_t stage = 0, phase = 0;
      uint32_t upos = 0, uwrap = 0;         // ring position / wrap count of the next unit to allocate (mirrors the conversion warps)
      uint32_t dw0 = 0, dw1 = 0;            // stages issued into D[0] / D[1]
      uint32_t xc0 = 0, xc1 = 0, xc2 = 0;   // layer-0 inputs consumed per X region (barrier phases persist across rounds)
      const uint32_t wring16 = (smem_base + lay.wring) >> 4, stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4;
      const uint32_t aring16 = (smem_base + lay.aring) >> 4, achunk16 = static_cast<uint32_t>(p.a_chunk_bytes) >> 4;
      constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);          // SBO = 128 B, descriptor version 1, no swizzle
      const uint32_t a_desc_lo_base = ((PIPE_ROWS * 16u) >> 4) << 16;  // LBO = 2048 B between the two K slabs of a chunk
      const uint32_t a_lo_off16 = (2u * PIPE_ROWS * 16u) >> 4;         // bf16-lo image of a chunk
      PTracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};
      for (int r = 0; r < n_rounds; ++r) {
        const int n_mt = round_count(r);
        if (n_mt == 0) continue;
        uint32_t ab0p = 0, ab0w = 0, ab1p = 0, ab1w = 0;   // first ring unit (position, wrap) of the pending A operand per slot (slot 0 in single mode)
        StageSeq seq; seq.init(n_mt * p.H, nL, n_mt);
        int j, l, b;
        int i;
        while (seq.next(j, l, b, i)) {
          if (TR) tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && j / n_mt == 2);
          const TcJob job = jobs[l];
          const uint32_t d = tmem_base + (b ? p.col_d1 : p.col_d0);
          {
            const uint32_t dw = b ? dw1 : dw0;
            if (dw > 0) mbar_wait(bar_drained + 8 * b, (dw - 1) & 1u, dbgp, 0x1000000u | (j << 8) | l);
            if (b) ++dw1; else ++dw0;
          }
          if (l == 0) {
            const uint32_t xc = i == 0 ? xc0 : (i == 1 ? xc1 : xc2);
            mbar_wait(bar_xfull + 8 * i, xc & 1u, dbgp, 0x1100000u | (j << 8));
            if (i == 0) ++xc0; else if (i == 1) ++xc1; else ++xc2;
          }
          const bool key1 = !seq.single && b;
          uint32_t pu = key1 ? ab1p : ab0p, wrap = key1 ? ab1w : ab0w;   // first ring unit of this stage's A operand
          if (l + 1 < nL) {   // ring units of the operand the conversion warps will produce from this stage's accumulator
            if (key1) { ab1p = upos; ab1w = uwrap; } else { ab0p = upos; ab0w = uwrap; }
            upos += (static_cast<uint32_t>(jobs[l + 1].nchunks) + 1u) >> 1;
            if (upos >= static_cast<uint32_t>(p.a_units)) { upos -= p.a_units; ++uwrap; }
          }
          fence_after_sync();
          tr.rec(0x1000u | (l << 8) | (b << 4));
          uint32_t acc = 0, u = 0;
          const uint32_t n_units = (job.nchunks + 1u) >> 1;
          const uint32_t xcol = tmem_base + p.col_x + i * p.x_w;
          // (Waiting once for the whole converted operand instead of per ring unit measured slower — 1.356 vs 1.237 ms per
          // C4 rollout: the stage then starts ~2k cycles later, after the last conversion step — and was removed.)
          // Readiness of the CURRENT unit's weight stage / ring unit, learnt from the probe issued with the previous unit
          // (ok_w only matters at the first unit of a group, ok_a at every unit of a layer >= 1 in single mode).
          uint32_t ok_w = 0, ok_a = 0;
          if (l > 0 && three) {
            // Layers >= 1: flat loop over the ring units, addresses advance by constants (the issuing warp shares its
            // scheduler with two conversion warps: every instruction of this loop is paid ~10 cycles).  A weight group
            // (ring stage) holds n >= 1 units: wide layers one, the output layer several.
            uint32_t alo = a_desc_lo_base | ((aring16 + 2 * pu * achunk16) & 0x3FFFu);
            const uint32_t alo0 = a_desc_lo_base | (aring16 & 0x3FFFu), astep = 2 * achunk16, bstep = 2 * job.chunk16;
            uint32_t gsz = job.gsz, n = gsz & 15u, kk = 0;
            uint32_t blo = job.desc_lo_base | ((wring16 + stage * stage16) & 0x3FFFu);
            for (; u < n_units; ++u) {
              if (kk == 0 && !ok_w) mbar_wait(bar_wfull + 8 * stage, phase, dbgp, 0x3000000u | (l << 8) | u);
              if (!ok_a) mbar_wait_poll(bar_afull + 8 * pu, wrap & 1u, dbgp, 0x2000000u | (j << 12) | (l << 8) | u);
              fence_after_sync();
              tr.rec(0x4200u | u);
              const bool last = (kk + 1 == n), more = u + 1 < n_units;
              uint32_t nstage = stage + 1, nphase = phase, pn = pu + 1, wn = wrap;
              if (nstage == static_cast<uint32_t>(p.n_wstages)) { nstage = 0; nphase ^= 1; }
              if (pn == static_cast<uint32_t>(p.a_units)) { pn = 0; ++wn; }
              uint32_t pw = 0, pa = 0;
              if (2 * u + 1 < job.nchunks)
                mma_unit_ss_probe_full(d, alo, blo, DESC_HI, a_lo_off16, job.lo_off16, achunk16, job.chunk16, job.idesc, acc,
                                       bar_afree + 8 * pu, last ? bar_wempty + 8 * stage : 0u, bar_wfull + 8 * nstage, nphase,
                                       bar_afull + 8 * pn, wn & 1u, pw, pa);
              else
                mma_unit_ss_probe_half(d, alo, blo, DESC_HI, a_lo_off16, job.lo_off16, job.idesc, acc, bar_afree + 8 * pu,
                                       last ? bar_wempty + 8 * stage : 0u, bar_wfull + 8 * nstage, nphase, bar_afull + 8 * pn, wn & 1u, pw, pa);
              ok_a = more ? pa : 0u;
              alo = pn ? alo + astep : alo0;
              pu = pn; wrap = wn; acc = 1u;
              if (last) {
                ok_w = more ? pw : 0u;
                stage = nstage; phase = nphase;
                gsz >>= 4; n = gsz & 15u; kk = 0;
                blo = job.desc_lo_base | ((wring16 + stage * stage16) & 0x3FFFu);
              } else {
                ++kk; blo += bstep;
              }
            }
          } else
          for (uint32_t g = 0; g < job.ngroups; ++g) {
            const uint32_t n = (job.gsz >> (4 * g)) & 15u;
            tr.rec(0x4000u | g);
            if (!ok_w) mbar_wait(bar_wfull + 8 * stage, phase, dbgp, 0x3000000u | (l << 8) | g);
            tr.rec(0x4100u | g);
            const uint32_t sbase = wring16 + stage * stage16;
            uint32_t nstage = stage + 1, nphase = phase;            // next group's weight stage (probed with this group's last unit)
            if (nstage == static_cast<uint32_t>(p.n_wstages)) { nstage = 0; nphase ^= 1; }
            for (uint32_t k = 0; k < n; ++k, ++u) {
              const bool two = 2 * u + 1 < job.nchunks;
              const uint32_t blo = job.desc_lo_base | ((sbase + 2 * k * job.chunk16) & 0x3FFFu);
              const uint64_t b0 = (static_cast<uint64_t>(DESC_HI) << 32) | blo;
              const bool last_of_group = (k + 1 == n);
              if (l == 0) {
                fence_after_sync();
                if (elect_one()) {
                  const uint32_t a0 = xcol + 32 * u;
                  mma_ts(d, a0, b0, job.idesc, acc);
                  if (three) { mma_ts(d, a0 + 8, b0, job.idesc, 1u); mma_ts(d, a0, b0 + job.lo_off16, job.idesc, 1u); }
                  if (two) {
                    const uint64_t b1 = b0 + job.chunk16;
                    mma_ts(d, a0 + 16, b1, job.idesc, 1u);
                    if (three) { mma_ts(d, a0 + 24, b1, job.idesc, 1u); mma_ts(d, a0 + 16, b1 + job.lo_off16, job.idesc, 1u); }
                  }
                  if (last_of_group) mma_commit(bar_wempty + 8 * stage);
                }
                __syncwarp();
                ok_w = 0;
              } else {
                if (!ok_a) mbar_wait_poll(bar_afull + 8 * pu, wrap & 1u, dbgp, 0x2000000u | (j << 12) | (l << 8) | u);
                fence_after_sync();
                tr.rec(0x4200u | u);
                uint32_t pn = pu + 1, wn = wrap;
                if (pn == static_cast<uint32_t>(p.a_units)) { pn = 0; ++wn; }
                const uint32_t alo = a_desc_lo_base | ((aring16 + 2 * pu * achunk16) & 0x3FFFu);
                const uint64_t a0 = (static_cast<uint64_t>(DESC_HI) << 32) | alo;
                // (single-pass precision only: layers >= 1 with three passes take the flat loop above)
                uint32_t pw = 0, pa = 0;
                mma_unit_ss_probe(d, a0, b0, a_lo_off16, job.lo_off16, achunk16, job.chunk16, job.idesc, acc, three ? 1u : 0u, two ? 1u : 0u,
                                  bar_afree + 8 * pu, last_of_group ? bar_wempty + 8 * stage : 0u,
                                  bar_wfull + 8 * nstage, nphase, bar_afull + 8 * pn, wn & 1u, pw, pa);
                ok_w = (last_of_group && g + 1 < job.ngroups) ? pw : 0u;
                ok_a = (u + 1 < n_units) ? pa : 0u;
                pu = pn; wrap = wn;
              }
              acc = 1u;
            }
            stage = nstage; phase = nphase;
          }
          if (elect_one()) mma_commit((l + 1 < nL ? bar_dfull : bar_dout) + 8 * b);
          __syncwarp();
          tr.rec(0x1800u | (l << 8) | (b << 4));
        }
      }
    }
  } else if (warp < 8) {
    setmaxnreg_dec<PIPE_REGS_PUB>();
    // ============================================================ publishers (ensembles only)
    // Output accumulator of this CTA's member -> exchange buffer of the tile in L2; the last of the 128 threads
    // to get there releases one count on the tile's flag.  Stateless: never blocked by a peer CTA.
    if (nM > 1) {
      const int q = warp & 3;
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const int row_in_tile = q * 32 + lane;
      uint32_t oc0 = 0, oc1 = 0;
      PTracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};
      for (int r = 0; r < n_rounds; ++r) {
        const int n_mt = round_count(r);
        if (n_mt == 0) continue;
        const int J = n_mt * p.H;
        const int b_single = (nL - 1) & 1;
        for (int j = 0; j < J; ++j) {
          const int i = j % n_mt, t = j / n_mt;
          tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && t == 2);
          const int b = (n_mt == 1) ? b_single : (j & 1);
          mbar_wait_poll_then_sleep(bar_dout + 8 * b, (b ? oc1 : oc0) & 1u, p.dfull_polls, dbgp, 0x5000000u | j);
          if (b) ++oc1; else ++oc0;
          fence_after_sync();
          tr.rec(0x30u);
          uint32_t rr[DS_T / 8][8];
          const uint32_t dcol = tmem_base + lane_off + (b ? p.col_d1 : p.col_d0);
#pragma unroll
          for (int c = 0; c < DS_T / 8; ++c) tmem_ld8(dcol + 8 * c, rr[c]);
          wait_ld();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_n(bar_drained + 8 * b, 2);
          const int xi = mt_tile(r, i) % p.xchg_tiles;
          float* base = p.xchg + ((static_cast<size_t>(xi) * 2 + (t & 1)) * nM + mt_member(i)) * (DS_T * PIPE_ROWS);
          float4* mine = reinterpret_cast<float4*>(base + static_cast<size_t>(row_in_tile) * DS_T);
#pragma unroll
          for (int c = 0; c < DS_T / 4; ++c)
            __stcg(mine + c, make_float4(__uint_as_float(rr[c / 2][4 * (c & 1)]), __uint_as_float(rr[c / 2][4 * (c & 1) + 1]),
                                         __uint_as_float(rr[c / 2][4 * (c & 1) + 2]), __uint_as_float(rr[c / 2][4 * (c & 1) + 3])));
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (warp == 4 && lane == 0)
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + xi) : "memory");
          tr.rec(0x31u);
        }
      }
    }
  } else if (warp < 12) {
    setmaxnreg_inc<PIPE_REGS_INT>();
    // ============================================================ integrators (one warp per TMEM lane quarter)
    // Per member-tile: state and return of 32 rows in registers.  Per job: wait for the tile's outputs of this step
    // (all members), sum in member order, process_output, reward, then process_input of the next step -> X region.
    const int q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const float inv_members = __frcp_rn(static_cast<float>(nM));
    const int kp0_chunks = M.layer[0].Kpad >> 4;
    uint32_t oc0 = 0, oc1 = 0;
    PTracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};

    // process_input: X = [norm(a) (DU_T slots) | norm(s) | 1 1 1 | 0...] -> TMEM region of member-tile i
    auto build_x = [&](const float (&s)[DS_T], const float (&a)[DU_T], int i) {
      constexpr int KP0_T = (DU_T + DS_T + BIAS_COLS + 15) / 16;
      const uint32_t xcol = tmem_base + lane_off + p.col_x + i * p.x_w;
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
            for (int jj = 0; jj < 4; ++jj) {
              const int k = 16 * c + j4 + jj;        // compile-time position -> register of a[] / s[]
              src[jj] = (k < DU_T) ? a[k < DU_T ? k : 0] : ((k - DU_T < DS_T) ? s[(k - DU_T < DS_T && k >= DU_T) ? k - DU_T : 0] : 0.0f);
            }
            const float x0 = fmaf(__fsub_rn(src[0], mn.x), rd.x, ad.x), x1 = fmaf(__fsub_rn(src[1], mn.y), rd.y, ad.y);
            const float x2 = fmaf(__fsub_rn(src[2], mn.z), rd.z, ad.z), x3 = fmaf(__fsub_rn(src[3], mn.w), rd.w, ad.w);
            split_bf16x2_packed(pk2(x0, x1), hi[j4 / 2], lo[j4 / 2]);
            split_bf16x2_packed(pk2(x2, x3), hi[j4 / 2 + 1], lo[j4 / 2 + 1]);
          }
          tmem_st8(xcol + 16 * c, hi);
          if (p.passes == 3) tmem_st8(xcol + 16 * c + 8, lo);
        }
      }
      wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_xfull + 8 * i);
    };
    auto load_actions = [&](float (&a)[DU_T], const float* arow_ptr, int t) {
#pragma unroll
      for (int k = 0; k < DU_T; ++k) a[k] = (k < p.dU) ? ldg_f32(arow_ptr + static_cast<size_t>(t) * p.dU + k) : 0.0f;
    };

    for (int r = 0; r < n_rounds; ++r) {
      const int n_mt = round_count(r);
      if (n_mt == 0) continue;
      const int J = n_mt * p.H;
      const int b_single = (nL - 1) & 1;
      // State and return of the member-tile a job belongs to live in registers while its boundary runs; with several
      // member-tiles per CTA they are parked in a per-CTA scratch in L2 between two of its jobs ([i][k/4][row][4] floats,
      // coalesced; the load is issued at the top of the boundary, ahead of the exchange wait).
      float4* const park = reinterpret_cast<float4*>(p.park) + static_cast<size_t>(cta) * PIPE_MAX_MT * (DS_T / 4 + 1) * PIPE_ROWS;
      float s[DS_T];
      float ret = 0.0f;
      // initial state and the layer-0 operand of step 0
      for (int i = 0; i < n_mt; ++i) {
        const int row = mt_tile(r, i) * PIPE_ROWS + row_in_tile;
        const int arow = row < p.rows ? row : 0;
        const float* srow_ptr = p.states + static_cast<size_t>(arow % p.A) * p.dS;
#pragma unroll
        for (int k = 0; k < DS_T; ++k) s[k] = (k < p.dS) ? srow_ptr[k] : 0.0f;
        float a0[DU_T];
        load_actions(a0, p.actions + static_cast<size_t>(arow) * p.H * p.dU, 0);
        build_x(s, a0, i);
        if (n_mt > 1) {
          float4* pk = park + static_cast<size_t>(i) * (DS_T / 4 + 1) * PIPE_ROWS + row_in_tile;
#pragma unroll
          for (int g = 0; g < DS_T / 4; ++g) __stcg(pk + g * PIPE_ROWS, make_float4(s[4 * g], s[4 * g + 1], s[4 * g + 2], s[4 * g + 3]));
          __stcg(pk + (DS_T / 4) * PIPE_ROWS, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        }
      }
      // one step boundary of member-tile i
      auto boundary = [&](int i, int j, int t) {
        const int tile_i = mt_tile(r, i);
        const int row = tile_i * PIPE_ROWS + row_in_tile;
        const bool valid = row < p.rows;
        const int arow = valid ? row : 0;
        const float* arow_ptr = p.actions + static_cast<size_t>(arow) * p.H * p.dU;
        float a[DU_T], an[DU_T];
        load_actions(a, arow_ptr, t);
        if (t + 1 < p.H) load_actions(an, arow_ptr, t + 1);
        float4* pk = park + static_cast<size_t>(i) * (DS_T / 4 + 1) * PIPE_ROWS + row_in_tile;
        if (n_mt > 1) {
#pragma unroll
          for (int g = 0; g < DS_T / 4; ++g) {
            const float4 v = __ldcg(pk + g * PIPE_ROWS);
            s[4 * g] = v.x; s[4 * g + 1] = v.y; s[4 * g + 2] = v.z; s[4 * g + 3] = v.w;
          }
          ret = __ldcg(pk + (DS_T / 4) * PIPE_ROWS).x;
        }
        float s2[DS_T];
        tr.rec(0x40u | i);
        if (nM > 1) {
          const int xi = tile_i % p.xchg_tiles;
          const unsigned target = static_cast<unsigned>(nM) * (static_cast<unsigned>(tile_i / p.xchg_tiles) * p.H + t + 1);
          // one lane polls with relaxed loads and a short sleep in between (an acquire load per iteration costs an L1
          // invalidate, CCTL.IVALL, each time: 22 M of them per launch in the first version); one acquire at the end
          if (lane == 0) {
            unsigned seen = 0, spins = 0;
            for (;;) {
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + xi) : "memory");
              if (seen >= target) break;
              __nanosleep(100);
              if (++spins > (1u << 22)) {
                if (dbgp) { dbgp[8 * warp] = 0xF1A60000u | threadIdx.x; dbgp[8 * warp + 1] = xi; dbgp[8 * warp + 2] = seen; dbgp[8 * warp + 3] = target; __threadfence_system(); }
                asm volatile("trap;");
              }
            }
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + xi) : "memory");
          }
          __syncwarp();
          tr.rec(0x41u);
          const float* base = p.xchg + ((static_cast<size_t>(xi) * 2 + (t & 1)) * nM) * (DS_T * PIPE_ROWS) + static_cast<size_t>(row_in_tile) * DS_T;
          // members summed in member order (identical on every CTA of the tile), 8 columns (2 x 16 B per member) at a time
#pragma unroll
          for (int c = 0; c < DS_T / 8; ++c) {
            float4 v[2][8];
#pragma unroll
            for (int m2 = 0; m2 < 8; ++m2)
              if (m2 < nM) {
                const float4* src = reinterpret_cast<const float4*>(base + static_cast<size_t>(m2) * (DS_T * PIPE_ROWS)) + 2 * c;
                v[0][m2] = __ldcg(src); v[1][m2] = __ldcg(src + 1);
              }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float y0 = 0.0f, y1 = 0.0f, y2 = 0.0f, y3 = 0.0f;
#pragma unroll
              for (int m2 = 0; m2 < 8; ++m2)
                if (m2 < nM) { y0 = __fadd_rn(y0, v[h][m2].x); y1 = __fadd_rn(y1, v[h][m2].y); y2 = __fadd_rn(y2, v[h][m2].z); y3 = __fadd_rn(y3, v[h][m2].w); }
              for (int m2 = 8; m2 < nM; ++m2) {   // ensembles wider than 8 members (rare)
                const float4 w = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(m2) * (DS_T * PIPE_ROWS)) + 2 * c + h);
                y0 = __fadd_rn(y0, w.x); y1 = __fadd_rn(y1, w.y); y2 = __fadd_rn(y2, w.z); y3 = __fadd_rn(y3, w.w);
              }
              const int k = 8 * c + 4 * h;
              s2[k] = __fadd_rn(fmaf(__fmul_rn(y0, inv_members), st_den_t[k], st_mean_t[k]), s[k]);
              s2[k + 1] = __fadd_rn(fmaf(__fmul_rn(y1, inv_members), st_den_t[k + 1], st_mean_t[k + 1]), s[k + 1]);
              s2[k + 2] = __fadd_rn(fmaf(__fmul_rn(y2, inv_members), st_den_t[k + 2], st_mean_t[k + 2]), s[k + 2]);
              s2[k + 3] = __fadd_rn(fmaf(__fmul_rn(y3, inv_members), st_den_t[k + 3], st_mean_t[k + 3]), s[k + 3]);
            }
          }
          fence_after_sync();
        } else {
          const int b = (n_mt == 1) ? b_single : (j & 1);
          mbar_wait_poll_then_sleep(bar_dout + 8 * b, (b ? oc1 : oc0) & 1u, p.dfull_polls, dbgp, 0x5000000u | j);
          if (b) ++oc1; else ++oc0;
          fence_after_sync();
          tr.rec(0x41u);
          uint32_t rr[DS_T / 8][8];
          const uint32_t dcol = tmem_base + lane_off + (b ? p.col_d1 : p.col_d0);
#pragma unroll
          for (int c = 0; c < DS_T / 8; ++c) tmem_ld8(dcol + 8 * c, rr[c]);
          wait_ld();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_n(bar_drained + 8 * b, 2);
          const int oact = M.layer[nL - 1].act;
#pragma unroll
          for (int k = 0; k < DS_T; ++k) {
            float y = __uint_as_float(rr[k / 8][k % 8]);
            if (oact != BBMPC_ACT_NONE) y = act_fast(y, oact);
            s2[k] = __fadd_rn(fmaf(y, st_den_t[k], st_mean_t[k]), s[k]);   // tables hold (0, 1) without normalisation, (0, 0) beyond dS
          }
        }
        tr.rec(0x42u);
        if (p.traj && valid && mt_member(i) == 0) {   // user reward: dump the visited state (user_reward.cu)
          float* tp = p.traj + (static_cast<size_t>(row) * p.H + t) * p.dS;
#pragma unroll
          for (int k = 0; k < DS_T; ++k) if (k < p.dS) tp[k] = s2[k];
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
            for (int k = 0; k < DU_T; ++k) if (k < p.dU) ss = __fadd_rn(ss, __fmul_rn(a[k], a[k]));
            r_t = __fsub_rn(r_t, __fmul_rn(0.0f, ss));
          }
        } else if (p.reward_id == BBMPC_REWARD_PENDULUM || p.reward_id == BBMPC_REWARD_PENDULUM_GYM) {
          const float pi = 3.14159274101257324f, two_pi = 6.28318548202514648f;
          const float ang = __fsub_rn(floormod_f(__fadd_rn(atan2f(s[1], s[0]), pi), two_pi), pi);
          float ss = 0.0f;
          if (p.reward_id == BBMPC_REWARD_PENDULUM) {  // `actions` parameter receives next_state
#pragma unroll
            for (int k = 0; k < DS_T; ++k) if (k < p.dS) ss = __fadd_rn(ss, __fmul_rn(s2[k], s2[k]));
          } else {
#pragma unroll
            for (int k = 0; k < DU_T; ++k) if (k < p.dU) ss = __fadd_rn(ss, __fmul_rn(a[k], a[k]));
          }
          const float sc = __fadd_rn(__fmul_rn(ang, ang), __fmul_rn(0.1f, __fmul_rn(s[2], s[2])));
          r_t = __fsub_rn(-sc, __fmul_rn(0.001f, ss));
        }
        ret = __fadd_rn(ret, r_t);
#pragma unroll
        for (int k = 0; k < DS_T; ++k) s[k] = s2[k];
        if (t + 1 < p.H) {
          build_x(s, an, i);
          if (n_mt > 1) {
#pragma unroll
            for (int g = 0; g < DS_T / 4; ++g) __stcg(pk + g * PIPE_ROWS, make_float4(s[4 * g], s[4 * g + 1], s[4 * g + 2], s[4 * g + 3]));
            __stcg(pk + (DS_T / 4) * PIPE_ROWS, make_float4(ret, 0.0f, 0.0f, 0.0f));
          }
        } else if (valid && mt_member(i) == 0) {
          float rv = isnan(ret) ? -1e6f : ret;  // deterministic.py:75-77
          if (p.penalty) rv = __fsub_rn(rv, p.penalty[row]);
          p.returns[row] = rv;
        }
        tr.rec(0x43u);
      };
      for (int j = 0; j < J; ++j) {
        const int i = j % n_mt, t = j / n_mt;
        tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && t == 2);
        boundary(i, j, t);
      }
    }
  } else {
    setmaxnreg_inc<PIPE_REGS_CONV>();
    // ============================================================ conversion warps
    const int e = warp - 12;
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch (hardware: warp id % 4)
    const int sub = e >> 2;                 // 0 / 1: even / odd K-chunks
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const int* cv = reinterpret_cast<const int*>(smem + lay.conv);
    uint32_t hc0 = 0, hc1 = 0;              // hidden accumulators consumed from D[0] / D[1]
    uint32_t pu = 0, wrap = 0;              // ring position of the next conversion's first unit
    PTracer<TR> tr{p.trace ? p.trace + warp * 1024 : nullptr, 0u, false};
    for (int r = 0; r < n_rounds; ++r) {
      const int n_mt = round_count(r);
      if (n_mt == 0) continue;
      StageSeq seq; seq.init(n_mt * p.H, nL, n_mt);
      int j, l, b;
      while (seq.next(j, l, b)) {
        if (l + 1 >= nL) continue;
        tr.arm(p.trace && blockIdx.x == 0 && lane == 0 && j / n_mt == 2);
        const int Npad = cv[8 * l], N = cv[8 * l + 1], act = (p.xflags & 2) ? BBMPC_ACT_NONE : cv[8 * l + 2];
        const int n_a_chunks = cv[8 * l + 3];
        const float* tail_tab = reinterpret_cast<const float*>(smem + lay.conv + MAX_LAYERS * 32) + 64 * l;
        const uint32_t taddr = tmem_base + lane_off + (b ? p.col_d1 : p.col_d0);
        mbar_wait_poll_then_sleep(bar_dfull + 8 * b, (b ? hc1 : hc0) & 1u, p.dfull_polls, dbgp, 0x4000000u | (j << 8) | l);
        if (b) ++hc1; else ++hc0;
        fence_after_sync();
        tr.rec(0x20u | (l << 8) | (b << 12));
        const uint32_t aring = smem_base + lay.aring;
#define PIPE_CONV(ACT) pipe_convert<ACT, TR>(taddr, Npad, N, n_a_chunks, sub, p.passes, aring, p.a_chunk_bytes, p.a_units, pu, wrap, \
                                            bar_afull, bar_afree, bar_drained + 8 * b, row_in_tile, lane, tail_tab, p.stagger, p.dfull_polls, dbgp, tr)
        if (ACT_T >= 0) {
          PIPE_CONV((ACT_T >= 0 ? ACT_T : 0));
        } else {
          switch (act) {
            case BBMPC_ACT_TANH: PIPE_CONV(BBMPC_ACT_TANH); break;
            case BBMPC_ACT_RELU: PIPE_CONV(BBMPC_ACT_RELU); break;
            case BBMPC_ACT_SIGMOID: PIPE_CONV(BBMPC_ACT_SIGMOID); break;
            default: PIPE_CONV(BBMPC_ACT_NONE); break;
          }
        }
#undef PIPE_CONV
        pu += (static_cast<uint32_t>(n_a_chunks) + 1u) >> 1;
        if (pu >= static_cast<uint32_t>(p.a_units)) { pu -= p.a_units; ++wrap; }
        tr.rec(0x21u | (l << 8) | (b << 12));
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
static int pipe_launch_t(bbmpc_ctx* ctx, const PipeParams& p, int grid, size_t smem_bytes, bool coop, cudaStream_t st) {
  bool all_tanh = !(p.xflags & 2);
  for (int l = 0; l + 1 < p.mlp.n_layers; ++l) all_tanh = all_tanh && p.mlp.layer[l].act == BBMPC_ACT_TANH;
  auto kern = (p.trace || p.dbg) ? (all_tanh ? rollout_pipe_kernel<DS_T, DU_T, true, BBMPC_ACT_TANH> : rollout_pipe_kernel<DS_T, DU_T, true, -1>)
                      : (all_tanh ? rollout_pipe_kernel<DS_T, DU_T, false, BBMPC_ACT_TANH> : rollout_pipe_kernel<DS_T, DU_T, false, -1>);
  BB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  if (coop) {
    void* args[] = {const_cast<PipeParams*>(&p)};
    const cudaError_t ce = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(PIPE_THREADS), args, smem_bytes, st);
    if (ce == cudaErrorCooperativeLaunchTooLarge) { cudaGetLastError(); return -100; }
    BB_CUDA(ctx, ce);
  } else {
    kern<<<grid, PIPE_THREADS, smem_bytes, st>>>(p);
  }
  BB_LAUNCH_CHECK(ctx);
  if (p.dbg) {  // BBMPC_DEBUG=1: synchronise and dump the watchdog record of a starved pipeline
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      const uint32_t* h = static_cast<const uint32_t*>(ctx->dbg_host);
      for (int b = 0; b < 2; ++b)
        for (int w = 0; w < PIPE_WARPS; ++w) {
          const uint32_t* r = h + 256 * b + 8 * w;
          if (r[0]) fprintf(stderr, "[bbmpc pipe watchdog] blk%%2=%d warp=%d tid=%u bar=+%u parity=%u tag=%08x\n", b, w, r[0] & 0xFFFF,
                            r[1] - 0u, r[2], r[3]);
        }
      return fail(ctx, BBMPC_ECUDA, "rollout_pipe_kernel: %s", cudaGetErrorString(e));
    }
  }
  if (p.trace) {
    cudaStreamSynchronize(st);
    static std::vector<uint32_t> h(32 * 1024);
    cudaMemcpy(h.data(), p.trace, h.size() * 4, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("BBMPC_TC_TRACE"), "w")) {
      for (int w = 0; w < PIPE_WARPS; ++w)
        for (int i = 0; i < 500; ++i)
          if (h[w * 1024 + 2 * i]) fprintf(f, "%d %x %u\n", w, h[w * 1024 + 2 * i], h[w * 1024 + 2 * i + 1]);
      fclose(f);
    }
    cudaMemset(p.trace, 0, h.size() * 4);
  }
  return BBMPC_OK;
}

// TMEM / shared-memory budget of the pipelined kernel for this model; false = use rollout_tc_kernel instead.
bool pipe_supported(const bbmpc_ctx* ctx, int passes, PipeParams* out) {
  const ModelHost& m = ctx->model;
  const MlpDev& M = m.mlp;
  const int nL = M.n_layers;
  if (!m.tc_ok || M.solo_jobs_per_step != nL) return false;   // N-split images are not supported here
  for (int l = 0; l < nL; ++l) if (M.layer[l].nsplit) return false;
  int buf_w = (M.layer[nL - 1].Npad + 15) / 16 * 16;
  if (buf_w < 32) buf_w = 32;
  int max_units = 1;
  for (int l = 0; l + 1 < nL; ++l) {
    if (M.layer[l].Npad > buf_w) buf_w = M.layer[l].Npad;
    const int u = ((M.layer[l + 1].Kpad >> 4) + 1) >> 1;
    if (u > max_units) max_units = u;
  }
  const int x_w = M.layer[0].Kpad;
  int max_mt = (512 - 2 * buf_w) / x_w;
  if (max_mt > PIPE_MAX_MT) max_mt = PIPE_MAX_MT;
  if (max_mt < 1) return false;
  PipeParams p{};
  p.col_d0 = 0; p.col_d1 = buf_w; p.col_x = 2 * buf_w; p.x_w = x_w; p.max_mt = max_mt;
  p.stage_bytes = M.stage_bytes;
  p.a_chunk_bytes = PIPE_ROWS * 16 * 2 * (passes == 3 ? 2 : 1);
  p.n_table = M.solo_groups_per_step;
  const size_t budget = 227 * 1024;
  // ring sizes: at least 3 weight stages and one layer + 1 unit of activations; spare shared memory goes to the weights
  int a_units = max_units + 1;
  if (const char* x = getenv("BBMPC_PIPE_AUNITS")) { const int v = atoi(x); if (v >= max_units && v <= PIPE_MAX_UNITS) a_units = v; }   // A/B: ring slack vs weight stages
  if (a_units > PIPE_MAX_UNITS) return false;
  int n_w = PIPE_MAX_WSTAGES;
  while (n_w > 2 && pipe_layout(p.stage_bytes, n_w, a_units, p.a_chunk_bytes, p.n_table, nL).total > budget) --n_w;
  if (pipe_layout(p.stage_bytes, n_w, a_units, p.a_chunk_bytes, p.n_table, nL).total > budget) return false;
  if (n_w < 3 && nL > 1) return false;
  p.n_wstages = n_w; p.a_units = a_units;
  if (out) *out = p;
  return true;
}

int launch_rollout_pipe(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                        const float* penalty, int rows, int A, int H, int passes, cudaStream_t st) {
  const ModelHost& m = ctx->model;
  PipeParams p{};
  if (!pipe_supported(ctx, passes, &p)) return -100;
  p.mlp = m.mlp; p.norm = m.norm; p.reward_id = ctx->reward_id; p.dS = m.dS; p.dU = m.dU;
  p.states = states; p.actions = actions; p.returns = returns; p.penalty = penalty;
  p.rows = rows; p.A = A; p.H = H; p.passes = passes; p.traj = ctx->traj_cur;
  p.n_tiles = (rows + PIPE_ROWS - 1) / PIPE_ROWS;
  p.jobs = m.mlp.solo_jobs; p.table = m.mlp.solo_table;
  if (const char* x = getenv("BBMPC_TC_X")) p.xflags = atoi(x);
  p.stagger = 0;
  p.dfull_polls = 0;    // measured: 0 -> 1.315, 128 -> 1.330, 1024 -> 1.346 ms per C4 rollout (polling warps take issue slots from the converting ones)
  if (const char* x = getenv("BBMPC_PIPE_DFULL_POLLS")) p.dfull_polls = atoi(x);
  if (const char* x = getenv("BBMPC_PIPE_STAGGER")) p.stagger = atoi(x);
  if (const char* x = getenv("BBMPC_PIPE_MT")) { const int v = atoi(x); if (v >= 1 && v < p.max_mt) p.max_mt = v; }
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
  const int nM = m.mlp.n_members;
  const long long ids = static_cast<long long>(p.n_tiles) * nM;
  // Launches that give every CTA at most one member-tile have nothing to interleave; there the one-tile-per-CTA kernel
  // wins (its conversion writes the next operand in place in TMEM, chunk by chunk behind the MMAs, without the shared-memory
  // ring and its handoffs — measured 0.59 vs 0.46 ms for a 1 250-row C4 shard; the SS-mode MMAs themselves run at the
  // TS-mode rate, tools/probe/dual_issue_probe.cu).  BBMPC_TC_PIPE=1 forces the pipelined kernel for every shape (tests, A/B).
  {
    const char* force = getenv("BBMPC_TC_PIPE");
    if (ids <= ctx->sm_count && !(force && force[0] == '1')) return -100;
  }
  int grid = ids < ctx->sm_count ? static_cast<int>(ids) : ctx->sm_count;
  if (nM > ctx->sm_count) return -100;
  // rounds are tile-aligned: at most max_mt member-tiles per CTA and round
  p.tiles_per_round = (p.max_mt * grid) / nM;
  if (p.tiles_per_round < 1) p.tiles_per_round = 1;
  p.xchg_tiles = p.n_tiles < 2 * p.tiles_per_round ? p.n_tiles : 2 * p.tiles_per_round;
  const size_t smem_bytes = pipe_layout(p.stage_bytes, p.n_wstages, p.a_units, p.a_chunk_bytes, p.n_table, m.mlp.n_layers).total;
  const int ds_t = m.dU > 8 ? 32 : (m.dS <= 8 ? 8 : (m.dS <= 24 ? 24 : 32));
  if (nM > 1) {
    const size_t need = static_cast<size_t>(p.xchg_tiles) * 2 * nM * PIPE_ROWS * ds_t;
    if (ctx->tc_xchg_floats < need) {
      BB_CUDA(ctx, cudaStreamSynchronize(st));
      cudaFree(ctx->tc_xchg);
      BB_CUDA(ctx, cudaMalloc(&ctx->tc_xchg, need * sizeof(float)));
      ctx->tc_xchg_floats = need;
    }
    if (ctx->tc_flags_n < p.xchg_tiles) {
      BB_CUDA(ctx, cudaStreamSynchronize(st));
      cudaFree(ctx->tc_flags);
      BB_CUDA(ctx, cudaMalloc(&ctx->tc_flags, p.xchg_tiles * sizeof(unsigned)));
      ctx->tc_flags_n = p.xchg_tiles;
    }
    BB_CUDA(ctx, cudaMemsetAsync(ctx->tc_flags, 0, p.xchg_tiles * sizeof(unsigned), st));
    p.xchg = ctx->tc_xchg; p.flags = ctx->tc_flags;
  }
  {
    const size_t need = static_cast<size_t>(grid) * PIPE_MAX_MT * (ds_t / 4 + 1) * PIPE_ROWS * 4;
    if (ctx->pipe_park_floats < need) {
      BB_CUDA(ctx, cudaStreamSynchronize(st));
      cudaFree(ctx->pipe_park);
      BB_CUDA(ctx, cudaMalloc(&ctx->pipe_park, need * sizeof(float)));
      ctx->pipe_park_floats = need;
    }
    p.park = ctx->pipe_park;
  }
  const bool coop = nM > 1;
  if (m.dS <= 8 && m.dU <= 8) return pipe_launch_t<8, 8>(ctx, p, grid, smem_bytes, coop, st);
  if (m.dS <= 24 && m.dU <= 8) return pipe_launch_t<24, 8>(ctx, p, grid, smem_bytes, coop, st);
  if (m.dU <= 8) return pipe_launch_t<32, 8>(ctx, p, grid, smem_bytes, coop, st);
  return pipe_launch_t<32, 16>(ctx, p, grid, smem_bytes, coop, st);
}

}  // namespace bbmpc
