// tc_epi.cuh — conversion-warp arithmetic shared by the tensor-core rollout kernels (rollout_tc.cu,
// rollout_pipe.cu): non-blocking mbarrier probes, the pre-scaled tanh variants and the activation of a
// 16-column accumulator chunk.  DESIGN.md section 4.1 has the measurements behind these forms.
#pragma once
#include "device_fns.cuh"
#include "tc05.cuh"

namespace bbmpc {
using namespace tc05;

#ifndef BBMPC_POLL_SPINS
#define BBMPC_POLL_SPINS (1u << 25)
#endif
#ifndef BBMPC_RCP_MUFU
#define BBMPC_RCP_MUFU 1   // 0: reciprocal on the FMA pipe (measured slower, see BBMPC_LO_CVT)
#endif
#ifndef BBMPC_PACKED_TANH
#define BBMPC_PACKED_TANH 1
#endif
constexpr bool PACKED_TANH = BBMPC_PACKED_TANH != 0;   // quad-shared reciprocal + packed fp32x2 epilogue math

__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {  // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// Busy-polling wait (non-blocking test_wait): a suspended try_wait wakes up ~500 cycles after the phase
// completes (measured with the in-kernel tracer); a single polling warp on the critical path reacts within
// one probe latency (~100 cycles) at the price of a few issue slots.
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity, volatile uint32_t* dbg = nullptr, uint32_t tag = 0) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++spins > BBMPC_POLL_SPINS) {
      if (dbg) {
        const uint32_t slot = 8u * (threadIdx.x >> 5) + 256u * (blockIdx.x & 1);
        dbg[slot + 0] = 0xDEAD0000u | threadIdx.x; dbg[slot + 1] = bar; dbg[slot + 2] = parity; dbg[slot + 3] = tag;
        __threadfence_system();
      }
      asm volatile("trap;");
    }
  }
}

// Medium waits on the critical path (an accumulator the tensor pipe is finishing): poll for a bounded time, then sleep.
__device__ __forceinline__ void mbar_wait_poll_then_sleep(uint32_t bar, uint32_t parity, uint32_t polls, volatile uint32_t* dbg = nullptr, uint32_t tag = 0) {
  for (uint32_t i = 0; i < polls; ++i)
    if (mbar_test_wait(bar, parity)) return;
  mbar_wait_sleep(bar, parity, dbg, tag);
}

// tanh of two pre-activations that arrive PRE-SCALED by 2 log2(e) (the scale is folded into the
// layer's weight image, see pack_tc_kernel): w = 2^-|t| in (0,1], tanh|x| = 2/(1+w) - 1.  The two
// reciprocals share ONE MUFU.RCP: r = 1/((1+w0)(1+w1)) (product <= 4, no overflow), 1/(1+w0) =
// r (1+w1).  3 MUFU per pair instead of 4; absolute error <= ~3e-7, far below the 2^-17 relative
// error of the bf16 hi+lo operand split downstream.  NaN propagates through ex2.
__device__ __forceinline__ void tanh_pair_prescaled(float& x0, float& x1) {
  float w0, w1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(x0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(x1)));
  const float d0 = w0 + 1.0f, d1 = w1 + 1.0f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  const float y0 = fmaf(r * d1, 2.0f, -1.0f), y1 = fmaf(r * d0, 2.0f, -1.0f);
  // copysign: (y & ~sign) | (x & sign) in one LOP3
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(x0) : "f"(y0), "f"(x0));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(x1) : "f"(y1), "f"(x1));
}

// tanh of FOUR pre-scaled pre-activations with 4 MUFU.EX2 + ONE MUFU.RCP, then bf16 hi/lo split, all FMA-pipe
// work as packed fp32x2: d_i = 1 + 2^-|t_i| in (1, 2]; r = 1/(d0 d1 d2 d3) (product <= 16); 1/d0 = r (d1 d3) d2,
// 1/d1 = r (d0 d2) d3, 1/d2 = r (d1 d3) d0, 1/d3 = r (d0 d2) d1.  One MUFU operation per element
// (the shared reciprocal runs on the FMA pipe).
__device__ __forceinline__ void tanh_split_quad(float x0, float x1, float x2, float x3, uint32_t& hi01, uint32_t& lo01,
                                                uint32_t& hi23, uint32_t& lo23) {
  float w0, w1, w2, w3, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(x0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(x1)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w2) : "f"(-fabsf(x2)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w3) : "f"(-fabsf(x3)));
  const uint64_t one = pk2(1.0f, 1.0f);
  const uint64_t d01 = add2(pk2(w0, w1), one), d23 = add2(pk2(w2, w3), one);
  const uint64_t p = mul2(d01, d23);                 // (d0 d2, d1 d3)
  float px, py;
  upk2(p, px, py);
  const float P = px * py;
#if BBMPC_RCP_MUFU
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(P));
  r = fmaf(r, fmaf(-P, r, 1.0f), r);                  // one Newton step: the shared reciprocal feeds four results
#else
  // reciprocal of P in [1, 16] on the FMA pipe: exponent-flip first guess (relative error <= 0.051), three Newton
  // steps (5.1e-2 -> 2.6e-3 -> 6.6e-6 -> < 1e-7); the MUFU pipe keeps the four ex2 only
  r = __uint_as_float(0x7EF311C7u - __float_as_uint(P));
  r = fmaf(r, fmaf(-P, r, 1.0f), r);
  r = fmaf(r, fmaf(-P, r, 1.0f), r);
  r = fmaf(r, fmaf(-P, r, 1.0f), r);
#endif
  const uint64_t t = mul2(pk2(r, r), pk2(py, px));    // (r d1 d3, r d0 d2)
  const uint64_t two = pk2(2.0f, 2.0f), m1 = pk2(-1.0f, -1.0f);
  uint64_t y01 = fma2(mul2(t, d23), two, m1);         // 2/d0 - 1, 2/d1 - 1
  uint64_t y23 = fma2(mul2(t, d01), two, m1);         // 2/d2 - 1, 2/d3 - 1
  float y0, y1, y2, y3;
  upk2(y01, y0, y1); upk2(y23, y2, y3);
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y0) : "f"(y0), "f"(x0));   // copysign
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y1) : "f"(y1), "f"(x1));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y2) : "f"(y2), "f"(x2));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y3) : "f"(y3), "f"(x3));
  split_bf16x2_packed(pk2(y0, y1), hi01, lo01);
  split_bf16x2_packed(pk2(y2, y3), hi23, lo23);
}

template <int ACT>
__device__ __forceinline__ void act16(float (&v)[16]) {
  if (ACT == BBMPC_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) tanh_pair_prescaled(v[j], v[j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_fast_t<ACT>(v[j]);
  }
}

// A trailing chunk of a hidden layer (real features | the ones-columns that meet the bias rows of the next layer's
// weights | zero padding) -> bf16 hi/lo words: v = act(D) * mask + add with per-column mask / add vectors prepared in
// shared memory.  n_real = real features in this chunk (<= 0: none).  tanh layers: quads of four real features take the
// fused tanh + split of the full chunks, quads without real features are constants (no MUFU work for ones / padding:
// a 3x200 layer's 13th chunk has 8 real features), only a quad that mixes both takes the generic form.
template <int ACT>
__device__ __forceinline__ void tail16(const uint32_t (&r)[16], bool has_data, int n_real, const float* mk_ad, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  if constexpr (ACT == BBMPC_ACT_TANH && PACKED_TANH) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 ad = *reinterpret_cast<const float4*>(mk_ad + 16 + 4 * q);
      if (has_data && 4 * q + 4 <= n_real) {
        tanh_split_quad(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]),
                        hi[2 * q], lo[2 * q], hi[2 * q + 1], lo[2 * q + 1]);
      } else if (!has_data || 4 * q >= n_real) {   // 0 / 1 constants: exact in bf16, no residual
        asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi[2 * q]) : "r"(__float_as_uint(ad.x)), "r"(__float_as_uint(ad.y)));
        asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi[2 * q + 1]) : "r"(__float_as_uint(ad.z)), "r"(__float_as_uint(ad.w)));
        lo[2 * q] = 0u; lo[2 * q + 1] = 0u;
      } else {
        const float4 mk = *reinterpret_cast<const float4*>(mk_ad + 4 * q);
        float v0 = __uint_as_float(r[4 * q]), v1 = __uint_as_float(r[4 * q + 1]), v2 = __uint_as_float(r[4 * q + 2]), v3 = __uint_as_float(r[4 * q + 3]);
        tanh_pair_prescaled(v0, v1);
        tanh_pair_prescaled(v2, v3);
        split_bf16x2_veltkamp(fmaf(v0, mk.x, ad.x), fmaf(v1, mk.y, ad.y), hi[2 * q], lo[2 * q]);
        split_bf16x2_veltkamp(fmaf(v2, mk.z, ad.z), fmaf(v3, mk.w, ad.w), hi[2 * q + 1], lo[2 * q + 1]);
      }
    }
  } else {
    float v[16];
    if (has_data) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      act16<ACT>(v);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 mk = *reinterpret_cast<const float4*>(mk_ad + j), ad = *reinterpret_cast<const float4*>(mk_ad + 16 + j);
      v[j] = fmaf(v[j], mk.x, ad.x); v[j + 1] = fmaf(v[j + 1], mk.y, ad.y);
      v[j + 2] = fmaf(v[j + 2], mk.z, ad.z); v[j + 3] = fmaf(v[j + 3], mk.w, ad.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16x2_veltkamp(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  }
}

}  // namespace bbmpc
