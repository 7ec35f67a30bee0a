// device_fns.cuh — counter-based sampling, analytical dynamics and reward device functions.
// Reference behaviour restated per function (paths relative to /root/reference/).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/bbmpc.h"

namespace bbmpc {

// ----------------------------------------------------------------------------- Philox4x32-10
// Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Checked against the
// Random123 known-answer vectors in tests/test_cabi_cpu.py through bbmpc_philox4x32_host.
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

__host__ __device__ inline Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = Philox4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// Stream ids: which random tensor of the reference a draw stands for.
enum : uint32_t {
  STREAM_SAMPLES = 1,    // cem.py:90 / pi2.py:65 / random_search.py:40 / spsa.py:73 / cma_es.py:139
  STREAM_EXPLORE = 2,    // optimizer_base.py:83
  STREAM_PSO_R = 3,      // pso.py:108-109
  STREAM_PSO_POS = 4,    // pso.py:121,147
  STREAM_PSO_VEL = 5,    // pso.py:130,151
};

// Counter layout: (element/4, global row, iteration | stream<<16, act-call index).  The key is the
// context seed.  Element e of row r is word (e & 3) of block (e >> 2): independent of sharding.
__device__ inline Philox4 draw_block(uint64_t seed, uint32_t act_call, uint32_t stream, uint32_t iter,
                                     uint32_t row, uint32_t block) {
  return philox4x32_10(Philox4{block, row, (stream << 16) | (iter & 0xFFFFu), act_call},
                       static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

__device__ inline float u01_open(uint32_t x) {   // (0,1): 24-bit, never 0 or 1
  return (static_cast<float>(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
__device__ inline float u01_halfopen(uint32_t x) {  // [0,1)  (tf.random.uniform)
  return static_cast<float>(x >> 8) * (1.0f / 16777216.0f);
}
// Standard normal truncated to [-2, 2] by inverse CDF: same distribution as TF's
// redraw-until-inside tf.random.truncated_normal [TF], one uniform per sample, no loop.
__device__ inline float std_truncnorm(uint32_t x) {
  constexpr float PHI_M2 = 0.022750131948179195f;          // Phi(-2)
  constexpr float SPAN = 0.9544997361036416f;              // Phi(2) - Phi(-2)
  const float z = normcdfinvf(fmaf(u01_open(x), SPAN, PHI_M2));
  return fminf(fmaxf(z, -2.0f), 2.0f);
}
__device__ inline float std_normal(uint32_t x) { return normcdfinvf(u01_open(x)); }

// ----------------------------------------------------------------------------- activations
__device__ inline float act_exact(float v, int act) {  // parity-grade (libm) versions
  switch (act) {
    case BBMPC_ACT_TANH: return tanhf(v);
    case BBMPC_ACT_RELU: return fmaxf(v, 0.0f);
    case BBMPC_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}
// 5-instruction tanh: 1 - 2/(2^(2x log2 e) + 1); absolute error <= ~2e-7 over the whole range,
// saturates correctly to +-1, propagates NaN.  Used by the tensor-core epilogue.
__device__ __forceinline__ float tanh_fast(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ float act_fast(float v, int act) {
  switch (act) {
    case BBMPC_ACT_TANH: return tanh_fast(v);
    case BBMPC_ACT_RELU: return fmaxf(v, 0.0f);
    case BBMPC_ACT_SIGMOID: return fmaf(0.5f, tanh_fast(0.5f * v), 0.5f);
    default: return v;
  }
}

template <int ACT>
__device__ __forceinline__ float act_fast_t(float v) {
  if (ACT == BBMPC_ACT_TANH) return tanh_fast(v);
  if (ACT == BBMPC_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == BBMPC_ACT_SIGMOID) return fmaf(0.5f, tanh_fast(0.5f * v), 0.5f);
  return v;
}

// ----------------------------------------------------------------------------- analytical dynamics
// utils/pendulum.py:78-92.  x = [cos th, sin th, thdot, u]; writes the DEVIATION new - s.
// g=10, m=l=1, dt=.05; thdot clipped to +-8 AFTER newth is formed; u is not clipped.
__device__ inline void pendulum_deviation(const float* s, const float* a, float* dev) {
  const float u = a[0], thdot = s[2];
  const float theta = atan2f(s[1], s[0]);
  const float pi = 3.14159274101257324f;  // float(np.pi)
  // (-3*g/(2*l) * sin(theta+pi) + 3/(m*l**2) * u) * dt, evaluated left to right in fp32
  const float k1 = __fdiv_rn(-3.0f * 10.0f, 2.0f * 1.0f);
  const float k2 = __fdiv_rn(3.0f, 1.0f * 1.0f);
  const float acc = __fadd_rn(__fmul_rn(k1, sinf(__fadd_rn(theta, pi))), __fmul_rn(k2, u));
  float newthdot = __fadd_rn(thdot, __fmul_rn(acc, 0.05f));
  const float newth = __fadd_rn(theta, __fmul_rn(newthdot, 0.05f));
  newthdot = fminf(fmaxf(newthdot, -8.0f), 8.0f);
  dev[0] = __fsub_rn(cosf(newth), s[0]);
  dev[1] = __fsub_rn(sinf(newth), s[1]);
  dev[2] = __fsub_rn(newthdot, s[2]);
}

// ----------------------------------------------------------------------------- rewards
__device__ inline float floormod_f(float x, float m) {  // tf `%` on floats: result has m's sign
  const float r = fmodf(x, m);
  return (r != 0.0f && ((r < 0.0f) != (m < 0.0f))) ? r + m : r;
}
// utils/pendulum.py:10-35 with `third` = whatever lands in its `actions` parameter.
__device__ inline float pendulum_reward_core(const float* s, const float* third, int n_third) {
  const float pi = 3.14159274101257324f;
  const float two_pi = 6.28318548202514648f;
  const float ang = __fsub_rn(floormod_f(__fadd_rn(atan2f(s[1], s[0]), pi), two_pi), pi);
  float ss = 0.0f;
  for (int i = 0; i < n_third; ++i) ss = __fadd_rn(ss, __fmul_rn(third[i], third[i]));
  const float state_cost = __fadd_rn(__fmul_rn(ang, ang), __fmul_rn(0.1f, __fmul_rn(s[2], s[2])));
  return __fsub_rn(-state_cost, __fmul_rn(0.001f, ss));
}
// tutorials/mujoco/cost_func.py:5-22 (the 0.0 * sum(a^2) term is kept: it turns inf/NaN actions
// into NaN exactly as the reference does).
__device__ inline float halfcheetah_reward(const float* s, const float* a, const float* s2, int dU) {
  float r = 0.0f;
  if (s[5] >= 0.2f) r += -10.0f;
  if (s[6] >= 0.0f) r += -10.0f;
  if (s[7] >= 0.0f) r += -10.0f;
  r = __fadd_rn(r, __fdiv_rn(__fsub_rn(s2[17], s[17]), 0.01f));
  float ss = 0.0f;
  for (int i = 0; i < dU; ++i) ss = __fadd_rn(ss, __fmul_rn(a[i], a[i]));
  return __fsub_rn(r, __fmul_rn(0.0f, ss));
}
// reward_function(current_state, actions, next_state) as invoked by deterministic.py:65-66,126-127
__device__ inline float reward_dispatch(int reward_id, const float* s, const float* a, const float* s2,
                                        int dS, int dU) {
  switch (reward_id) {
    case BBMPC_REWARD_PENDULUM: return pendulum_reward_core(s, s2, dS);      // arg-order quirk
    case BBMPC_REWARD_PENDULUM_GYM: return pendulum_reward_core(s, a, dU);
    case BBMPC_REWARD_HALFCHEETAH: return halfcheetah_reward(s, a, s2, dU);
    default: return 0.0f;
  }
}

}  // namespace bbmpc
