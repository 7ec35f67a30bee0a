// epi_cost.cu — throughput of the instruction classes of the epilogue (clocks per 16-element chunk per
// warp) with 1..4 warps resident per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I blackbox_mpc_b200/csrc tools/probe/epi_cost.cu -o tools/probe/epi_cost
#include <cstdio>
#include "tc05.cuh"
using namespace tc05;


// ---- packed fp32x2 helpers (sm_100: FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// pair tanh (prescaled input) + Veltkamp hi/lo split, packed math.  LO_CVT: lo via cvt (XU) or Veltkamp (FMA pipe)
template <bool LO_CVT>
__device__ __forceinline__ void tanh_split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  float w0, w1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(x0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(x1)));
  const uint64_t one = pk(1.0f, 1.0f);
  const uint64_t d = add2(pk(w0, w1), one);
  float d0, d1; upk(d, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  uint64_t y = fma2(mul2(pk(r, r), pk(d1, d0)), pk(2.0f, 2.0f), pk(-1.0f, -1.0f));
  float y0, y1; upk(y, y0, y1);
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y0) : "f"(y0), "f"(x0));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y1) : "f"(y1), "f"(x1));
  y = pk(y0, y1);
  const uint64_t k = pk(65537.0f, 65537.0f);
  const uint64_t c = mul2(y, k);
  const uint64_t h = sub2(c, sub2(c, y));
  const uint64_t l = sub2(y, h);
  float h0, h1, l0, l1; upk(h, h0, h1); upk(l, l0, l1);
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(__float_as_uint(h0)), "r"(__float_as_uint(h1)));
  if (LO_CVT) {
    lo = pack_bf16x2(l0, l1);
  } else {
    const uint64_t c2 = mul2(l, k);
    const uint64_t lh = sub2(c2, sub2(c2, l));
    float a, b; upk(lh, a, b);
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(lo) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
  }
}

template <int V>
__global__ void probe(float* io, unsigned long long* out, int iters) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = io[threadIdx.x * 16 + j];
  __syncthreads();
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    if (V == 0) {  // 16 ex2
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[j]) : "f"(v[j]));
    } else if (V == 1) {  // 16 rcp
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[j]) : "f"(v[j]));
    } else if (V == 2) {  // 16 cvt.rn.bf16x2 (8 hi + 8 lo style)
      uint32_t h[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(v[j]), "f"(v[(j + 1) & 15]));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(h[j] & 0x3F803F80u);
    } else if (V == 3) {  // 64 FFMA
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[j]) : "f"(0.999f), "f"(0.001f));
    } else if (V == 4) {  // 64 LOP3/ALU
      uint32_t h[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = __float_as_uint(v[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h[j]) : "r"(h[(j + 1) & 15]), "r"(0x1234567u + k));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(h[j] & 0x3FFFFFFFu);
    } else if (V == 5) {  // the real thing: pair tanh + split (no TMEM)
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        float w0, w1, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(v[j])));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(v[j + 1])));
        const float d0 = w0 + 1.0f, d1 = w1 + 1.0f;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
        const float y0 = fmaf(r * d1, 2.0f, -1.0f), y1 = fmaf(r * d0, 2.0f, -1.0f);
        v[j] = __uint_as_float((__float_as_uint(y0) & 0x7FFFFFFFu) | (__float_as_uint(v[j]) & 0x80000000u));
        v[j + 1] = __uint_as_float((__float_as_uint(y1) & 0x7FFFFFFFu) | (__float_as_uint(v[j + 1]) & 0x80000000u));
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[2 * j] += __uint_as_float(hi[j]); v[2 * j + 1] += __uint_as_float(lo[j]); }
    } else if (V == 7 || V == 8) {  // packed-math pair tanh + split; 7: lo via cvt, 8: lo via Veltkamp
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tanh_split_pair<V == 7>(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[2 * j] += __uint_as_float(hi[j]); v[2 * j + 1] += __uint_as_float(lo[j]); }
    } else if (V == 6) {  // split only
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[2 * j] = __uint_as_float(hi[j] & 0x3F803F80u); v[2 * j + 1] = __uint_as_float(lo[j] & 0x3F803F80u) ; }
    }
  }
  const unsigned long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc += v[j];
  io[threadIdx.x] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int V>
void run(const char* name, float* io, unsigned long long* d) {
  for (int wps : {1, 2, 3, 4}) {
    probe<V><<<1, 128 * wps>>>(io, d, 2000);
    cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-36s warps/SMSP=%d: %7.1f clk/iter/warp  -> %7.1f clk per chunk per SMSP\n", name, wps, double(h) / 2000, double(h) / 2000 / wps);
  }
}

int main() {
  float* io; unsigned long long* d;
  cudaMalloc(&io, 512 * 16 * 4); cudaMemset(io, 0, 512 * 16 * 4); cudaMalloc(&d, 64);
  run<0>("16 ex2", io, d);
  run<1>("16 rcp", io, d);
  run<2>("16 cvt.bf16x2", io, d);
  run<3>("64 ffma", io, d);
  run<4>("64 lop3", io, d);
  run<5>("pair-tanh16 + split", io, d);
  run<6>("split16", io, d);
  run<7>("packed pair-tanh16+split(cvt lo)", io, d);
  run<8>("packed pair-tanh16+split(velt lo)", io, d);
  return 0;
}
