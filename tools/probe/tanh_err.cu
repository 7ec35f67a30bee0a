// tanh_err.cu — accuracy of the epilogue tanh variants (pre-scaled input) against double precision.
#include <cstdio>
#include <cmath>
#include <vector>
#include "common.cuh"
#include "device_fns.cuh"
#include "tc05.cuh"
using namespace tc05;
namespace bbmpc {
#define BBMPC_PROBE_ONLY
}
// copy of the two variants under test (kept in sync with rollout_tc.cu by hand)
__device__ __forceinline__ void tanh_pair_prescaled(float& x0, float& x1) {
  float w0, w1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(x0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(x1)));
  const float d0 = w0 + 1.0f, d1 = w1 + 1.0f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  const float y0 = fmaf(r * d1, 2.0f, -1.0f), y1 = fmaf(r * d0, 2.0f, -1.0f);
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(x0) : "f"(y0), "f"(x0));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(x1) : "f"(y1), "f"(x1));
}
__device__ __forceinline__ void tanh_quad(float x0, float x1, float x2, float x3, float* y) {
  float w0, w1, w2, w3, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(x0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(x1)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w2) : "f"(-fabsf(x2)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w3) : "f"(-fabsf(x3)));
  const uint64_t one = pk2(1.0f, 1.0f);
  const uint64_t d01 = add2(pk2(w0, w1), one), d23 = add2(pk2(w2, w3), one);
  const uint64_t p = mul2(d01, d23);
  float px, py; upk2(p, px, py);
  const float P = px * py;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(P));
  r = fmaf(r, fmaf(-P, r, 1.0f), r);
  const uint64_t t = mul2(pk2(r, r), pk2(py, px));
  const uint64_t two = pk2(2.0f, 2.0f), m1 = pk2(-1.0f, -1.0f);
  uint64_t y01 = fma2(mul2(t, d23), two, m1), y23 = fma2(mul2(t, d01), two, m1);
  float y0, y1, y2, y3; upk2(y01, y0, y1); upk2(y23, y2, y3);
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y[0]) : "f"(y0), "f"(x0));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y[1]) : "f"(y1), "f"(x1));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y[2]) : "f"(y2), "f"(x2));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=f"(y[3]) : "f"(y3), "f"(x3));
}
__global__ void k(const float* x, float* yp, float* yq, uint32_t* hq, uint32_t* lq, int n) {
  const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i + 3 >= n) return;
  float a = x[i], b = x[i + 1], c = x[i + 2], d = x[i + 3];
  tanh_pair_prescaled(a, b); tanh_pair_prescaled(c, d);
  yp[i] = a; yp[i + 1] = b; yp[i + 2] = c; yp[i + 3] = d;
  tanh_quad(x[i], x[i + 1], x[i + 2], x[i + 3], yq + i);
  uint32_t h, l;
  split_bf16x2_packed(pk2(yq[i], yq[i + 1]), h, l); hq[i / 2] = h; lq[i / 2] = l;
  split_bf16x2_packed(pk2(yq[i + 2], yq[i + 3]), h, l); hq[i / 2 + 1] = h; lq[i / 2 + 1] = l;
}
static float bf(uint32_t v) { uint32_t u = v << 16; float f; memcpy(&f, &u, 4); return f; }
int main() {
  const int n = 1 << 16;
  std::vector<float> x(n);
  for (int i = 0; i < n; ++i) x[i] = (float)((i * 2654435761u % 200001) / 100000.0 - 1.0) * ((i % 7 == 0) ? 12.0f : 1.5f);
  float *dx, *dp, *dq; uint32_t *dh, *dl;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dp, n * 4); cudaMalloc(&dq, n * 4); cudaMalloc(&dh, n * 2); cudaMalloc(&dl, n * 2);
  cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
  k<<<n / 4 / 128, 128>>>(dx, dp, dq, dh, dl, n);
  std::vector<float> yp(n), yq(n); std::vector<uint32_t> h(n / 2), l(n / 2);
  cudaMemcpy(yp.data(), dp, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(yq.data(), dq, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h.data(), dh, n * 2, cudaMemcpyDeviceToHost); cudaMemcpy(l.data(), dl, n * 2, cudaMemcpyDeviceToHost);
  double ep = 0, eq = 0, es = 0; int worst = 0;
  for (int i = 0; i < n; ++i) {
    const double ref = tanh((double)x[i] / 2.8853900817779268);
    ep = fmax(ep, fabs(yp[i] - ref));
    if (fabs(yq[i] - ref) > eq) { eq = fabs(yq[i] - ref); worst = i; }
    const uint32_t hh = h[i / 2], ll = l[i / 2];
    const double rec = (i & 1) ? (double)bf(hh >> 16) + bf(ll >> 16) : (double)bf(hh & 0xFFFF) + bf(ll & 0xFFFF);
    const double e1 = fabs(rec - yq[i]) / fmax(fabs((double)yq[i]), 1e-30);
    if (e1 > es) { es = e1; printf("  i=%d x=%g y=%.9g hi=%08x lo=%08x rec=%.9g\n", i, x[i], yq[i], hh, ll, rec); }
  }
  printf("max abs err: pair %.3e  quad %.3e (worst x=%g yq=%g ref=%g)  split rel err %.3e\n", ep, eq, x[worst], yq[worst],
         tanh((double)x[worst] / 2.8853900817779268), es);
  return 0;
}
