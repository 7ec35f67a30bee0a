#include <cstdio>
#include <cstring>
#include "tc05.cuh"
using namespace tc05;
__global__ void k(float y0, float y1, float* o) {
  const uint64_t y = pk2(y0, y1);
  const uint64_t c = mul2(y, pk2(65537.0f, 65537.0f));
  const uint64_t t = sub2(c, y);
  const uint64_t h = sub2(c, t);
  const uint64_t l = sub2(y, h);
  upk2(c, o[0], o[1]); upk2(t, o[2], o[3]); upk2(h, o[4], o[5]); upk2(l, o[6], o[7]);
  const float cs = __fmul_rn(y0, 65537.0f), ts = __fsub_rn(cs, y0), hs = __fsub_rn(cs, ts), ls = __fsub_rn(y0, hs);
  o[8] = cs; o[9] = ts; o[10] = hs; o[11] = ls;
  const uint64_t a = add2(pk2(1.0f, 2.0f), pk2(10.0f, 20.0f)); upk2(a, o[12], o[13]);
  const uint64_t f = fma2(pk2(1.5f, 2.5f), pk2(2.0f, 4.0f), pk2(-1.0f, 1.0f)); upk2(f, o[14], o[15]);
}
int main() {
  float* d; cudaMalloc(&d, 64); k<<<1, 1>>>(0.7234567f, -0.0123456f, d);
  float h[16]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 16; ++i) { uint32_t u; memcpy(&u, &h[i], 4); printf("o[%2d] = %.9g (%08x)\n", i, h[i], u); }
}
