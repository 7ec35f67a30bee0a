// umma_probe.cu — hardware bring-up probe for the tcgen05 building blocks used by the rollout
// kernel (tools/, not product code).  One CTA computes D[128,N] = A[128,K] * W[N,K]^T in bf16
// with fp32 accumulation through one of several operand paths and checks it against the host.
//
//   umma_probe <variant> [N] [K]
//     variant bit0: A operand  0 = shared memory (SS)      1 = tensor memory (TS, tcgen05.st)
//     variant bit1: A packing  0 = even k in low half       1 = odd k in low half      (TS only)
//     variant bit2: descriptor 0 = lbo:K-step, sbo:8-row    1 = swapped
//     variant bit3: W staging  0 = st.shared by threads     1 = cp.async.bulk + mbarrier tx
//
// Each variant must run in its own process: a watchdog trap poisons the context.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "../../blackbox_mpc_b200/csrc/tc05.cuh"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

using namespace tc05;

// smem image of a K-major operand with R rows: [K/8][R][8] bf16 (16-byte vectors, row-contiguous)
__device__ __host__ inline size_t img_off(int r, int k, int R) {
  return (static_cast<size_t>(k >> 3) * R + r) * 8 + (k & 7);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* __restrict__ A,     // [128][K] row-major
             const __nv_bfloat16* __restrict__ Wimg,  // pre-formatted [K/8][N][8]
             float* __restrict__ D,                   // [128][N]
             int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mma, bar_tx;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool a_in_tmem = variant & 1, swap_pack = variant & 2, swap_desc = variant & 4,
             bulk = variant & 8;

  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sA = sW + static_cast<size_t>(N) * K;
  const uint32_t w_bytes = static_cast<uint32_t>(N) * K * 2;

  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar_mma), 1);
    mbar_init(smem_u32(&bar_tx), 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_d = tmem_base;        // columns [0, N)
  const uint32_t tmem_a = tmem_base + 256;  // columns [256, 256 + K/2)

  // ---- stage W
  if (bulk) {
    if (tid == 0) {
      mbar_arrive_expect_tx(smem_u32(&bar_tx), w_bytes);
      bulk_g2s(smem_u32(sW), Wimg, w_bytes, smem_u32(&bar_tx));
    }
    mbar_wait(smem_u32(&bar_tx), 0);
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(Wimg);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int i = tid; i < static_cast<int>(w_bytes / 16); i += blockDim.x) dst[i] = src[i];
  }
  // ---- stage A
  if (!a_in_tmem) {
    for (int k8 = 0; k8 < K / 8; ++k8) {  // thread = row
      uint4 v = *reinterpret_cast<const uint4*>(A + static_cast<size_t>(tid) * K + k8 * 8);
      *reinterpret_cast<uint4*>(sA + img_off(tid, k8 * 8, 128)) = v;
    }
  } else {
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int c8 = 0; c8 < K / 16; ++c8) {  // 8 columns = 16 bf16 per store
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint16_t e0 = __bfloat16_as_ushort(A[static_cast<size_t>(tid) * K + c8 * 16 + 2 * j]);
        const uint16_t e1 = __bfloat16_as_ushort(A[static_cast<size_t>(tid) * K + c8 * 16 + 2 * j + 1]);
        r[j] = swap_pack ? (static_cast<uint32_t>(e0) << 16 | e1) : (static_cast<uint32_t>(e1) << 16 | e0);
      }
      tmem_st8(tmem_a + lane_base + c8 * 8, r);
    }
    wait_st();
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // ---- MMA (one thread)
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N);
    const uint32_t w_kstep = static_cast<uint32_t>(N) * 16;  // bytes between k8 slabs of W
    const uint32_t a_kstep = 128u * 16;
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint32_t wl = swap_desc ? 128u : w_kstep, ws = swap_desc ? w_kstep : 128u;
      const uint64_t bdesc = smem_desc_kmajor_noswz(smem_u32(sW) + kk * 2 * w_kstep, wl, ws);
      if (a_in_tmem) {
        mma_ts(tmem_d, tmem_a + kk * 8, bdesc, idesc, kk > 0);
      } else {
        const uint32_t al = swap_desc ? 128u : a_kstep, as = swap_desc ? a_kstep : 128u;
        const uint64_t adesc = smem_desc_kmajor_noswz(smem_u32(sA) + kk * 2 * a_kstep, al, as);
        mma_ss(tmem_d, adesc, bdesc, idesc, kk > 0);
      }
    }
    mma_commit(smem_u32(&bar_mma));
  }
  mbar_wait(smem_u32(&bar_mma), 0);
  fence_after_sync();

  // ---- epilogue: thread = row (lane of TMEM)
  {
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_d + lane_base + c, r);
      wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[static_cast<size_t>(tid) * N + c + j] = __uint_as_float(r[j]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  (void)lane;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = argc > 2 ? atoi(argv[2]) : 208;
  const int K = argc > 3 ? atoi(argv[3]) : 208;
  if (N % 16 || K % 16 || N > 256 || K > 256) { printf("bad N/K\n"); return 2; }
  std::vector<__nv_bfloat16> hA(128 * K), hW(static_cast<size_t>(N) * K), hWimg(static_cast<size_t>(N) * K);
  std::vector<float> fA(128 * K), fW(static_cast<size_t>(N) * K);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return static_cast<int>((s >> 24) % 7) - 3; };
  for (int i = 0; i < 128 * K; ++i) { fA[i] = static_cast<float>(rnd()); hA[i] = __float2bfloat16(fA[i]); }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float v = static_cast<float>(rnd());
      fW[static_cast<size_t>(n) * K + k] = v;
      hW[static_cast<size_t>(n) * K + k] = __float2bfloat16(v);
      hWimg[img_off(n, k, N)] = __float2bfloat16(v);
    }
  __nv_bfloat16 *dA, *dW; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dW, hWimg.size() * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hWimg.data(), hWimg.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4));
  const size_t smem = static_cast<size_t>(N) * K * 2 + 128 * K * 2;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  probe_kernel<<<1, 128, smem>>>(dA, dW, dD, N, K, variant);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += fA[m * K + k] * fW[static_cast<size_t>(n) * K + k];
      const float got = hD[m * N + n];
      const double e = fabs(static_cast<double>(got) - ref);
      if (!(e <= 1e-3)) { if (bad < 6) printf("  mismatch m=%d n=%d got=%g ref=%g\n", m, n, got, ref); ++bad; }
      if (e > maxerr) maxerr = e;
    }
  printf("PROBE variant=%d N=%d K=%d  bad=%d/%d maxerr=%g  %s\n", variant, N, K, bad, 128 * N, maxerr,
         bad ? "FAIL" : "PASS");
  return bad ? 1 : 0;
}
