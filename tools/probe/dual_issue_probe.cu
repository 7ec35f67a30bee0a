// dual_issue_probe.cu — does a second MMA-issuing warp raise tensor throughput when the issuer's own instruction stream
// (barrier waits, commits, probes between the MMAs of a ring unit) is the limiter?
// A "unit" = 6 SS MMAs (M128 x N208 x K16) + `waits` successful mbarrier waits + 2 commits, as in rollout_pipe_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I blackbox_mpc_b200/csrc tools/probe/dual_issue_probe.cu -o tools/probe/dual_issue_probe
#include <cstdio>
#include "tc05.cuh"
using namespace tc05;

struct Args { int issuers, waits, iters, N, mode; unsigned long long* out; };
constexpr uint32_t A_LBO = 128 * 16, A_CHUNK = 2 * A_LBO;

__device__ __forceinline__ uint32_t wait_flavour(int mode, uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  switch (mode) {
    case 0: while (!mbar_try_wait(bar, parity)) {} return 1;
    case 1: do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok); return 1;
    case 2: do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok); return 1;
    case 3: do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok); return 1;
  }
  return ok;
}

__global__ void __launch_bounds__(128, 1) probe(Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[40];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16 * A_CHUNK + 12 * 208 * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8 * i, 1);
    for (int i = 4; i < 40; ++i) mbar_init(bar0 + 8 * i, 1000000);
    fence_mbar_init();
    mbar_arrive(bar0 + 8 * 2);      // bars[2]: phase 0 complete -> waits on parity 0 succeed at once
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = idesc_bf16_f32(128, a.N);
  const uint64_t bdesc = smem_desc_kmajor_noswz(smem_u32(smem + 16 * A_CHUNK), a.N * 16, 128);
  const uint64_t adesc0 = smem_desc_kmajor_noswz(smem_u32(smem), A_LBO, 128);
  const uint64_t achunk = A_CHUNK >> 4;
  if (warp >= 1 && warp <= a.issuers) {
    const int w = warp - 1;
    const uint32_t d = tm + 208 * w;
    const unsigned long long t0 = clock64();
    for (int i = 0; i < a.iters; ++i) {
      if (a.mode == 4) {   // the kernel's fused block: two probes (results unused here), 6 MMAs, 2 commits in one asm statement
        fence_after_sync();
        const uint64_t ah = adesc0 + (4 * ((i + 4 * w) & 3)) * achunk;
        const uint64_t bh = bdesc + (uint64_t)((4 * ((i + w) % 3)) * (a.N * 32) >> 4);
        uint32_t o0, o1;
        mma_unit_ss_probe(d, ah, bh, (uint32_t)achunk, (uint32_t)((a.N * 32) >> 4), (uint32_t)(2 * achunk), (uint32_t)(2 * ((a.N * 32) >> 4)), idesc, 1u, 1u, 1u,
                          bar0 + 8 * (4 + 18 * w + (i % 9)), bar0 + 8 * (4 + 18 * w + 9 + (i % 9)), bar0 + 8 * 2, 0u, bar0 + 8 * 2, 0u, o0, o1);
        if (a.waits && !(o0 & o1)) break;
        continue;
      }
      for (int k = 0; k < a.waits; ++k) wait_flavour(a.mode, bar0 + 8 * 2, 0);
      fence_after_sync();
      if (elect_one()) {
        const uint64_t ah = adesc0 + (4 * ((i + 4 * w) & 3)) * achunk, al = ah + achunk;
        const uint64_t bh = bdesc + (uint64_t)((4 * ((i + w) % 3)) * (a.N * 32) >> 4), bl = bh + (uint64_t)((a.N * 32) >> 4);
        mma_ss(d, ah, bh, idesc, 1u); mma_ss(d, al, bh, idesc, 1u); mma_ss(d, ah, bl, idesc, 1u);
        mma_ss(d, ah + 2 * achunk, bh + 2 * (uint64_t)((a.N * 32) >> 4), idesc, 1u); mma_ss(d, al + 2 * achunk, bh + 2 * (uint64_t)((a.N * 32) >> 4), idesc, 1u);
        mma_ss(d, ah + 2 * achunk, bl + 2 * (uint64_t)((a.N * 32) >> 4), idesc, 1u);
        mma_commit(bar0 + 8 * (4 + 18 * w + (i % 9)));
        mma_commit(bar0 + 8 * (4 + 18 * w + 9 + (i % 9)));
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(bar0 + 8 * w);
    __syncwarp();
    mbar_wait(bar0 + 8 * w, 0);
    const unsigned long long t1 = clock64();
    if ((tid & 31) == 0) a.out[w] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  const int smem_bytes = 16 * A_CHUNK + 12 * 208 * 32;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const char* mn[] = {"try_wait", "test_wait", "try_wait.relaxed", "test_wait.relaxed", "fused probe block"};
  for (int N : {208, 32})
   for (int mode : {0, 1, 2, 3, 4})
    for (int waits : {0, 2})
      for (int issuers : {1, 2}) {
        if (mode > 0 && mode < 4 && waits == 0) continue;
        Args a{issuers, waits, 400, N, mode, d};
        cudaMemset(d, 0, 64);
        probe<<<1, 128, smem_bytes>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double t = double(h[0] > h[1] ? h[0] : h[1]);
        printf("N=%3d %-18s waits/unit=%d issuers=%d: %s  %.0f clk per unit and issuer, %.0f clk per unit overall (tensor floor %d)\n", N, mn[mode], waits, issuers,
               e == cudaSuccess ? "ok" : cudaGetErrorString(e), t / a.iters, t / a.iters / issuers, 6 * 128 * N / 256);
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
