// issue_cost.cu — cost of each instruction class in the tcgen05 issuer loop (clocks per iteration).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I blackbox_mpc_b200/csrc tools/probe/issue_cost.cu -o tools/probe/issue_cost
#include <cstdio>
#include "tc05.cuh"
using namespace tc05;

// bit flags
constexpr int F_WARP = 1;      // warp-wide loop with elect_one + __syncwarp (else lane 0 only)
constexpr int F_COMMIT = 2;    // tcgen05.commit every iteration
constexpr int F_FENCE = 4;     // tcgen05.fence::after_thread_sync every iteration
constexpr int F_WAIT = 8;      // successful mbarrier.try_wait (all lanes) every iteration
constexpr int F_WAIT1 = 16;    // successful mbarrier.try_wait by lane 0 only (+ __syncwarp)
constexpr int F_MMA3 = 32;     // 3 MMAs per iteration instead of 1
constexpr int F_NOMMA = 64;    // no MMA at all
constexpr int F_SAMED = 128;   // all MMAs accumulate into the same D (dependent chain)
constexpr int F_TESTWAIT = 256;  // test_wait (non-blocking) instead of try_wait

template <int F>
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, unsigned long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (tid == 0) { for (int i = 0; i < 32; ++i) mbar_init(bar0 + 8 * i, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) mbar_arrive(bar0 + 8 * 20);  // barrier 20: phase 0 complete -> wait(parity 0) always succeeds
  __syncthreads();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = idesc_bf16_f32(128, N);
  const uint64_t bdesc = smem_desc_kmajor_noswz(smem_u32(smem), N * 16, 128);
  if (warp == 1) {
    unsigned long long t0 = 0, t1 = 0;
    const bool active = (F & F_WARP) ? true : lane == 0;
    if (active) {
      t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tm + ((F & F_SAMED) ? 256u : 256u + 64u * (i & 3)), a = tm + 16u * (i & 7);
        if (F & F_WAIT) {
          if (F & F_TESTWAIT) {
            uint32_t ok = 0;
            while (!ok)
              asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar0 + 160), "r"(0) : "memory");
          } else mbar_wait(bar0 + 8 * 20, 0);
        }
        if (F & F_WAIT1) { if (lane == 0) mbar_wait(bar0 + 8 * 20, 0); __syncwarp(); }
        if (F & F_FENCE) fence_after_sync();
        if (F & F_WARP) {
          if (elect_one()) {
            if (!(F & F_NOMMA)) {
              mma_ts(d, a, bdesc, idesc, 1u);
              if (F & F_MMA3) { mma_ts(d, a + 8, bdesc, idesc, 1u); mma_ts(d, a, bdesc + 64, idesc, 1u); }
            }
            if (F & F_COMMIT) mma_commit(bar0 + 8 * (1 + (i & 15)));
          }
          __syncwarp();
        } else {
          if (!(F & F_NOMMA)) {
            mma_ts(d, a, bdesc, idesc, 1u);
            if (F & F_MMA3) { mma_ts(d, a + 8, bdesc, idesc, 1u); mma_ts(d, a, bdesc + 64, idesc, 1u); }
          }
          if (F & F_COMMIT) mma_commit(bar0 + 8 * (1 + (i & 15)));
        }
      }
      if (lane == 0) { mma_commit(bar0); mbar_wait(bar0, 0); }
      t1 = clock64();
    }
    if (lane == 0) out[0] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int F>
void run(const char* name, int N, unsigned long long* d) {
  const int iters = 512;
  cudaFuncSetAttribute(probe<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<F><<<1, 128, 64 * 1024>>>(N, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d %-44s %s %7.1f clk/iter\n", N, name, e == cudaSuccess ? "ok" : cudaGetErrorString(e), double(h) / iters);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  for (int N : {32, 208}) {
    run<0>("lane0: 1 mma", N, d);
    run<F_SAMED>("lane0: 1 mma same D", N, d);
    run<F_MMA3>("lane0: 3 mma", N, d);
    run<F_MMA3 | F_SAMED>("lane0: 3 mma same D", N, d);
    run<F_COMMIT>("lane0: 1 mma + commit", N, d);
    run<F_NOMMA | F_COMMIT>("lane0: commit only", N, d);
    run<F_WARP>("warp: 1 mma", N, d);
    run<F_WARP | F_NOMMA>("warp: elect+syncwarp only", N, d);
    run<F_WARP | F_COMMIT>("warp: 1 mma + commit", N, d);
    run<F_WARP | F_FENCE>("warp: 1 mma + fence", N, d);
    run<F_WARP | F_WAIT>("warp: 1 mma + try_wait(all lanes)", N, d);
    run<F_WARP | F_WAIT | F_TESTWAIT>("warp: 1 mma + test_wait(all lanes)", N, d);
    run<F_WARP | F_WAIT1>("warp: 1 mma + try_wait(lane0)+syncwarp", N, d);
    run<F_WARP | F_NOMMA | F_WAIT>("warp: try_wait only", N, d);
    run<F_WAIT>("lane0: 1 mma + try_wait", N, d);
    run<F_NOMMA | F_WAIT>("lane0: try_wait only", N, d);
    run<F_WARP | F_MMA3 | F_SAMED | F_WAIT | F_FENCE>("warp: 3 mma sameD + wait + fence", N, d);
    run<F_WARP | F_MMA3 | F_SAMED | F_WAIT | F_FENCE | F_COMMIT>("warp: 3 mma sameD + wait + fence + commit", N, d);
    run<F_MMA3 | F_SAMED | F_WAIT | F_FENCE>("lane0: 3 mma sameD + wait + fence", N, d);
    run<F_MMA3 | F_SAMED | F_WAIT | F_FENCE | F_COMMIT>("lane0: 3 mma sameD + wait + fence + commit", N, d);
  }
  return 0;
}
