// mma_rate.cu — micro-benchmark of the tcgen05 issue loop used by rollout_tc_kernel.
// Measures clocks per K-chunk for TS MMAs (A in TMEM, B in SMEM) under different issue patterns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I blackbox_mpc_b200/csrc tools/probe/mma_rate.cu -o tools/probe/mma_rate
#include <cstdio>
#include <cstdlib>
#include "tc05.cuh"
using namespace tc05;

struct Args { int variant, N, iters, per_chunk; unsigned long long* out; };

__global__ void __launch_bounds__(128, 1) probe(Args a) {
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
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = idesc_bf16_f32(128, a.N);
  const uint32_t kstep = a.N * 16;
  const uint64_t bdesc = smem_desc_kmajor_noswz(smem_u32(smem), kstep, 128);
  if (warp == 1) {
    unsigned long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    if (a.variant == 0) {  // back-to-back, one commit at the end
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < a.iters; ++i)
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256, tm + 16 * (i % 13), bdesc, idesc, 1u);
        mma_commit(bar0);
        mbar_wait(bar0, 0);
        t1 = clock64();
      }
    } else if (a.variant == 1) {  // commit per chunk onto a ring of barriers, single thread, no waits
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < a.iters; ++i) {
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256, tm + 16 * (i % 13), bdesc, idesc, 1u);
          mma_commit(bar0 + 8 * (1 + (i % 16)));
        }
        mma_commit(bar0);
        mbar_wait(bar0, 0);
        t1 = clock64();
      }
    } else if (a.variant == 2) {  // warp-wide loop: elect + syncwarp, commit per chunk
      t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        fence_after_sync();
        if (elect_one()) {
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256, tm + 16 * (i % 13), bdesc, idesc, 1u);
          mma_commit(bar0 + 8 * (1 + (i % 16)));
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(bar0);
      __syncwarp();
      mbar_wait(bar0, 0);
      t1 = clock64();
    } else if (a.variant == 3) {  // like 2, but waits for the commit of chunk i-8 before issuing chunk i (ring reuse)
      t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        if (i >= 8) mbar_wait(bar0 + 8 * (1 + ((i - 8) % 16)), (((i - 8) / 16) & 1));
        fence_after_sync();
        if (elect_one()) {
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256, tm + 16 * (i % 13), bdesc, idesc, 1u);
          mma_commit(bar0 + 8 * (1 + (i % 16)));
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(bar0);
      __syncwarp();
      mbar_wait(bar0, 0);
      t1 = clock64();
    } else if (a.variant == 4) {  // SS: A from smem too
      const uint64_t adesc = smem_desc_kmajor_noswz(smem_u32(smem), 2048, 128);
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < a.iters; ++i)
          for (int k = 0; k < a.per_chunk; ++k) mma_ss(tm + 256, adesc, bdesc, idesc, 1u);
        mma_commit(bar0);
        mbar_wait(bar0, 0);
        t1 = clock64();
      }

    } else if (a.variant >= 6 && a.variant <= 9) {
      // 6: wait (all lanes) without the tcgen05 fence; 7: lane-0-only wait + syncwarp + fence;
      // 8: test_wait polling instead of try_wait; 9: two waits per chunk
      t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        if (i >= 8) {
          const uint32_t b = bar0 + 8 * (1 + ((i - 8) % 16)), par = (((i - 8) / 16) & 1);
          if (a.variant == 6) { mbar_wait(b, par); }
          else if (a.variant == 7) { if (lane == 0) mbar_wait(b, par); __syncwarp(); fence_after_sync(); }
          else if (a.variant == 8) {
            uint32_t ok = 0;
            while (!ok) {
              asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            }
            fence_after_sync();
          } else { mbar_wait(b, par); mbar_wait(b, par); fence_after_sync(); }
        }
        if (elect_one()) {
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256, tm + 16 * (i % 13), bdesc, idesc, 1u);
          mma_commit(bar0 + 8 * (1 + (i % 16)));
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(bar0);
      __syncwarp();
      mbar_wait(bar0, 0);
      t1 = clock64();
    } else if (a.variant == 10) {  // waits but NO commit per chunk (commit every 4 chunks onto the ring)
      t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        if (i >= 8 && (i & 3) == 0) { mbar_wait(bar0 + 8 * (1 + (((i - 8) >> 2) % 16)), ((((i - 8) >> 2) / 16) & 1)); fence_after_sync(); }
        if (elect_one()) {
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256, tm + 16 * (i % 13), bdesc, idesc, 1u);
          if ((i & 3) == 3) mma_commit(bar0 + 8 * (1 + ((i >> 2) % 16)));
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(bar0);
      __syncwarp();
      mbar_wait(bar0, 0);
      t1 = clock64();
    } else if (a.variant == 5) {  // back-to-back TS, alternating two accumulators (no D dependency between neighbours)
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < a.iters; ++i)
          for (int k = 0; k < a.per_chunk; ++k) mma_ts(tm + 256 * (k & 1 ? 0 : 1) + (k & 1 ? 224 : 0) * 0, tm + 16 * (i % 13), bdesc, idesc, 1u);
        mma_commit(bar0);
        mbar_wait(bar0, 0);
        t1 = clock64();
      }
    }
    (void)ph;
    if (lane == 0) { a.out[0] = t1 - t0; }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int Ns[] = {208, 32};
  for (int N : Ns)
    for (int variant : {2, 3, 6, 7, 8, 9, 10})
      for (int per_chunk : {1, 3}) {
        Args a{variant, N, 416, per_chunk, d};
        probe<<<1, 128, 64 * 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d variant=%d mma/chunk=%d: %s  %.1f clk/chunk  (%.1f clk/mma; floor %d)\n", N, variant, per_chunk,
               e == cudaSuccess ? "ok" : cudaGetErrorString(e), double(h) / a.iters, double(h) / a.iters / per_chunk, 128 * N / 256);
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
