// utccp_probe.cu — can the A operand be staged SMEM -> TMEM with tcgen05.cp (UTCCP) so that the MMAs run in TS mode?
//   part 1 (layout): a K16 bf16 chunk in the kernel's K-major no-swizzle layout (two 16-byte K-slabs, LBO apart; 8-row
//           groups SBO = 128 B apart) copied with tcgen05.cp.128x256b, read back with tcgen05.ld and compared;
//   part 2 (rate): clocks per K-chunk (3 MMAs, M128 x N x K16) for  TS (A resident) / SS / cp+TS / cp only / hi in TMEM + lo SS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I blackbox_mpc_b200/csrc tools/probe/utccp_probe.cu -o tools/probe/utccp_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "tc05.cuh"
using namespace tc05;

struct Args { int variant, N, iters; unsigned long long* out; uint32_t* dump; };

__device__ __forceinline__ void utccp_128x256b(uint32_t dst_tmem, uint64_t src_desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(dst_tmem), "l"(src_desc) : "memory");
}

constexpr uint32_t A_LBO = 128 * 16;   // second K-slab of a chunk (rows x 16 B)
constexpr uint32_t A_CHUNK = 2 * A_LBO;

__global__ void __launch_bounds__(128, 1) probe(Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* A = smem;                  // 16 chunks of A (variants < 7 use 4: hi, lo, hi', lo')
  uint8_t* B = smem + 16 * A_CHUNK;   // 12 K16 x N chunks of B (variants < 7 use one)
  for (int i = tid; i < (16 * A_CHUNK + 12 * 208 * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  __syncthreads();
  // element (r, k) of chunk c = 0x(c)(r)(k): unique 16-bit patterns
  for (int i = tid; i < 4 * 128 * 16; i += 128) {
    const int c = i / (128 * 16), r = (i / 16) % 128, k = i % 16;
    const uint32_t off = c * A_CHUNK + (k / 8) * A_LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<uint16_t*>(A + off) = static_cast<uint16_t>((c << 11) | (r << 4) | k);
  }
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8 * i, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = idesc_bf16_f32(128, a.N);
  const uint64_t bdesc = smem_desc_kmajor_noswz(smem_u32(B), a.N * 16, 128);
  const uint64_t adesc0 = smem_desc_kmajor_noswz(smem_u32(A), A_LBO, 128);
  const uint64_t achunk = A_CHUNK >> 4;
  if (a.variant == 100) {   // layout check
    if (warp == 1 && lane == 0) {
      utccp_128x256b(tm + 448, adesc0 + 2 * achunk);     // chunk 2 -> columns 448..455
      utccp_128x256b(tm + 456, adesc0 + 3 * achunk);     // chunk 3 -> columns 456..463
      mma_commit(bar0);
    }
    mbar_wait(bar0, 0);
    fence_after_sync();
    uint32_t r[16];
    tmem_ld16(tm + 448 + ((warp * 32u) << 16), r);
    wait_ld();
    for (int j = 0; j < 16; ++j) a.dump[(warp * 32 + lane) * 16 + j] = r[j];
  } else if (warp == 1) {
    unsigned long long t0 = 0, t1 = 0;
    if (lane == 0) {
      // the TS variants read A from columns 448.. (whatever they hold: timing only)
      t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        const uint32_t st = tm + 448 + 16 * (i & 1);
        const uint64_t ad = adesc0 + (2 * (i & 1)) * achunk;
        switch (a.variant) {
          case 0: mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st + 8, bdesc, idesc, 1u); break;
          case 1: mma_ss(tm, ad, bdesc, idesc, 1u); mma_ss(tm, ad, bdesc, idesc, 1u); mma_ss(tm, ad + achunk, bdesc, idesc, 1u); break;
          case 2: utccp_128x256b(st, ad); utccp_128x256b(st + 8, ad + achunk);
                  mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st + 8, bdesc, idesc, 1u); break;
          case 3: utccp_128x256b(st, ad); utccp_128x256b(st + 8, ad + achunk); break;
          case 4: utccp_128x256b(st, ad);
                  mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st, bdesc, idesc, 1u); mma_ss(tm, ad + achunk, bdesc, idesc, 1u); break;
          case 5: // cp issued one chunk ahead of its MMAs
                  utccp_128x256b(tm + 448 + 16 * ((i + 1) & 1), ad); utccp_128x256b(tm + 448 + 16 * ((i + 1) & 1) + 8, ad + achunk);
                  mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st, bdesc, idesc, 1u); mma_ts(tm, st + 8, bdesc, idesc, 1u); break;
          case 7: {  // SS x3 with the kernel's addressing: A hi/lo of K-chunk i%8 from a 16-chunk ring, B hi/lo from a 6-stage ring
                  const uint64_t ah = adesc0 + (2 * (i & 7)) * achunk, al = ah + achunk;
                  const uint64_t bh = bdesc + (uint64_t)((2 * (i % 6)) * (a.N * 32) >> 4), bl = bh + (uint64_t)((a.N * 32) >> 4);
                  mma_ss(tm, ah, bh, idesc, 1u); mma_ss(tm, al, bh, idesc, 1u); mma_ss(tm, ah, bl, idesc, 1u); break; }
          case 8: {  // TS x3, same B addressing, A hi/lo resident in TMEM columns
                  const uint32_t ah = tm + 256 + 16 * (i % 13), al = ah + 8;
                  const uint64_t bh = bdesc + (uint64_t)((2 * (i % 6)) * (a.N * 32) >> 4), bl = bh + (uint64_t)((a.N * 32) >> 4);
                  mma_ts(tm, ah, bh, idesc, 1u); mma_ts(tm, al, bh, idesc, 1u); mma_ts(tm, ah, bl, idesc, 1u); break; }
          case 9: {  // cp + TS x3 with the same addressing
                  const uint64_t ah = adesc0 + (2 * (i & 7)) * achunk, al = ah + achunk;
                  const uint64_t bh = bdesc + (uint64_t)((2 * (i % 6)) * (a.N * 32) >> 4), bl = bh + (uint64_t)((a.N * 32) >> 4);
                  utccp_128x256b(st, ah); utccp_128x256b(st + 8, al);
                  mma_ts(tm, st, bh, idesc, 1u); mma_ts(tm, st + 8, bh, idesc, 1u); mma_ts(tm, st, bl, idesc, 1u); break; }
          case 10: {  // TS x3 into three independent accumulators (column offsets 0 / 64 / 128 for N <= 64; 0 / 208 / 0 otherwise)
                  const uint32_t ah = tm + 448, al = ah + 8;
                  const uint32_t d1 = a.N <= 64 ? 64 : 208, d2 = a.N <= 64 ? 128 : 0;
                  mma_ts(tm, ah, bdesc, idesc, 1u); mma_ts(tm + d1, al, bdesc, idesc, 1u); mma_ts(tm + d2, ah, bdesc, idesc, 1u); break; }
          case 11: {  // SS x3 into three independent accumulators
                  const uint32_t d1 = a.N <= 64 ? 64 : 208, d2 = a.N <= 64 ? 128 : 0;
                  mma_ss(tm, ad, bdesc, idesc, 1u); mma_ss(tm + d1, ad + achunk, bdesc, idesc, 1u); mma_ss(tm + d2, ad, bdesc, idesc, 1u); break; }
          case 12: {  // SS x2: hi x [Whi | Wlo] as ONE MMA of width 2N, lo x Whi as the second (N <= 64 only)
                  const uint32_t id2 = idesc_bf16_f32(128, 2 * a.N <= 256 ? 2 * a.N : a.N);
                  mma_ss(tm, ad, bdesc, id2, 1u); mma_ss(tm + 128, ad + achunk, bdesc, idesc, 1u); break; }
          case 6: // alternate two accumulators (two tiles in flight)
                  utccp_128x256b(st, ad); utccp_128x256b(st + 8, ad + achunk);
                  mma_ts(tm + 208 * (i & 1), st, bdesc, idesc, 1u); mma_ts(tm + 208 * (i & 1), st, bdesc, idesc, 1u); mma_ts(tm + 208 * (i & 1), st + 8, bdesc, idesc, 1u); break;
        }
      }
      mma_commit(bar0);
      mbar_wait(bar0, 0);
      t1 = clock64();
      a.out[0] = t1 - t0;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  unsigned long long* d; uint32_t* dump;
  cudaMalloc(&d, 64); cudaMalloc(&dump, 128 * 16 * 4);
  const int smem_bytes = 16 * A_CHUNK + 12 * 208 * 32;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  {
    Args a{100, 208, 0, d, dump};
    probe<<<1, 128, smem_bytes>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("layout probe: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> h(128 * 16);
    cudaMemcpy(h.data(), dump, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int j = 0; j < 16; ++j) {
        const int c = 2 + j / 8, k = 2 * (j % 8);
        const uint32_t want = uint32_t((c << 11) | (r << 4) | k) | (uint32_t((c << 11) | (r << 4) | (k + 1)) << 16);
        if (h[r * 16 + j] != want) { if (bad < 12) printf("  lane %3d col %2d: got %08x want %08x\n", r, j, h[r * 16 + j], want); ++bad; }
      }
    printf("layout: tcgen05.cp.128x256b of a K-major no-swizzle K16 chunk -> TMEM A operand layout: %s (%d mismatches)\n", bad ? "MISMATCH" : "MATCH", bad);
  }
  const char* names[] = {"TS x3 (A resident in TMEM)", "SS x3", "cp hi+lo, TS x3", "cp hi+lo only", "cp hi, TS x2 + SS lo", "cp one chunk ahead, TS x3", "cp + TS x3, two accumulators", "SS x3, ring addressing", "TS x3, ring addressing", "cp + TS x3, ring addressing", "TS x3, 3 accumulators", "SS x3, 3 accumulators", "SS x2 (N-concat hi|lo)"};
  for (int N : {208, 32, 64, 16})
    for (int v : {0, 1, 7, 8, 10, 11, 12}) {
      Args a{v, N, 416, d, dump};
      probe<<<1, 128, smem_bytes>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h = 0;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d %-32s: %s  %.1f clk/K-chunk\n", N, names[v], e == cudaSuccess ? "ok" : cudaGetErrorString(e), double(h) / a.iters);
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
