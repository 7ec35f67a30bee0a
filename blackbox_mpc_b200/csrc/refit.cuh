// refit.cuh — block-level selection primitives shared by the optimizer refit kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bbmpc {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX_K = 1024;

// Order-preserving float -> uint map (ascending).  Larger reward <=> larger key.
__device__ __forceinline__ uint32_t f2key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Exclusive prefix sum of one int per thread over a 1024-thread block; returns the block total
// through `total`.  `scratch` needs 33 ints.
__device__ inline int block_exclusive_scan(int v, int* scratch, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = scratch[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    scratch[lane] = w;
    if (lane == 31) scratch[32] = w;
  }
  __syncthreads();
  const int base = warp ? scratch[warp - 1] : 0;
  total = scratch[32];
  __syncthreads();
  return base + incl - v;
}

// (key desc, idx asc) comparison: does (ka, ia) come strictly before (kb, ib)?
__device__ __forceinline__ bool before(uint32_t ka, int ia, uint32_t kb, int ib) {
  return ka > kb || (ka == kb && ia < ib);
}

// In-place bitonic sort of n2 (power of two, <= 1024) (key, idx) pairs in shared memory into
// (key desc, idx asc) order.  All 1024 threads must call.
__device__ inline void bitonic_sort_desc(uint32_t* keys, int* idx, int n2) {
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      const int t = threadIdx.x;
      if (t < n2) {
        const int partner = t ^ stride;
        if (partner > t) {
          const bool up = ((t & size) == 0);  // this run sorted in "before" order
          const uint32_t ka = keys[t], kb = keys[partner];
          const int ia = idx[t], ib = idx[partner];
          const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
          if (swap) { keys[t] = kb; keys[partner] = ka; idx[t] = ib; idx[partner] = ia; }
        }
      }
    }
  }
  __syncthreads();
}

// Block-wide exact top-k of vals[i*stride] (i < n) with tf.nn.top_k's tie rule (lowest index
// first) [TF].  Output: out_keys/out_idx[0..k) in shared memory, sorted (value desc, index asc);
// slots beyond min(k, n) hold key 0 / idx INT_MAX.  k <= 1024.  smem scratch provided by caller:
//   hist[256], misc[40], out_keys[1024], out_idx[1024].
struct SelectScratch {
  int* hist; int* misc; uint32_t* out_keys; int* out_idx;
};
// `kcache` (nullable, n words of shared memory): the keys are read from global memory once and the five
// further passes run out of shared memory.
__device__ inline void block_topk(const float* __restrict__ vals, int n, int stride, int k,
                                  const SelectScratch& sc, uint32_t* kcache = nullptr) {
  const int tid = threadIdx.x;
  int k_eff = k < n ? k : n;
  if (kcache) {
    for (int i = tid; i < n; i += SEL_THREADS) kcache[i] = f2key(vals[static_cast<size_t>(i) * stride]);
    __syncthreads();
  }
  auto key_at = [&](int i) -> uint32_t { return kcache ? kcache[i] : f2key(vals[static_cast<size_t>(i) * stride]); };
  for (int i = tid; i < SEL_MAX_K; i += SEL_THREADS) { sc.out_keys[i] = 0u; sc.out_idx[i] = 0x7FFFFFFF; }
  if (k_eff == 0) { __syncthreads(); return; }
  // --- radix select the k_eff-th largest key
  uint32_t prefix = 0, mask = 0;
  int need = k_eff;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += SEL_THREADS) sc.hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) {
      const uint32_t ky = key_at(i);
      if ((ky & mask) == prefix) atomicAdd(&sc.hist[(ky >> shift) & 255], 1);
    }
    __syncthreads();
    if (tid < 32) {
      // bucket b with  sum(hist[b+1..255]) < need <= sum(hist[b..255])  (b = 0 if the sum never reaches need), by one warp:
      // lane L owns buckets 255-8L .. 248-8L (descending), suffix sums across the lanes with shuffles
      int h[8], mine = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { h[q] = sc.hist[255 - 8 * tid - q]; mine += h[q]; }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      int cum = incl - mine;                          // buckets above this lane's
      const bool here = cum < need && incl >= need;   // the crossing lies in this lane's eight buckets
      const unsigned ball = __ballot_sync(0xffffffffu, here);
      if (ball == 0u) { if (tid == 31) { sc.misc[34] = 0; sc.misc[35] = need - (incl - h[7]); } }   // never reached: bucket 0 takes the rest
      else if (here) {
        int b = 255 - 8 * tid;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (cum + h[q] >= need) break;
          cum += h[q]; --b;
        }
        sc.misc[34] = b; sc.misc[35] = need - cum;
      }
    }
    __syncthreads();
    prefix |= static_cast<uint32_t>(sc.misc[34]) << shift;
    mask |= 255u << shift;
    need = sc.misc[35];
    __syncthreads();
  }
  const uint32_t T = prefix;          // threshold key; take all > T and the first `need` == T
  const int n_greater = k_eff - need;
  if (tid == 0) sc.misc[36] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += SEL_THREADS) {
    const uint32_t ky = key_at(i);
    if (ky > T) { const int slot = atomicAdd(&sc.misc[36], 1); sc.out_keys[slot] = ky; sc.out_idx[slot] = i; }
  }
  // ties at the threshold.  Usually every key equal to T is taken (a unique threshold value: need == 1 == number of ties):
  // then the slot order does not matter (the final sort is by (key, index)) and one atomic counter places them.
  if (tid == 0) sc.misc[37] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += SEL_THREADS)
    if (key_at(i) == T) atomicAdd(&sc.misc[37], 1);
  __syncthreads();
  const int n_ties = sc.misc[37];
  __syncthreads();
  if (n_ties <= need) {
    if (tid == 0) sc.misc[37] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS)
      if (key_at(i) == T) { const int slot = n_greater + atomicAdd(&sc.misc[37], 1); sc.out_keys[slot] = T; sc.out_idx[slot] = i; }
  }
  // more ties than places: ordered scan so the lowest indices win
  int taken = 0;
  for (int base = 0; n_ties > need && base < n && taken < need; base += SEL_THREADS) {
    const int i = base + tid;
    const int flag = (i < n && key_at(i) == T) ? 1 : 0;
    int total;
    const int rank = block_exclusive_scan(flag, sc.misc, total);
    if (flag && taken + rank < need) { sc.out_keys[n_greater + taken + rank] = T; sc.out_idx[n_greater + taken + rank] = i; }
    taken += total;
  }
  __syncthreads();
  int n2 = 1;
  while (n2 < k_eff) n2 <<= 1;
  bitonic_sort_desc(sc.out_keys, sc.out_idx, n2);
}

}  // namespace bbmpc
