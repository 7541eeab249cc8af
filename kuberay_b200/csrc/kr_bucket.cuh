// kr_bucket.cuh — bucketing pods by RayCluster in List order: count / scan / place (fast pipeline) and the stable LSD radix sort (general pipeline).
// Part of the sm_100a kernel set of the batched reconcile engine; see kr_kernels.cuh for the pipeline overview.
#pragma once

#include "kr_match.cuh"

namespace kr {

// ------------------------------------------------------------------------------------------------ fast pipeline: scan + place
// Exclusive scan of the per-cluster pod counts (bucket n_clusters = orphans) -> cstart[0 .. n_clusters+1].
// Flags buckets too large for the in-warp sort (the engine then re-runs the pass on the radix pipeline).
#define KR_FAST_MAX_BUCKET 1024u
// Chained multi-block exclusive scan: block `chunk` scans 8192 consecutive counters (8 per thread), waits for the inclusive
// carry of block chunk-1, adds it and publishes its own.  Blocks are dispatched in index order, so a waiting block's
// predecessor is always running or done.  v[] returns this thread's 8 exclusive prefixes; chain = {ready flag, carry} pairs,
// zeroed before the launch.
static constexpr uint32_t kScanChunk = 8192;
__device__ __forceinline__ uint32_t chained_scan_chunk(const uint32_t *__restrict__ in, uint32_t n, uint32_t chunk, uint32_t *chain,
                                                       uint32_t big_limit, bool &big, uint32_t (&excl)[8], uint32_t *s_warp, uint32_t *s_prefix) {
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint32_t i0 = chunk * kScanChunk + t * 8;
  uint32_t v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { sum += v[k]; big |= (i0 + k < big_limit) && v[k] > KR_FAST_MAX_BUCKET; }
  uint32_t x = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  uint32_t wv = s_warp[lane], wx = wv;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wx, d); if (lane >= d) wx += y; }
  const uint32_t woff = __shfl_sync(0xFFFFFFFFu, wx - wv, w), total = __shfl_sync(0xFFFFFFFFu, wx, 31);
  if (t == 0) {
    uint32_t prefix = 0;
    if (chunk > 0) {
      volatile uint32_t *prev = chain + 2 * (size_t)(chunk - 1);
      while (prev[0] == 0) {}
      __threadfence();
      prefix = prev[1];
    }
    chain[2 * (size_t)chunk + 1] = prefix + total;
    __threadfence();
    reinterpret_cast<volatile uint32_t *>(chain)[2 * (size_t)chunk] = 1;
    *s_prefix = prefix;
  }
  __syncthreads();
  uint32_t run = *s_prefix + woff + x - sum;
#pragma unroll
  for (int k = 0; k < 8; k++) { excl[k] = run; run += v[k]; }
  return *s_prefix + total;  // inclusive carry after this chunk
}

// Bucket starts: exclusive scan of the per-cluster pod counts (blocks [0, nchunks_c)) and of the per-tile orphan counts
// (the remaining blocks).  Flags real clusters too large for the in-warp sort (the orphan bucket is exempt: it is never sorted).
__global__ void __launch_bounds__(1024) k_scan_counts(const uint32_t *__restrict__ ccount, uint32_t *__restrict__ cstart, uint32_t nb, uint32_t nchunks_c,
                                                      uint32_t *__restrict__ tile_orph, uint32_t ntiles, uint32_t *chain, uint32_t *totals) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_prefix;
  uint32_t excl[8];
  bool big = false;
  if (blockIdx.x < nchunks_c) {
    const uint32_t chunk = blockIdx.x;
    uint32_t carry = chained_scan_chunk(ccount, nb, chunk, chain, nb - 1, big, excl, s_warp, &s_prefix);
    const uint32_t i0 = chunk * kScanChunk + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) if (i0 + k < nb) cstart[i0 + k] = excl[k];
    if (chunk == nchunks_c - 1 && threadIdx.x == 0) cstart[nb] = carry;
    if (big) KR_MARK_ATTEMPT_VOID(totals);
  } else {
    const uint32_t chunk = blockIdx.x - nchunks_c;
    chained_scan_chunk(tile_orph, ntiles, chunk, chain + 2 * (size_t)nchunks_c, 0, big, excl, s_warp, &s_prefix);
    __syncthreads();  // every thread of the block has read its inputs (in place)
    const uint32_t i0 = chunk * kScanChunk + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) if (i0 + k < ntiles) tile_orph[i0 + k] = excl[k];
  }
}

// pod -> its slot in the cluster's bucket: cstart[cluster] + arrival rank (order inside a bucket is fixed up by the
// in-warp sort in k_decide, so the result does not depend on the order the atomics landed in).
__global__ void __launch_bounds__(256) k_place(const uint32_t *__restrict__ key, const uint32_t *__restrict__ rank,
                                               const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ tile_orph,
                                               uint32_t *__restrict__ out, uint32_t n, uint32_t n_clusters) {
  uint32_t p = blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; k++, p += 256)
    if (p < n) {
      uint32_t c = __ldg(&key[p]);
      uint32_t pos = __ldg(&cstart[c]) + __ldg(&rank[p]);
      if (c == n_clusters) pos += __ldg(&tile_orph[p / kMatchTile]);  // orphans: already in List order, bucket of any size
      out[pos] = p;
    }
}

// Bitonic sort of 32*K values held K per lane (element g = lane*K + k); ascending.
template <int K>
__device__ __forceinline__ void warp_bitonic_sort(uint32_t (&v)[K], uint32_t lane) {
#pragma unroll
  for (int size = 2; size <= 32 * K; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= K) {
        const int ls = stride / K;
        const bool lower = (lane & ls) == 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
          uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v[k], ls);
          bool asc = ((lane * K + k) & size) == 0;
          v[k] = (asc == lower) ? min(v[k], o) : max(v[k], o);
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; k++) {
          if ((k & stride) == 0) {
            bool asc = ((lane * K + k) & size) == 0;
            uint32_t lo = min(v[k], v[k + stride]), hi = max(v[k], v[k + stride]);
            v[k] = asc ? lo : hi; v[k + stride] = asc ? hi : lo;
          }
        }
      }
    }
  }
}

// Sort one bucket of pod indices ascending (= informer List order): in[0..P) -> out[0..P), P <= 32*K.
template <int K>
__device__ __forceinline__ void warp_sort_bucket(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t P, uint32_t lane) {
  uint32_t v[K];
#pragma unroll
  for (int k = 0; k < K; k++) { uint32_t g = lane * K + k; v[k] = g < P ? in[g] : 0xFFFFFFFFu; }
  warp_bitonic_sort<K>(v, lane);
#pragma unroll
  for (int k = 0; k < K; k++) { uint32_t g = lane * K + k; if (g < P) out[g] = v[k]; }
}

__device__ __forceinline__ void warp_sort_dispatch(const uint32_t *in, uint32_t *out, uint32_t P, uint32_t lane) {
  if (P <= 32) warp_sort_bucket<1>(in, out, P, lane);
  else if (P <= 128) warp_sort_bucket<4>(in, out, P, lane);
  else if (P <= 256) warp_sort_bucket<8>(in, out, P, lane);
  else warp_sort_bucket<32>(in, out, P, lane);
}

// ------------------------------------------------------------------------------------------------ radix sort (stable LSD)

__global__ void __launch_bounds__(kSortThreads) k_hist(const uint32_t *__restrict__ keys, uint32_t *__restrict__ hist, uint32_t n, int shift) {
  __shared__ uint32_t s_hist[kRadix];
  const uint32_t tile = blockIdx.x, ntiles = gridDim.x;
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = tile * kSortTile + threadIdx.x;
#pragma unroll
  for (int it = 0; it < kSortItems; it++) {
    uint32_t i = base + it * kSortThreads;
    if (i < n) atomicAdd(&s_hist[(__ldg(&keys[i]) >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * ntiles + tile] = s_hist[threadIdx.x];
}

// Exclusive scan along each digit row of hist[256][ntiles] in place (block d = digit d) + the row total.
// k_scatter turns the 256 row totals into digit bases itself, so no single-block scan sits on the critical path.
static constexpr int kRowScanThreads = 128;
__global__ void __launch_bounds__(kRowScanThreads) k_scan_rows(uint32_t *__restrict__ hist, uint32_t *__restrict__ row_total, uint32_t ntiles) {
  __shared__ uint32_t s_warp[kRowScanThreads / 32];
  __shared__ uint32_t s_carry;
  uint32_t *row = hist + (size_t)blockIdx.x * ntiles;
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += kRowScanThreads * 4) {
    uint32_t i0 = base + t * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i0 + k < ntiles) ? row[i0 + k] : 0u;
    uint32_t sum = v[0] + v[1] + v[2] + v[3], x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kRowScanThreads / 32; k++) { uint32_t wv = s_warp[k]; if (k < (int)w) woff += wv; total += wv; }
    uint32_t run = s_carry + woff + x - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (i0 + k < ntiles) row[i0 + k] = run; run += v[k]; }
    __syncthreads();
    if (t == 0) s_carry += total;
    __syncthreads();
  }
  if (t == 0) row_total[blockIdx.x] = s_carry;
}

// Stable scatter of one tile: warp-match ranking keeps equal digits in original order.
// first_pass: values are the identity (pod index == position). write_keys: needed unless the consumer only wants values.
__global__ void __launch_bounds__(kSortThreads) k_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                          uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                          const uint32_t *__restrict__ hist, const uint32_t *__restrict__ row_total,
                                                          uint32_t n, int shift, int first_pass) {
  __shared__ uint32_t s_cnt[kSortThreads / 32][kRadix];
  __shared__ uint32_t s_base[kRadix];
  __shared__ uint32_t s_wsum[kSortThreads / 32];
  const uint32_t tile = blockIdx.x, ntiles = gridDim.x;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < kRadix; i += 32) s_cnt[warp][i] = 0;
  {  // digit base = exclusive scan of the 256 row totals (thread d owns digit d) + this tile's offset inside the row
    uint32_t tot = __ldg(&row_total[threadIdx.x]), x = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int k = 0; k < kSortThreads / 32; k++) if (k < (int)warp) woff += s_wsum[k];
    s_base[threadIdx.x] = woff + x - tot + __ldg(&hist[threadIdx.x * ntiles + tile]);
  }
  __syncwarp();
  const uint32_t base = tile * kSortTile + warp * (32 * kSortItems) + lane;
  uint32_t key[kSortItems], rank[kSortItems];
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int it = 0; it < kSortItems; it++) {
    uint32_t i = base + it * 32;
    bool valid = i < n;
    key[it] = valid ? __ldg(&keys_in[i]) : 0u;
    uint32_t d = valid ? ((key[it] >> shift) & (kRadix - 1)) : kRadix;  // sentinel digit for the ragged tail
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
    uint32_t prefix = __popc(peers & lt);
    uint32_t old = 0;
    if (valid) old = s_cnt[warp][d];
    __syncwarp();
    if (valid && prefix == 0) s_cnt[warp][d] = old + __popc(peers);
    __syncwarp();
    rank[it] = old + prefix;
  }
  __syncthreads();
  {  // per digit: exclusive scan over the 8 warps, add the tile's global base
    uint32_t d = threadIdx.x, run = s_base[d];
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; w++) { uint32_t v = s_cnt[w][d]; s_cnt[w][d] = run; run += v; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; it++) {
    uint32_t i = base + it * 32;
    if (i >= n) continue;
    uint32_t d = (key[it] >> shift) & (kRadix - 1);
    uint32_t dst = s_cnt[warp][d] + rank[it];
    keys_out[dst] = key[it];
    vals_out[dst] = first_pass ? i : __ldg(&vals_in[i]);
  }
}

}  // namespace kr
