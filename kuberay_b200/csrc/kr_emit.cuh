// kr_emit.cuh — what follows the decisions: create offsets + replica indices, the compact action list, the fused variants, incremental pod-row patches, the RayJob roll-up.
// Part of the sm_100a kernel set of the batched reconcile engine; see kr_kernels.cuh for the pipeline overview.
#pragma once

#include "kr_decide.cuh"

namespace kr {

// ------------------------------------------------------------------------------------------------ creates

// exclusive scan of the dense n_create array -> groups[].create_off, total in totals[0] (chained multi-block scan).
__global__ void __launch_bounds__(1024) k_scan_creates(ResDev r, const uint32_t *__restrict__ gcreate, uint32_t n_groups, uint32_t *chain) {
  if (KR_ATTEMPT_VOID(r.totals)) return;
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_prefix;
  uint32_t excl[8];
  bool big = false;
  const uint32_t chunk = blockIdx.x;
  uint32_t carry = chained_scan_chunk(gcreate, n_groups, chunk, chain, 0, big, excl, s_warp, &s_prefix);
  const uint32_t i0 = chunk * kScanChunk + threadIdx.x * 8;
#pragma unroll
  for (int k = 0; k < 8; k++) if (i0 + k < n_groups) r.groups[i0 + k].create_off = excl[k];
  if (chunk == gridDim.x - 1 && threadIdx.x == 0) r.totals[0] = carry;
}

// Lowest free ray.io/worker-group-replica-index values for the pods to create (raycluster_controller.go:854-881).
// One warp per group; candidate indices are swept in windows of 1024 bits held in shared memory.
__device__ __forceinline__ void create_fill_group(const SnapDev &s, const ScratchDev &sc, const ResDev &r, const kr_flags &f, uint32_t g,
                                                  uint32_t create_off, uint32_t create_cap, uint32_t *s_bits /* [32] per warp */, uint32_t lane) {
  const kr_group_result gr = r.groups[g];
  if (gr.n_create == 0) return;
  const bool mh = (gr.flags & KR_GR_MULTIHOST) != 0;  // multi-host: in-use indices = label of the first pod of every valid replica (:1067-1077)
  if ((uint64_t)create_off + gr.n_create > create_cap) return;  // host reports KR_E_CAPACITY from totals[0]
  int32_t *out = r.create_idx + create_off;
  if (!f.gate_multihost_indexing) {  // createWorkerPod without an index (:884-889)
    for (uint32_t k = lane; k < gr.n_create; k += 32) out[k] = -1;
    return;
  }
  const uint32_t c = s.g_cluster_idx[g];
  const uint32_t slot = g - s.c_group_off[c];
  const kr_cluster_result *cr = &r.clusters[c];
  const uint32_t seg0 = cr->pod_start, seg1 = seg0 + (uint32_t)cr->n_pods;
  const uint64_t bound = (uint64_t)gr.n_running + gr.n_create;  // the n_create lowest free indices all lie below this
  uint32_t written = 0;
  for (uint64_t w0 = 0; w0 < bound && written < gr.n_create; w0 += 1024) {
    s_bits[lane] = 0;
    __syncwarp();
    for (uint32_t b = seg0; b < seg1; b += 32) {
      uint32_t i = b + lane;
      if (i < seg1) {
        const uint4 row = sc.rows[r.sorted_pod_idx[i]];
        // this group's pods first (mh_head is only written for them), then runningPods: listed and not deleted by name
        if ((row.w >> 16) == slot && (row.w & KR_PP_HAS_REPLICA_IDX) && (mh ? sc.mh_head[i] != 0 : r.sorted_action[i] == KR_ACT_KEEP)) {
          int32_t idx = (int32_t)row.z;
          if (idx >= 0 && (uint64_t)idx >= w0 && (uint64_t)idx < w0 + 1024 && (uint64_t)idx < bound)
            atomicOr(&s_bits[(idx - w0) >> 5], 1u << ((idx - w0) & 31));
        }
      }
    }
    __syncwarp();
    uint32_t word = s_bits[lane];
    uint64_t wbase = w0 + 32ull * lane;
    uint32_t freeb = ~word;
    if (wbase >= bound) freeb = 0;
    else if (bound - wbase < 32) freeb &= (1u << (uint32_t)(bound - wbase)) - 1;
    uint32_t cnt = __popc(freeb), x = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    uint32_t pos = written + x - cnt;
    while (freeb && pos < gr.n_create) {
      uint32_t bit = __ffs(freeb) - 1;
      freeb &= freeb - 1;
      out[pos++] = (int32_t)(wbase + bit);
    }
    written += __shfl_sync(0xFFFFFFFFu, x, 31);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) k_create_fill(SnapDev s, ScratchDev sc, ResDev r, Sizes n, kr_flags f, uint32_t create_cap) {
  if (KR_ATTEMPT_VOID(r.totals)) return;
  __shared__ uint32_t s_bits[4][32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * 4 + warp;
  if (g >= n.n_groups) return;
  create_fill_group(s, sc, r, f, g, r.groups[g].create_off, create_cap, s_bits[warp], lane);
}

// Compact action list of one cluster: (pod idx, action) of every pod whose action != KEEP, List order kept (one warp).
__device__ __forceinline__ void compact_cluster_actions(const ResDev &r, const ScratchDev &sc, uint32_t c, uint32_t dst, uint32_t cnt, uint32_t lane) {
  // the decide warp left the cluster's cnt actions compacted at its pod_start (List order): move them to their place in the list
  const size_t src = r.clusters[c].pod_start;
  for (uint32_t k = lane; k < cnt; k += 32) { r.act_pod_idx[dst + k] = sc.act_tmp_idx[src + k]; r.act_code[dst + k] = sc.act_tmp_code[src + k]; }
}

// unfused path: starts of the per-cluster action lists (chained scan) ...
__global__ void __launch_bounds__(1024) k_scan_actions(ResDev r, const uint32_t *__restrict__ cact, uint32_t n_clusters, uint32_t *chain) {
  if (KR_ATTEMPT_VOID(r.totals)) return;
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_prefix;
  uint32_t excl[8];
  bool big = false;
  const uint32_t chunk = blockIdx.x;
  uint32_t carry = chained_scan_chunk(cact, n_clusters, chunk, chain, 0, big, excl, s_warp, &s_prefix);
  const uint32_t i0 = chunk * kScanChunk + threadIdx.x * 8;
#pragma unroll
  for (int k = 0; k < 8; k++) if (i0 + k < n_clusters) { r.act_start[i0 + k] = excl[k]; r.act_cnt[i0 + k] = cact[i0 + k]; }
  if (chunk == gridDim.x - 1 && threadIdx.x == 0) r.act_start[n_clusters] = carry;
}
// ... and the lists themselves, one warp per cluster
__global__ void __launch_bounds__(128) k_compact_actions(ResDev r, ScratchDev sc, uint32_t n_clusters) {
  if (KR_ATTEMPT_VOID(r.totals)) return;
  uint32_t c = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= n_clusters) return;
  const uint32_t cnt = sc.cact[c];
  if (cnt == 0) return;
  compact_cluster_actions(r, sc, c, r.act_start[c], cnt, threadIdx.x & 31);
}

// ---- fused variants for snapshots whose per-cluster / per-group counters fit in shared memory: every block scans the counters
// itself (a few tens of KB out of L2) instead of waiting for a scan kernel, which removes two ~10 us stages from the chain.
static constexpr uint32_t kFusedMaxCounters = 48 * 1024;  // 192 KB of shared memory

// exclusive scan of in[0..n) into shared memory by the whole block (any block size that is a multiple of 32, <= 1024).
// 8 counters per thread per trip keeps the 1024-thread CTAs at 32 registers: with 16 the CTA no longer fits beside the blocks
// of the previous kernel and the programmatic early launch turns into a wait (k_place_fused started 25 us late)
__device__ __forceinline__ uint32_t block_scan_to_smem(const uint32_t *__restrict__ in, uint32_t n, uint32_t *out_sm, uint32_t big_limit, bool &big,
                                                       uint32_t *s_warp, uint32_t *s_carry) {
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
  if (t == 0) *s_carry = 0;
  __syncthreads();
  constexpr int V = 8;
  for (uint32_t base = 0; base < n; base += blockDim.x * V) {
    uint32_t i0 = base + t * V;
    uint32_t v[V];
#pragma unroll
    for (int k = 0; k < V; k++) v[k] = (i0 + k < n) ? __ldg(&in[i0 + k]) : 0u;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < V; k++) { sum += v[k]; big |= (i0 + k < big_limit) && v[k] > KR_FAST_MAX_BUCKET; }
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    uint32_t wv = lane < nw ? s_warp[lane] : 0u, wx = wv;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wx, d); if (lane >= d) wx += y; }
    uint32_t woff = __shfl_sync(0xFFFFFFFFu, wx - wv, w), total = __shfl_sync(0xFFFFFFFFu, wx, 31);
    uint32_t run = *s_carry + woff + x - sum;
#pragma unroll
    for (int k = 0; k < V; k++) { if (i0 + k < n) out_sm[i0 + k] = run; run += v[k]; }
    __syncthreads();
    if (t == 0) *s_carry += total;
    __syncthreads();
  }
  const uint32_t grand_total = *s_carry;
  __syncthreads();  // every thread has read the total before a following call resets the carry
  return grand_total;
}

// Two independent exclusive scans (a[0..na) -> outa, b[0..nb) -> outb) sharing every trip: both rounds of loads are in flight
// together and the barriers are paid once (k_creates_fused: 9 us of two back-to-back scans -> one pass).
__device__ __forceinline__ void block_scan2_to_smem(const uint32_t *__restrict__ a, uint32_t na, uint32_t *outa, const uint32_t *__restrict__ b, uint32_t nb,
                                                    uint32_t *outb, uint32_t *s_warp /* [64] */, uint32_t *s_carry /* [2] */, uint32_t &tota, uint32_t &totb) {
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
  if (t < 2) s_carry[t] = 0;
  __syncthreads();
  const uint32_t nmax = na > nb ? na : nb;
  for (uint32_t base = 0; base < nmax; base += blockDim.x * 8) {
    const uint32_t i0 = base + t * 8;
    uint32_t va[8], vb[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { va[k] = (i0 + k < na) ? __ldg(&a[i0 + k]) : 0u; vb[k] = (i0 + k < nb) ? __ldg(&b[i0 + k]) : 0u; }
    uint32_t sa = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { sa += va[k]; sb += vb[k]; }
    uint32_t xa = sa, xb = sb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t ya = __shfl_up_sync(0xFFFFFFFFu, xa, d), yb = __shfl_up_sync(0xFFFFFFFFu, xb, d);
      if (lane >= d) { xa += ya; xb += yb; }
    }
    if (lane == 31) { s_warp[w] = xa; s_warp[32 + w] = xb; }
    __syncthreads();
    uint32_t wa = lane < nw ? s_warp[lane] : 0u, wb = lane < nw ? s_warp[32 + lane] : 0u, pa = wa, pb = wb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t ya = __shfl_up_sync(0xFFFFFFFFu, pa, d), yb = __shfl_up_sync(0xFFFFFFFFu, pb, d);
      if (lane >= d) { pa += ya; pb += yb; }
    }
    const uint32_t offa = __shfl_sync(0xFFFFFFFFu, pa - wa, w), offb = __shfl_sync(0xFFFFFFFFu, pb - wb, w);
    const uint32_t tta = __shfl_sync(0xFFFFFFFFu, pa, 31), ttb = __shfl_sync(0xFFFFFFFFu, pb, 31);
    uint32_t ra = s_carry[0] + offa + xa - sa, rb = s_carry[1] + offb + xb - sb;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (i0 + k < na) outa[i0 + k] = ra;
      if (i0 + k < nb) outb[i0 + k] = rb;
      ra += va[k]; rb += vb[k];
    }
    __syncthreads();
    if (t == 0) { s_carry[0] += tta; s_carry[1] += ttb; }
    __syncthreads();
  }
  tota = s_carry[0]; totb = s_carry[1];
  __syncthreads();
}

// bucket starts + placement in one persistent kernel (replaces k_scan_counts + k_place)
__global__ void __launch_bounds__(1024) k_place_fused(const uint32_t *__restrict__ key, const uint32_t *__restrict__ rank, const uint32_t *__restrict__ ccount,
                                                      uint32_t *__restrict__ cstart, const uint32_t *__restrict__ tile_orph, uint32_t *__restrict__ out,
                                                      uint32_t n, uint32_t n_clusters, uint32_t ntiles, uint32_t *totals) {
  KR_TL(2);
  extern __shared__ uint32_t sm_dyn[];
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  pdl_wait(); pdl_trigger();
  uint32_t *sm_start = sm_dyn;                    // [n_clusters + 2]
  uint32_t *sm_orph = sm_dyn + n_clusters + 2;    // [ntiles]
  const uint32_t nb = n_clusters + 1;
  bool big = false, dummy = false;
  uint32_t tot = block_scan_to_smem(ccount, nb, sm_start, nb - 1, big, s_warp, &s_carry);
  if (threadIdx.x == 0) sm_start[nb] = tot;
  block_scan_to_smem(tile_orph, ntiles, sm_orph, 0, dummy, s_warp, &s_carry);
  __syncthreads();
  if (blockIdx.x == 0) {
    for (uint32_t i = threadIdx.x; i <= nb; i += blockDim.x) cstart[i] = sm_start[i];
    if (big) KR_MARK_ATTEMPT_VOID(totals);
  }
  // four pods per thread per trip: all eight loads are in flight before the first dependent shared-memory lookup
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < n; p0 += 4 * stride) {
    uint32_t c[4], rk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t p = p0 + k * stride;
      c[k] = p < n ? __ldg(&key[p]) : 0u;
      rk[k] = p < n ? __ldg(&rank[p]) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t p = p0 + k * stride;
      if (p >= n) break;
      uint32_t pos = sm_start[c[k]] + rk[k];
      if (c[k] == n_clusters) pos += sm_orph[p / kMatchTile];
      out[pos] = p;
    }
  }
}

// create offsets + replica-index allocation in one persistent kernel (replaces k_scan_creates + k_create_fill)
__global__ void __launch_bounds__(1024) k_creates_fused(SnapDev s, ScratchDev sc, ResDev r, Sizes n, kr_flags f, uint32_t create_cap) {
  KR_TL(6);
  extern __shared__ uint32_t sm_dyn[];
  __shared__ uint32_t s_warp[64];
  __shared__ uint32_t s_carry[2];
  __shared__ uint32_t s_bits[32][32];
  pdl_wait(); pdl_trigger();
  if (KR_ATTEMPT_VOID(r.totals)) return;  // (a void attempt never wrote the counters scanned below)
  uint32_t *sm_off = sm_dyn;               // [n_groups] create offsets
  uint32_t *sm_act = sm_dyn + n.n_groups;  // [n_clusters + 1] action-list starts
  uint32_t tot, tot_act;
  block_scan2_to_smem(sc.gcreate, n.n_groups, sm_off, sc.cact, n.n_clusters, sm_act, s_warp, s_carry, tot, tot_act);
  if (threadIdx.x == 0) sm_act[n.n_clusters] = tot_act;
  __syncthreads();
  KR_TL_POINT(10);
  if (blockIdx.x == 0) {
    for (uint32_t g = threadIdx.x; g < n.n_groups; g += blockDim.x) r.groups[g].create_off = sm_off[g];
    for (uint32_t c = threadIdx.x; c <= n.n_clusters; c += blockDim.x) { r.act_start[c] = sm_act[c]; if (c < n.n_clusters) r.act_cnt[c] = sm_act[c + 1] - sm_act[c]; }
    if (threadIdx.x == 0) r.totals[0] = tot;
  }
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (uint32_t g = blockIdx.x * nw + warp; g < n.n_groups; g += gridDim.x * nw)
    if (__ldg(&sc.gcreate[g])) create_fill_group(s, sc, r, f, g, sm_off[g], create_cap, s_bits[warp], lane);
  KR_TL_POINT(11);
  for (uint32_t c = blockIdx.x * nw + warp; c < n.n_clusters; c += gridDim.x * nw)
    if (sm_act[c + 1] != sm_act[c]) compact_cluster_actions(r, sc, c, sm_act[c], sm_act[c + 1] - sm_act[c], lane);
}

// ------------------------------------------------------------------------------------------------ k_patch_pods
// Incremental epoch: copy n updated pod rows from the pinned host arena (mapped, read over PCIe in 32-B sectors — the host never
// gathers them) into the resident columns.  Only the row list is staged.
struct PodCols { uint32_t *c[7]; };
__global__ void __launch_bounds__(256) k_patch_pods(const uint32_t *__restrict__ rows, uint32_t n, PodCols host, PodCols dev) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = rows[i];
  uint32_t v[7];
#pragma unroll
  for (int k = 0; k < 7; k++) v[k] = __ldcv(host.c[k] + p);  // volatile-cached: never served from a stale L2 line
#pragma unroll
  for (int k = 0; k < 7; k++) dev.c[k][p] = v[k];
}

// Journal-style incremental epoch: the rows arrive in one contiguous staging buffer ([n row indices][n x 7 values]); scatter them
// into the resident columns.  (Writing them through to the mapped pinned arena as well was measured: 70 k four-byte PCIe
// writes cost as much as the sector pulls of k_patch_pods, ~180 us — the caller keeps its arenas current itself.)
__global__ void __launch_bounds__(256) k_patch_pod_values(const uint32_t *__restrict__ stage, uint32_t n, PodCols dev) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = stage[i];
  const uint32_t *v = stage + n + 7 * (size_t)i;
#pragma unroll
  for (int k = 0; k < 7; k++) dev.c[k][p] = v[k];
}

// ------------------------------------------------------------------------------------------------ k_jobs
// RayJob roll-up (rayjob_controller.go:203-216, 343, 885): join by (namespace, status.rayClusterName).
__global__ void __launch_bounds__(256) k_jobs(SnapDev s, ScratchDev sc, ResDev r, Sizes n) {
  KR_TL(8);
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n.n_jobs) return;
  kr_job_result jr; jr.cluster_idx = -1; jr.cluster_state = 0; jr.not_ready = 0; jr.status_changed = 0; jr.reserved = 0;
  uint32_t c;
  if (cl_lookup(sc, s.j_ns_id[j], s.j_cluster_name_id[j], c)) {
    jr.cluster_idx = (int32_t)c;
    jr.cluster_state = s.c_old_state[c];
    jr.not_ready = s.c_old_state[c] != KR_STATE_READY;
    jr.status_changed = s.j_summary_id[j] != s.c_summary_id[c];
  }
  r.jobs[j] = jr;
}

}  // namespace kr
