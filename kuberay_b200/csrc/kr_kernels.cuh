// kr_kernels.cuh — sm_100a kernels of the batched reconcile engine.
//
// Integer / hash work: no tensor cores.  What matters here: coalesced SoA column streaming,
// shared-memory staging (hash chunks, per-tile digit counters, per-warp group accumulators), warp-ballot /
// match_any group-by, grids sized in multiples of the SM count.
//
// Pipeline of one pass (engine stream M unless noted); one CUDA graph, programmatic dependent launch along the chain:
//   k_clear          per-pass clears (hash tables, workersToDelete resolutions, totals, bucket counters) in one launch
//   k_build_tables   cluster table (ns,name)->idx (+ per-cluster group record), workersToDelete-name table, head-aux table
//   k_match          per pod: label/selector match -> cluster idx + group slot, 16-byte pod row, bucket rank
//   k_place_fused    bucket starts (scan in shared memory) + pod -> slot of its cluster's bucket   [large: k_scan_counts + k_place]
//   k_decide_small   one warp per RayCluster (<= 256 pods): in-register bitonic sort (List order), warp-ballot group-by,
//                    head / group decisions, ordered deletes, status roll-up        | k_decide: general path, side stream
//   k_hash2          (stream H, concurrent) SHA-1 + base32hex of every muted-spec JSON
//   k_decide phase 1 clusters whose Recreate gate needs the hash
//   k_creates_fused  create offsets + lowest free replica indices + compact action list   [large: k_scan_* + k_create_fill ...]
//   k_jobs           RayJob -> RayCluster status roll-up join
//   k_patch_pods     (copy stream, incremental epochs) rewritten pod rows pulled from the mapped pinned arena
//   radix pipeline   (k_match<radix>, k_hist, k_scan_rows, k_scatter): stable LSD sort, taken when a RayCluster has > 1024 pods
//
// Reference semantics restated here are cited per function (paths relative to
// ray-operator/controllers/ray/ in ray-project/kuberay).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kr_engine.h"

namespace kr {

// ------------------------------------------------------------------------------------------------ device views

struct SnapDev {  // device mirror of kr_snapshot_bufs
  const uint32_t *c_ns_id, *c_name_id;
  const uint64_t *c_uid_hash;
  const uint32_t *c_flags;
  const uint8_t *c_suspend_status, *c_ext_err_kind;
  const uint32_t *c_ext_err_msg_id, *c_group_off, *c_group_cnt;
  const uint64_t *c_json_off;
  const uint32_t *c_json_len;
  const uint8_t *c_old_state;
  const int32_t *c_old_counts;
  const uint8_t *c_old_cond_status, *c_old_cond_variant;
  const uint32_t *c_old_cond_reason_id, *c_old_cond_msg_id, *c_old_head_ids;
  const uint8_t *c_svc_count, *c_svc_ip_kind;
  const uint32_t *c_svc_ip_id, *c_svc_name_id;
  const uint32_t *g_cluster_idx, *g_name_id;
  const int32_t *g_replicas, *g_min, *g_max, *g_num_hosts;
  const uint32_t *g_flags, *g_wtd_off, *g_wtd_cnt;
  const uint32_t *w_name_id;
  const uint32_t *p_ns_id, *p_cluster_name_id, *p_group_name_id, *p_name_id, *p_packed;
  const int32_t *p_replica_index;
  const uint32_t *p_replica_name_id;
  const uint32_t *h_pod_idx;
  const uint8_t *h_ready_status;
  const uint32_t *h_ready_reason_id, *h_ready_msg_id, *h_pod_ip_id;
  const uint8_t *h_annot_state, *h_version_state, *h_annot_hash;
  const uint32_t *j_ns_id, *j_cluster_name_id, *j_summary_id, *c_summary_id;
  const uint8_t *json;
};

struct ResDev {  // device results arena
  kr_cluster_result *clusters;
  char *hash;
  kr_group_result *groups;
  uint32_t *wtd_pod_idx;  // unsigned for atomicMin; 0xFFFFFFFF == -1 == NotFound
  uint32_t *sorted_pod_idx;
  uint8_t *sorted_action;
  int32_t *create_idx;
  kr_job_result *jobs;
  uint32_t *act_start;    // [n_clusters + 1]
  uint32_t *act_pod_idx;  // [n_pods] capacity; n_actions used
  uint8_t *act_code;
  uint32_t *totals;  // [0]=n_create_total [1]=n_orphans [2]=n_actions [3]=error flags [4]=clusters deferred to decide phase 1
};

struct ScratchDev {
  uint4 *cl_slots; uint32_t cl_mask;                           // cluster table: {key lo, key hi, cluster idx, -} per 16-byte slot
  uint4 *cl_rec;                                               // [n_clusters] {group_off, group_cnt, name id of group 0, bit0 = has a multi-host group}
  uint64_t *wt_keys; uint32_t *wt_head; uint32_t *wt_next; uint32_t wt_mask;  // workersToDelete-name table
  uint32_t *aux_keys; uint32_t *aux_vals; uint32_t aux_mask;   // pod idx -> head-aux row
  uint4 *rows;                                                 // 16-byte pod rows, original order
  uint32_t *keys[2]; uint32_t *vals[2];                        // radix ping-pong
  uint32_t *hist;                                              // [256 * ntiles] digit-major
  uint32_t *row_total;                                         // [256] per-digit totals of the current pass
  uint32_t *gcreate;                                           // [n_groups] dense n_create (input of the creates scan)
  uint32_t *cact;                                              // [n_clusters] pods with an action per cluster (input of the action-list scan)
  uint32_t *mh_rep, *mh_name, *mh_meta, *mh_cnt, *mh_flg;      // multi-host scratch, indexed by sorted position
  uint8_t *mh_act, *mh_head;                                   // per position: action of a multi-host pod / first pod of a valid replica
  uint32_t *tile_orph;                                         // fast pipeline: orphans per k_match tile -> exclusive prefix
  uint32_t *chain;                                             // chained-scan hand-off cells {ready, carry} (zeroed with ccount)
  uint32_t *ccount, *cstart;                                   // fast pipeline: pods per cluster bucket [n_clusters+1], bucket starts [n_clusters+2]
  uint32_t *deferred_list;                                     // clusters left for decide phase 1 (count in totals[4])
  int32_t *gacc;                                               // [4 * n_groups] spill accumulators (clusters with > KR_SMEM_GROUPS groups)
};

struct Sizes { uint32_t n_clusters, n_groups, n_wtd, n_pods, n_heads, n_jobs; };

// row.w layout: low 16 bits = p_packed low bits (+ KR_ROW_WTD_OWN), high 16 bits = group slot inside the cluster
#define KR_ROW_WTD_OWN (1u << 11)   // named by its own group's scaleStrategy.workersToDelete
#define KR_ROW_NO_GROUP 0xFFFFu
#define KR_TOTALS_BIG_BUCKET 2u        // fast pipeline only: some cluster (or the orphan bucket) holds more pods than the in-warp sort takes


static constexpr int kSortThreads = 256;
static constexpr int kSortItems = 8;
static constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per tile
static constexpr int kMatchItems = 2;                        // fast pipeline: pods per thread in k_match (tile = 512 pods; occupancy beats per-thread ILP here: 8/4/2/1 items -> 47/40/33/33 us at C3)
static constexpr int kMatchTile = kSortThreads * kMatchItems;
static constexpr int kRadixBits = 8;
static constexpr int kRadix = 1 << kRadixBits;

// ------------------------------------------------------------------------------------------------ small helpers
#ifdef KR_TIMELINE
// Development aid (tools/timeline.py, built with -DKR_TIMELINE into a separate library): every kernel stamps the earliest
// block start and the latest block end it sees (%globaltimer, ns) so the gaps between the kernels of one graph replay show.
__device__ unsigned long long g_tl[64];
struct TlScope {
  int id;
  __device__ __forceinline__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
  __device__ __forceinline__ explicit TlScope(int i) : id(i) { if (threadIdx.x == 0) atomicMin(&g_tl[2 * id], now()); }
  __device__ __forceinline__ ~TlScope() { if (threadIdx.x == 0) atomicMax(&g_tl[2 * id + 1], now()); }
};
#define KR_TL(id) TlScope tl_scope_(id)
#else
#define KR_TL(id)
#endif

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t key2(uint32_t a, uint32_t b) { return ((uint64_t)a << 32) | b; }
// slot hash of an (a, b) id pair: two 32-bit multiplies + one finalizer (the tables are small and 2x over-provisioned)
__device__ __forceinline__ uint32_t hash_pair(uint32_t a, uint32_t b) { return mix32(a * 0x9E3779B1u ^ (b * 0x85EBCA77u + 0x165667B1u)); }
#define KR_EMPTY64 0xFFFFFFFFFFFFFFFFull
#define KR_EMPTY32 0xFFFFFFFFu

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may be scheduled while
// its stream predecessor is still running; it must not touch the predecessor's outputs before pdl_wait().  Both are no-ops
// for ordinary launches.  pdl_trigger() lets the NEXT kernel in the chain be scheduled early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ uint32_t pp_node_type(uint32_t f) { return (f >> KR_PP_NODE_TYPE_SHIFT) & 3u; }
__device__ __forceinline__ uint32_t pp_phase(uint32_t f) { return (f >> KR_PP_PHASE_SHIFT) & 7u; }
__device__ __forceinline__ uint32_t pp_ready(uint32_t f) { return (f >> KR_PP_READY_SHIFT) & 3u; }

// shouldDeletePod (raycluster_controller.go:1181-1231)
__device__ __forceinline__ bool should_delete(uint32_t f) {
  uint32_t ph = pp_phase(f);
  return ph == KR_PHASE_FAILED || ph == KR_PHASE_SUCCEEDED ||
         (ph == KR_PHASE_RUNNING && (f & KR_PP_RAY_TERMINATED) && (f & KR_PP_RESTART_NEVER));
}

// utils.GetWorkerGroupDesiredReplicas (utils/util.go:386-404); int32 multiply wraps like Go's
__device__ __forceinline__ int32_t desired_replicas(int32_t replicas, int32_t mn, int32_t mx, int32_t hosts, uint32_t gf) {
  int32_t minr = (gf & KR_GF_MIN_NIL) ? 0 : mn;
  int32_t maxr = (gf & KR_GF_MAX_NIL) ? INT32_MAX : mx;
  if (gf & KR_GF_SUSPEND) return 0;
  int32_t w;
  if ((gf & KR_GF_REPLICAS_NIL) || replicas < minr) w = minr;
  else if (replicas > maxr) w = maxr;
  else w = replicas;
  return (int32_t)((uint32_t)w * (uint32_t)hosts);
}

__device__ __forceinline__ bool cl_lookup(const ScratchDev &sc, uint32_t ns, uint32_t name, uint32_t &out) {
  if (name == 0) return false;
  uint32_t i = hash_pair(ns, name) & sc.cl_mask;
  while (true) {
    uint4 sl = __ldg(&sc.cl_slots[i]);  // one 16-byte load: key and value together
    if (sl.x == name && sl.y == ns) { out = sl.z; return true; }
    if (sl.x == KR_EMPTY32 && sl.y == KR_EMPTY32) return false;
    i = (i + 1) & sc.cl_mask;
  }
}

// ------------------------------------------------------------------------------------------------ k_clear
// One launch for the per-pass clears (hash tables to 0xFF, workersToDelete resolutions to -1, totals and bucket counters
// to 0) instead of four memset nodes at the head of the graph.
struct ClearArgs { uint32_t *ptr[4]; uint32_t words[4]; uint32_t value[4]; };
__global__ void __launch_bounds__(256) k_clear(ClearArgs a) {
  KR_TL(9);
  const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    uint32_t *p = a.ptr[r];
    const uint32_t v = a.value[r], nw = a.words[r];
    uint4 *p4 = reinterpret_cast<uint4 *>(p);  // every region starts 256-byte aligned
    for (uint32_t i = t0; i < nw / 4; i += stride) p4[i] = make_uint4(v, v, v, v);
    for (uint32_t i = (nw & ~3u) + t0; i < nw; i += stride) p[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ k_build_tables
// One thread per cluster / workersToDelete entry / head-aux row.  Tables were memset to 0xFF.

__global__ void __launch_bounds__(256) k_build_tables(SnapDev s, ScratchDev sc, ResDev r, Sizes n) {
  KR_TL(0);
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n.n_clusters) {
    uint32_t ns = s.c_ns_id[t], name = s.c_name_id[t];
    uint32_t i = hash_pair(ns, name) & sc.cl_mask;
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(sc.cl_slots);  // [2*i] = key (x = name, y = ns), [2*i+1] low word = idx
    const unsigned long long kk = ((unsigned long long)ns << 32) | name;
    while (true) {
      unsigned long long prev = atomicCAS(&slots[2 * (size_t)i], KR_EMPTY64, kk);
      if (prev == KR_EMPTY64 || prev == kk) { atomicMin(reinterpret_cast<uint32_t *>(&slots[2 * (size_t)i + 1]), t); break; }  // duplicate (ns,name): lowest index wins
      i = (i + 1) & sc.cl_mask;
    }
    uint32_t g0 = s.c_group_off[t], G = s.c_group_cnt[t], mh = 0;
    for (uint32_t gi = 0; gi < G; gi++) mh |= (s.g_num_hosts[g0 + gi] > 1) ? 1u : 0u;
    sc.cl_rec[t] = make_uint4(g0, G, G ? s.g_name_id[g0] : 0u, mh);  // .w bit 0: some worker group has numOfHosts > 1
    return;
  }
  t -= n.n_clusters;
  if (t < n.n_groups) {
    // every workersToDelete name of this group: Delete(ns of the cluster, name) (raycluster_controller.go:817-822)
    uint32_t c = s.g_cluster_idx[t];
    uint32_t ns = s.c_ns_id[c];
    uint32_t off = s.g_wtd_off[t], cnt = s.g_wtd_cnt[t];
    for (uint32_t w = 0; w < cnt; w++) {
      uint32_t e = off + w;
      uint64_t k = key2(ns, s.w_name_id[e]);
      uint32_t i = hash_pair(ns, s.w_name_id[e]) & sc.wt_mask;
      while (true) {
        unsigned long long prev = atomicCAS((unsigned long long *)&sc.wt_keys[i], KR_EMPTY64, k);
        if (prev == KR_EMPTY64 || prev == k) {
          // push e on the slot's chain
          uint32_t old = atomicExch(&sc.wt_head[i], e);
          sc.wt_next[e] = old;  // KR_EMPTY32 terminates (wt_head memset to 0xFF)
          break;
        }
        i = (i + 1) & sc.wt_mask;
      }
    }
    return;
  }
  t -= n.n_groups;
  if (t < n.n_heads) {
    uint32_t p = s.h_pod_idx[t];
    uint32_t i = mix32(p) & sc.aux_mask;
    while (true) {
      uint32_t prev = atomicCAS(&sc.aux_keys[i], KR_EMPTY32, p);
      if (prev == KR_EMPTY32 || prev == p) { atomicMin(&sc.aux_vals[i], t); break; }
      i = (i + 1) & sc.aux_mask;
    }
  }
}

__device__ __forceinline__ int32_t aux_lookup(const ScratchDev &sc, uint32_t p) {
  uint32_t i = mix32(p) & sc.aux_mask;
  while (true) {
    uint32_t k = sc.aux_keys[i];
    if (k == p) return (int32_t)sc.aux_vals[i];
    if (k == KR_EMPTY32) return -1;
    i = (i + 1) & sc.aux_mask;
  }
}

// ------------------------------------------------------------------------------------------------ k_match
// The selector match (common/association.go:83-130): pod -> RayCluster by (namespace, ray.io/cluster), then
// ray.io/group against the cluster's worker groups.  Streams 7 coalesced columns (28 B/pod), writes one 16-byte row
// + 4-byte sort key per pod, and the pass-0 digit histogram of its tile.

// kFast: the count/place/sort-in-warp pipeline (per-cluster arrival rank by a returning atomic, no radix histogram).
// The loop is phased — all column loads, then all table probes, then all record loads — so that each thread keeps
// 8 independent memory requests in flight per phase instead of walking one pod's dependent chain at a time.
template <bool kFast, int kItems>
__global__ void __launch_bounds__(kSortThreads) k_match(SnapDev s, ScratchDev sc, ResDev r, Sizes n, int has_wtd) {
  KR_TL(1);
  __shared__ uint32_t s_hist[kRadix];
  pdl_wait(); pdl_trigger();
  const uint32_t tile = blockIdx.x, ntiles = gridDim.x;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!kFast) { s_hist[threadIdx.x] = 0; __syncthreads(); }
  const uint32_t base = tile * (kSortThreads * kItems) + warp * (32 * kItems) + lane;
  uint32_t ns[kItems], cn[kItems], gn[kItems], nm[kItems], pk[kItems], rn[kItems], ri[kItems];
  // phase A: 7 coalesced column loads per pod
#pragma unroll
  for (int it = 0; it < kItems; it++) {
    uint32_t p = base + it * 32;
    bool v = p < n.n_pods;
    ns[it] = v ? __ldg(&s.p_ns_id[p]) : 0u; cn[it] = v ? __ldg(&s.p_cluster_name_id[p]) : 0u;
    gn[it] = v ? __ldg(&s.p_group_name_id[p]) : 0u; nm[it] = v ? __ldg(&s.p_name_id[p]) : 0u;
    pk[it] = v ? __ldg(&s.p_packed[p]) : 0u; ri[it] = v ? (uint32_t)__ldg(&s.p_replica_index[p]) : 0u;
    rn[it] = v ? __ldg(&s.p_replica_name_id[p]) : 0u;
  }
  // phase B: hash-join probe (namespace, ray.io/cluster) -> cluster idx; first slot of every pod in flight together
  uint32_t c[kItems], pi[kItems];
  uint4 sl[kItems];
#pragma unroll
  for (int it = 0; it < kItems; it++) {
    pi[it] = hash_pair(ns[it], cn[it]) & sc.cl_mask;
    sl[it] = __ldg(&sc.cl_slots[pi[it]]);
  }
#pragma unroll
  for (int it = 0; it < kItems; it++) {
    c[it] = n.n_clusters;
    if (cn[it] != 0) {
      uint4 q = sl[it];
      uint32_t i = pi[it];
      while (true) {
        if (q.x == cn[it] && q.y == ns[it]) { c[it] = q.z; break; }
        if (q.x == KR_EMPTY32 && q.y == KR_EMPTY32) break;
        i = (i + 1) & sc.cl_mask;
        q = __ldg(&sc.cl_slots[i]);
      }
    }
  }
  // phase C: the cluster's group record
  uint4 rec[kItems];
#pragma unroll
  for (int it = 0; it < kItems; it++) rec[it] = (c[it] < n.n_clusters) ? __ldg(&sc.cl_rec[c[it]]) : make_uint4(0, 0, 0, 0);
  // phase C': first probe of the (tiny, cache-resident) workersToDelete-name table for every pod, and the bucket ranks
  uint32_t wi[kItems];
  uint64_t wk[kItems];
  if (has_wtd) {
#pragma unroll
    for (int it = 0; it < kItems; it++) { wi[it] = hash_pair(ns[it], nm[it]) & sc.wt_mask; wk[it] = __ldg(&sc.wt_keys[wi[it]]); }
  }
  uint32_t rank[kItems], orank[kItems], woff = 0;
  if (kFast) {
#pragma unroll
    for (int it = 0; it < kItems; it++)  // arrival rank inside the cluster's bucket; 8 atomics in flight
      rank[it] = (base + it * 32 < n.n_pods && c[it] < n.n_clusters) ? atomicAdd(&sc.ccount[c[it]], 1u) : 0u;
    // Orphans (no RayCluster) need no decision, only List order, and their bucket has no size bound: give them a STABLE rank
    // right here — thread order inside a tile is pod order (warp, then item, then lane) — plus the tile's orphan count, which
    // k_scan_counts turns into a per-tile prefix.  No atomics, no sort.
    __shared__ uint32_t s_worph[kSortThreads / 32];
    uint32_t wcount = 0;
    const uint32_t ltm = lanemask_lt();
#pragma unroll
    for (int it = 0; it < kItems; it++) {  // (kept apart from rank[]: nothing here may wait for the atomics above)
      bool orph = (base + it * 32 < n.n_pods) && c[it] == n.n_clusters;
      uint32_t bal = __ballot_sync(0xFFFFFFFFu, orph);
      orank[it] = wcount + __popc(bal & ltm);
      wcount += __popc(bal);
    }
    if (lane == 0) s_worph[warp] = wcount;
    __syncthreads();
    uint32_t ttot = 0;
#pragma unroll
    for (int w2 = 0; w2 < kSortThreads / 32; w2++) { uint32_t v = s_worph[w2]; if (w2 < (int)warp) woff += v; ttot += v; }
    if (threadIdx.x == 0) { sc.tile_orph[tile] = ttot; if (ttot) atomicAdd(&sc.ccount[n.n_clusters], ttot); }
  }
  // phase D: ray.io/group against the cluster's worker groups, workersToDelete-name intersection, outputs
#pragma unroll
  for (int it = 0; it < kItems; it++) {
    uint32_t p = base + it * 32;
    if (p >= n.n_pods) continue;
    uint32_t slot = KR_ROW_NO_GROUP, g0 = rec[it].x;
    if (c[it] < n.n_clusters && gn[it] != 0) {
      if (rec[it].y && rec[it].z == gn[it]) slot = 0;
      else
        for (uint32_t gi = 1; gi < rec[it].y; gi++)
          if (__ldg(&s.g_name_id[g0 + gi]) == gn[it]) { slot = gi; break; }  // group names are unique (pkg/webhooks/v1/raycluster_webhook.go:74)
    }
    uint32_t flags = pk[it] & (0x7FFu | KR_PP_TOMBSTONE);  // bit 11 of the row word is KR_ROW_WTD_OWN
    if (has_wtd) {
      const uint64_t k = key2(ns[it], nm[it]);
      uint32_t i = wi[it];
      uint64_t kk = wk[it];
      while (kk != KR_EMPTY64) {
        if (kk == k) {
          for (uint32_t e = sc.wt_head[i]; e != KR_EMPTY32; e = sc.wt_next[e]) {
            atomicMin(&r.wtd_pod_idx[e], p);
            if (slot != KR_ROW_NO_GROUP) {  // is e one of this pod's own group's names?
              uint32_t g = g0 + slot;
              uint32_t off = __ldg(&s.g_wtd_off[g]);
              if (e >= off && e < off + __ldg(&s.g_wtd_cnt[g])) flags |= KR_ROW_WTD_OWN;
            }
          }
          break;
        }
        i = (i + 1) & sc.wt_mask;
        kk = __ldg(&sc.wt_keys[i]);
      }
    }
    sc.rows[p] = make_uint4(nm[it], rn[it], ri[it], (slot << 16) | flags);
    sc.keys[0][p] = c[it];
    if (kFast) sc.keys[1][p] = (c[it] == n.n_clusters) ? orank[it] + woff : rank[it];
    else atomicAdd(&s_hist[c[it] & (kRadix - 1)], 1u);
  }
  if (!kFast) {
    __syncthreads();
    sc.hist[threadIdx.x * ntiles + tile] = s_hist[threadIdx.x];
  }
}

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
    if (big) atomicOr(&totals[3], KR_TOTALS_BIG_BUCKET);
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

// ------------------------------------------------------------------------------------------------ k_decide

#define KR_SMEM_GROUPS 32  // clusters with more worker groups than this spill their accumulators to global scratch
static constexpr int kDecideWarps = 4;

// group processing modes (internal)
enum { GM_UNPROCESSED = 0, GM_SKIP = 1, GM_SUSPENDED = 2, GM_UNHEALTHY = 3, GM_NORMAL = 4, GM_MULTIHOST = 5 };

// first index i in [0,n) with a[i] >= v; warp-cooperative 32-ary search, result uniform across the warp
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t *__restrict__ a, uint32_t n, uint32_t v, uint32_t lane) {
  uint32_t lo = 0, hi = n;  // answer in [lo, hi]
  while (hi - lo > 32) {
    uint32_t step = (hi - lo + 31) / 32;  // probe points lo + (l+1)*step - 1
    uint32_t idx = lo + (lane + 1) * step - 1;
    bool ge = (idx >= hi) ? true : (__ldg(&a[idx]) >= v);
    uint32_t b = __ballot_sync(0xFFFFFFFFu, ge);
    if (b == 0) return hi;  // every probe (the last one sits at hi-1 at the earliest) is < v: the answer is hi itself
    uint32_t first = __ffs(b) - 1;
    uint32_t nlo = lo + first * step, nhi = min(hi, lo + (first + 1) * step - 1);
    lo = nlo; hi = nhi;
  }
  uint32_t idx = lo + lane;
  bool ge = (idx >= hi) ? true : (__ldg(&a[idx]) >= v);
  uint32_t b = __ballot_sync(0xFFFFFFFFu, ge);
  return b ? lo + (__ffs(b) - 1) : hi;
}

struct DecideArgs {
  SnapDev s; ScratchDev sc; ResDev r; Sizes n; kr_flags f;
  const uint32_t *sorted_keys;  // radix pipeline: cluster idx per sorted position
  const uint32_t *unsorted;     // fast pipeline: pods bucketed by cluster in arrival order (sorted per bucket here)
  int fast;
  int phase;                    // 0: everything that does not need the hash; 1: only clusters deferred by phase 0
};

// calculateStatus (raycluster_controller.go:1552-1719) + InconsistentRayClusterStatus (utils/consistency.go:16-34).
// Scalar code, executed by lane 0 only.
__device__ void status_rollup(const DecideArgs &a, uint32_t c, kr_cluster_result &cr, uint32_t P, uint32_t n_heads, int32_t head_pod,
                              uint32_t head_name_id, int32_t ready, int32_t available, bool all_running) {
  const SnapDev &s = a.s;
  const uint32_t cf = s.c_flags[c];
  const bool gate = a.f.gate_status_conditions != 0;
  const bool reconcile_err = cr.err_kind != KR_ERR_NONE;
  const uint8_t ek = s.c_ext_err_kind[c];
  uint8_t cst[KR_NUM_CONDS], cvr[KR_NUM_CONDS];
#pragma unroll
  for (int k = 0; k < KR_NUM_CONDS; k++) { cst[k] = s.c_old_cond_status[5 * (size_t)c + k]; cvr[k] = s.c_old_cond_variant[5 * (size_t)c + k]; }
  uint32_t hpr_reason = s.c_old_cond_reason_id[c], hpr_msg = s.c_old_cond_msg_id[2 * (size_t)c], rf_msg = s.c_old_cond_msg_id[2 * (size_t)c + 1];
  if (gate) {  // :1563-1577
    if (reconcile_err) {
      if (ek >= KR_EXT_ERR_FAILED_DELETE_ALL_PODS && ek <= KR_EXT_ERR_FAILED_CREATE_WORKER_POD) {
        cst[KR_COND_REPLICA_FAILURE] = KR_COND_TRUE; cvr[KR_COND_REPLICA_FAILURE] = ek; rf_msg = s.c_ext_err_msg_id[c];
      }
    } else {
      cst[KR_COND_REPLICA_FAILURE] = KR_COND_ABSENT; cvr[KR_COND_REPLICA_FAILURE] = KR_CV_NONE; rf_msg = 0;
    }
  }
  int32_t desired = 0, minr = 0; long long maxr = 0;  // utils/util.go:407-442
  const uint32_t G = s.c_group_cnt[c], g0 = s.c_group_off[c];
  for (uint32_t gi = 0; gi < G; gi++) {
    uint32_t g = g0 + gi, gf = s.g_flags[g];
    int32_t hosts = s.g_num_hosts[g];
    desired = (int32_t)((uint32_t)desired + (uint32_t)desired_replicas(s.g_replicas[g], s.g_min[g], s.g_max[g], hosts, gf));
    if (gf & KR_GF_SUSPEND) continue;
    int32_t mn = (gf & KR_GF_MIN_NIL) ? 0 : s.g_min[g];
    int32_t mx = (gf & KR_GF_MAX_NIL) ? INT32_MAX : s.g_max[g];
    minr = (int32_t)((uint32_t)minr + (uint32_t)mn * (uint32_t)hosts);
    maxr += (long long)mx * (long long)hosts;
  }
  int32_t maxc = maxr > INT32_MAX ? INT32_MAX : (maxr < INT32_MIN ? INT32_MIN : (int32_t)maxr);  // utils/util.go:284-292

  cr.n_pods = (int32_t)P; cr.n_heads = (int32_t)n_heads; cr.head_pod_idx = head_pod;
  uint8_t serr = KR_SERR_NONE;  // :1608-1611, :1785-1806, :1721-1745
  if (n_heads > 1) serr = KR_SERR_MULTIPLE_HEADS;
  else if (s.c_svc_count[c] == 0) serr = KR_SERR_NO_HEAD_SERVICE;
  else if (s.c_svc_count[c] > 1) serr = KR_SERR_MULTIPLE_HEAD_SERVICES;
  else if (s.c_svc_ip_kind[c] == KR_SVCIP_EMPTY) serr = KR_SERR_EMPTY_SERVICE_IP;
  cr.status_err = serr;
  if (serr != KR_SERR_NONE) return;

  const uint8_t old_state = s.c_old_state[c];
  uint8_t new_state = old_state;
  bool reason_cleared = false;
  if (!reconcile_err && (long long)P == (long long)desired + 1 && all_running) { new_state = KR_STATE_READY; reason_cleared = true; }  // :1599-1604

  uint32_t head_pod_ip = 0, head_pod_name = 0;
  int32_t aux = -1;
  if (n_heads == 1) {
    aux = aux_lookup(a.sc, (uint32_t)head_pod);
    head_pod_ip = aux >= 0 ? s.h_pod_ip_id[aux] : 0;
    head_pod_name = head_name_id;
  }
  if (gate) {
    if (n_heads == 0) {  // :1613-1619
      cst[KR_COND_HEAD_POD_READY] = KR_COND_FALSE; cvr[KR_COND_HEAD_POD_READY] = KR_CV_HEAD_NOT_FOUND;
      hpr_reason = a.f.id_head_not_found_reason; hpr_msg = a.f.id_head_not_found_msg;
    } else {             // :1621-1622
      cst[KR_COND_HEAD_POD_READY] = aux >= 0 ? s.h_ready_status[aux] : (uint8_t)KR_COND_FALSE;
      cvr[KR_COND_HEAD_POD_READY] = KR_CV_HEAD_FROM_POD;
      hpr_reason = aux >= 0 ? s.h_ready_reason_id[aux] : 0; hpr_msg = aux >= 0 ? s.h_ready_msg_id[aux] : 0;
    }
    const uint8_t ss = s.c_suspend_status[c];
    if (cst[KR_COND_PROVISIONED] != KR_COND_TRUE && ss != KR_SUSPEND_SUSPENDED) {  // :1625-1644
      if (all_running) { cst[KR_COND_PROVISIONED] = KR_COND_TRUE; cvr[KR_COND_PROVISIONED] = KR_CV_PROV_ALL_READY; }
      else { cst[KR_COND_PROVISIONED] = KR_COND_FALSE; cvr[KR_COND_PROVISIONED] = KR_CV_PROV_PROVISIONING; }
    }
    if (ss == KR_SUSPEND_SUSPENDING) {  // :1646-1693
      if (P == 0) {
        cst[KR_COND_PROVISIONED] = KR_COND_FALSE; cvr[KR_COND_PROVISIONED] = KR_CV_PROV_SUSPENDED;
        cst[KR_COND_SUSPENDING] = KR_COND_FALSE; cvr[KR_COND_SUSPENDING] = KR_CV_CANONICAL;
        cst[KR_COND_SUSPENDED] = KR_COND_TRUE; cvr[KR_COND_SUSPENDED] = KR_CV_CANONICAL;
      }
    } else if (ss == KR_SUSPEND_SUSPENDED) {
      if (cf & KR_CF_SUSPEND_SET_FALSE) { cst[KR_COND_SUSPENDED] = KR_COND_FALSE; cvr[KR_COND_SUSPENDED] = KR_CV_CANONICAL; }
    } else {
      cst[KR_COND_SUSPENDED] = KR_COND_FALSE; cvr[KR_COND_SUSPENDED] = KR_CV_CANONICAL;
      cst[KR_COND_SUSPENDING] = (cf & KR_CF_SUSPEND) ? KR_COND_TRUE : KR_COND_FALSE; cvr[KR_COND_SUSPENDING] = KR_CV_CANONICAL;
    }
  }
  if ((cf & KR_CF_SUSPEND) && P == 0) new_state = KR_STATE_SUSPENDED;  // :1696-1698

  uint32_t svc_ip = s.c_svc_ip_id[c];
  if (s.c_svc_ip_kind[c] == KR_SVCIP_NONE) svc_ip = (n_heads == 1) ? head_pod_ip : 0;  // :1732-1742
  uint32_t head_ids[4] = {head_pod_ip, svc_ip, head_pod_name, s.c_svc_name_id[c]};

  cr.new_state = new_state;
  cr.state_changed = new_state != old_state;
  cr.status_flags = (reason_cleared ? KR_SF_READY_BRANCH : 0u) | (all_running ? KR_SF_ALL_PODS_RUNNING : 0u);
  cr.counts[0] = ready; cr.counts[1] = available; cr.counts[2] = desired; cr.counts[3] = minr; cr.counts[4] = maxc;
#pragma unroll
  for (int k = 0; k < KR_NUM_CONDS; k++) { cr.cond_status[k] = cst[k]; cr.cond_variant[k] = cvr[k]; }
  cr.head_ready_reason_id = hpr_reason; cr.head_ready_msg_id = hpr_msg;
#pragma unroll
  for (int k = 0; k < 4; k++) cr.head_ids[k] = head_ids[k];

  bool inc = new_state != old_state;  // utils/consistency.go:16-34
  if (reason_cleared && (cf & KR_CF_OLD_REASON_NONEMPTY)) inc = true;
#pragma unroll
  for (int k = 0; k < 5; k++) if (s.c_old_counts[5 * (size_t)c + k] != cr.counts[k]) inc = true;
  if (cf & KR_CF_ENDPOINTS_CHANGED) inc = true;
#pragma unroll
  for (int k = 0; k < 4; k++) if (s.c_old_head_ids[4 * (size_t)c + k] != head_ids[k]) inc = true;
#pragma unroll
  for (int k = 0; k < KR_NUM_CONDS; k++) {
    uint8_t os = s.c_old_cond_status[5 * (size_t)c + k], ov = s.c_old_cond_variant[5 * (size_t)c + k];
    if (os != cst[k]) { inc = true; continue; }
    if (cst[k] == KR_COND_ABSENT) continue;
    if (k == KR_COND_HEAD_POD_READY) {
      if (s.c_old_cond_reason_id[c] != hpr_reason || s.c_old_cond_msg_id[2 * (size_t)c] != hpr_msg) inc = true;
    } else if (k == KR_COND_REPLICA_FAILURE) {
      if (ov != cvr[k] || s.c_old_cond_msg_id[2 * (size_t)c + 1] != rf_msg) inc = true;
    } else if (ov != cvr[k]) inc = true;
  }
  cr.needs_status_write = inc ? 1 : 0;
}

// reconcileMultiHostWorkerGroup (raycluster_controller.go:963-1125) for one worker group, by one warp.
// Replicas (pods sharing ray.io/worker-group-replica-name) are identified by the list position of their first pod, so
// "first appearance in list order" — the deterministic stand-in for the reference's Go-map iteration (SURVEY Appendix
// A.5) — is simply ascending position.  Cost O(pods x replicas / 32); every sweep is a coalesced 4-byte column read.
// Returns the KR_ERR_* kind (0 = nil).
#define KR_MH_NONE 0xFFFFFFFFu        // not a member of this group
#define KR_MH_UNASSIGNED 0xFFFFFFFEu  // member with a replica-name label, replica not identified yet
#define KR_MH_NOREP 0xFFFFFFFDu       // member without the label: belongs to no replica
#define KR_MHF_DELETED 1u
#define KR_MHF_WTD 2u
#define KR_MHF_SCALE 4u
__device__ int decide_multihost(const DecideArgs &a, uint32_t slot, uint32_t seg0, uint32_t seg1, int32_t expected, int32_t H,
                                bool delete_allowed, uint32_t wtd_cnt, kr_group_result &gr, int32_t &err_arg, uint32_t lane) {
  uint32_t *rep = a.sc.mh_rep, *name = a.sc.mh_name, *meta = a.sc.mh_meta, *cnt = a.sc.mh_cnt, *flg = a.sc.mh_flg;
  uint8_t *act = a.sc.mh_act, *headv = a.sc.mh_head;
  const uint32_t lt = lanemask_lt();
  // 0. per-position columns of this group (valid only while this group is being decided)
  for (uint32_t b = seg0; b < seg1; b += 32) {
    uint32_t i = b + lane;
    if (i < seg1) {
      uint4 row = a.sc.rows[a.r.sorted_pod_idx[i]];
      bool member = (row.w >> 16) == slot;
      rep[i] = member ? (row.y ? KR_MH_UNASSIGNED : KR_MH_NOREP) : KR_MH_NONE;
      if (member) { name[i] = row.y; meta[i] = row.w & 0xFFFFu; act[i] = KR_ACT_KEEP; headv[i] = 0; }
    }
  }
  __syncwarp();
  // 1. replicaMap (:967-972): peel replicas off in order of first appearance
  uint32_t cursor = seg0, first_incomplete = KR_MH_NONE, m_empty = KR_MH_NONE;
  int32_t incomplete_cnt = 0;
  while (true) {
    uint32_t m = KR_MH_NONE;
    for (uint32_t b = cursor; b < seg1; b += 32) {
      uint32_t i = b + lane;
      uint32_t bal = __ballot_sync(0xFFFFFFFFu, i < seg1 && rep[i] == KR_MH_UNASSIGNED);
      if (bal) { m = b + (__ffs(bal) - 1); break; }
    }
    if (m == KR_MH_NONE) break;
    const uint32_t nm = name[m];
    uint32_t count = 0;
    for (uint32_t b = m; b < seg1; b += 32) {
      uint32_t i = b + lane;
      bool hit = i < seg1 && rep[i] == KR_MH_UNASSIGNED && name[i] == nm;
      if (hit) rep[i] = m;
      count += __popc(__ballot_sync(0xFFFFFFFFu, hit));
    }
    if (lane == 0) { cnt[m] = count; flg[m] = 0; }
    if (nm == KR_ID_EMPTY_STRING) m_empty = m;
    if ((int64_t)count < (int64_t)H && first_incomplete == KR_MH_NONE) { first_incomplete = m; incomplete_cnt = (int32_t)count; }
    cursor = m + 1;
    __syncwarp();
  }
  // 2. incomplete replica groups (:975-984)
  if (first_incomplete != KR_MH_NONE) {
    for (uint32_t b = seg0; b < seg1; b += 32) { uint32_t i = b + lane; if (i < seg1 && rep[i] == first_incomplete) act[i] = KR_ACT_DELETE_MH_INCOMPLETE; }
    gr.flags |= KR_GR_ABORTED; err_arg = incomplete_cnt;
    __syncwarp();
    return KR_ERR_MH_INCOMPLETE;
  }
  // 3. unhealthy replica groups (:987-1007): a pod marks its replica; unlabelled pods resolve to the "" replica if one exists
  for (uint32_t b = seg0; b < seg1; b += 32) {
    uint32_t i = b + lane;
    if (i < seg1 && rep[i] != KR_MH_NONE && should_delete(meta[i])) {
      uint32_t r = rep[i] == KR_MH_NOREP ? m_empty : rep[i];
      if (r != KR_MH_NONE) atomicOr(&flg[r], KR_MHF_DELETED);
    }
  }
  __syncwarp();
  int32_t n_unh = 0;
  for (uint32_t b = seg0; b < seg1; b += 32) {
    uint32_t i = b + lane;
    bool hit = i < seg1 && rep[i] < KR_MH_NOREP && (flg[rep[i]] & KR_MHF_DELETED);
    if (hit) act[i] = KR_ACT_DELETE_MH_UNHEALTHY;
    n_unh += __popc(__ballot_sync(0xFFFFFFFFu, hit));
  }
  gr.n_unhealthy = n_unh;
  // 4. explicit deletions from the autoscaler (:1010-1038)
  if (wtd_cnt > 0) {
    for (uint32_t b = seg0; b < seg1; b += 32) {
      uint32_t i = b + lane;
      if (i < seg1 && rep[i] != KR_MH_NONE && (meta[i] & KR_ROW_WTD_OWN)) {
        uint32_t r = rep[i] == KR_MH_NOREP ? m_empty : rep[i];
        if (r != KR_MH_NONE) atomicOr(&flg[r], KR_MHF_WTD);
      }
    }
    __syncwarp();
    int32_t n_del = 0;
    for (uint32_t b = seg0; b < seg1; b += 32) {
      uint32_t i = b + lane;
      bool hit = i < seg1 && rep[i] < KR_MH_NOREP && (flg[rep[i]] & KR_MHF_WTD);
      if (hit && act[i] == KR_ACT_KEEP) act[i] = KR_ACT_DELETE_MH_WTD;
      n_del += __popc(__ballot_sync(0xFFFFFFFFu, hit));
    }
    gr.flags |= KR_GR_WTD_EXECUTED;
    if (n_del > 0) { gr.flags |= KR_GR_ABORTED; err_arg = n_del; __syncwarp(); return KR_ERR_MH_WTD; }
  }
  // 5. diff by replica (:1042-1064)
  int32_t running = 0;
  for (uint32_t b = seg0; b < seg1; b += 32) {
    uint32_t i = b + lane;
    bool ok = i < seg1 && rep[i] == i && !(flg[i] & KR_MHF_DELETED);  // first pod of a healthy, complete replica
    if (ok) headv[i] = 1;
    running += __popc(__ballot_sync(0xFFFFFFFFu, ok));
  }
  gr.n_running = running;
  if (expected % H != 0) { gr.flags |= KR_GR_ABORTED; err_arg = expected; __syncwarp(); return KR_ERR_MH_NOT_MULTIPLE; }
  const int32_t to_create = expected / H - running;
  gr.diff = to_create;
  if (to_create > 0) gr.n_create = (uint32_t)to_create;  // one replica index per new replica group; k_create_fill allocates them
  else if (to_create < 0) {
    if (delete_allowed) {  // :1104-1118 — the first -to_create valid replicas in first-appearance order
      int32_t seen = 0;
      const int32_t remove = -to_create;
      for (uint32_t b = seg0; b < seg1 && seen < remove; b += 32) {
        uint32_t i = b + lane;
        bool ok = i < seg1 && rep[i] == i && !(flg[i] & KR_MHF_DELETED);
        uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (ok && seen + (int32_t)__popc(bal & lt) < remove) flg[i] |= KR_MHF_SCALE;
        seen += __popc(bal);
      }
      __syncwarp();
      for (uint32_t b = seg0; b < seg1; b += 32) {
        uint32_t i = b + lane;
        if (i < seg1 && rep[i] < KR_MH_NOREP && (flg[rep[i]] & KR_MHF_SCALE)) act[i] = KR_ACT_DELETE_MH_SCALE_DOWN;
      }
    } else gr.flags |= KR_GR_RANDOM_DELETE_OFF;
  }
  __syncwarp();
  return KR_ERR_NONE;
}

#define LDG(x) __ldg(&(x))

// Bitonic sort of 32*K values held K per lane, STRIPED (element g = k*32 + lane), ascending.  Striped order is what the
// chunked scans of k_decide want: register k of lane l is list position k*32+l, and loads/stores are fully coalesced.
template <int K>
__device__ __forceinline__ void warp_bitonic_sort_striped(uint32_t (&v)[K], uint32_t lane) {
#pragma unroll
  for (int size = 2; size <= 32 * K; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride < 32) {
        const bool lower = (lane & stride) == 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
          uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v[k], stride);
          bool asc = ((k * 32 + lane) & size) == 0;
          v[k] = (asc == lower) ? min(v[k], o) : max(v[k], o);
        }
      } else {
        const int ks = stride / 32;
#pragma unroll
        for (int k = 0; k < K; k++) {
          if ((k & ks) == 0) {
            const bool asc = ((k * 32) & size) == 0;  // size >= 64 here: the lane bits do not reach it
            uint32_t lo = min(v[k], v[k + ks]), hi = max(v[k], v[k + ks]);
            v[k] = asc ? lo : hi; v[k + ks] = asc ? hi : lo;
          }
        }
      }
    }
  }
}

// reconcilePods (raycluster_controller.go:619-935) + calculateStatus for one RayCluster, by one warp.
// K > 0: the cluster's bucket (<= 32*K pods) is sorted and kept in registers — pod index pidx[k] and row word pw[k] of list
// position k*32+lane — so the two scans below touch no memory.  K == 0: positions are read from sorted_pod_idx / rows
// (radix pipeline, buckets larger than 256 pods, and phase 1).
template <int K, bool kMH>
__device__ __forceinline__ void decide_cluster(const DecideArgs &a, const uint32_t c, const uint32_t seg0, const uint32_t seg1,
                                               uint32_t (&pidx)[K ? K : 1], uint32_t (&pw)[K ? K : 1],
                                               int32_t (&s_acc)[4][KR_SMEM_GROUPS], int32_t (&s_mode)[2][KR_SMEM_GROUPS], const uint32_t lane) {
  const SnapDev &s = a.s;
  const uint32_t lt = lanemask_lt();
  const uint32_t P = seg1 - seg0;
  const uint32_t nchunks = (P + 31) / 32;
  // cluster scalars: independent read-only loads, issued together
  const uint32_t cf = LDG(s.c_flags[c]);
  const uint32_t G = LDG(s.c_group_cnt[c]), g0 = LDG(s.c_group_off[c]);
  const uint8_t suspend_status = LDG(s.c_suspend_status[c]);
  const uint8_t ext_err = LDG(s.c_ext_err_kind[c]);
  const uint8_t old_prov = LDG(s.c_old_cond_status[5 * (size_t)c + KR_COND_PROVISIONED]);
  const bool gate = a.f.gate_status_conditions != 0;

  // accumulators: shared memory for the common case, global scratch for clusters with many groups
  int32_t *acc_list, *acc_unh, *acc_wtd, *acc_rank, *g_mode, *g_prefix;
  if (G <= KR_SMEM_GROUPS) {
    acc_list = s_acc[0]; acc_unh = s_acc[1]; acc_wtd = s_acc[2]; acc_rank = s_acc[3];
    g_mode = s_mode[0]; g_prefix = s_mode[1];
    if (lane < KR_SMEM_GROUPS) { acc_list[lane] = 0; acc_unh[lane] = 0; acc_wtd[lane] = 0; acc_rank[lane] = 0; g_mode[lane] = GM_UNPROCESSED; g_prefix[lane] = 0; }
  } else {
    const uint32_t Ng = a.n.n_groups;
    acc_list = a.sc.gacc + g0; acc_unh = a.sc.gacc + Ng + g0; acc_wtd = a.sc.gacc + 2 * (size_t)Ng + g0; acc_rank = a.sc.gacc + 3 * (size_t)Ng + g0;
    g_mode = nullptr; g_prefix = nullptr;  // modes recycle the n_list / n_unhealthy cells once they are consumed
    for (uint32_t gi = lane; gi < G; gi += 32) { acc_list[gi] = 0; acc_unh[gi] = 0; acc_wtd[gi] = 0; acc_rank[gi] = 0; }
  }
  __syncwarp();

  // ---------------- scan 1: counts over the cluster's pods (list order)
  int32_t ready = 0, available = 0, n_heads = 0;
  bool all_running = P > 0;  // CheckAllPodsRunning (utils/util.go:584-603)
  int32_t head_pod = -1;     // first head in list order
#pragma unroll
  for (int k = 0; k < (K ? K : 1 << 30); k++) {
    if ((uint32_t)k >= nchunks) break;
    const uint32_t i = seg0 + k * 32 + lane;
    const bool valid = i < seg1;
    uint32_t pod, w;
    if (K) { pod = pidx[K ? k : 0]; w = pw[K ? k : 0]; }
    else { pod = valid ? LDG(a.r.sorted_pod_idx[i]) : 0u; w = valid ? reinterpret_cast<const uint32_t *>(a.sc.rows)[4 * (size_t)pod + 3] : 0u; }
    const uint32_t fl = w & 0xFFFFu, slot = valid ? (w >> 16) : KR_ROW_NO_GROUP;
    const uint32_t nt = pp_node_type(fl), ph = pp_phase(fl), rd = pp_ready(fl);
    const bool w_run = valid && nt == KR_NT_WORKER && ph == KR_PHASE_RUNNING;
    available += __popc(__ballot_sync(0xFFFFFFFFu, w_run));
    ready += __popc(__ballot_sync(0xFFFFFFFFu, w_run && rd == KR_COND_TRUE));
    const bool not_ok = valid && (ph != KR_PHASE_RUNNING || rd == KR_COND_FALSE || rd == KR_COND_UNKNOWN);
    if (__any_sync(0xFFFFFFFFu, not_ok)) all_running = false;
    const uint32_t hb = __ballot_sync(0xFFFFFFFFu, valid && nt == KR_NT_HEAD);
    if (hb) { if (n_heads == 0) head_pod = (int32_t)__shfl_sync(0xFFFFFFFFu, pod, __ffs(hb) - 1); n_heads += __popc(hb); }
    // warp-ballot group-by on the group slot
    const uint32_t gkey = (slot < G) ? slot : KR_ROW_NO_GROUP;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, gkey);
    if (gkey != KR_ROW_NO_GROUP) {
      const uint32_t ub = __ballot_sync(peers, should_delete(fl));
      const uint32_t wb = __ballot_sync(peers, (fl & KR_ROW_WTD_OWN) != 0);
      if ((peers & lt) == 0) {  // leader of its group in this chunk
        acc_list[gkey] += __popc(peers);
        acc_unh[gkey] += __popc(ub & peers);
        acc_wtd[gkey] += __popc(wb & peers);
      }
    }
    __syncwarp();
  }

  uint32_t head_flags = 0, head_name = 0;
  if (n_heads > 0) { uint4 hrow = __ldg(&a.sc.rows[head_pod]); head_flags = hrow.w & 0xFFFFu; head_name = hrow.x; }

  // ---------------- scalar decisions (uniform across the warp)
  kr_cluster_result cr;
  {
    uint32_t *z = reinterpret_cast<uint32_t *>(&cr);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(cr) / 4); k++) z[k] = 0;
  }
  cr.head_pod_idx = -1; cr.stop_after_group = -1; cr.pod_start = seg0;
  uint8_t all_action = KR_ACT_KEEP;  // action applied to every pod of the cluster (delete-all paths)
  bool head_delete = false;
  bool run_groups = false;

  if (cf & KR_CF_SKIP) {
    cr.path = KR_PATH_SKIPPED;
  } else if (ext_err != KR_EXT_ERR_NONE) {
    cr.path = KR_PATH_SKIPPED;  // :308-314
    cr.err_kind = ext_err == KR_EXT_ERR_STATUS_ONLY_NIL ? KR_ERR_NONE : KR_ERR_EXTERNAL;
  } else if (suspend_status == KR_SUSPEND_SUSPENDING || (!gate && (cf & KR_CF_SUSPEND))) {
    cr.path = KR_PATH_SUSPENDING_DELETE_ALL; all_action = KR_ACT_DELETE_ALL_SUSPEND;  // :629-644
  } else if (gate && (suspend_status == KR_SUSPEND_SUSPENDED || (cf & KR_CF_SUSPEND))) {
    cr.path = KR_PATH_SUSPENDED_NOOP;  // :646-654
  } else {
    bool recreate = false;
    if ((cf & KR_CF_UPGRADE_RECREATE) && n_heads > 0) {  // shouldRecreatePodsForUpgrade :1132-1171
      int32_t aux = aux_lookup(a.sc, (uint32_t)head_pod);
      uint8_t ver = aux >= 0 ? s.h_version_state[aux] : (uint8_t)KR_VER_EMPTY;
      uint8_t ast = aux >= 0 ? s.h_annot_state[aux] : (uint8_t)KR_ANNOT_EMPTY;
      if (ver == KR_VER_DIFFERENT) cr.head_update_annotations = 1;
      else if (ast == KR_ANNOT_OTHER) recreate = true;
      else if (ast == KR_ANNOT_HASH32 && !a.f.skip_hash) {
        if (a.phase == 0) {  // the hash kernel runs concurrently on another stream: decide this cluster in phase 1
          if (lane == 0) a.sc.deferred_list[atomicAdd(&a.r.totals[4], 1u)] = c;
          return;
        }
        const uint8_t *ah = s.h_annot_hash + 32 * (size_t)aux;
        const char *hh = a.r.hash + 32 * (size_t)c;
        bool ne = ah[lane] != (uint8_t)hh[lane];
        recreate = __any_sync(0xFFFFFFFFu, ne);
      }
    }
    if (recreate) {
      cr.path = KR_PATH_RECREATE_DELETE_ALL; all_action = KR_ACT_DELETE_ALL_RECREATE;  // :657-670
    } else {
      cr.path = KR_PATH_NORMAL;
      // head (:673-748)
      if (!(cf & KR_CF_HEAD_EXPECT_OK)) { cr.head_action = KR_HEAD_EXPECT_PENDING; run_groups = true; }
      else if (n_heads == 1) {
        if (should_delete(head_flags)) { cr.head_action = KR_HEAD_DELETE; cr.err_kind = KR_ERR_HEAD_DELETED; head_delete = true; }
        else run_groups = true;
      } else if (n_heads == 0) {
        if (old_prov == KR_COND_TRUE && (cf & KR_CF_SKIP_HEAD_RESTART)) cr.head_action = KR_HEAD_SKIP_RESTART;
        else { cr.head_action = KR_HEAD_CREATE; run_groups = true; }
      } else {
        cr.head_action = KR_HEAD_MULTIPLE; cr.err_kind = KR_ERR_MULTIPLE_HEADS; cr.err_arg = n_heads;
      }
    }
  }

  // worker groups in spec order (:751-933): O(1) per group from the scan-1 counters
  if (run_groups) {
    const bool autoscaling = (cf & KR_CF_AUTOSCALING) != 0;
    cr.stop_after_group = (int32_t)G;
    for (uint32_t gi = 0; gi < G; gi++) {
      const uint32_t g = g0 + gi, gf = LDG(s.g_flags[g]);
      const int32_t hosts = LDG(s.g_num_hosts[g]), g_rep = LDG(s.g_replicas[g]), g_mn = LDG(s.g_min[g]), g_mx = LDG(s.g_max[g]);
      kr_group_result gr;
      gr.expected = 0; gr.n_list = 0; gr.n_unhealthy = 0; gr.n_running = 0; gr.diff = 0; gr.n_create = 0; gr.create_off = 0;
      gr.flags = KR_GR_PROCESSED;
      int32_t mode = GM_SKIP, prefix = 0;
      bool abort_here = false;
      if (!(gf & KR_GF_EXPECT_OK)) {
        gr.flags |= KR_GR_EXPECT_PENDING;
      } else {
        const int32_t expected = desired_replicas(g_rep, g_mn, g_mx, hosts, gf);
        const int32_t n_list = acc_list[gi], n_unh = acc_unh[gi], n_wtd = acc_wtd[gi];
        gr.expected = expected; gr.n_list = n_list;
        if (gf & KR_GF_SUSPEND) { gr.flags |= KR_GR_SUSPENDED; mode = GM_SUSPENDED; }
        else if (kMH && hosts > 1 && a.f.gate_multihost_indexing) {  // :777-784 (clusters with such groups never reach the <.., false> instantiations)
          gr.flags |= KR_GR_MULTIHOST; mode = GM_MULTIHOST;
          int32_t earg = 0;
          int ek = decide_multihost(a, gi, seg0, seg1, expected, hosts, !autoscaling || a.f.env_random_pod_delete, LDG(s.g_wtd_cnt[g]), gr, earg, lane);
          if (ek != KR_ERR_NONE) { cr.err_kind = (uint8_t)ek; cr.err_arg = earg; abort_here = true; }
        }
        else if (n_unh > 0) {  // :786-812
          gr.n_unhealthy = n_unh; gr.flags |= KR_GR_ABORTED; mode = GM_UNHEALTHY;
          cr.err_kind = KR_ERR_UNHEALTHY_WORKERS; cr.err_arg = n_unh; abort_here = true;
        } else {
          gr.flags |= KR_GR_WTD_EXECUTED; mode = GM_NORMAL;  // :814-849
          const int32_t running = n_list - n_wtd;
          const int32_t diff = expected - running;
          gr.n_running = running; gr.diff = diff;
          if (diff > 0) gr.n_create = (uint32_t)diff;
          else if (diff < 0) {
            if (!autoscaling || a.f.env_random_pod_delete) {  // :898-928
              long long remove = -(long long)diff;
              if (remove > running) {  // expected < 0: the Go loop would index past runningPods (:917)
                prefix = running; gr.flags |= KR_GR_ABORTED;
                cr.err_kind = KR_ERR_NEGATIVE_EXPECTED; cr.err_arg = expected; abort_here = true;
              } else prefix = (int32_t)remove;
            } else gr.flags |= KR_GR_RANDOM_DELETE_OFF;
          }
        }
      }
      __syncwarp();  // every lane has read this group's counters before lane 0 recycles their cells
      if (lane == 0) {
        if (g_mode) { g_mode[gi] = mode; g_prefix[gi] = prefix; }
        else { a.sc.gacc[g] = mode; a.sc.gacc[a.n.n_groups + g] = prefix; }
        a.r.groups[g] = gr;
        a.sc.gcreate[g] = gr.n_create;
      }
      if (abort_here) { cr.stop_after_group = (int32_t)gi; break; }
    }
  }
  // groups never reached keep an all-zero record
  {
    const int32_t reached = ((cf & KR_CF_SKIP) || !run_groups) ? 0 : (cr.stop_after_group == (int32_t)G ? (int32_t)G : cr.stop_after_group + 1);
    for (uint32_t gi = reached + lane; gi < G; gi += 32) {
      kr_group_result z; z.expected = 0; z.n_list = 0; z.n_unhealthy = 0; z.n_running = 0; z.diff = 0; z.n_create = 0; z.create_off = 0; z.flags = 0;
      a.r.groups[g0 + gi] = z;
      a.sc.gcreate[g0 + gi] = 0;
      if (!g_mode) { a.sc.gacc[g0 + gi] = GM_UNPROCESSED; a.sc.gacc[a.n.n_groups + g0 + gi] = 0; }
    }
  }
  __syncwarp();
  const int32_t *mode_arr = g_mode ? g_mode : a.sc.gacc + g0;
  const int32_t *prefix_arr = g_prefix ? g_prefix : a.sc.gacc + a.n.n_groups + g0;

  // ---------------- scan 2: per-pod actions in list order
  uint32_t n_act = 0;
#pragma unroll
  for (int k = 0; k < (K ? K : 1 << 30); k++) {
    if ((uint32_t)k >= nchunks) break;
    const uint32_t i = seg0 + k * 32 + lane;
    const bool valid = i < seg1;
    uint32_t pod, w = 0;
    if (K) { pod = pidx[K ? k : 0]; w = pw[K ? k : 0]; }
    else {
      pod = valid ? LDG(a.r.sorted_pod_idx[i]) : 0u;
      if (valid && run_groups) w = reinterpret_cast<const uint32_t *>(a.sc.rows)[4 * (size_t)pod + 3];
    }
    uint8_t act = KR_ACT_KEEP;
    uint32_t gkey = KR_ROW_NO_GROUP;
    const uint32_t fl = w & 0xFFFFu;
    if (valid && run_groups && (w >> 16) < G) gkey = w >> 16;
    const int32_t mode = (gkey != KR_ROW_NO_GROUP) ? mode_arr[gkey] : GM_UNPROCESSED;
    bool candidate = false;  // running pod of a group in normal mode: subject to the ordered delete prefix
    if (all_action != KR_ACT_KEEP) act = valid ? all_action : (uint8_t)KR_ACT_KEEP;
    else if (head_delete) { if (valid && (int32_t)pod == head_pod) act = KR_ACT_DELETE_HEAD; }
    else if (mode == GM_SUSPENDED) act = KR_ACT_DELETE_GROUP_SUSPEND;
    else if (kMH && mode == GM_MULTIHOST) act = a.sc.mh_act[i];
    else if (mode == GM_UNHEALTHY) { if (should_delete(fl)) act = KR_ACT_DELETE_UNHEALTHY; }
    else if (mode == GM_NORMAL) {
      if (fl & KR_ROW_WTD_OWN) act = KR_ACT_DELETE_WTD;
      else candidate = true;
    }
    // stable rank among the running pods of the same group: ballot group-by + per-group cursor
    const uint32_t ckey = candidate ? gkey : KR_ROW_NO_GROUP;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ckey);
    if (candidate) {
      const int32_t cur = acc_rank[ckey];
      __syncwarp(peers);
      const int32_t rank = cur + __popc(peers & lt);
      if ((peers & lt) == 0) acc_rank[ckey] = cur + __popc(peers);
      if (rank < prefix_arr[ckey]) act = KR_ACT_DELETE_RANDOM;  // runningPods.Items[0 .. -diff) (:916-919)
    }
    __syncwarp();
    if (valid) a.r.sorted_action[i] = act;
    n_act += __popc(__ballot_sync(0xFFFFFFFFu, valid && act != KR_ACT_KEEP));
  }

  // ---------------- status roll-up + record
  if (lane == 0) {
    if (!(cf & KR_CF_SKIP))
      status_rollup(a, c, cr, P, (uint32_t)n_heads, head_pod, head_name, ready, available, all_running);
    a.r.clusters[c] = cr;
    a.sc.cact[c] = n_act;
    if (n_act) atomicAdd(&a.r.totals[2], n_act);
  }
}

// Sort a bucket of <= 32*K pod indices in registers, publish it (sorted_pod_idx), gather the row words, decide.
template <int K>
__device__ __forceinline__ void decide_cluster_regs(const DecideArgs &a, uint32_t c, uint32_t seg0, uint32_t seg1,
                                                    int32_t (&s_acc)[4][KR_SMEM_GROUPS], int32_t (&s_mode)[2][KR_SMEM_GROUPS], uint32_t lane) {
  uint32_t pidx[K], pw[K];
  const uint32_t P = seg1 - seg0;
#pragma unroll
  for (int k = 0; k < K; k++) { uint32_t g = k * 32 + lane; pidx[k] = g < P ? LDG(a.unsorted[seg0 + g]) : 0xFFFFFFFFu; }
  warp_bitonic_sort_striped<K>(pidx, lane);
#pragma unroll
  for (int k = 0; k < K; k++) {
    uint32_t g = k * 32 + lane;
    pw[k] = 0;
    if (g < P) { a.r.sorted_pod_idx[seg0 + g] = pidx[k]; pw[k] = reinterpret_cast<const uint32_t *>(a.sc.rows)[4 * (size_t)pidx[k] + 3]; }
    else pidx[k] = 0;
  }
  __syncwarp();  // sorted_pod_idx of this bucket is visible to the whole warp (decide_multihost re-reads it)
  decide_cluster<K, false>(a, c, seg0, seg1, pidx, pw, s_acc, s_mode, lane);
}

// Is this cluster decided by k_decide_small (bucket sorted and kept in registers)?  Fast pipeline, phase 0, at most 256
// pods, no multi-host worker group (those need the memory-resident sweeps of decide_multihost).
__device__ __forceinline__ bool small_path(const DecideArgs &a, uint32_t c, uint32_t P) {
  return a.fast && a.phase == 0 && P <= 256 && !(a.f.gate_multihost_indexing && (__ldg(&a.sc.cl_rec[c]).w & 1u));
}

// Common case: one warp per RayCluster with <= 256 pods, everything after the bucket load stays in registers.
__global__ void __launch_bounds__(kDecideWarps * 32, 8) k_decide_small(DecideArgs a) {
  KR_TL(3);
  __shared__ int32_t s_acc[kDecideWarps][4][KR_SMEM_GROUPS];  // n_list, n_unhealthy, n_wtd_own, running-rank cursor
  __shared__ int32_t s_mode[kDecideWarps][2][KR_SMEM_GROUPS]; // mode, delete-prefix length
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = blockIdx.x * kDecideWarps + warp;
  pdl_wait(); pdl_trigger();
  if (c >= a.n.n_clusters) return;
  const uint32_t seg0 = LDG(a.sc.cstart[c]), seg1 = LDG(a.sc.cstart[c + 1]);
  const uint32_t P = seg1 - seg0;
  if (!small_path(a, c, P)) return;
  if (P <= 128) decide_cluster_regs<4>(a, c, seg0, seg1, s_acc[warp], s_mode[warp], lane);
  else decide_cluster_regs<8>(a, c, seg0, seg1, s_acc[warp], s_mode[warp], lane);
}

// General case: radix pipeline (all clusters), big buckets, clusters with multi-host groups, phase 1, the orphan bucket.
__global__ void __launch_bounds__(kDecideWarps * 32) k_decide(DecideArgs a) {
  KR_TL(4 + a.phase);
  __shared__ int32_t s_acc[kDecideWarps][4][KR_SMEM_GROUPS];
  __shared__ int32_t s_mode[kDecideWarps][2][KR_SMEM_GROUPS];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t Nc = a.n.n_clusters, Np = a.n.n_pods;
  uint32_t c = blockIdx.x * kDecideWarps + warp;
  if (a.phase == 1) {  // compact list of the clusters phase 0 deferred
    if (c >= a.r.totals[4]) return;
    c = a.sc.deferred_list[c];
  } else if (c > Nc) return;

  uint32_t seg0, seg1;
  if (a.fast) { seg0 = LDG(a.sc.cstart[c]); seg1 = LDG(a.sc.cstart[c + 1]); }
  else {
    seg0 = warp_lower_bound(a.sorted_keys, Np, c, lane);
    seg1 = (c == Nc) ? Np : warp_lower_bound(a.sorted_keys, Np, c + 1, lane);
  }
  const uint32_t P = seg1 - seg0;
  if (a.phase == 0) {
    // The orphans' segment: pods whose (namespace, ray.io/cluster) names no RayCluster in the snapshot, and the free rows of an
    // incrementally maintained arena (KR_PP_TOMBSTONE; there can be many).  Every warp of the grid labels a strided share.
    const uint32_t o0 = a.fast ? LDG(a.sc.cstart[Nc]) : warp_lower_bound(a.sorted_keys, Np, Nc, lane);
    const uint32_t gw = blockIdx.x * kDecideWarps + warp, nw = gridDim.x * kDecideWarps;
    uint32_t real = 0;
    for (uint32_t i = o0 + gw * 32 + lane; i < Np; i += nw * 32) {
      const uint32_t pod = a.fast ? LDG(a.unsorted[i]) : a.r.sorted_pod_idx[i];  // k_match/k_place put this segment in List order already
      if (a.fast) a.r.sorted_pod_idx[i] = pod;
      const bool tomb = reinterpret_cast<const uint32_t *>(a.sc.rows)[4 * (size_t)pod + 3] & KR_PP_TOMBSTONE;
      a.r.sorted_action[i] = tomb ? KR_ACT_TOMBSTONE : KR_ACT_ORPHAN;
      real += tomb ? 0u : 1u;
    }
    real = __reduce_add_sync(0xFFFFFFFFu, real);
    if (lane == 0 && real) atomicAdd(&a.r.totals[1], real);
  }
  if (c == Nc) return;
  if (small_path(a, c, P)) return;  // k_decide_small owns it
  // fast pipeline, phase 0: informer List order inside the bucket = ascending pod index (phase 1 finds it already sorted)
  if (a.fast && a.phase == 0 && P <= KR_FAST_MAX_BUCKET) { warp_sort_dispatch(a.unsorted + seg0, a.r.sorted_pod_idx + seg0, P, lane); __syncwarp(); }
  uint32_t d0[1] = {0}, d1[1] = {0};
  decide_cluster<0, true>(a, c, seg0, seg1, d0, d1, s_acc[warp], s_mode[warp], lane);
}

// ------------------------------------------------------------------------------------------------ creates

// exclusive scan of the dense n_create array -> groups[].create_off, total in totals[0] (chained multi-block scan).
__global__ void __launch_bounds__(1024) k_scan_creates(ResDev r, const uint32_t *__restrict__ gcreate, uint32_t n_groups, uint32_t *chain) {
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
      if (i < seg1 && (mh ? sc.mh_head[i] != 0 : r.sorted_action[i] == KR_ACT_KEEP)) {  // runningPods: listed and not deleted by name
        uint4 row = sc.rows[r.sorted_pod_idx[i]];
        if ((row.w >> 16) == slot && (row.w & KR_PP_HAS_REPLICA_IDX)) {
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
  __shared__ uint32_t s_bits[4][32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * 4 + warp;
  if (g >= n.n_groups) return;
  create_fill_group(s, sc, r, f, g, r.groups[g].create_off, create_cap, s_bits[warp], lane);
}

// Compact action list of one cluster: (pod idx, action) of every pod whose action != KEEP, List order kept (one warp).
__device__ __forceinline__ void compact_cluster_actions(const ResDev &r, uint32_t c, uint32_t dst, uint32_t lane) {
  const kr_cluster_result *cr = &r.clusters[c];
  const uint32_t seg0 = cr->pod_start;
  // n_pods is only filled when calculateStatus ran; a cluster with actions always has it
  const uint32_t seg1 = seg0 + (uint32_t)cr->n_pods;
  const uint32_t lt = lanemask_lt();
  for (uint32_t b = seg0; b < seg1; b += 32) {
    uint32_t i = b + lane;
    uint8_t act = i < seg1 ? r.sorted_action[i] : (uint8_t)KR_ACT_KEEP;
    uint32_t bal = __ballot_sync(0xFFFFFFFFu, act != KR_ACT_KEEP);
    if (act != KR_ACT_KEEP) { uint32_t o = dst + __popc(bal & lt); r.act_pod_idx[o] = r.sorted_pod_idx[i]; r.act_code[o] = act; }
    dst += __popc(bal);
  }
}

// unfused path: starts of the per-cluster action lists (chained scan) ...
__global__ void __launch_bounds__(1024) k_scan_actions(ResDev r, const uint32_t *__restrict__ cact, uint32_t n_clusters, uint32_t *chain) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_prefix;
  uint32_t excl[8];
  bool big = false;
  const uint32_t chunk = blockIdx.x;
  uint32_t carry = chained_scan_chunk(cact, n_clusters, chunk, chain, 0, big, excl, s_warp, &s_prefix);
  const uint32_t i0 = chunk * kScanChunk + threadIdx.x * 8;
#pragma unroll
  for (int k = 0; k < 8; k++) if (i0 + k < n_clusters) r.act_start[i0 + k] = excl[k];
  if (chunk == gridDim.x - 1 && threadIdx.x == 0) r.act_start[n_clusters] = carry;
}
// ... and the lists themselves, one warp per cluster
__global__ void __launch_bounds__(128) k_compact_actions(ResDev r, const uint32_t *__restrict__ cact, uint32_t n_clusters) {
  const uint32_t c = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= n_clusters || cact[c] == 0) return;
  compact_cluster_actions(r, c, r.act_start[c], threadIdx.x & 31);
}

// ---- fused variants for snapshots whose per-cluster / per-group counters fit in shared memory: every block scans the counters
// itself (a few tens of KB out of L2) instead of waiting for a scan kernel, which removes two ~10 us stages from the chain.
static constexpr uint32_t kFusedMaxCounters = 48 * 1024;  // 192 KB of shared memory

// exclusive scan of in[0..n) into shared memory by the whole block (any block size that is a multiple of 32, <= 1024)
__device__ __forceinline__ uint32_t block_scan_to_smem(const uint32_t *__restrict__ in, uint32_t n, uint32_t *out_sm, uint32_t big_limit, bool &big,
                                                       uint32_t *s_warp, uint32_t *s_carry) {
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
  if (t == 0) *s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += blockDim.x * 8) {
    uint32_t i0 = base + t * 8;
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (i0 + k < n) ? __ldg(&in[i0 + k]) : 0u;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { sum += v[k]; big |= (i0 + k < big_limit) && v[k] > KR_FAST_MAX_BUCKET; }
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
    for (int k = 0; k < 8; k++) { if (i0 + k < n) out_sm[i0 + k] = run; run += v[k]; }
    __syncthreads();
    if (t == 0) *s_carry += total;
    __syncthreads();
  }
  return *s_carry;
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
    if (big) atomicOr(&totals[3], KR_TOTALS_BIG_BUCKET);
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
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_bits[32][32];
  pdl_wait(); pdl_trigger();
  uint32_t *sm_off = sm_dyn;               // [n_groups] create offsets
  uint32_t *sm_act = sm_dyn + n.n_groups;  // [n_clusters + 1] action-list starts
  bool dummy = false;
  uint32_t tot = block_scan_to_smem(sc.gcreate, n.n_groups, sm_off, 0, dummy, s_warp, &s_carry);
  uint32_t tot_act = block_scan_to_smem(sc.cact, n.n_clusters, sm_act, 0, dummy, s_warp, &s_carry);
  if (threadIdx.x == 0) sm_act[n.n_clusters] = tot_act;
  __syncthreads();
  if (blockIdx.x == 0) {
    for (uint32_t g = threadIdx.x; g < n.n_groups; g += blockDim.x) r.groups[g].create_off = sm_off[g];
    for (uint32_t c = threadIdx.x; c <= n.n_clusters; c += blockDim.x) r.act_start[c] = sm_act[c];
    if (threadIdx.x == 0) r.totals[0] = tot;
  }
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (uint32_t g = blockIdx.x * nw + warp; g < n.n_groups; g += gridDim.x * nw)
    if (__ldg(&sc.gcreate[g])) create_fill_group(s, sc, r, f, g, sm_off[g], create_cap, s_bits[warp], lane);
  for (uint32_t c = blockIdx.x * nw + warp; c < n.n_clusters; c += gridDim.x * nw)
    if (sm_act[c + 1] != sm_act[c]) compact_cluster_actions(r, c, sm_act[c], lane);
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

// ------------------------------------------------------------------------------------------------ k_hash
// base32hex(sha1(json)) per RayCluster (utils/util.go:628-640).  One lane per message (SHA-1 is a serial chain per
// message); the warp stages 128 bytes of each of its 32 messages per step with coalesced 16-byte loads into an
// XOR-swizzled shared tile, so the per-lane reads are conflict-free LDS.128.

__device__ __forceinline__ uint32_t rol(uint32_t x, int k) { return __funnelshift_l(x, x, k); }
__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// ------------------------------------------------------------------------------------------------ k_hash2
// Second-generation hash kernel.  Same one-lane-per-message mapping, but
//  (1) the 128-byte chunks are fetched with cp.async (LDGSTS) straight into a double-buffered, XOR-swizzled shared tile:
//      no register staging, and the fetch of chunk i+1 is in flight during the 160 rounds of chunk i by construction;
//  (2) each round is written so that the only operation on the serial a->a chain is rol5(a)+s (one LEA.HI); s = f+e+K+w is
//      formed off the chain;
//  (3) VARIANT 1 forms s with IMADs (multiply by an opaque 1 from the constant bank) so those adds issue on the FMA pipe
//      while LOP3/SHF/LEA keep the ALU pipe.  Measured on B200 (tools/hash_bench.cu, profiles/r1_hash_variants.txt): with one
//      warp per scheduler (10k messages) VARIANT 0 wins (82 us vs 94 us; 8 ALU-pipe instructions per round at 2 cycles
//      each is the floor), with many warps per scheduler (100k messages) VARIANT 1 wins (1.04 vs 0.95 TB/s).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t mad1(uint32_t a, uint32_t one, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
  return d;
}

// rol(x, n) on the FMA pipe: x * 2^n as a 64-bit product puts x << n in the low word and x >> (32 - n) in the high word; the
// two halves have no bit in common, so lo * 1 + hi is the rotation.  `pow2` and `one` are opaque (derived from a kernel
// parameter), otherwise ptxas strength-reduces both back to ALU-pipe shifts.
__device__ __forceinline__ uint32_t rol_fma(uint32_t x, uint32_t pow2, uint32_t one) {
  uint32_t r;
  asm("{\n\t.reg .u64 t;\n\t.reg .u32 lo, hi;\n\tmul.wide.u32 t, %1, %2;\n\tmov.b64 {lo, hi}, t;\n\tmad.lo.u32 %0, lo, %3, hi;\n\t}" : "=r"(r) : "r"(x), "r"(pow2), "r"(one));
  return r;
}

// VARIANT: 0 = every round operation on the ALU pipe; 1 = s formed by two IMADs; 2..5 = experiments that move off-chain work
// to the FMA pipe (5: w+K; 2: w+K and rol30(b); 3: w+K and the schedule's rol1; 4: all three) hoping a lone warp would
// alternate pipes.  It does not pay: at 10k messages 0 -> 82 us, 5 -> 96, 2 -> 121, 3 -> 125, 4 -> 143 us; at 100k messages
// only VARIANT 1 beats 0 (345 vs 375 us).  The engine uses 0 (latency regime) and 1 (throughput regime); 2..5 stay for
// tools/hash_bench.cu, which reproduces the table (profiles/r1_hash_variants.txt).
template <int VARIANT>
__device__ __forceinline__ void sha1_rounds2(uint32_t (&w)[16], uint32_t (&h)[5], uint32_t one) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
  const uint32_t two = one << 1, two30 = one << 30;
  constexpr bool kFmaRol30 = VARIANT == 2 || VARIANT == 4, kFmaRol1 = VARIANT == 3 || VARIANT == 4;
#pragma unroll
  for (int i = 0; i < 80; i++) {
    uint32_t wi;
    if (i < 16) wi = w[i];
    else {
      const uint32_t x = w[(i - 3) & 15] ^ w[(i - 8) & 15] ^ w[(i - 14) & 15] ^ w[i & 15];
      wi = kFmaRol1 ? rol_fma(x, two, one) : rol(x, 1);
      w[i & 15] = wi;
    }
    const uint32_t k = i < 20 ? 0x5A827999u : (i < 40 ? 0x6ED9EBA1u : (i < 60 ? 0x8F1BBCDCu : 0xCA62C1D6u));
    uint32_t f;
    if (i < 20) f = (b & c) | (~b & d);
    else if (i < 40) f = b ^ c ^ d;
    else if (i < 60) f = (b & c) | (b & d) | (c & d);
    else f = b ^ c ^ d;
    uint32_t s;
    if (VARIANT == 0) s = f + e + (wi + k);            // lone warp per scheduler (latency regime): fewest instructions wins
    else if (VARIANT == 1) s = mad1(f, one, mad1(e, one, wi + k));  // many warps per scheduler (throughput regime): adds on the FMA pipe
    else s = f + e + mad1(wi, one, k);                  // w + K is far off the chain: FMA pipe
    asm volatile("" : "+r"(s));  // keep s a value of its own: the a->a chain below is then a single rol5(a)+s
    uint32_t t = rol(a, 5) + s;
    e = d; d = c; c = kFmaRol30 ? rol_fma(b, two30, one) : rol(b, 30); b = a; a = t;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

template <int WARPS, int VARIANT>
__global__ void __launch_bounds__(WARPS * 32) k_hash2(const uint8_t *__restrict__ bytes, const uint64_t *__restrict__ off,
                                                      const uint32_t *__restrict__ len32, const uint64_t *__restrict__ off_end,
                                                      uint32_t n, char *__restrict__ out, uint32_t one = 1) {
  KR_TL(7);
  __shared__ uint4 s_tile[2][WARPS][32][8];  // [buffer][warp][message lane][16-byte piece ^ (lane & 7)]
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid-stride over groups of WARPS*32 messages: the engine caps the grid for large n so that the hash leaves room on every
  // SM for the main chain's blocks (each warp owns its shared tile, so the trips need no block-wide barrier)
  for (uint32_t grp_i = blockIdx.x; (uint64_t)grp_i * (WARPS * 32) < n; grp_i += gridDim.x) {
  const uint32_t m = (grp_i * WARPS + warp) * 32 + lane;
  const bool have = m < n;
  uint64_t moff = 0;
  uint32_t mlen = 0;
  if (have) { moff = off[m]; mlen = len32 ? len32[m] : (uint32_t)(off_end[m] - moff); }
  const uint32_t nblocks = have ? (mlen + 8) / 64 + 1 : 0;
  uint32_t max_blocks = nblocks;
#pragma unroll
  for (int d = 16; d; d >>= 1) max_blocks = max(max_blocks, __shfl_xor_sync(0xFFFFFFFFu, max_blocks, d));
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  const uint32_t sub = lane & 7, grp = lane >> 3;
  // this lane fetches piece `sub` of messages 4r+grp, r = 0..7: keep their base pointers and padded lengths
  const uint8_t *src[8];
  uint32_t lim[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    uint32_t sl = 4 * r + grp;
    uint64_t o = __shfl_sync(0xFFFFFFFFu, moff, sl);
    uint32_t l = __shfl_sync(0xFFFFFFFFu, mlen, sl);
    src[r] = bytes + o + sub * 16;
    lim[r] = (l + 15) & ~15u;  // the arena pads every message to 16 bytes
  }
  auto fetch = [&](uint32_t chunk, int buf) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t sl = 4 * r + grp;
      uint32_t pos = chunk * 128 + sub * 16;
      bool in = pos < lim[r];
      cp_async16(&s_tile[buf][warp][sl][sub ^ (sl & 7)], in ? (const void *)(src[r] + (size_t)chunk * 128) : (const void *)bytes, in ? 16u : 0u);
    }
    cp_async_commit();
  };
  const uint32_t nchunks = (max_blocks + 1) / 2;
  if (nchunks) fetch(0, 0);
  for (uint32_t chunk = 0; chunk < nchunks; chunk++) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) { fetch(chunk + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; half++) {
      uint32_t blk = chunk * 2 + half;
      if (blk >= nblocks) continue;
      uint32_t w[16];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint4 v = s_tile[buf][warp][lane][(half * 4 + q) ^ (lane & 7)];
        w[4 * q] = bswap(v.x); w[4 * q + 1] = bswap(v.y); w[4 * q + 2] = bswap(v.z); w[4 * q + 3] = bswap(v.w);
      }
      const uint32_t bstart = blk * 64;
      if (bstart + 64 > mlen) {  // tail block(s): 0x80, zero fill, 64-bit big-endian bit length (FIPS 180-4 §5.1.1)
#pragma unroll
        for (int q = 0; q < 16; q++) {
          uint32_t wpos = bstart + 4 * q;
          uint32_t v = w[q];
          if (wpos >= mlen) v = (wpos == mlen) ? 0x80000000u : 0u;
          else if (wpos + 4 > mlen) {
            uint32_t keep = mlen - wpos;  // 1..3 message bytes in this word
            v = (v & (0xFFFFFFFFu << (8 * (4 - keep)))) | (0x80u << (8 * (3 - keep)));
          }
          w[q] = v;
        }
        if (blk == nblocks - 1) { w[14] = mlen >> 29; w[15] = mlen << 3; }
      }
      sha1_rounds2<VARIANT>(w, h, one);
    }
    __syncwarp();  // every lane is done reading this buffer before the fetch two iterations ahead overwrites it
  }
  if (have) {
  uint32_t o32[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint64_t v;
    switch (j) {
      case 0: v = ((uint64_t)h[0] << 8) | (h[1] >> 24); break;
      case 1: v = ((uint64_t)(h[1] & 0xFFFFFFu) << 16) | (h[2] >> 16); break;
      case 2: v = ((uint64_t)(h[2] & 0xFFFFu) << 24) | (h[3] >> 8); break;
      default: v = ((uint64_t)(h[3] & 0xFFu) << 32) | h[4]; break;
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      uint32_t cc = (uint32_t)(v >> (35 - 5 * kk)) & 31u;
      uint32_t ch = cc < 10 ? ('0' + cc) : ('A' + cc - 10);
      if (kk < 4) lo |= ch << (8 * kk); else hi |= ch << (8 * (kk - 4));
    }
    o32[2 * j] = lo; o32[2 * j + 1] = hi;
  }
  uint4 *dst = reinterpret_cast<uint4 *>(out + 32 * (size_t)m);
  dst[0] = make_uint4(o32[0], o32[1], o32[2], o32[3]);
  dst[1] = make_uint4(o32[4], o32[5], o32[6], o32[7]);
  }
  __syncwarp();
  }  // next group of messages
}

}  // namespace kr
