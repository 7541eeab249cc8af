// kr_incr.cuh — device-side incremental epochs on top of the bucket pipeline (kr_bucket2.cuh).
// Part of the sm_100a kernel set of the batched reconcile engine; see kr_kernels.cuh for the pipeline overview.
//
// A full pass leaves everything a later pass needs resident on the device: the join tables, every RayCluster's bucket of pod
// records and its {count, first head} cell, the 128-byte input records, the digests and the result records.  The informer
// events of an epoch (raycluster_controller.go:1525-1533 watches RayClusters and the Pods they own) then arrive as
//   * pod rows     kr_snapshot_commit_pod_rows / _values   -> k_inc_retire runs on the rows' OLD values before the patch lands:
//                                                             it takes each row out of the orphan count / workersToDelete
//                                                             resolutions, stamps it and marks the RayCluster it was in dirty;
//   * object rows  kr_snapshot_commit_parts(KR_PART_OBJECTS) -> uploaded beside the resident tables and diffed by k_inc_objects:
//                                                             changed RayCluster / group / head-aux rows mark their cluster
//                                                             dirty, a changed table key or CSR offset makes the epoch
//                                                             "structural" (the next pass is a full one).
// The pass then touches only what changed:
//   k_inc_admit    (one thread per touched row) runs the selector match of k_match2 on the row's NEW values and appends the
//                  record (marked KR_ROW_FRESH) to its RayCluster's bucket — marking that RayCluster dirty as well;
//   k_inc_refresh  (launched by the object commit itself, behind its diff kernel) rewrites the 128-byte input record of the RayClusters whose
//                  RayCluster / group rows changed;
//   k_decide2<K>   phase 2: the decide kernel over the dirty list.  Each warp first drops the records of stamped rows from its bucket (stored back compacted, arrival order
//                  kept) and recomputes the first head; then decides as in a full pass (digests are resident, so the Recreate
//                  gate is decided in place); a cluster keeps its places in the action list / create arena while they suffice;
//                  and packs the cluster's changed records for one small D2H copy (IncStage).
// Results are bit-identical to a full pass over the same state (tests/test_live_arena.py, tests/test_packer.py run every epoch
// against the oracle); anything the resident state cannot absorb — structural object changes, a bucket or arena overflow — voids
// the attempt and the engine takes the full pass instead.
#pragma once

#include "kr_bucket2.cuh"

namespace kr {

__device__ __forceinline__ uint32_t inc_epoch(const ScratchDev &sc) { return inc_epoch_of(sc); }

__device__ __forceinline__ void mark_dirty(const ScratchDev &sc, uint32_t c, uint32_t epoch) {
  if (atomicExch(&sc.dirty_flag[c], epoch) != epoch) sc.dirty_list[atomicAdd(&sc.inc[KR_INC_DIRTY], 1u)] = c;
}

// (namespace, ray.io/cluster) -> cluster idx, table flags and the name of worker group 0 (the probe of k_match2)
__device__ __forceinline__ bool cl_probe(const ScratchDev &sc, uint32_t ns, uint32_t name, uint32_t &c, uint32_t &cflags, uint32_t &gname0) {
  if (name == 0) return false;
  uint32_t i = hash_pair(ns, name) & sc.cl_mask;
  while (true) {
    const uint4 q = __ldcg(&sc.cl_slots[i]);
    if (q.x == name && q.y == ns) { c = q.w >> 2; cflags = q.w & 3u; gname0 = q.z; return true; }
    if (q.x == KR_EMPTY32 && q.y == KR_EMPTY32) return false;
    i = (i + 1) & sc.cl_mask;
  }
}

// ------------------------------------------------------------------------------------------------ k_inc_retire
// One thread per committed row, launched in front of the patch kernel: the device columns still hold the row's previous
// values.  n_resident = pod rows that existed at the last pass (rows appended since hold nothing to retire).
__global__ void __launch_bounds__(256) k_inc_retire(const uint32_t *rows, uint32_t n_rows, SnapDev s, ScratchDev sc, ResDev r, Sizes n, uint32_t n_resident, int has_wtd) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const uint32_t p = rows[i];
  const uint32_t epoch = inc_epoch(sc);
  if (atomicExch(&sc.stamp[p], epoch) == epoch) return;  // committed twice since the last pass: retired already
  const uint32_t ti = atomicAdd(&sc.inc[KR_INC_TOUCHED], 1u);
  sc.touched[ti] = p; sc.touched_old[ti] = KR_EMPTY32;
  if (p >= n_resident) return;
  const uint32_t ns = s.p_ns_id[p], cn = s.p_cluster_name_id[p], nm = s.p_name_id[p], pk = s.p_packed[p];
  uint32_t c = 0, cflags, gname0;
  if (cl_probe(sc, ns, cn, c, cflags, gname0)) { mark_dirty(sc, c, epoch); sc.touched_old[ti] = c; }  // (k_inc_admit rewrites the record in place if the row stays)
  else if (!(pk & KR_PP_TOMBSTONE)) atomicSub(&r.totals[1], 1u);  // it was an orphan
  if (has_wtd) {  // names that resolved to this row: (namespace, name) is unique among live Pods, so nothing else holds them
    const uint32_t hk = hash_pair(ns, nm);
    if ((__ldcg(&sc.wt_bits[(hk & sc.wt_bits_mask) >> 5]) & (1u << (hk & 31))) && (__ldcg(&sc.wt_bits[(bloom2(hk) & sc.wt_bits_mask) >> 5]) & (1u << (bloom2(hk) & 31)))) {
      const uint64_t k = key2(ns, nm);
      uint32_t j = hk & sc.wt_mask;
      uint64_t kk = __ldcg(&sc.wt_keys[j]);
      while (kk != KR_EMPTY64) {
        if (kk == k) {
          for (uint32_t e = sc.wt_head[j]; e != KR_EMPTY32; e = sc.wt_next[e]) atomicCAS(&r.wtd_pod_idx[e], p, 0xFFFFFFFFu);
          break;
        }
        j = (j + 1) & sc.wt_mask;
        kk = __ldcg(&sc.wt_keys[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ k_inc_objects
// Diff of a KR_PART_OBJECTS upload (staged beside the resident tables) against the resident copy, then the copy into place.
enum { KR_OC_COPY = 0, KR_OC_STRUCT = 1, KR_OC_CLUSTER = 2, KR_OC_GROUP = 3, KR_OC_HEAD = 4, KR_OC_HEADKEY = 5 };
static constexpr int kMaxObjCols = 48;
struct ObjDiffArgs {
  const uint8_t *src[kMaxObjCols];   // staged (new) column
  const uint32_t *rowlist[kMaxObjCols];  // NULL: staged row k is resident row k (a whole-table upload); else staged row k is resident row rowlist[k]
                                         // (kr_snapshot_commit_object_rows: only the rewritten rows travel, packed)
  uint8_t *dst[kMaxObjCols];         // resident column
  uint32_t first[kMaxObjCols + 1];   // flat index of the column's first row (prefix sums of the new row counts)
  uint32_t rows_old[kMaxObjCols];    // rows the resident column held
  uint16_t row_bytes[kMaxObjCols];
  uint8_t cls[kMaxObjCols];
  int n_cols;
  const uint32_t *g_cluster_idx_new;  // staged g_cluster_idx (group row -> RayCluster)
  const uint32_t *h_pod_idx_new;      // staged h_pod_idx
  const uint32_t *h_pod_idx_old;      // resident h_pod_idx (read before this launch's copy of that column: it is diffed LAST)
  uint32_t n_heads_old;
};

__device__ __forceinline__ void mark_pod_cluster_dirty(const SnapDev &s, const ScratchDev &sc, const Sizes &n, uint32_t p, uint32_t epoch) {
  if (p >= n.n_pods) return;
  uint32_t c = 0, cflags, gname0;
  if (cl_probe(sc, s.p_ns_id[p], s.p_cluster_name_id[p], c, cflags, gname0)) mark_dirty(sc, c, epoch);
}

__global__ void __launch_bounds__(256) k_inc_objects(ObjDiffArgs a, SnapDev s, ScratchDev sc, Sizes n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.first[a.n_cols]) return;
  int lo = 0, hi = a.n_cols - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.first[mid] <= t) lo = mid; else hi = mid - 1; }
  const int col = lo;
  const uint32_t k_st = t - a.first[col], rb = a.row_bytes[col];
  const uint32_t row = a.rowlist[col] ? a.rowlist[col][k_st] : k_st;
  const uint8_t *src = a.src[col] + (size_t)k_st * rb;
  uint8_t *dst = a.dst[col] + (size_t)row * rb;
  bool differ = row >= a.rows_old[col];
  if (!differ) {
    if ((rb & 3u) == 0) { for (uint32_t k = 0; k < rb; k += 4) differ |= *reinterpret_cast<const uint32_t *>(src + k) != *reinterpret_cast<const uint32_t *>(dst + k); }
    else for (uint32_t k = 0; k < rb; k++) differ |= src[k] != dst[k];
  }
  if (!differ) return;
  const uint32_t epoch = inc_epoch(sc);
  switch (a.cls[col]) {
    case KR_OC_STRUCT: sc.inc[KR_INC_STRUCTURAL] = 1u; break;
    case KR_OC_CLUSTER: if (row < n.n_clusters) { sc.obj_flag[row] = epoch; mark_dirty(sc, row, epoch); } break;
    case KR_OC_GROUP: { const uint32_t c = a.g_cluster_idx_new[k_st]; if (c < n.n_clusters) { sc.obj_flag[c] = epoch; mark_dirty(sc, c, epoch); } break; }
    case KR_OC_HEADKEY:  // (the host compared the keys as well and rebuilds the pod -> row table); both pods' clusters see a different head-aux row now
    case KR_OC_HEAD:
      mark_pod_cluster_dirty(s, sc, n, a.h_pod_idx_new[k_st], epoch);
      if (row < a.n_heads_old) mark_pod_cluster_dirty(s, sc, n, a.h_pod_idx_old[row], epoch);
      break;
    default: break;
  }
  if (a.cls[col] == KR_OC_HEADKEY) return;  // (copied by k_inc_objects_keys once every head row has read the old key)
  if ((rb & 3u) == 0) { for (uint32_t k = 0; k < rb; k += 4) *reinterpret_cast<uint32_t *>(dst + k) = *reinterpret_cast<const uint32_t *>(src + k); }
  else for (uint32_t k = 0; k < rb; k++) dst[k] = src[k];
}

// second step of the object diff: h_pod_idx into place (every head row of k_inc_objects read the old keys first)
__global__ void __launch_bounds__(256) k_inc_objects_keys(const uint32_t *src, uint32_t *dst, uint32_t n, const uint32_t *rowlist) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[rowlist ? rowlist[t] : t] = src[t];
}

// ------------------------------------------------------------------------------------------------ head-aux table rebuild
// pod idx -> head-aux row, when an epoch changed h_pod_idx (a head Pod came or went): clear, then insert (k_build_tables' third part)
__global__ void __launch_bounds__(256) k_inc_aux_clear(ScratchDev sc) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= sc.aux_mask; i += gridDim.x * blockDim.x) { sc.aux_keys[i] = KR_EMPTY32; sc.aux_vals[i] = KR_EMPTY32; }
}
__global__ void __launch_bounds__(256) k_inc_aux_insert(SnapDev s, ScratchDev sc, Sizes n) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n.n_heads; t += gridDim.x * blockDim.x) {
    const uint32_t p = s.h_pod_idx[t];
    uint32_t i = mix32(p) & sc.aux_mask;
    while (true) {
      const uint32_t prev = atomicCAS(&sc.aux_keys[i], KR_EMPTY32, p);
      if (prev == KR_EMPTY32 || prev == p) { atomicMin(&sc.aux_vals[i], t); break; }
      i = (i + 1) & sc.aux_mask;
    }
  }
}

// every RayCluster whose Recreate gate reads a digest, when the spec JSON was committed again (the digests are being recomputed)
__global__ void __launch_bounds__(256) k_inc_mark_recreate(SnapDev s, ScratchDev sc, Sizes n) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n.n_clusters && (s.c_flags[c] & KR_CF_UPGRADE_RECREATE)) mark_dirty(sc, c, inc_epoch(sc));
}

// ------------------------------------------------------------------------------------------------ k_inc_refresh
// Input records (cl_in) of the RayClusters one of whose RayCluster / group rows an object commit changed (grid-stride over the
// dirty list as the commits left it; clusters k_inc_admit adds later had no object change).
__global__ void __launch_bounds__(256) k_inc_refresh(SnapDev s, ScratchDev sc) {
  const uint32_t n_dirty = __ldcg(&sc.inc[KR_INC_DIRTY]);
  const uint32_t epoch = inc_epoch(sc);
  if (__ldcg(&sc.inc[KR_INC_STRUCTURAL])) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_dirty; i += gridDim.x * blockDim.x) {
    const uint32_t c = sc.dirty_list[i];
    if (__ldcg(&sc.obj_flag[c]) == epoch) write_cl_in(s, sc, c, s.c_group_off[c], s.c_group_cnt[c]);
  }
}

// ------------------------------------------------------------------------------------------------ k_inc_admit
// The selector match of k_match2 for the touched rows' new values (grid-stride over the touched list).
__global__ void __launch_bounds__(256) k_inc_admit(SnapDev s, ScratchDev sc, ResDev r, Sizes n, int has_wtd) {
  const uint32_t n_touched = __ldcg(&sc.inc[KR_INC_TOUCHED]);
  const uint32_t epoch = inc_epoch(sc);
  if (__ldcg(&sc.inc[KR_INC_STRUCTURAL])) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_touched; i += gridDim.x * blockDim.x) {
    const uint32_t p = sc.touched[i], c_old = sc.touched_old[i];
    if (p >= n.n_pods) continue;
    const uint32_t ns = s.p_ns_id[p], cn = s.p_cluster_name_id[p], gn = s.p_group_name_id[p], nm = s.p_name_id[p], pk = s.p_packed[p];
    const uint32_t ri = (uint32_t)s.p_replica_index[p];
    uint32_t c = n.n_clusters, cflags = 0, gname0 = 0;
    const bool matched = cl_probe(sc, ns, cn, c, cflags, gname0);
    uint32_t slot = KR_ROW_NO_GROUP, g0 = 0xFFFFFFFFu;
    if (matched && gn != 0) {
      if (gname0 == gn) slot = 0;
      else if (cflags & KR_CL_MULTI) {
        const uint4 rec = sc.cl_rec[c];
        g0 = rec.x;
        for (uint32_t gi = 1; gi < rec.y; gi++)
          if (s.g_name_id[g0 + gi] == gn) { slot = gi; break; }
      }
    }
    uint32_t flags = pk & (0x7FFu | KR_PP_TOMBSTONE);
    if (should_delete(pk)) flags |= KR_ROW_UNHEALTHY;
    if (has_wtd) {
      const uint32_t hk = hash_pair(ns, nm);
      if ((__ldcg(&sc.wt_bits[(hk & sc.wt_bits_mask) >> 5]) & (1u << (hk & 31))) && (__ldcg(&sc.wt_bits[(bloom2(hk) & sc.wt_bits_mask) >> 5]) & (1u << (bloom2(hk) & 31)))) {
        const uint64_t k = key2(ns, nm);
        uint32_t j = hk & sc.wt_mask;
        uint64_t kk = __ldcg(&sc.wt_keys[j]);
        while (kk != KR_EMPTY64) {
          if (kk == k) {
            for (uint32_t e = sc.wt_head[j]; e != KR_EMPTY32; e = sc.wt_next[e]) {
              atomicMin(&r.wtd_pod_idx[e], p);
              if (slot != KR_ROW_NO_GROUP) {
                if (g0 == 0xFFFFFFFFu) g0 = s.c_group_off[c];
                const uint32_t g = g0 + slot, off = s.g_wtd_off[g];
                if (e >= off && e < off + s.g_wtd_cnt[g]) flags |= KR_ROW_WTD_OWN;
              }
            }
            break;
          }
          j = (j + 1) & sc.wt_mask;
          kk = __ldcg(&sc.wt_keys[j]);
        }
      }
    }
    if (matched && c == c_old) {
      // the usual event — a status update: the row stays in its RayCluster, its record is rewritten where it sits and the row's
      // stamp is lifted (nothing of this cluster has to be dropped on its account)
      sc.bucket[(size_t)c * sc.bucket_stride + sc.pos[p]] = make_uint4(p, (slot << 16) | flags, ri, nm);
      sc.stamp[p] = 0u;
      continue;  // (k_inc_retire marked the cluster dirty)
    }
    if (c_old != KR_EMPTY32) sc.cl_dyn[c_old].y = epoch;  // the old cluster lost this row: its decide warp drops the stale record (the row stays stamped)
    if (!matched) {
      if (!(pk & KR_PP_TOMBSTONE)) atomicAdd(&r.totals[1], 1u);
      continue;
    }
    mark_dirty(sc, c, epoch);  // the row joined this cluster: a fresh record at the end of its bucket
    const uint32_t rank = atomicAdd(&sc.cl_dyn[c].x, 1u);
    if (rank < sc.bucket_stride) { sc.bucket[(size_t)c * sc.bucket_stride + rank] = make_uint4(p, (slot << 16) | flags | KR_ROW_FRESH, ri, nm); sc.pos[p] = rank; }
    else sc.inc[KR_INC_VOID] = 1u;
  }
}

// ------------------------------------------------------------------------------------------------ k_inc_finish
// (The changed records are packed for the host by the decide warps themselves: IncStage in kr_bucket2.cuh.)

// closes the epoch (after the host's copy of the counters was enqueued): next epoch's stamps differ from every stamp written so far
__global__ void k_inc_finish(ScratchDev sc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    sc.inc[KR_INC_TOUCHED] = 0; sc.inc[KR_INC_DIRTY] = 0; sc.inc[KR_INC_STRUCTURAL] = 0; sc.inc[KR_INC_HEADS] = 0; sc.inc[KR_INC_VOID] = 0; sc.inc[KR_INC_GROUPS] = 0;
    sc.inc[KR_INC_EPOCH] += 1u;
  }
}

}  // namespace kr
