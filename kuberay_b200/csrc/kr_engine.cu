// kr_engine.cu — C ABI (include/kr_engine.h) of the batched reconcile engine: arenas, streams, kernel schedule.
//
// One engine = one device and four streams: M (clear -> tables -> match -> place -> decide -> creates), H (hash),
// G (general decide kernel beside the small one) and a copy stream; the pass is one CUDA graph joined by events.
// Inputs live in ONE pinned host arena and ONE device arena with identical layouts computed per snapshot from
// kr_sizes: a full commit is two contiguous asynchronous H2D copies (columns, then the spec-JSON arena), partial and
// per-row commits upload less (kr_snapshot_commit_parts / kr_snapshot_commit_pod_rows).  Results come back with the
// exact sizes read from a 32-byte totals record and a single host wait (fetch_results).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "kr_kernels.cuh"
#include "kr_incr.cuh"

using namespace kr;

// kr_specjson.cpp
int kr_specjson_emit_string(const uint8_t *spec_json, uint64_t len, bool muted, long max_groups, std::string &out, long *n_groups);

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }
inline uint32_t pow2_at_least(uint64_t x) { uint32_t p = 16; while (p < x) p <<= 1; return p; }

struct InLayout {  // offsets of every input column inside the snapshot arena
  size_t off[64];
  size_t total;
};

// (element size, per-row multiplicity, dimension index) in the order of kr_snapshot_bufs
enum { D_CLUSTERS, D_GROUPS, D_WTD, D_PODS, D_HEADS, D_JOBS, D_JSON };
struct ColDesc { uint8_t elem, mult, dim; };
constexpr ColDesc kCols[] = {
    {4, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS}, {8, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS}, {1, 1, D_CLUSTERS}, {1, 1, D_CLUSTERS},
    {4, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS}, {8, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS},
    {1, 1, D_CLUSTERS}, {4, 5, D_CLUSTERS}, {1, 5, D_CLUSTERS}, {1, 5, D_CLUSTERS}, {4, 1, D_CLUSTERS}, {4, 2, D_CLUSTERS},
    {4, 4, D_CLUSTERS}, {1, 1, D_CLUSTERS}, {1, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS}, {4, 1, D_CLUSTERS},
    {4, 1, D_GROUPS}, {4, 1, D_GROUPS}, {4, 1, D_GROUPS}, {4, 1, D_GROUPS}, {4, 1, D_GROUPS}, {4, 1, D_GROUPS},
    {4, 1, D_GROUPS}, {4, 1, D_GROUPS}, {4, 1, D_GROUPS},
    {4, 1, D_WTD},
    {4, 1, D_PODS}, {4, 1, D_PODS}, {4, 1, D_PODS}, {4, 1, D_PODS}, {4, 1, D_PODS}, {4, 1, D_PODS}, {4, 1, D_PODS},
    {4, 1, D_HEADS}, {1, 1, D_HEADS}, {4, 1, D_HEADS}, {4, 1, D_HEADS}, {4, 1, D_HEADS}, {1, 1, D_HEADS}, {1, 1, D_HEADS}, {1, 32, D_HEADS},
    {4, 1, D_JOBS}, {4, 1, D_JOBS}, {4, 1, D_JOBS}, {4, 1, D_CLUSTERS},
    {1, 1, D_JSON},
};
constexpr int kNumCols = sizeof(kCols) / sizeof(kCols[0]);
constexpr int kFirstPodCol = 32;  // p_ns_id ... p_replica_name_id are columns 32..38
static_assert(kCols[kFirstPodCol].dim == D_PODS && kCols[kFirstPodCol - 1].dim != D_PODS && kCols[kFirstPodCol + 6].dim == D_PODS && kCols[kFirstPodCol + 7].dim != D_PODS,
              "kFirstPodCol must point at the seven per-pod columns");
static_assert(sizeof(kr_snapshot_bufs) == kNumCols * sizeof(void *), "kCols must mirror kr_snapshot_bufs");
static_assert(sizeof(SnapDev) == kNumCols * sizeof(void *), "SnapDev must mirror kr_snapshot_bufs");
static_assert(sizeof(kr_cluster_result) == 96 && sizeof(kr_group_result) == 32 && sizeof(kr_job_result) == 8, "result record sizes");

// how a changed row of each object column is treated by the on-device diff of an object commit (kr_incr.cuh, KR_OC_*)
static const uint8_t kObjClass[kNumCols] = {
        KR_OC_STRUCT, KR_OC_STRUCT, KR_OC_COPY, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_STRUCT, KR_OC_STRUCT, KR_OC_COPY, KR_OC_COPY,
        KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER, KR_OC_CLUSTER,
        KR_OC_STRUCT, KR_OC_STRUCT, KR_OC_GROUP, KR_OC_GROUP, KR_OC_GROUP, KR_OC_STRUCT, KR_OC_GROUP, KR_OC_STRUCT, KR_OC_STRUCT,
        KR_OC_STRUCT,
        0, 0, 0, 0, 0, 0, 0,
        KR_OC_HEADKEY, KR_OC_HEAD, KR_OC_HEAD, KR_OC_HEAD, KR_OC_HEAD, KR_OC_HEAD, KR_OC_HEAD, KR_OC_HEAD,
        KR_OC_COPY, KR_OC_COPY, KR_OC_COPY, KR_OC_COPY,
        0};
constexpr int kHeadKeyCol = 39, kGroupClusterCol = 22;
static_assert(kCols[kHeadKeyCol].dim == D_HEADS && kCols[kHeadKeyCol - 1].dim == D_PODS && kCols[kGroupClusterCol].dim == D_GROUPS && kCols[kGroupClusterCol - 1].dim == D_CLUSTERS, "column indices of the object diff");


void dims_of(const kr_sizes &n, uint64_t d[7]) {
  d[D_CLUSTERS] = n.n_clusters; d[D_GROUPS] = n.n_groups; d[D_WTD] = n.n_wtd; d[D_PODS] = n.n_pods;
  d[D_HEADS] = n.n_heads; d[D_JOBS] = n.n_jobs; d[D_JSON] = n.json_bytes;
}

InLayout in_layout(const kr_sizes &n) {
  InLayout L;
  uint64_t d[7];
  dims_of(n, d);
  size_t o = 0;
  for (int i = 0; i < kNumCols; i++) {
    L.off[i] = o;
    o = align_up(o + (size_t)kCols[i].elem * kCols[i].mult * d[kCols[i].dim]);
  }
  L.total = o;
  return L;
}

struct OutLayout {  // results arena: [small fixed part | full pod lists | variable-length lists]
  size_t totals, clusters, hash, groups, wtd, jobs, act_start, act_cnt, small_total, sorted_idx, sorted_act, act_idx, act_code, create, total;
};
OutLayout out_layout(const kr_sizes &n, uint32_t create_cap) {
  OutLayout L;
  size_t o = 0;
  L.totals = o; o = align_up(o + 256);  // 8 counters + (128 bytes in) the void-attempt word
  L.clusters = o; o = align_up(o + sizeof(kr_cluster_result) * (size_t)n.n_clusters);
  L.hash = o; o = align_up(o + 32 * (size_t)n.n_clusters);
  L.groups = o; o = align_up(o + sizeof(kr_group_result) * (size_t)n.n_groups);
  L.wtd = o; o = align_up(o + 4 * (size_t)n.n_wtd);
  L.jobs = o; o = align_up(o + sizeof(kr_job_result) * (size_t)n.n_jobs);
  L.act_start = o; o = align_up(o + 4 * ((size_t)n.n_clusters + 1));
  L.act_cnt = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.small_total = o;  // everything above comes back in ONE copy
  L.sorted_idx = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.sorted_act = o; o = align_up(o + (size_t)n.n_pods);
  L.act_idx = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.act_code = o; o = align_up(o + (size_t)n.n_pods);
  L.create = o; o = align_up(o + 4 * (size_t)create_cap);
  L.total = o;
  return L;
}

struct ScratchLayout {
  // 0xFF-initialised region first
  size_t cl_slots_off, wt_keys, wt_head, aux_keys, aux_vals, ff_total;
  size_t cl_rec, wt_next, rows, keys0, keys1, vals0, vals1, hist, row_total, gacc, gcreate, deferred_list, cact, ccount, chain, wt_bits, cl_dyn, cstart, tile_orph, mh_rep, mh_name, mh_meta, mh_cnt, mh_flg, mh_act, mh_head, act_tmp_idx, act_tmp_code, cl_in, bucket, inc_zero, stamp, dirty_flag, obj_flag, inc, inc_zero_end, touched, touched_old, pos, dirty_list, act_res, cre_res, total;
  uint32_t cl_slots, wt_slots, aux_slots, ntiles, mtiles;  // radix tiles (2048 keys) / k_match tiles of the fast pipeline
  uint32_t wt_bits_n;      // bits of the workersToDelete Bloom bitmap
  size_t bucket_entries;   // capacity of the bucket arena of the bucket pipeline (0: that pipeline is off for this engine)
};
ScratchLayout scratch_layout(const kr_sizes &n) {
  ScratchLayout L;
  L.cl_slots = pow2_at_least(2ull * n.n_clusters);
  L.wt_slots = pow2_at_least(2ull * n.n_wtd);
  L.aux_slots = pow2_at_least(2ull * n.n_heads);
  L.ntiles = (uint32_t)((n.n_pods + kSortTile - 1) / kSortTile);
  if (L.ntiles == 0) L.ntiles = 1;
  L.mtiles = (uint32_t)((n.n_pods + kMatchTile - 1) / kMatchTile);
  if (L.mtiles == 0) L.mtiles = 1;
  size_t o = 0;
  L.cl_slots_off = o; o = align_up(o + 16 * (size_t)L.cl_slots);
  L.wt_keys = o; o = align_up(o + 8 * (size_t)L.wt_slots);
  L.wt_head = o; o = align_up(o + 4 * (size_t)L.wt_slots);
  L.aux_keys = o; o = align_up(o + 4 * (size_t)L.aux_slots);
  L.aux_vals = o; o = align_up(o + 4 * (size_t)L.aux_slots);
  L.ff_total = o;
  L.cl_rec = o; o = align_up(o + 16 * (size_t)n.n_clusters);
  L.wt_next = o; o = align_up(o + 4 * (size_t)n.n_wtd);
  L.rows = o; o = align_up(o + 16 * (size_t)n.n_pods);
  L.keys0 = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.keys1 = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.vals0 = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.vals1 = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.hist = o; o = align_up(o + 4 * (size_t)kRadix * L.ntiles);
  L.row_total = o; o = align_up(o + 4 * (size_t)kRadix);
  L.gacc = o; o = align_up(o + 16 * (size_t)n.n_groups);
  L.gcreate = o; o = align_up(o + 4 * (size_t)n.n_groups + 32);
  L.deferred_list = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.cact = o; o = align_up(o + 4 * ((size_t)n.n_clusters + 8));
  L.ccount = o; o = align_up(o + 4 * ((size_t)n.n_clusters + 2));
  L.chain = o;  // directly after ccount: k_clear zeroes both as one region
  o = align_up(o + 8 * (((size_t)n.n_clusters + 2) / 8192 + (size_t)L.mtiles / 8192 + (size_t)n.n_groups / 8192 + (size_t)n.n_clusters / 8192 + 10));
  {
    static const uint64_t per_name = [] { const char *g = getenv("KR_BLOOM_BITS"); return g && atoi(g) > 0 ? (uint64_t)atoi(g) : 64ull; }();
    uint64_t bits = 1024;
    while (bits < per_name * n.n_wtd && bits < (1ull << 17)) bits <<= 1;  // 64 bits per name up to 16 KB (shared-memory copy per k_match2 CTA)
    L.wt_bits_n = (uint32_t)bits;
  }
  L.wt_bits = o; o = align_up(o + L.wt_bits_n / 8);   // zeroed with ccount and chain (one region up to cstart)
  L.cl_dyn = o; o = align_up(o + 16 * (size_t)n.n_clusters);
  L.cstart = o; o = align_up(o + 4 * ((size_t)n.n_clusters + 2));
  L.tile_orph = o; o = align_up(o + 4 * ((size_t)L.mtiles + 8));
  L.mh_rep = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.mh_name = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.mh_meta = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.mh_cnt = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.mh_flg = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.mh_act = o; o = align_up(o + (size_t)n.n_pods);
  L.mh_head = o; o = align_up(o + (size_t)n.n_pods);
  L.act_tmp_idx = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.act_tmp_code = o; o = align_up(o + (size_t)n.n_pods);
  // bucket pipeline: fixed-stride buckets of 16-byte records.  Room for a stride of >= 4x the mean cluster size, at least 64
  // records per cluster; snapshots of very many tiny clusters (64 records per cluster would dwarf the pods) do without.
  L.bucket_entries = std::max<size_t>(4 * (size_t)n.n_pods, 64 * (size_t)n.n_clusters);
  L.bucket_entries = std::max<size_t>(L.bucket_entries, std::min<size_t>(256 * (size_t)n.n_clusters, (size_t)4 << 20));  // small snapshots: room for the widest stride
  if (64 * (size_t)n.n_clusters > 8 * (size_t)n.n_pods + (4u << 20)) L.bucket_entries = 0;
  L.cl_in = o; o = align_up(o + 128 * (size_t)n.n_clusters);
  L.bucket = o; o = align_up(o + 16 * L.bucket_entries);
  // incremental epochs (kr_incr.cuh): [stamps | dirty flags | counters] start out zero (one memset when the layout moves)
  L.inc_zero = o;
  L.stamp = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.dirty_flag = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.obj_flag = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.inc = o; o = align_up(o + 64);
  L.inc_zero_end = o;
  L.touched = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.touched_old = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.pos = o; o = align_up(o + 4 * (size_t)n.n_pods);
  L.dirty_list = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.act_res = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.cre_res = o; o = align_up(o + 4 * (size_t)n.n_clusters);
  L.total = o;
  return L;
}

}  // namespace

struct kr_engine {
  kr_config cfg{};
  cudaStream_t sm = nullptr, sh = nullptr, sg = nullptr, scopy = nullptr;
  cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr, ev_cols = nullptr, ev_json = nullptr;  // commit: copy start, columns landed, JSON landed
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr, ev_fork3 = nullptr, ev_join3 = nullptr;
  cudaEvent_t ev_inc = nullptr;  // an incremental pass's counters have reached the host
  cudaEvent_t ev_fork = nullptr, ev_hash = nullptr, ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
  cudaEvent_t ev_k[KR_MAX_KERNEL_TIMES + 1]{};
  uint8_t *h_in = nullptr, *d_in = nullptr, *d_scratch = nullptr, *d_out = nullptr, *h_out = nullptr;
  size_t in_cap = 0, scratch_cap = 0, out_cap = 0;
  kr_sizes sizes{};
  InLayout il{};
  OutLayout ol{};
  ScratchLayout sl{};
  bool begun = false, committed = false, ran = false;
  kr_flags last_flags{};      // flags of the last pass (kr_results_fetch honours fetch_pod_lists)
  bool committed_full = false;  // every part of the current layout has been uploaded at least once
  bool fixed_layout = false;    // KR_OPT_FIXED_LAYOUT: arenas laid out for the capacities, live counts in `sizes`
  uint32_t n_recreate = 0;  // clusters with KR_CF_UPGRADE_RECREATE (decide phase 1 needed)
  kr_profile prof{};
  std::string err;
  // kr_hash_batch staging
  uint8_t *hb_h = nullptr, *hb_d = nullptr;
  size_t hb_cap = 0;
  uint8_t *h_in_dev = nullptr;               // device-side address of h_in
  uint8_t *pr_h = nullptr, *pr_d = nullptr;  // incremental pod commits: staging
  cudaEvent_t ev_pr = nullptr;               // staging buffer consumed by the copy stream
  bool pr_busy = false;
  size_t pr_cap = 0;
  int sm_count = 148;
  // the whole pass (both streams) captured once per (layout, flags, n_recreate) and replayed
  cudaGraphExec_t gexec = nullptr;
  kr_flags gflags{};
  bool gvalid = false;
  bool use_graph = true;
  bool use_pdl = true;        // KR_NO_PDL=1 disables programmatic dependent launch
  int hash_ctas_per_sm = 2;   // KR_HASH_CTAS: resident hash CTAs per SM in the throughput regime
  int place_ctas = 1;         // k_place_fused CTAs per SM (KR_PLACE_CTAS; measured at C3: 1 -> 25 us, 2 -> 35 us next to the hash)
  // pipeline choice: fast = count/place/in-warp sort (every bucket <= 1024 pods); radix = general stable LSD sort.
  bool force_radix = false;   // sticky per layout: set when a pass met a bucket the fast pipeline cannot sort
  bool ran_fast = false;
  bool h2d_timed = true;      // h2d_ms of the last commit has been read back from its events
  bool no_fuse = false;       // KR_NO_FUSE=1: always take the separate scan kernels (tests; large snapshots take them anyway)
  bool env_radix = false;     // KR_FORCE_RADIX=1: always take the general pipeline (tests)
  uint32_t *h_totals = nullptr;  // pinned copy of the device totals (pipeline fallback check)
  // bucket pipeline (kr_bucket2.cuh): taken when the caller does not fetch the full pod lists and the snapshot qualifies
  bool no_bucket = false;       // KR_NO_BUCKET=1: never take it (tests of the sort pipeline)
  bool hash_spin = true;        // bucket pipeline: Recreate gates wait for their digest inside k_decide2 instead of a second decide phase
                                // (KR_NO_HASH_SPIN=1, or a pass in which a warp gave up waiting, turns it off)
  uint64_t recreate_sig = 0;    // which RayClusters carry KR_CF_UPGRADE_RECREATE (their messages lead the hash order)
  std::vector<uint8_t> recreate_bit;  // ... per cluster row, as of the last object commit (kr_snapshot_commit_object_rows checks against it)
  uint8_t *orow_h = nullptr, *orow_d = nullptr; size_t orow_cap = 0;  // kr_snapshot_commit_object_rows staging
  cudaEvent_t ev_orow = nullptr; bool orow_busy = false;
  uint32_t bstride = 0;         // bucket stride of this layout (64 / 128 / 256); 0 = the layout does not qualify (a cluster outgrew 256 pods, ...)
  bool snap_has_mh = false;     // some worker group has numOfHosts > 1
  uint32_t snap_max_groups = 0; // most worker groups in one RayCluster
  bool ran_bucket = false;
  uint64_t h2d_accum = 0;       // bytes uploaded by the commits since the last pass (kr_profile.h2d_bytes)
  // hash order: message ids by descending SHA-1 block count, rebuilt at every commit from c_json_len
  uint32_t *h_order = nullptr, *d_order = nullptr;
  cudaEvent_t ev_order = nullptr;  // the upload of h_order has left the pinned buffer
  bool order_pending = false;
  std::vector<uint32_t> row_stamp;  // kr_snapshot_commit_pod_values: duplicate-row detection (epoch-stamped)
  uint32_t row_epoch = 0;
  // device-side incremental epochs (kr_incr.cuh)
  bool no_incr = false;          // KR_NO_INCR=1: every pass is a full pass (tests)
  bool inc_valid = false;        // the resident buckets / tables / results describe the committed snapshot up to the commits since the last pass
  bool inc_zero_needed = true;   // the stamp / dirty-flag / counter region of this layout has not been zeroed yet
  kr_flags inc_flags{};          // flags of the pass that left the resident state
  uint32_t inc_n_pods = 0, inc_n_heads = 0;  // rows resident at the last pass
  uint32_t res_n_heads = 0;                  // head-aux rows the resident device columns hold (object commits move it)
  std::vector<uint32_t> prev_h_pod_idx;      // ... and their keys: a change means the pod -> head-aux row table must be rebuilt
  bool heads_rebuild = false;
  bool hash_dirty = false;       // spec JSON (or a JSON range) committed since the digests were computed
  bool ran_inc = false;          // the last pass was an incremental one
  bool host_results_stale = false;  // an incremental pass went unfetched: the host copy misses its records, the next fetch copies everything
  bool fetched = true;              // the last pass's results have been copied to the host arena
  uint32_t inc_n_dirty = 0;      // changed RayClusters of the last incremental pass
  bool inc_gathered = false;     // ... and their records sit packed in the staging buffer
  bool inc_hash_ran = false;
  std::vector<uint64_t> prev_json_off; std::vector<uint32_t> prev_json_len;  // JSON ranges the digests were computed from
  uint8_t *d_obj_stage = nullptr; size_t obj_stage_cap = 0;   // KR_PART_OBJECTS uploads land here while the state is resident
  uint8_t *d_inc_stage = nullptr, *h_inc_stage = nullptr; size_t inc_stage_cap = 0; uint32_t inc_stage_clusters = 0, inc_stage_groups = 0;
  uint32_t *h_inc = nullptr;     // pinned copy of the epoch counters (16 words) + the changed-cluster list
  uint32_t *h_changed = nullptr; size_t h_changed_cap = 0;
};

namespace {

int fail(kr_engine *e, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return code;
}

#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) return fail(e, KR_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

kr_sizes cap_sizes(const kr_config &c) {
  kr_sizes n;
  n.n_clusters = c.max_clusters; n.n_groups = c.max_groups; n.n_wtd = c.max_wtd; n.n_pods = c.max_pods;
  n.n_heads = c.max_heads; n.n_jobs = c.max_jobs; n.json_bytes = c.max_json_bytes;
  return n;
}

void bind_in(const InLayout &L, uint8_t *base, void *struct_of_ptrs) {
  void **p = reinterpret_cast<void **>(struct_of_ptrs);
  for (int i = 0; i < kNumCols; i++) p[i] = base + L.off[i];
}

ResDev bind_out(const OutLayout &L, uint8_t *base) {
  ResDev r;
  r.totals = reinterpret_cast<uint32_t *>(base + L.totals);
  r.clusters = reinterpret_cast<kr_cluster_result *>(base + L.clusters);
  r.hash = reinterpret_cast<char *>(base + L.hash);
  r.groups = reinterpret_cast<kr_group_result *>(base + L.groups);
  r.wtd_pod_idx = reinterpret_cast<uint32_t *>(base + L.wtd);
  r.sorted_pod_idx = reinterpret_cast<uint32_t *>(base + L.sorted_idx);
  r.sorted_action = base + L.sorted_act;
  r.jobs = reinterpret_cast<kr_job_result *>(base + L.jobs);
  r.act_start = reinterpret_cast<uint32_t *>(base + L.act_start);
  r.act_cnt = reinterpret_cast<uint32_t *>(base + L.act_cnt);
  r.act_pod_idx = reinterpret_cast<uint32_t *>(base + L.act_idx);
  r.act_code = base + L.act_code;
  r.create_idx = reinterpret_cast<int32_t *>(base + L.create);
  return r;
}

ScratchDev bind_scratch(const ScratchLayout &L, uint8_t *b) {
  ScratchDev s;
  s.cl_slots = reinterpret_cast<uint4 *>(b + L.cl_slots_off); s.cl_mask = L.cl_slots - 1;
  s.cl_rec = reinterpret_cast<uint4 *>(b + L.cl_rec);
  s.wt_keys = reinterpret_cast<uint64_t *>(b + L.wt_keys); s.wt_head = reinterpret_cast<uint32_t *>(b + L.wt_head);
  s.wt_next = reinterpret_cast<uint32_t *>(b + L.wt_next); s.wt_mask = L.wt_slots - 1;
  s.aux_keys = reinterpret_cast<uint32_t *>(b + L.aux_keys); s.aux_vals = reinterpret_cast<uint32_t *>(b + L.aux_vals); s.aux_mask = L.aux_slots - 1;
  s.rows = reinterpret_cast<uint4 *>(b + L.rows);
  s.keys[0] = reinterpret_cast<uint32_t *>(b + L.keys0); s.keys[1] = reinterpret_cast<uint32_t *>(b + L.keys1);
  s.vals[0] = reinterpret_cast<uint32_t *>(b + L.vals0); s.vals[1] = reinterpret_cast<uint32_t *>(b + L.vals1);
  s.hist = reinterpret_cast<uint32_t *>(b + L.hist);
  s.row_total = reinterpret_cast<uint32_t *>(b + L.row_total);
  s.gacc = reinterpret_cast<int32_t *>(b + L.gacc);
  s.gcreate = reinterpret_cast<uint32_t *>(b + L.gcreate);
  s.deferred_list = reinterpret_cast<uint32_t *>(b + L.deferred_list);
  s.cact = reinterpret_cast<uint32_t *>(b + L.cact);
  s.ccount = reinterpret_cast<uint32_t *>(b + L.ccount);
  s.chain = reinterpret_cast<uint32_t *>(b + L.chain);
  s.cstart = reinterpret_cast<uint32_t *>(b + L.cstart);
  s.tile_orph = reinterpret_cast<uint32_t *>(b + L.tile_orph);
  s.mh_rep = reinterpret_cast<uint32_t *>(b + L.mh_rep); s.mh_name = reinterpret_cast<uint32_t *>(b + L.mh_name);
  s.mh_meta = reinterpret_cast<uint32_t *>(b + L.mh_meta); s.mh_cnt = reinterpret_cast<uint32_t *>(b + L.mh_cnt);
  s.mh_flg = reinterpret_cast<uint32_t *>(b + L.mh_flg); s.mh_act = b + L.mh_act; s.mh_head = b + L.mh_head;
  s.act_tmp_idx = reinterpret_cast<uint32_t *>(b + L.act_tmp_idx); s.act_tmp_code = b + L.act_tmp_code;
  s.bucket = reinterpret_cast<uint4 *>(b + L.bucket); s.bucket_stride = 0;
  s.wt_bits = reinterpret_cast<uint32_t *>(b + L.wt_bits); s.wt_bits_mask = L.wt_bits_n - 1;
  s.cl_in = reinterpret_cast<uint32_t *>(b + L.cl_in);
  s.cl_dyn = reinterpret_cast<uint4 *>(b + L.cl_dyn);
  s.stamp = reinterpret_cast<uint32_t *>(b + L.stamp); s.touched = reinterpret_cast<uint32_t *>(b + L.touched);
  s.touched_old = reinterpret_cast<uint32_t *>(b + L.touched_old); s.pos = reinterpret_cast<uint32_t *>(b + L.pos);
  s.obj_flag = reinterpret_cast<uint32_t *>(b + L.obj_flag);
  s.dirty_flag = reinterpret_cast<uint32_t *>(b + L.dirty_flag); s.dirty_list = reinterpret_cast<uint32_t *>(b + L.dirty_list);
  s.act_res = reinterpret_cast<uint32_t *>(b + L.act_res); s.cre_res = reinterpret_cast<uint32_t *>(b + L.cre_res);
  s.inc = reinterpret_cast<uint32_t *>(b + L.inc);
  return s;
}

// Kernel launch with the programmatic-dependent-launch attribute (see pdl_wait / pdl_trigger in kr_common.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Launches the whole pass.  profile: serialise everything on stream M and bracket each kernel with events.
int launch_pass(kr_engine *e, const kr_flags &f, bool profile, bool capturing = false) {
  const kr_sizes &n = e->sizes;
  SnapDev s;
  bind_in(e->il, e->d_in, &s);
  ResDev r = bind_out(e->ol, e->d_out);
  ScratchDev sc = bind_scratch(e->sl, e->d_scratch);
  Sizes z{n.n_clusters, n.n_groups, n.n_wtd, n.n_pods, n.n_heads, n.n_jobs};
  cudaStream_t M = e->sm, H = profile ? e->sm : e->sh;
  int k = 0;
  auto mark = [&](const char *name) {
    if (profile && k < KR_MAX_KERNEL_TIMES) { e->prof.kernel_name[k] = name; cudaEventRecord(e->ev_k[k], M); }
    k++;
  };
  e->prof.n_kernels = 0;

  // --- stream H: hash (only needs the committed snapshot)
  const bool do_hash = !f.skip_hash && n.n_clusters > 0;
  // the committed snapshot: columns gate stream M, the JSON arena gates the hash
  const unsigned wflag = capturing ? cudaEventWaitExternal : cudaEventWaitDefault;
  // (the fork comes first so the hash can start while the columns are still landing — an incremental pod-row epoch leaves the JSON untouched)
  // bucket pipeline (kr_bucket2.cuh): the caller does not fetch the full pod lists, no multi-host group is in play, every
  // RayCluster has few worker groups and (checked on the device) at most `bstride` pods
  const bool bucket = !e->no_bucket && !f.fetch_pod_lists && e->bstride != 0 && !e->force_radix && e->snap_max_groups <= KR_SMEM_GROUPS &&
                      !(e->snap_has_mh && f.gate_multihost_indexing) && (size_t)n.n_clusters * e->bstride <= e->sl.bucket_entries;
  // ... and there the clusters whose Recreate gate reads a digest wait for it inside the decide kernel (the hash runs beside it)
  const bool spin = bucket && !profile && do_hash && e->hash_spin && e->n_recreate > 0;
  auto launch_hash = [&]() {
    // messages are taken in e->d_order (descending SHA-1 block count, built at commit): length-homogeneous warps, longest first
    const uint32_t ngroups = (n.n_clusters + 31) / 32;
    if (ngroups <= (uint32_t)e->sm_count * 4) {
      // latency regime (C3: 313 groups): the pass waits for the longest message's serial chain — warp-specialised pairs,
      // two per SM so that every hash warp has a scheduler to itself
      const uint32_t G = std::min<uint32_t>(ngroups, (uint32_t)e->sm_count * 2);
      k_hash3<1, 0><<<G, 64, sizeof(H3Smem), H>>>(s.json, s.c_json_off, s.c_json_len, e->d_order, n.n_clusters, r.hash);
    } else {
      // throughput regime: resident CTAs of four one-lane-per-message warps walk the groups, round adds on the FMA pipe
      // (tools/hash_bench, 100 k messages: 1.58 TB/s against 1.40 TB/s for the pairs)
      uint32_t blocks = std::min<uint32_t>((n.n_clusters + 127) / 128, (uint32_t)e->sm_count * e->hash_ctas_per_sm);
      k_hash2<4, 1><<<blocks, 128, 0, H>>>(s.json, s.c_json_off, s.c_json_len, e->d_order, n.n_clusters, r.hash, 1u);
    }
  };
  auto start_hash_stream = [&]() -> int {
    if (profile) return KR_OK;
    CK(cudaEventRecord(e->ev_fork, M)); CK(cudaStreamWaitEvent(H, e->ev_fork, 0));
    CK(cudaStreamWaitEvent(H, e->ev_json, wflag));
    if (do_hash && spin && n.n_clusters) CK(cudaMemsetAsync(r.hash, 0, 32 * (size_t)n.n_clusters, H));  // the digests' last words are "ready" marks (k_decide2)
    if (do_hash) { launch_hash(); k++; }  // (counted among the pass's kernels: kr_profile.n_kernels)
    else if (n.n_clusters) CK(cudaMemsetAsync(r.hash, 0, 32 * (size_t)n.n_clusters, H));
    CK(cudaEventRecord(e->ev_hash, H));
    return KR_OK;
  };
  // The hash goes first: its 313 one-warp CTAs must be resident before the main chain fills the SMs (launched after
  // k_build_tables instead, they queue behind the chain's blocks and the hash takes 250 us instead of 100).
  { int rc = start_hash_stream(); if (rc) return rc; }

  // --- stream M
  e->ran_bucket = bucket;
  sc.bucket_stride = e->bstride;
  const bool pdl = !profile && e->use_pdl;
  bool fuse_place_done = false;    // k_decide_small directly follows k_place_fused on stream M
  bool creates_after_kernel = false;  // k_creates_fused directly follows a kernel on stream M (no event wait in between)
  {
    ClearArgs ca{};
    ca.ptr[0] = reinterpret_cast<uint32_t *>(e->d_scratch); ca.words[0] = (uint32_t)(e->sl.ff_total / 4); ca.value[0] = 0xFFFFFFFFu;
    ca.ptr[1] = r.wtd_pod_idx; ca.words[1] = n.n_wtd; ca.value[1] = 0xFFFFFFFFu;
    ca.ptr[2] = r.totals; ca.words[2] = 64; ca.value[2] = 0;  // the 8 counters and, 128 bytes in, the void-attempt word (the block is 256 bytes)
    // per-cluster counts + the chained-scan cells + the workersToDelete Bloom bitmap + the bucket fill counters / first-head
    // cells (one region)
    ca.ptr[3] = sc.ccount; ca.words[3] = (uint32_t)((e->sl.cstart - e->sl.ccount) / 4); ca.value[3] = 0;
    mark("k_clear");
    k_clear<<<e->sm_count * 2, 256, 0, M>>>(ca);
  }
  CK(cudaStreamWaitEvent(M, e->ev_cols, wflag));  // the scratch clears above overlap the tail of the upload
  {
    uint32_t items = n.n_clusters + n.n_groups + n.n_heads;
    if (items) { mark("k_build_tables"); k_build_tables<<<(items + 255) / 256, 256, 0, M>>>(s, sc, r, z); }
  }
  if (bucket) {
    e->ran_fast = false;
    const uint32_t mtiles = e->sl.mtiles;
    if (n.n_pods) {
      mark("k_match2");
      CK(launch_pdl(k_match2<kMatchItems>, dim3(mtiles), dim3(kSortThreads), n.n_wtd ? e->sl.wt_bits_n / 8 : 0, M, pdl, s, sc, r, z, n.n_wtd ? 1 : 0));
    }
    Decide2Args da{s, sc, r, z, f, IncStage{}, e->cfg.max_creates, spin ? 1 : 0, 0};
    auto launch_decide2 = [&](dim3 grid, bool with_pdl) -> cudaError_t {
      if (e->bstride <= 64) return launch_pdl(k_decide2<2>, grid, dim3(kD2Warps * 32), 0, M, with_pdl, da);
      if (e->bstride <= 128) return launch_pdl(k_decide2<4>, grid, dim3(kD2Warps * 32), 0, M, with_pdl, da);
      return launch_pdl(k_decide2<8>, grid, dim3(kD2Warps * 32), 0, M, with_pdl, da);
    };
    if (n.n_clusters) {
      mark("k_decide2");
      CK(launch_decide2(dim3((n.n_clusters + kD2Warps - 1) / kD2Warps), pdl && n.n_pods != 0));
    } else CK(cudaMemsetAsync(r.act_start, 0, 4, M));
    if (n.n_jobs) { mark("k_jobs"); k_jobs<<<(n.n_jobs + 255) / 256, 256, 0, M>>>(s, sc, r, z); }
    if (profile) {
      if (do_hash) { mark("k_hash"); launch_hash(); }
      else if (n.n_clusters) CK(cudaMemsetAsync(r.hash, 0, 32 * (size_t)n.n_clusters, M));
    } else {
      CK(cudaStreamWaitEvent(M, e->ev_hash, 0));
    }
    if (e->n_recreate > 0 && do_hash && !spin) {  // clusters whose Recreate gate needs the digest: decided again, in the places phase 0 reserved
      da.phase = 1;
      mark("k_decide2_phase1");
      CK(launch_decide2(dim3((e->n_recreate + kD2Warps - 1) / kD2Warps), false));
    }
  } else {
  const uint32_t ntiles = e->sl.ntiles;
  const bool fast = !e->force_radix;
  e->ran_fast = fast;
  const uint32_t *sorted_keys = sc.keys[0];
  if (n.n_pods && fast) {
    mark("k_match");
    const uint32_t mtiles = e->sl.mtiles;
    CK(launch_pdl(k_match<true, kMatchItems>, dim3(mtiles), dim3(kSortThreads), 0, M, pdl, s, sc, r, z, n.n_wtd ? 1 : 0));
    const bool fuse_place = !e->no_fuse && (uint64_t)n.n_clusters + 2 + mtiles <= kFusedMaxCounters;
    if (fuse_place) {
      fuse_place_done = true;
      mark("k_place_fused");
      size_t smem = 4 * ((size_t)n.n_clusters + 2 + mtiles);
      CK(launch_pdl(k_place_fused, dim3(e->sm_count * e->place_ctas), dim3(1024), smem, M, pdl, (const uint32_t *)sc.keys[0], (const uint32_t *)sc.keys[1], (const uint32_t *)sc.ccount, sc.cstart,
                    (const uint32_t *)sc.tile_orph, sc.vals[0], n.n_pods, n.n_clusters, mtiles, r.totals));
    } else {
    mark("k_scan_counts");
    const uint32_t nch_c = (n.n_clusters + 1 + kScanChunk - 1) / kScanChunk, nch_t = (mtiles + kScanChunk - 1) / kScanChunk;
    k_scan_counts<<<nch_c + nch_t, 1024, 0, M>>>(sc.ccount, sc.cstart, n.n_clusters + 1, nch_c, sc.tile_orph, mtiles, sc.chain, r.totals);
    mark("k_place");
    k_place<<<(n.n_pods + 1023) / 1024, 256, 0, M>>>(sc.keys[0], sc.keys[1], sc.cstart, sc.tile_orph, sc.vals[0], n.n_pods, n.n_clusters);
    }
  } else if (n.n_pods) {
    uint32_t bits = 1;
    while ((1ull << bits) <= n.n_clusters) bits++;  // keys are in [0, n_clusters]
    const int passes = (int)((bits + kRadixBits - 1) / kRadixBits);
    mark("k_match");
    k_match<false, kSortItems><<<ntiles, kSortThreads, 0, M>>>(s, sc, r, z, n.n_wtd ? 1 : 0);
    int cur = 0;
    for (int p = 0; p < passes; p++) {
      if (p > 0) { mark("k_hist"); k_hist<<<ntiles, kSortThreads, 0, M>>>(sc.keys[cur], sc.hist, n.n_pods, p * kRadixBits); }
      mark("k_scan_rows");
      k_scan_rows<<<kRadix, kRowScanThreads, 0, M>>>(sc.hist, sc.row_total, ntiles);
      mark("k_scatter");
      uint32_t *vout = (p == passes - 1) ? r.sorted_pod_idx : sc.vals[cur ^ 1];
      k_scatter<<<ntiles, kSortThreads, 0, M>>>(sc.keys[cur], sc.vals[cur], sc.keys[cur ^ 1], vout, sc.hist, sc.row_total, n.n_pods, p * kRadixBits, p == 0);
      cur ^= 1;
    }
    sorted_keys = sc.keys[cur];
  } else if (fast) {
    CK(cudaMemsetAsync(sc.cstart, 0, 4 * ((size_t)n.n_clusters + 2), M));
  }
  DecideArgs da{s, sc, r, z, f, sorted_keys, sc.vals[0], fast ? 1 : 0, 0};
  {
    // fast pipeline: k_decide_small (buckets kept in registers) and the general k_decide run side by side
    cudaStream_t G2 = (fast && !profile) ? e->sg : M;
    if (fast && !profile) { CK(cudaEventRecord(e->ev_fork2, M)); CK(cudaStreamWaitEvent(G2, e->ev_fork2, 0)); }
    uint32_t warps = n.n_clusters + 1;
    mark("k_decide");
    k_decide<<<(warps + kDecideWarps - 1) / kDecideWarps, kDecideWarps * 32, 0, G2>>>(da);
    if (fast && n.n_clusters) {
      mark("k_decide_small");
      CK(launch_pdl(k_decide_small, dim3((n.n_clusters + kDecideWarps - 1) / kDecideWarps), dim3(kDecideWarps * 32), 0, M, pdl && fuse_place_done, da));
    }
    if (fast && !profile) { CK(cudaEventRecord(e->ev_join2, G2)); CK(cudaStreamWaitEvent(M, e->ev_join2, 0)); }
  }
  if (n.n_jobs) { mark("k_jobs"); k_jobs<<<(n.n_jobs + 255) / 256, 256, 0, M>>>(s, sc, r, z); }
  if (profile) {
    if (do_hash) { mark("k_hash"); launch_hash(); }
    else if (n.n_clusters) CK(cudaMemsetAsync(r.hash, 0, 32 * (size_t)n.n_clusters, M));
  } else {
    CK(cudaStreamWaitEvent(M, e->ev_hash, 0));
  }
  if (e->n_recreate > 0 && do_hash) {
    da.phase = 1;
    uint32_t warps = e->n_recreate;  // upper bound on the deferred list
    const dim3 grid1((warps + kDecideWarps - 1) / kDecideWarps), block1(kDecideWarps * 32);
    // like phase 0: the register-resident kernel on M for the small clusters, the general one beside it on G for the rest
    cudaStream_t G1 = (fast && !profile) ? e->sg : M;
    if (fast && !profile) { CK(cudaEventRecord(e->ev_fork3, M)); CK(cudaStreamWaitEvent(G1, e->ev_fork3, 0)); }
    mark("k_decide_phase1");
    k_decide<<<grid1, block1, 0, G1>>>(da);
    if (fast) { mark("k_decide_small_phase1"); k_decide_small<<<grid1, block1, 0, M>>>(da); }
    if (fast && !profile) { CK(cudaEventRecord(e->ev_join3, G1)); CK(cudaStreamWaitEvent(M, e->ev_join3, 0)); }
    creates_after_kernel = !(fast && !profile);
  }
  if (!e->no_fuse && (uint64_t)n.n_groups + n.n_clusters + 1 <= kFusedMaxCounters) {
    mark("k_creates_fused");
    CK(launch_pdl(k_creates_fused, dim3(e->sm_count), dim3(1024), 4 * ((size_t)n.n_groups + n.n_clusters + 1), M, pdl && creates_after_kernel, s, sc, r, z, f, e->cfg.max_creates));
  } else {
    const uint32_t nch_a = (n.n_clusters + kScanChunk - 1) / kScanChunk;
    uint32_t *achain = sc.chain + 2 * ((size_t)(n.n_clusters + 1 + kScanChunk - 1) / kScanChunk + (e->sl.mtiles + kScanChunk - 1) / kScanChunk + (n.n_groups + kScanChunk - 1) / kScanChunk + 1);
    if (n.n_clusters) {
      if (e->force_radix) CK(cudaMemsetAsync(achain, 0, 8 * (size_t)nch_a, M));
      mark("k_scan_actions");
      k_scan_actions<<<nch_a, 1024, 0, M>>>(r, sc.cact, n.n_clusters, achain);
      mark("k_compact_actions");
      k_compact_actions<<<(n.n_clusters + 3) / 4, 128, 0, M>>>(r, sc, n.n_clusters);
    } else CK(cudaMemsetAsync(r.act_start, 0, 4, M));
  }
  if (e->no_fuse || (uint64_t)n.n_groups + n.n_clusters + 1 > kFusedMaxCounters) if (n.n_groups) {
    mark("k_scan_creates");
    const uint32_t nch_g = (n.n_groups + kScanChunk - 1) / kScanChunk;
    uint32_t *gchain = sc.chain + 2 * ((size_t)(n.n_clusters + 1 + kScanChunk - 1) / kScanChunk + (e->sl.mtiles + kScanChunk - 1) / kScanChunk);
    if (e->force_radix) CK(cudaMemsetAsync(gchain, 0, 8 * (size_t)nch_g, M));  // (the fast pipeline cleared the cells together with ccount)
    k_scan_creates<<<nch_g, 1024, 0, M>>>(r, sc.gcreate, n.n_groups, gchain);
    mark("k_create_fill");
    k_create_fill<<<(n.n_groups + 3) / 4, 128, 0, M>>>(s, sc, r, z, f, e->cfg.max_creates);
  }
  }  // sort / radix pipelines
  if (profile && k <= KR_MAX_KERNEL_TIMES) cudaEventRecord(e->ev_k[k < KR_MAX_KERNEL_TIMES ? k : KR_MAX_KERNEL_TIMES], M);
  e->prof.n_kernels = (uint32_t)k;
  CK(cudaGetLastError());
  return KR_OK;
}

// Replays the captured CUDA graph of the pass (captures it first when the layout / flags changed).
int run_pass_once(kr_engine *e, const kr_flags &f) {
  if (!e->use_graph) return launch_pass(e, f, false);
  kr_flags fk = f;  // (fetch_pod_lists selects the pipeline: part of the key)
  if (!e->gvalid || memcmp(&e->gflags, &fk, sizeof fk) != 0) {
    e->gvalid = false;
    CK(cudaStreamBeginCapture(e->sm, cudaStreamCaptureModeThreadLocal));
    int rc = launch_pass(e, f, false, true);
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(e->sm, &g);
    if (rc != KR_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
    if (ce != cudaSuccess) return fail(e, KR_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    // New row counts / pointers with the same kernel chain are a parameter update of the instantiated graph (cheap);
    // a different chain (fast <-> radix, fused <-> unfused, phase 1 or RayJobs appearing) is instantiated afresh.
    bool updated = false;
    if (e->gexec) {
      cudaGraphExecUpdateResultInfo info;
      updated = cudaGraphExecUpdate(e->gexec, g, &info) == cudaSuccess;
      if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
    }
    if (!updated) ce = cudaGraphInstantiate(&e->gexec, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(e, KR_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
    e->gflags = fk;
    e->gvalid = true;
  }
#ifdef KR_TIMELINE
  {
    unsigned long long init[64];
    for (int i = 0; i < 64; i++) init[i] = (i & 1) ? 0ull : ~0ull;
    CK(cudaMemcpyToSymbolAsync(g_tl, init, sizeof init, 0, cudaMemcpyHostToDevice, e->sm));
  }
#endif
  CK(cudaGraphLaunch(e->gexec, e->sm));
  return KR_OK;
}


// a bucket-pipeline pass leaves everything an incremental epoch needs on the device
void after_full_pass(kr_engine *e, const kr_flags &f) {
  e->inc_valid = e->ran_bucket && !e->no_incr;
  e->inc_flags = f; e->inc_n_pods = e->sizes.n_pods; e->inc_n_heads = e->sizes.n_heads;
  e->host_results_stale = false; e->inc_n_dirty = 0; e->fetched = false; e->ran_inc = false; e->heads_rebuild = false;
  if (!f.skip_hash) e->hash_dirty = false;
}

// One incremental pass over the resident state (kr_incr.cuh).  Returns KR_OK with *done_inc = true when its results stand;
// *done_inc = false means the attempt was void (structural object change, bucket / arena overflow) and a full pass must follow.
int run_pass_inc(kr_engine *e, const kr_flags &f, cudaEvent_t done, bool profile, bool *done_inc) {
  *done_inc = false;
  const kr_sizes &n = e->sizes;
  SnapDev s;
  bind_in(e->il, e->d_in, &s);
  ResDev r = bind_out(e->ol, e->d_out);
  ScratchDev sc = bind_scratch(e->sl, e->d_scratch);
  sc.bucket_stride = e->bstride;
  Sizes z{n.n_clusters, n.n_groups, n.n_wtd, n.n_pods, n.n_heads, n.n_jobs};
  cudaStream_t M = e->sm, H = profile ? e->sm : e->sh;
  int k = 0;
  auto mark = [&](const char *name) {
    if (profile && k < KR_MAX_KERNEL_TIMES) { e->prof.kernel_name[k] = name; cudaEventRecord(e->ev_k[k], M); }
    k++;
  };
  e->prof.n_kernels = 0;
  CK(cudaStreamWaitEvent(M, e->ev_cols, 0));
  const bool do_hash = e->hash_dirty && !f.skip_hash && n.n_clusters > 0;
  if (do_hash) {  // the spec JSON was committed again: every digest is recomputed (on its own stream), every Recreate gate re-read
    if (!profile) { CK(cudaEventRecord(e->ev_fork, M)); CK(cudaStreamWaitEvent(H, e->ev_fork, 0)); }
    CK(cudaStreamWaitEvent(H, e->ev_json, 0));
    if (profile) mark("k_hash");
    const uint32_t ngroups = (n.n_clusters + 31) / 32;
    if (ngroups <= (uint32_t)e->sm_count * 4)
      k_hash3<1, 0><<<std::min<uint32_t>(ngroups, (uint32_t)e->sm_count * 2), 64, sizeof(H3Smem), H>>>(s.json, s.c_json_off, s.c_json_len, e->d_order, n.n_clusters, r.hash);
    else
      k_hash2<4, 1><<<std::min<uint32_t>((n.n_clusters + 127) / 128, (uint32_t)e->sm_count * e->hash_ctas_per_sm), 128, 0, H>>>(s.json, s.c_json_off, s.c_json_len, e->d_order, n.n_clusters, r.hash, 1u);
    if (!profile) CK(cudaEventRecord(e->ev_hash, H));
    if (e->n_recreate) { mark("k_inc_mark_recreate"); k_inc_mark_recreate<<<(n.n_clusters + 255) / 256, 256, 0, M>>>(s, sc, z); }
  }
  const int grid = e->sm_count * 2;
  if (e->heads_rebuild) {  // a head Pod came or went since the table was built (the commit compared the keys on the host)
    mark("k_inc_aux_rebuild");
    k_inc_aux_clear<<<std::min<uint32_t>(grid, (e->sl.aux_slots + 255) / 256), 256, 0, M>>>(sc);
    k_inc_aux_insert<<<std::min<uint32_t>(grid, (n.n_heads + 255) / 256 + 1), 256, 0, M>>>(s, sc, z);
  }
  // (k_inc_refresh ran behind the object commits' diff kernels: the input records are current)
  mark("k_inc_admit");
  k_inc_admit<<<grid, 256, 0, M>>>(s, sc, r, z, n.n_wtd ? 1 : 0);
  if (do_hash && !profile) CK(cudaStreamWaitEvent(M, e->ev_hash, 0));
  // staging for the changed records: up to a quarter of the RayClusters (beyond that the whole record arrays are as cheap to move)
  IncStage st{};
  {
    const uint32_t capc = std::max<uint32_t>(64, n.n_clusters / 4), capg = (uint32_t)std::min<uint64_t>((uint64_t)capc * KR_SMEM_GROUPS, (uint64_t)n.n_groups + 1);
    const size_t need = align_up(32 * (size_t)capc) + align_up(sizeof(kr_cluster_result) * (size_t)capc) + sizeof(kr_group_result) * (size_t)capg + 1024;
    if (need > e->inc_stage_cap) {
      if (e->d_inc_stage) cudaFree(e->d_inc_stage);
      if (e->h_inc_stage) cudaFreeHost(e->h_inc_stage);
      e->d_inc_stage = nullptr; e->h_inc_stage = nullptr; e->inc_stage_cap = 0;
      CK(cudaMalloc((void **)&e->d_inc_stage, need));
      CK(cudaHostAlloc((void **)&e->h_inc_stage, need, cudaHostAllocDefault));
      e->inc_stage_cap = need;
    }
    e->inc_stage_clusters = capc; e->inc_stage_groups = capg;
    st.meta = reinterpret_cast<uint32_t *>(e->d_inc_stage);
    st.clusters = reinterpret_cast<kr_cluster_result *>(e->d_inc_stage + align_up(32 * (size_t)capc));
    st.groups = reinterpret_cast<kr_group_result *>(e->d_inc_stage + align_up(32 * (size_t)capc) + align_up(sizeof(kr_cluster_result) * (size_t)capc));
    st.cap_clusters = capc; st.cap_groups = capg;
  }
  if (n.n_clusters) {
    Decide2Args da{s, sc, r, z, f, st, e->cfg.max_creates, 0, 2};
    const dim3 dgrid((n.n_clusters + kD2Warps - 1) / kD2Warps), dblock(kD2Warps * 32);
    mark("k_decide2_dirty");
    if (e->bstride <= 64) k_decide2<2, true><<<dgrid, dblock, 0, M>>>(da);
    else if (e->bstride <= 128) k_decide2<4, true><<<dgrid, dblock, 0, M>>>(da);
    else k_decide2<8, true><<<dgrid, dblock, 0, M>>>(da);
  }
  if (n.n_jobs) { mark("k_jobs"); k_jobs<<<(n.n_jobs + 255) / 256, 256, 0, M>>>(s, sc, r, z); }
  if (profile && k <= KR_MAX_KERNEL_TIMES) cudaEventRecord(e->ev_k[k < KR_MAX_KERNEL_TIMES ? k : KR_MAX_KERNEL_TIMES], M);
  e->prof.n_kernels = (uint32_t)k;
  if (done) CK(cudaEventRecord(done, M));
  CK(cudaMemcpyAsync(e->h_totals, e->d_out + e->ol.totals, 48, cudaMemcpyDeviceToHost, M));
  CK(cudaMemcpyAsync(e->h_inc, sc.inc, 64, cudaMemcpyDeviceToHost, M));
  if (!e->ev_inc) CK(cudaEventCreateWithFlags(&e->ev_inc, cudaEventDisableTiming));
  CK(cudaEventRecord(e->ev_inc, M));
  k_inc_finish<<<1, 32, 0, M>>>(sc);  // (the host does not wait for it: whatever comes next is ordered behind it on this stream)
  CK(cudaGetLastError());
  CK(cudaEventSynchronize(e->ev_inc));
  e->order_pending = false;
  e->h2d_accum = 0;
  if (!e->h2d_timed) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev_h2d0, e->ev_h2d1) == cudaSuccess) e->prof.h2d_ms = ms;
    e->h2d_timed = true;
  }
  e->heads_rebuild = false;  // (rebuilt here, or about to be rebuilt by the full pass)
  if (e->h_inc[KR_INC_VOID] || e->h_inc[KR_INC_STRUCTURAL]) return KR_OK;  // the caller takes the full pass
  if (!e->fetched) e->host_results_stale = true;  // the previous pass's records never reached the host copy
  e->fetched = false;
  e->inc_n_dirty = e->h_inc[KR_INC_DIRTY];
  e->inc_gathered = e->inc_n_dirty <= e->inc_stage_clusters && e->h_inc[KR_INC_GROUPS] <= e->inc_stage_groups;
  e->inc_hash_ran = do_hash;
  if (do_hash) e->hash_dirty = false;
  e->ran_inc = true;
  *done_inc = true;
  return KR_OK;
}

// Runs the pass; if the fast pipeline met a bucket it cannot sort (> 1024 pods in one RayCluster or among the orphans),
// switches this layout to the radix pipeline and runs again.  Leaves the stream synchronised (after an incremental pass only its one-thread epoch-closing kernel may still be in flight: it touches the epoch counters, nothing a reader of the results sees).
int run_pass(kr_engine *e, const kr_flags &f, cudaEvent_t done) {
  e->last_flags = f;
  if (e->inc_valid && !e->no_incr && memcmp(&e->inc_flags, &f, sizeof f) == 0) {
    bool ok = false;
    if (int rc = run_pass_inc(e, f, done, false, &ok)) return rc;
    if (ok) { e->inc_n_pods = e->sizes.n_pods; e->inc_n_heads = e->sizes.n_heads; return KR_OK; }
  }
  e->inc_valid = false; e->ran_inc = false;
  if (e->inc_zero_needed) {  // first pass on this layout: stamps, dirty flags and epoch counters start from zero
    CK(cudaMemsetAsync(e->d_scratch + e->sl.inc_zero, 0, e->sl.inc_zero_end - e->sl.inc_zero, e->sm));
    e->inc_zero_needed = false;
  }
  for (int attempt = 0; attempt < 5; attempt++) {
    int rc = run_pass_once(e, f);
    if (rc) return rc;
    if (done) CK(cudaEventRecord(done, e->sm));
    CK(cudaMemcpyAsync(e->h_totals, e->d_out + e->ol.totals, 48, cudaMemcpyDeviceToHost, e->sm));
    CK(cudaStreamSynchronize(e->sm));
    e->order_pending = false;
    e->h2d_accum = 0;  // (kr_profile.h2d_bytes keeps the sum of the commits that fed this pass)
    if (!e->h2d_timed) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, e->ev_h2d0, e->ev_h2d1) == cudaSuccess) e->prof.h2d_ms = ms;
      e->h2d_timed = true;
    }
    if (e->h_totals[3] & KR_TOTALS_HASH_WAIT) {  // a decide warp gave up waiting for its digest: rerun on the two-phase schedule
      e->hash_spin = false; e->gvalid = false;
      continue;
    }
    if (!(e->h_totals[3] & KR_TOTALS_BIG_BUCKET)) { after_full_pass(e, f); return KR_OK; }
    // some RayCluster outgrew what this pipeline holds per bucket: bucket pipeline -> wider stride -> sort pipeline -> radix pipeline
    if (e->ran_bucket) {
      const uint32_t wider = e->bstride * 2;
      e->bstride = (wider <= 256 && (size_t)e->sizes.n_clusters * wider <= e->sl.bucket_entries) ? wider : 0;
    } else if (e->ran_fast) e->force_radix = true;
    else break;
    e->gvalid = false;
  }
  return fail(e, KR_E_STATE, "internal: radix pipeline flagged a big bucket");
}

// Results back to the pinned host arena.  run_pass already brought the 32-byte totals over, so every copy is issued with its
// exact size up front and the host waits once: [small fixed part] (+ the full pod lists when asked for) + the compact action
// list + the replica-index arena.
int fetch_results(kr_engine *e, kr_results_view *out) {
  const kr_sizes &n = e->sizes;
  const uint32_t *tot = e->h_totals;
  // bucket pipeline: the two arenas can hold reserved-but-unused places (totals[9] / totals[8] are their extents, [6] / [2] the counts)
  const uint32_t n_create = e->ran_bucket ? tot[9] : tot[0], n_actions = e->ran_bucket ? tot[8] : tot[2];
  const bool full = e->last_flags.fetch_pod_lists != 0;
  CK(cudaEventRecord(e->ev_b, e->sm));
  if (n_create > e->cfg.max_creates) {
    CK(cudaEventRecord(e->ev_c, e->sm));
    return fail(e, KR_E_CAPACITY, "pods to create (%u) exceed kr_config.max_creates (%u)", n_create, e->cfg.max_creates);
  }
  const bool inc = e->ran_inc && !e->host_results_stale;
  const bool packed = inc && e->inc_gathered;
  uint64_t bytes = 0;
  const uint32_t nd = e->inc_n_dirty, ngr = inc ? e->h_inc[KR_INC_GROUPS] : 0;
  const size_t st_cl = align_up(32 * (size_t)e->inc_stage_clusters), st_gr = st_cl + align_up(sizeof(kr_cluster_result) * (size_t)e->inc_stage_clusters);
  if (packed) {
    // incremental pass: the changed cluster / group records come back packed (k_inc_gather) and are scattered into the host
    // arena below; the flat arrays an epoch can touch anywhere (name resolutions, RayJob rows, digests when they were
    // recomputed) are small and come back whole
    if (nd) {
      CK(cudaMemcpyAsync(e->h_inc_stage, e->d_inc_stage, 32 * (size_t)nd, cudaMemcpyDeviceToHost, e->sm));
      CK(cudaMemcpyAsync(e->h_inc_stage + st_cl, e->d_inc_stage + st_cl, sizeof(kr_cluster_result) * (size_t)nd, cudaMemcpyDeviceToHost, e->sm));
      if (ngr) CK(cudaMemcpyAsync(e->h_inc_stage + st_gr, e->d_inc_stage + st_gr, sizeof(kr_group_result) * (size_t)ngr, cudaMemcpyDeviceToHost, e->sm));
      bytes += (32 + sizeof(kr_cluster_result)) * (uint64_t)nd + sizeof(kr_group_result) * (uint64_t)ngr;
    }
    CK(cudaMemcpyAsync(e->h_out + e->ol.totals, e->d_out + e->ol.totals, 256, cudaMemcpyDeviceToHost, e->sm));
    if (n.n_wtd) { CK(cudaMemcpyAsync(e->h_out + e->ol.wtd, e->d_out + e->ol.wtd, 4 * (size_t)n.n_wtd, cudaMemcpyDeviceToHost, e->sm)); bytes += 4ull * n.n_wtd; }
    if (n.n_jobs) { CK(cudaMemcpyAsync(e->h_out + e->ol.jobs, e->d_out + e->ol.jobs, sizeof(kr_job_result) * (size_t)n.n_jobs, cudaMemcpyDeviceToHost, e->sm)); bytes += sizeof(kr_job_result) * (uint64_t)n.n_jobs; }
    if (e->inc_hash_ran && n.n_clusters) { CK(cudaMemcpyAsync(e->h_out + e->ol.hash, e->d_out + e->ol.hash, 32 * (size_t)n.n_clusters, cudaMemcpyDeviceToHost, e->sm)); bytes += 32ull * n.n_clusters; }
  } else {
    bytes = e->ol.small_total;
    CK(cudaMemcpyAsync(e->h_out, e->d_out, e->ol.small_total, cudaMemcpyDeviceToHost, e->sm));
  }
  if (inc && nd) {  // the changed-cluster list itself
    if ((size_t)nd > e->h_changed_cap) {
      if (e->h_changed) cudaFreeHost(e->h_changed);
      e->h_changed = nullptr; e->h_changed_cap = 0;
      const size_t cap = std::max<size_t>(1024, (size_t)e->cfg.max_clusters);
      CK(cudaHostAlloc((void **)&e->h_changed, 4 * cap, cudaHostAllocDefault));
      e->h_changed_cap = cap;
    }
    CK(cudaMemcpyAsync(e->h_changed, e->d_scratch + e->sl.dirty_list, 4 * (size_t)nd, cudaMemcpyDeviceToHost, e->sm));
    bytes += 4ull * nd;
  }
  if (full && n.n_pods) {
    if (!e->fixed_layout) {
      size_t span = e->ol.act_idx - e->ol.sorted_idx;  // sorted_pod_idx + sorted_action, contiguous
      CK(cudaMemcpyAsync(e->h_out + e->ol.sorted_idx, e->d_out + e->ol.sorted_idx, span, cudaMemcpyDeviceToHost, e->sm));
      bytes += span;
    } else {  // capacity slack between the two arrays: copy the live prefixes
      CK(cudaMemcpyAsync(e->h_out + e->ol.sorted_idx, e->d_out + e->ol.sorted_idx, 4 * (size_t)n.n_pods, cudaMemcpyDeviceToHost, e->sm));
      CK(cudaMemcpyAsync(e->h_out + e->ol.sorted_act, e->d_out + e->ol.sorted_act, (size_t)n.n_pods, cudaMemcpyDeviceToHost, e->sm));
      bytes += 5 * (size_t)n.n_pods;
    }
  }
  if (n_actions) {
    CK(cudaMemcpyAsync(e->h_out + e->ol.act_idx, e->d_out + e->ol.act_idx, 4 * (size_t)n_actions, cudaMemcpyDeviceToHost, e->sm));
    CK(cudaMemcpyAsync(e->h_out + e->ol.act_code, e->d_out + e->ol.act_code, (size_t)n_actions, cudaMemcpyDeviceToHost, e->sm));
    bytes += 5ull * n_actions;
  }
  if (n_create) {
    CK(cudaMemcpyAsync(e->h_out + e->ol.create, e->d_out + e->ol.create, 4 * (size_t)n_create, cudaMemcpyDeviceToHost, e->sm));
    bytes += 4ull * n_create;
  }
  CK(cudaEventRecord(e->ev_c, e->sm));
  CK(cudaStreamSynchronize(e->sm));
  if (packed && nd) {  // scatter the packed records into the host arena
    ResDev hr = bind_out(e->ol, e->h_out);
    const uint32_t *meta = reinterpret_cast<const uint32_t *>(e->h_inc_stage);
    const kr_cluster_result *scl = reinterpret_cast<const kr_cluster_result *>(e->h_inc_stage + st_cl);
    const kr_group_result *sgr = reinterpret_cast<const kr_group_result *>(e->h_inc_stage + st_gr);
    for (uint32_t i = 0; i < nd; i++) {
      const uint32_t *m = meta + 8 * (size_t)i;
      const uint32_t c = m[0];
      hr.clusters[c] = scl[i]; hr.act_start[c] = m[1]; hr.act_cnt[c] = m[2];
      if (m[4]) memcpy(&hr.groups[m[3]], &sgr[m[5]], sizeof(kr_group_result) * (size_t)m[4]);
    }
  }
  e->fetched = true; e->host_results_stale = false;
  if (out) {
    ResDev hr = bind_out(e->ol, e->h_out);
    out->clusters = hr.clusters; out->hash = hr.hash; out->groups = hr.groups;
    out->wtd_pod_idx = reinterpret_cast<const int32_t *>(hr.wtd_pod_idx);
    out->sorted_pod_idx = full ? hr.sorted_pod_idx : nullptr; out->sorted_action = full ? hr.sorted_action : nullptr;
    out->create_idx = hr.create_idx; out->jobs = hr.jobs;
    out->act_start = hr.act_start; out->act_cnt = hr.act_cnt; out->act_pod_idx = hr.act_pod_idx; out->act_code = hr.act_code;
    out->n_create_total = e->ran_bucket ? tot[6] : n_create; out->n_orphans = tot[1]; out->n_actions = tot[2];
    out->create_extent = n_create; out->act_extent = n_actions;
    out->n_changed = inc ? nd : n.n_clusters;
    out->changed_clusters = (inc && nd) ? e->h_changed : nullptr;
  }
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e->ev_b, e->ev_c) == cudaSuccess) e->prof.d2h_ms = ms;
  e->prof.d2h_bytes = bytes;
  return KR_OK;
}

}  // namespace

// =================================================================================================== C ABI

extern "C" {

#ifdef KR_TIMELINE
// development aid: (first block start, last block end) in ns of %globaltimer for kernel ids 0..31 of the last graph replay
int kr_debug_timeline(kr_engine *e, unsigned long long *out64) {
  CK(cudaStreamSynchronize(e->sm));
  CK(cudaMemcpyFromSymbol(out64, g_tl, 64 * sizeof(unsigned long long)));
  return KR_OK;
}
#endif

int kr_engine_set_option(kr_engine *e, uint32_t option, uint64_t value) {
  if (!e) return KR_E_INVALID;
  if (option == KR_OPT_FIXED_LAYOUT) {
    if (e->begun) return fail(e, KR_E_STATE, "KR_OPT_FIXED_LAYOUT must be set before the first kr_snapshot_begin");
    e->fixed_layout = value != 0;
    return KR_OK;
  }
  if (option == KR_OPT_INCREMENTAL) {
    e->no_incr = value == 0;
    if (e->no_incr) e->inc_valid = false;
    return KR_OK;
  }
  return fail(e, KR_E_INVALID, "unknown option %u", option);
}

int kr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return KR_E_NO_DEVICE; }
  return n;
}

int kr_engine_create(const kr_config *cfg, kr_engine **out) {
  if (!cfg || !out) return KR_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return KR_E_NO_DEVICE; }
  if (cfg->device < 0 || cfg->device >= ndev) return KR_E_INVALID;
  kr_engine *e = new kr_engine();
  e->cfg = *cfg;
  auto bail = [&](int code) { kr_engine_destroy(e); return code; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) return bail(KR_E_CUDA);
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  kr_sizes cap = cap_sizes(*cfg);
  e->in_cap = in_layout(cap).total;
  e->scratch_cap = scratch_layout(cap).total;
  e->out_cap = out_layout(cap, cfg->max_creates).total;
  // Block-scheduling priorities (kept by the captured graph nodes): the short general-decide kernel on G goes ahead of the
  // k_decide_small blocks still queued on M, and both go ahead of the hash, which is never on the critical path of the chain.
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  const int prio_m = prio_greatest < prio_least ? prio_greatest + 1 : prio_greatest;
  if (cudaStreamCreateWithPriority(&e->sm, cudaStreamNonBlocking, prio_m) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaStreamCreateWithPriority(&e->sh, cudaStreamNonBlocking, prio_least) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaStreamCreateWithPriority(&e->sg, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaStreamCreateWithFlags(&e->scopy, cudaStreamNonBlocking) != cudaSuccess) return bail(KR_E_CUDA);
  cudaEventCreate(&e->ev_h2d0); cudaEventCreate(&e->ev_h2d1); cudaEventCreateWithFlags(&e->ev_pr, cudaEventDisableTiming); cudaEventCreate(&e->ev_cols); cudaEventCreate(&e->ev_json);
  cudaEventCreateWithFlags(&e->ev_fork2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_fork3, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_join3, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_hash, cudaEventDisableTiming);
  cudaEventCreate(&e->ev_a); cudaEventCreate(&e->ev_b); cudaEventCreate(&e->ev_c);
  for (auto &ev : e->ev_k) cudaEventCreate(&ev);
  if (cudaHostAlloc((void **)&e->h_in, e->in_cap, cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaHostAlloc((void **)&e->h_out, e->out_cap, cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaMalloc((void **)&e->d_in, e->in_cap) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaMalloc((void **)&e->d_scratch, e->scratch_cap) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaMalloc((void **)&e->d_out, e->out_cap) != cudaSuccess) return bail(KR_E_CUDA);
  cudaMemset(e->d_out, 0, e->out_cap);  // the alignment padding between the result arrays travels with the single D2H copy
  cudaMemset(e->d_scratch, 0, e->scratch_cap);  // bucket records past a cluster's count are loaded speculatively (and masked): never garbage
  cudaFuncSetAttribute(k_place_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (int)kFusedMaxCounters);
  cudaFuncSetAttribute(k_creates_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (int)kFusedMaxCounters);
  {
    // One shared-memory carveout for every kernel of the pass.  An SM can only change its L1 / shared-memory split while it is
    // idle, and the hash kernel keeps every SM busy for the first half of the pass: with per-kernel defaults the main chain
    // either inherits whatever split the previous pass left behind (k_match then runs with a minimal L1: 60 us instead of
    // 38 us) or waits for the hash to drain (split 0 / 25: k_place_fused starts 40-150 us late).  50 % holds the largest
    // shared-memory user (k_creates_fused, 80 KB) and leaves 114 KB of L1 for the table probes.  Measured with
    // tools/timeline.py at C3: pass 190 us (default) -> 166 us (50); KR_CARVEOUT=<percent> overrides.
    int pct = 50;
    if (const char *g = getenv("KR_CARVEOUT")) pct = atoi(g);
    const void *ks[] = {(const void *)k_build_tables, (const void *)k_match<true, kMatchItems>, (const void *)k_place_fused, (const void *)k_decide_small,
                        (const void *)k_decide, (const void *)k_creates_fused, (const void *)k_jobs, (const void *)k_hash2<1, 0>, (const void *)k_hash2<4, 1>,
                        (const void *)k_clear, (const void *)k_match<false, kSortItems>, (const void *)k_hist, (const void *)k_scan_rows, (const void *)k_scatter,
                        (const void *)k_scan_counts, (const void *)k_place, (const void *)k_scan_creates, (const void *)k_create_fill, (const void *)k_scan_actions,
                        (const void *)k_compact_actions, (const void *)k_patch_pods, (const void *)k_patch_pod_values,
                        (const void *)k_match2<kMatchItems>, (const void *)k_decide2<2>, (const void *)k_decide2<4>, (const void *)k_decide2<8>, (const void *)k_hash3<1, 0>,
                        (const void *)k_inc_retire, (const void *)k_inc_objects, (const void *)k_inc_objects_keys, (const void *)k_inc_aux_clear, (const void *)k_inc_aux_insert,
                        (const void *)k_inc_mark_recreate, (const void *)k_decide2<2, true>, (const void *)k_decide2<4, true>, (const void *)k_decide2<8, true>, (const void *)k_inc_refresh, (const void *)k_inc_admit,
                        (const void *)k_inc_finish};
    for (const void *k : ks) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  if (const char *g = getenv("KR_NO_GRAPH")) e->use_graph = !(g[0] == '1');
  if (const char *g = getenv("KR_FORCE_RADIX")) e->env_radix = (g[0] == '1');
  if (const char *g = getenv("KR_NO_FUSE")) e->no_fuse = (g[0] == '1');
  if (const char *g = getenv("KR_NO_PDL")) e->use_pdl = !(g[0] == '1');
  if (const char *g = getenv("KR_HASH_CTAS")) e->hash_ctas_per_sm = atoi(g) > 0 ? atoi(g) : 2;
  if (const char *g = getenv("KR_PLACE_CTAS")) e->place_ctas = atoi(g) > 0 ? atoi(g) : 1;
  e->force_radix = e->env_radix;
  if (cudaHostAlloc((void **)&e->h_totals, 64, cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
  if (const char *g = getenv("KR_NO_BUCKET")) e->no_bucket = (g[0] == '1');
  if (const char *g = getenv("KR_NO_INCR")) e->no_incr = (g[0] == '1');
  if (const char *g = getenv("KR_NO_HASH_SPIN")) e->hash_spin = !(g[0] == '1');
  if (cudaHostAlloc((void **)&e->h_inc, 64, cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
  {  // buffers of the incremental path, sized for the capacities up front (a pinned allocation inside an epoch costs milliseconds)
    const InLayout capl = in_layout(cap);
    const size_t objs = capl.off[kFirstPodCol] + (capl.off[kNumCols - 1] - capl.off[kFirstPodCol + 7]);
    if (cudaMalloc((void **)&e->d_obj_stage, objs) != cudaSuccess) return bail(KR_E_CUDA);
    e->obj_stage_cap = objs;
    const uint32_t capc = std::max<uint32_t>(64, cfg->max_clusters / 4), capg = (uint32_t)std::min<uint64_t>((uint64_t)capc * KR_SMEM_GROUPS, (uint64_t)cfg->max_groups + 1);
    const size_t need = align_up(32 * (size_t)capc) + align_up(sizeof(kr_cluster_result) * (size_t)capc) + sizeof(kr_group_result) * (size_t)capg + 1024;
    if (cudaMalloc((void **)&e->d_inc_stage, need) != cudaSuccess) return bail(KR_E_CUDA);
    if (cudaHostAlloc((void **)&e->h_inc_stage, need, cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
    e->inc_stage_cap = need;
    const size_t chg = std::max<size_t>(1024, (size_t)cfg->max_clusters);
    if (cudaHostAlloc((void **)&e->h_changed, 4 * chg, cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
    e->h_changed_cap = chg;
  }
  if (cudaHostAlloc((void **)&e->h_order, 4 * ((size_t)cfg->max_clusters + 1), cudaHostAllocDefault) != cudaSuccess) return bail(KR_E_CUDA);
  if (cudaMalloc((void **)&e->d_order, 4 * ((size_t)cfg->max_clusters + 1)) != cudaSuccess) return bail(KR_E_CUDA);
  cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming);
  cudaFuncSetAttribute(k_hash3<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(H3Smem));
  *out = e;
  return KR_OK;
}

void kr_engine_destroy(kr_engine *e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->sm) cudaStreamSynchronize(e->sm);
  if (e->sh) cudaStreamSynchronize(e->sh);
  if (e->h_in) cudaFreeHost(e->h_in);
  if (e->h_out) cudaFreeHost(e->h_out);
  if (e->hb_h) cudaFreeHost(e->hb_h);
  if (e->h_totals) cudaFreeHost(e->h_totals);
  if (e->h_order) cudaFreeHost(e->h_order);
  if (e->h_inc) cudaFreeHost(e->h_inc);
  if (e->orow_h) cudaFreeHost(e->orow_h);
  if (e->orow_d) cudaFree(e->orow_d);
  if (e->ev_orow) cudaEventDestroy(e->ev_orow);
  if (e->h_changed) cudaFreeHost(e->h_changed);
  if (e->h_inc_stage) cudaFreeHost(e->h_inc_stage);
  if (e->d_inc_stage) cudaFree(e->d_inc_stage);
  if (e->d_obj_stage) cudaFree(e->d_obj_stage);
  if (e->d_order) cudaFree(e->d_order);
  if (e->ev_order) cudaEventDestroy(e->ev_order);
  if (e->d_in) cudaFree(e->d_in);
  if (e->d_scratch) cudaFree(e->d_scratch);
  if (e->d_out) cudaFree(e->d_out);
  if (e->hb_d) cudaFree(e->hb_d);
  if (e->pr_h) cudaFreeHost(e->pr_h);
  if (e->pr_d) cudaFree(e->pr_d);
  if (e->gexec) cudaGraphExecDestroy(e->gexec);
  for (auto ev : {e->ev_fork, e->ev_hash, e->ev_a, e->ev_b, e->ev_c}) if (ev) cudaEventDestroy(ev);
  for (auto ev : e->ev_k) if (ev) cudaEventDestroy(ev);
  if (e->sm) cudaStreamDestroy(e->sm);
  if (e->sh) cudaStreamDestroy(e->sh);
  if (e->sg) cudaStreamDestroy(e->sg);
  if (e->scopy) { cudaStreamSynchronize(e->scopy); cudaStreamDestroy(e->scopy); }
  for (auto ev : {e->ev_h2d0, e->ev_h2d1, e->ev_cols, e->ev_json, e->ev_pr}) if (ev) cudaEventDestroy(ev);
  if (e->ev_inc) cudaEventDestroy(e->ev_inc);
  if (e->ev_fork2) cudaEventDestroy(e->ev_fork2);
  if (e->ev_join2) cudaEventDestroy(e->ev_join2);
  if (e->ev_fork3) cudaEventDestroy(e->ev_fork3);
  if (e->ev_join3) cudaEventDestroy(e->ev_join3);
  delete e;
}

int kr_snapshot_begin(kr_engine *e, const kr_sizes *sizes, kr_snapshot_bufs *out) {
  if (!e || !sizes || !out) return KR_E_INVALID;
  const kr_config &c = e->cfg;
  if (sizes->n_clusters > c.max_clusters || sizes->n_groups > c.max_groups || sizes->n_wtd > c.max_wtd || sizes->n_pods > c.max_pods ||
      sizes->n_heads > c.max_heads || sizes->n_jobs > c.max_jobs || sizes->json_bytes > c.max_json_bytes)
    return fail(e, KR_E_CAPACITY, "snapshot exceeds the engine capacities given to kr_engine_create");
  if (sizes->n_clusters >= 0xFFFFFFF0u || sizes->n_pods >= 0xFFFFFFF0u) return fail(e, KR_E_CAPACITY, "too many rows");
  CK(cudaSetDevice(c.device));
  CK(cudaStreamSynchronize(e->scopy));
  CK(cudaStreamSynchronize(e->sm));  // previous results are invalidated from here on
  if (memcmp(&e->sizes, sizes, sizeof *sizes) != 0) {  // row counts (and, without KR_OPT_FIXED_LAYOUT, every column address) change
    e->gvalid = false;
    // The resident state of the incremental path survives new live counts under a fixed layout as long as the object tables keep
    // their shape: pod rows appended (they arrive as committed rows), head-aux rows come and go, the JSON arena grows.
    const bool keep = e->inc_valid && e->fixed_layout && sizes->n_clusters == e->sizes.n_clusters && sizes->n_groups == e->sizes.n_groups &&
                      sizes->n_wtd == e->sizes.n_wtd && sizes->n_jobs == e->sizes.n_jobs && sizes->n_pods >= e->sizes.n_pods;
    if (!e->fixed_layout) { e->committed_full = false; e->inc_zero_needed = true; }
    if (!keep) {
      e->inc_valid = false;
      e->force_radix = e->env_radix;
      // bucket stride: a power of two with 25 % head room over the mean cluster size (a cluster that outgrows it voids the
      // attempt; the pass then widens the stride, up to 256, or leaves the bucket pipeline for this layout)
      uint32_t st = 64;
      const uint64_t want = sizes->n_clusters ? ((uint64_t)sizes->n_pods * 5 / 4 + sizes->n_clusters - 1) / sizes->n_clusters : 0;
      while (st < want && st < 512) st <<= 1;
      e->bstride = st <= 256 ? st : 0;
    }
  }
  e->sizes = *sizes;
  const kr_sizes lay = e->fixed_layout ? cap_sizes(c) : *sizes;
  e->il = in_layout(lay);
  e->ol = out_layout(lay, c.max_creates);
  e->sl = scratch_layout(lay);
  if (e->fixed_layout) {  // offsets and table sizes from the capacities, tile counts (grid sizes) from the live pod count
    const ScratchLayout live = scratch_layout(*sizes);
    e->sl.ntiles = live.ntiles; e->sl.mtiles = live.mtiles;
  }
  if (e->il.total > e->in_cap || e->ol.total > e->out_cap || e->sl.total > e->scratch_cap)
    return fail(e, KR_E_CAPACITY, "internal: layout exceeds arena");
  bind_in(e->il, e->h_in, out);
  e->begun = true; e->committed = false; e->ran = false;
  return KR_OK;
}

int kr_snapshot_commit(kr_engine *e) { return kr_snapshot_commit_parts(e, KR_PART_ALL); }

int kr_snapshot_commit_parts(kr_engine *e, uint32_t parts) {
  if (!e || !e->begun) return e ? fail(e, KR_E_STATE, "kr_snapshot_commit before kr_snapshot_begin") : KR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  // cheap host-side checks of the invariants the kernels rely on
  kr_snapshot_bufs hb;
  bind_in(e->il, e->h_in, &hb);
  const kr_sizes &n = e->sizes;
  uint32_t n_recreate = 0, max_groups = 0;
  bool has_mh = false;
  uint64_t goff = 0, woff = 0;
  for (uint32_t c = 0; c < n.n_clusters; c++) {
    if (hb.c_group_off[c] != goff) return fail(e, KR_E_INVALID, "cluster %u: groups must be stored in cluster order (group_off %u != %llu)", c, hb.c_group_off[c], (unsigned long long)goff);
    if (hb.c_group_cnt[c] >= 0xFFFFu) return fail(e, KR_E_CAPACITY, "cluster %u has %u worker groups (limit 65534)", c, hb.c_group_cnt[c]);
    if (goff + hb.c_group_cnt[c] > n.n_groups) return fail(e, KR_E_INVALID, "cluster %u: groups run past n_groups", c);
    // the kernels trust these indices: a shim bug must come back as KR_E_INVALID, not as out-of-bounds device writes
    for (uint64_t g = goff; g < goff + hb.c_group_cnt[c]; g++) {
      if (hb.g_cluster_idx[g] != c) return fail(e, KR_E_INVALID, "group %llu: g_cluster_idx %u != owning cluster %u", (unsigned long long)g, hb.g_cluster_idx[g], c);
      if (hb.g_wtd_off[g] != woff) return fail(e, KR_E_INVALID, "group %llu: workersToDelete names must be stored in group order (wtd_off %u != %llu)", (unsigned long long)g, hb.g_wtd_off[g], (unsigned long long)woff);
      woff += hb.g_wtd_cnt[g];
      if (woff > n.n_wtd) return fail(e, KR_E_INVALID, "group %llu: workersToDelete names run past n_wtd", (unsigned long long)g);
      has_mh |= hb.g_num_hosts[g] > 1;
    }
    max_groups = std::max(max_groups, hb.c_group_cnt[c]);
    goff += hb.c_group_cnt[c];
    if (hb.c_json_off[c] & 15) return fail(e, KR_E_INVALID, "cluster %u: json offset not 16-byte aligned", c);
    if (hb.c_json_off[c] + hb.c_json_len[c] > n.json_bytes) return fail(e, KR_E_INVALID, "cluster %u: json range outside arena", c);
    if (hb.c_flags[c] & KR_CF_UPGRADE_RECREATE) n_recreate++;
  }
  if (parts & KR_PART_COLUMNS) e->inc_valid = false;  // pod columns uploaded wholesale: the resident buckets no longer describe them
  if (parts & KR_PART_JSON) e->hash_dirty = true;
  bool ranges_moved = e->prev_json_off.size() != n.n_clusters;  // some RayCluster's JSON range differs from the one the digests / the hash order were computed from
  for (uint32_t c = 0; c < n.n_clusters && !ranges_moved; c++)
    ranges_moved = e->prev_json_off[c] != hb.c_json_off[c] || e->prev_json_len[c] != hb.c_json_len[c];
  if (goff != n.n_groups) return fail(e, KR_E_INVALID, "sum of group_cnt (%llu) != n_groups (%u)", (unsigned long long)goff, n.n_groups);
  if (woff != n.n_wtd) return fail(e, KR_E_INVALID, "sum of g_wtd_cnt (%llu) != n_wtd (%u)", (unsigned long long)woff, n.n_wtd);
  for (uint32_t h = 0; h < n.n_heads; h++)
    if (hb.h_pod_idx[h] >= n.n_pods) return fail(e, KR_E_INVALID, "head-aux row %u: h_pod_idx %u >= n_pods %u", h, hb.h_pod_idx[h], n.n_pods);
  if (n_recreate != e->n_recreate || has_mh != e->snap_has_mh || (max_groups > KR_SMEM_GROUPS) != (e->snap_max_groups > KR_SMEM_GROUPS)) e->gvalid = false;  // launch shape / pipeline depend on them
  e->n_recreate = n_recreate; e->snap_has_mh = has_mh; e->snap_max_groups = max_groups;
  // hash order: message ids by descending SHA-1 block count (counting sort; the kernels run length-homogeneous warps, longest first)
  if (e->order_pending) { CK(cudaEventSynchronize(e->ev_order)); e->order_pending = false; }  // a previous upload may still be reading h_order
  // The RayClusters whose Recreate gate compares a digest lead the order: their digests are ready when the decide kernel, running
  // beside the hash, gets to them (k_decide2 waits for a digest's last word otherwise).
  uint64_t rsig = 0x9E3779B97F4A7C15ull * (n_recreate + 1);
  for (uint32_t c = 0; c < n.n_clusters; c++) if (hb.c_flags[c] & KR_CF_UPGRADE_RECREATE) rsig = (rsig ^ c) * 0x100000001B3ull;
  const bool order_stale = ranges_moved || rsig != e->recreate_sig;
  if (order_stale) {  // (unchanged lengths and gates: the resident order stands — an object / pod epoch does not pay for it)
    uint32_t maxb = 0;
    for (uint32_t c = 0; c < n.n_clusters; c++) maxb = std::max(maxb, (hb.c_json_len[c] + 8) / 64 + 1);
    auto blocks_of = [&](uint32_t c) { return (hb.c_json_len[c] + 8) / 64 + 1; };
    auto lead = [&](uint32_t c) { return (hb.c_flags[c] & KR_CF_UPGRADE_RECREATE) ? 0u : 1u; };
    if (maxb <= (1u << 20)) {  // counting sort on (not Recreate, descending block count): bucket 0 = the longest Recreate message
      std::vector<uint32_t> start(2 * ((size_t)maxb + 1) + 1, 0);
      auto key = [&](uint32_t c) { return lead(c) * (maxb + 1) + (maxb - blocks_of(c)); };
      for (uint32_t c = 0; c < n.n_clusters; c++) start[key(c) + 1]++;
      for (size_t b = 0; b + 1 < start.size(); b++) start[b + 1] += start[b];
      for (uint32_t c = 0; c < n.n_clusters; c++) e->h_order[start[key(c)]++] = c;
    } else {
      for (uint32_t c = 0; c < n.n_clusters; c++) e->h_order[c] = c;
      std::stable_sort(e->h_order, e->h_order + n.n_clusters, [&](uint32_t a, uint32_t b) { return lead(a) != lead(b) ? lead(a) < lead(b) : blocks_of(a) > blocks_of(b); });
    }
  }
  // Asynchronous, in two parts on the copy stream: every column first, the spec-JSON arena (the larger half) second.
  // The pass waits on the two events, so match/place/decide run while the JSON is still crossing PCIe and only the hash
  // (and what depends on it) waits for the second part.  Nothing here blocks the host.
  CK(cudaStreamSynchronize(e->sm));  // a pass still reading the previous snapshot must finish before it is overwritten
  const size_t json_off = e->il.off[kNumCols - 1];
  if ((parts & KR_PART_ALL) != KR_PART_ALL && !e->committed_full)
    return fail(e, KR_E_STATE, "a partial commit needs a full commit of this layout first");
  size_t bytes = 0;
  CK(cudaEventRecord(e->ev_h2d0, e->scopy));
  const size_t a1 = e->il.off[kFirstPodCol], b0 = e->il.off[kFirstPodCol + 7];
  // While the incremental state is resident, an object commit lands beside the resident tables and is diffed against them on
  // the device (k_inc_objects): changed rows mark their RayCluster dirty, a changed key makes the next pass a full one.
  const bool stage_objects = e->inc_valid && !e->no_incr && (parts & KR_PART_OBJECTS) && !(parts & KR_PART_COLUMNS);
  if (stage_objects && a1 + (json_off - b0) > e->obj_stage_cap) {
    if (e->d_obj_stage) cudaFree(e->d_obj_stage);
    e->d_obj_stage = nullptr; e->obj_stage_cap = 0;
    CK(cudaMalloc((void **)&e->d_obj_stage, a1 + (json_off - b0)));
    e->obj_stage_cap = a1 + (json_off - b0);
  }
  auto stage_of = [&](size_t off) { return off < a1 ? off : a1 + (off - b0); };
  auto up = [&](size_t off, size_t len) -> int {
    if (!len) return KR_OK;
    uint8_t *dst = (stage_objects && off < json_off) ? e->d_obj_stage + stage_of(off) : e->d_in + off;
    CK(cudaMemcpyAsync(dst, e->h_in + off, len, cudaMemcpyHostToDevice, e->scopy));
    bytes += len;
    return KR_OK;
  };
  if ((parts & KR_PART_COLUMNS) && !e->fixed_layout) { if (int rc = up(0, json_off)) return rc; }
  else if (parts & (KR_PART_COLUMNS | KR_PART_OBJECTS)) {
    // the (small) columns on either side of the seven per-pod ones; then, for KR_PART_COLUMNS under a fixed layout, the live
    // prefix of each per-pod column (the capacity slack between the columns is not worth moving)
    if (int rc = up(0, a1)) return rc;
    if (int rc = up(b0, json_off - b0)) return rc;
    if (parts & KR_PART_COLUMNS)
      for (int k = 0; k < 7; k++)
        if (int rc = up(e->il.off[kFirstPodCol + k], 4 * (size_t)n.n_pods)) return rc;
  }
  if (stage_objects) {
    uint64_t dn[7];
    dims_of(n, dn);
    ObjDiffArgs oa{};
    int nc = 0;
    uint32_t first = 0;
    for (int i = 0; i < kNumCols - 1; i++) {
      if (kCols[i].dim == D_PODS) continue;
      oa.src[nc] = e->d_obj_stage + stage_of(e->il.off[i]);
      oa.dst[nc] = e->d_in + e->il.off[i];
      oa.first[nc] = first;
      oa.rows_old[nc] = kCols[i].dim == D_HEADS ? e->res_n_heads : (uint32_t)dn[kCols[i].dim];
      oa.row_bytes[nc] = (uint16_t)(kCols[i].elem * kCols[i].mult);
      oa.cls[nc] = kObjClass[i];
      first += (uint32_t)dn[kCols[i].dim];
      nc++;
    }
    oa.first[nc] = first; oa.n_cols = nc;
    oa.g_cluster_idx_new = reinterpret_cast<const uint32_t *>(e->d_obj_stage + stage_of(e->il.off[kGroupClusterCol]));
    oa.h_pod_idx_new = reinterpret_cast<const uint32_t *>(e->d_obj_stage + stage_of(e->il.off[kHeadKeyCol]));
    oa.h_pod_idx_old = reinterpret_cast<const uint32_t *>(e->d_in + e->il.off[kHeadKeyCol]);
    oa.n_heads_old = e->res_n_heads;
    SnapDev sd;
    bind_in(e->il, e->d_in, &sd);
    ScratchDev scd = bind_scratch(e->sl, e->d_scratch);
    Sizes zz{n.n_clusters, n.n_groups, n.n_wtd, n.n_pods, n.n_heads, n.n_jobs};
    if (first) k_inc_objects<<<(first + 255) / 256, 256, 0, e->scopy>>>(oa, sd, scd, zz);
    if (n.n_heads) k_inc_objects_keys<<<(n.n_heads + 255) / 256, 256, 0, e->scopy>>>(oa.h_pod_idx_new, const_cast<uint32_t *>(sd.h_pod_idx), n.n_heads, nullptr);
    // the input records (cl_in) of the RayClusters the diff found changed, now that every column of theirs is in place
    if (n.n_clusters) k_inc_refresh<<<std::min<uint32_t>((uint32_t)e->sm_count * 2, (n.n_clusters + 255) / 256 + 1), 256, 0, e->scopy>>>(sd, scd);
    CK(cudaGetLastError());
  }
  if (parts & (KR_PART_COLUMNS | KR_PART_OBJECTS)) {
    e->recreate_bit.resize(n.n_clusters);
    for (uint32_t c = 0; c < n.n_clusters; c++) e->recreate_bit[c] = (hb.c_flags[c] & KR_CF_UPGRADE_RECREATE) ? 1 : 0;
    if (e->prev_h_pod_idx.size() != n.n_heads || (n.n_heads && memcmp(e->prev_h_pod_idx.data(), hb.h_pod_idx, 4 * (size_t)n.n_heads) != 0)) {
      e->heads_rebuild = true;
      e->prev_h_pod_idx.assign(hb.h_pod_idx, hb.h_pod_idx + n.n_heads);
    }
    e->res_n_heads = n.n_heads;
  }
  CK(cudaEventRecord(e->ev_cols, e->scopy));
  if (ranges_moved) {  // digests of moved ranges are stale
    e->hash_dirty = true;
    e->prev_json_off.assign(hb.c_json_off, hb.c_json_off + n.n_clusters); e->prev_json_len.assign(hb.c_json_len, hb.c_json_len + n.n_clusters);
  }
  if (order_stale) {  // the new order travels with this commit: from here on the recorded ranges / gates are the ones it was built from
    e->recreate_sig = rsig;
    if (n.n_clusters) {
      CK(cudaMemcpyAsync(e->d_order, e->h_order, 4 * (size_t)n.n_clusters, cudaMemcpyHostToDevice, e->scopy)); bytes += 4 * (size_t)n.n_clusters;
      CK(cudaEventRecord(e->ev_order, e->scopy));
      e->order_pending = true;
    }
  }
  if (parts & KR_PART_JSON) { if (int rc = up(json_off, e->fixed_layout ? (size_t)n.json_bytes : e->il.total - json_off)) return rc; }
  CK(cudaEventRecord(e->ev_json, e->scopy));
  CK(cudaEventRecord(e->ev_h2d1, e->scopy));
  if ((parts & KR_PART_ALL) == KR_PART_ALL) e->committed_full = true;
  e->h2d_timed = false;
  e->prof.h2d_bytes = (e->h2d_accum += bytes);
  e->committed = true;
  return KR_OK;
}

// Shared by the two incremental pod commits: stage the row list (and, journal style, the 7 values per row), upload, scatter.
static int commit_pod_patch(kr_engine *e, const uint32_t *rows, const uint32_t *values, uint32_t n, bool rows_known_distinct = false) {
  if (!e || (!rows && n)) return KR_E_INVALID;
  if (!e->committed_full) return fail(e, KR_E_STATE, "an incremental pod commit needs a full commit of this layout first");
  if (n == 0) return KR_OK;
  CK(cudaSetDevice(e->cfg.device));
  const size_t bytes = (values ? 32 : 4) * (size_t)n;  // row list (+ 7 values per row); kr_snapshot_commit_pod_rows lets the device pull the rows
  if (e->pr_busy) { CK(cudaEventSynchronize(e->ev_pr)); e->pr_busy = false; }  // a previous patch may still be reading the staging buffer
  if (bytes > e->pr_cap) {
    if (e->pr_h) cudaFreeHost(e->pr_h);
    if (e->pr_d) cudaFree(e->pr_d);
    e->pr_h = nullptr; e->pr_d = nullptr; e->pr_cap = 0;
    size_t cap = bytes + bytes / 2 + 4096;
    CK(cudaHostAlloc((void **)&e->pr_h, cap, cudaHostAllocDefault));
    CK(cudaMalloc((void **)&e->pr_d, cap));
    e->pr_cap = cap;
  }
  for (uint32_t i = 0; i < n; i++)
    if (rows[i] >= e->sizes.n_pods) return fail(e, KR_E_INVALID, "pod row %u out of range", rows[i]);
  if (values && !rows_known_distinct) {  // the scatter kernel writes one thread per entry: two entries for one row would race
    if (e->row_stamp.size() < e->sizes.n_pods) e->row_stamp.assign(e->sizes.n_pods, 0);
    if (++e->row_epoch == 0) { std::fill(e->row_stamp.begin(), e->row_stamp.end(), 0u); e->row_epoch = 1; }
    for (uint32_t i = 0; i < n; i++) {
      if (e->row_stamp[rows[i]] == e->row_epoch) return fail(e, KR_E_INVALID, "kr_snapshot_commit_pod_values: pod row %u appears twice", rows[i]);
      e->row_stamp[rows[i]] = e->row_epoch;
    }
  }
  memcpy(e->pr_h, rows, 4 * (size_t)n);
  if (values) memcpy(e->pr_h + 4 * (size_t)n, values, 28 * (size_t)n);
  kr_snapshot_bufs hb;
  bind_in(e->il, e->h_in, &hb);
  SnapDev s;
  bind_in(e->il, e->d_in, &s);
  PodCols hc, dc;
  const void *hsrc[7] = {hb.p_ns_id, hb.p_cluster_name_id, hb.p_group_name_id, hb.p_name_id, hb.p_packed, hb.p_replica_index, hb.p_replica_name_id};
  const void *dsrc[7] = {s.p_ns_id, s.p_cluster_name_id, s.p_group_name_id, s.p_name_id, s.p_packed, s.p_replica_index, s.p_replica_name_id};
  if (!e->h_in_dev) CK(cudaHostGetDevicePointer((void **)&e->h_in_dev, e->h_in, 0));  // device-side address of the pinned arena (mapped under UVA)
  for (int k = 0; k < 7; k++) {
    hc.c[k] = reinterpret_cast<uint32_t *>(e->h_in_dev + (static_cast<const uint8_t *>(hsrc[k]) - e->h_in));
    dc.c[k] = static_cast<uint32_t *>(const_cast<void *>(dsrc[k]));
  }
  CK(cudaStreamSynchronize(e->sm));  // a pass still reading the columns must finish first
  CK(cudaEventRecord(e->ev_h2d0, e->scopy));
  CK(cudaMemcpyAsync(e->pr_d, e->pr_h, bytes, cudaMemcpyHostToDevice, e->scopy));
  CK(cudaEventRecord(e->ev_pr, e->scopy));
  e->pr_busy = true;
  if (e->inc_valid && !e->no_incr) {  // the rows' previous values leave the resident state before the new ones land
    ScratchDev scd = bind_scratch(e->sl, e->d_scratch);
    ResDev rd = bind_out(e->ol, e->d_out);
    Sizes zz{e->sizes.n_clusters, e->sizes.n_groups, e->sizes.n_wtd, e->sizes.n_pods, e->sizes.n_heads, e->sizes.n_jobs};
    k_inc_retire<<<(n + 255) / 256, 256, 0, e->scopy>>>(reinterpret_cast<const uint32_t *>(e->pr_d), n, s, scd, rd, zz, e->inc_n_pods, e->sizes.n_wtd ? 1 : 0);
  }
  if (values) k_patch_pod_values<<<(n + 255) / 256, 256, 0, e->scopy>>>(reinterpret_cast<const uint32_t *>(e->pr_d), n, dc);
  else k_patch_pods<<<(n + 255) / 256, 256, 0, e->scopy>>>(reinterpret_cast<const uint32_t *>(e->pr_d), n, hc, dc);
  CK(cudaGetLastError());
  CK(cudaEventRecord(e->ev_h2d1, e->scopy));
  CK(cudaEventRecord(e->ev_cols, e->scopy));  // ev_json keeps pointing at the last JSON upload: the hash need not wait for the patch
  e->h2d_timed = false;
  e->prof.h2d_bytes = (e->h2d_accum += 32 * (size_t)n);  // row list + the 28-byte row payload (pulled one 32-byte sector per value in the rows-only variant)
  e->committed = true;
  return KR_OK;
}


int kr_snapshot_commit_object_rows(kr_engine *e, const uint32_t *cluster_rows, uint32_t n_cl, const uint32_t *head_rows, uint32_t n_hd) {
  if (!e || (!cluster_rows && n_cl) || (!head_rows && n_hd)) return KR_E_INVALID;
  if (!e->begun) return fail(e, KR_E_STATE, "kr_snapshot_commit_object_rows before kr_snapshot_begin");
  if (n_cl == 0 && n_hd == 0) return KR_OK;
  const kr_sizes &n = e->sizes;
  kr_snapshot_bufs hb;
  bind_in(e->il, e->h_in, &hb);
  // Only an optimisation of kr_snapshot_commit_parts(KR_PART_OBJECTS): whenever the resident state cannot take the rows as they
  // are — no resident state, a Recreate gate or a JSON range that changed (hash order / digests), head rows added or removed —
  // the whole object part is committed instead.
  bool whole = !e->inc_valid || e->no_incr || !e->committed_full || e->res_n_heads != n.n_heads || e->recreate_bit.size() != n.n_clusters ||
               e->prev_json_off.size() != n.n_clusters;
  for (uint32_t i = 0; i < n_cl && !whole; i++) {
    const uint32_t c = cluster_rows[i];
    if (c >= n.n_clusters) return fail(e, KR_E_INVALID, "cluster row %u out of range", c);
    whole = e->recreate_bit[c] != ((hb.c_flags[c] & KR_CF_UPGRADE_RECREATE) ? 1 : 0) || e->prev_json_off[c] != hb.c_json_off[c] || e->prev_json_len[c] != hb.c_json_len[c] ||
            (uint64_t)hb.c_group_off[c] + hb.c_group_cnt[c] > n.n_groups;
  }
  for (uint32_t i = 0; i < n_hd && !whole; i++) {
    if (head_rows[i] >= n.n_heads) return fail(e, KR_E_INVALID, "head-aux row %u out of range", head_rows[i]);
    if (hb.h_pod_idx[head_rows[i]] >= n.n_pods) return fail(e, KR_E_INVALID, "head-aux row %u: h_pod_idx %u >= n_pods %u", head_rows[i], hb.h_pod_idx[head_rows[i]], n.n_pods);
  }
  if (whole) return kr_snapshot_commit_parts(e, KR_PART_OBJECTS);
  CK(cudaSetDevice(e->cfg.device));
  // group rows of the named clusters
  std::vector<uint32_t> grows;
  for (uint32_t i = 0; i < n_cl; i++) for (uint32_t g = 0; g < hb.c_group_cnt[cluster_rows[i]]; g++) grows.push_back(hb.c_group_off[cluster_rows[i]] + g);
  const uint32_t cnt[7] = {n_cl, (uint32_t)grows.size(), 0, 0, n_hd, 0, 0};
  const uint32_t *lists[7] = {cluster_rows, grows.data(), nullptr, nullptr, head_rows, nullptr, nullptr};
  // staging: the three row lists, then every object column's rows packed in list order
  size_t need = 0, list_off[7] = {0};
  for (int d = 0; d < 7; d++) { list_off[d] = need; need = align_up(need + 4 * (size_t)cnt[d]); }
  size_t col_off[kNumCols] = {0};
  for (int i = 0; i < kNumCols - 1; i++) {
    const int d = kCols[i].dim;
    if (d == D_PODS || !cnt[d]) continue;
    col_off[i] = need; need = align_up(need + (size_t)kCols[i].elem * kCols[i].mult * cnt[d]);
  }
  if (e->orow_busy) { CK(cudaEventSynchronize(e->ev_orow)); e->orow_busy = false; }
  if (need > e->orow_cap) {
    if (e->orow_h) cudaFreeHost(e->orow_h);
    if (e->orow_d) cudaFree(e->orow_d);
    e->orow_h = nullptr; e->orow_d = nullptr; e->orow_cap = 0;
    const size_t cap = need + need / 2 + 65536;
    CK(cudaHostAlloc((void **)&e->orow_h, cap, cudaHostAllocDefault));
    CK(cudaMalloc((void **)&e->orow_d, cap));
    e->orow_cap = cap;
  }
  if (!e->ev_orow) CK(cudaEventCreateWithFlags(&e->ev_orow, cudaEventDisableTiming));
  for (int d = 0; d < 7; d++) if (cnt[d]) memcpy(e->orow_h + list_off[d], lists[d], 4 * (size_t)cnt[d]);
  ObjDiffArgs oa{};
  int nc = 0;
  uint32_t first = 0;
  for (int i = 0; i < kNumCols - 1; i++) {
    const int d = kCols[i].dim;
    if (d == D_PODS || !cnt[d]) continue;
    const size_t rb = (size_t)kCols[i].elem * kCols[i].mult;
    const uint8_t *col = e->h_in + e->il.off[i];
    uint8_t *dst = e->orow_h + col_off[i];
    // (row sizes are 1, 4 or 8 bytes for almost every column: fixed-size copies instead of ~10 k variable-length memcpy calls per epoch)
    const uint32_t *rl = lists[d];
    if (rb == 4) { const uint32_t *c4 = reinterpret_cast<const uint32_t *>(col); uint32_t *d4 = reinterpret_cast<uint32_t *>(dst); for (uint32_t k = 0; k < cnt[d]; k++) d4[k] = c4[rl[k]]; }
    else if (rb == 1) { for (uint32_t k = 0; k < cnt[d]; k++) dst[k] = col[rl[k]]; }
    else if (rb == 8) { const uint64_t *c8 = reinterpret_cast<const uint64_t *>(col); uint64_t *d8 = reinterpret_cast<uint64_t *>(dst); for (uint32_t k = 0; k < cnt[d]; k++) d8[k] = c8[rl[k]]; }
    else for (uint32_t k = 0; k < cnt[d]; k++) memcpy(dst + k * rb, col + (size_t)rl[k] * rb, rb);
    oa.src[nc] = e->orow_d + col_off[i];
    oa.rowlist[nc] = reinterpret_cast<const uint32_t *>(e->orow_d + list_off[d]);
    oa.dst[nc] = e->d_in + e->il.off[i];
    oa.first[nc] = first;
    oa.rows_old[nc] = d == D_HEADS ? e->res_n_heads : (d == D_CLUSTERS ? n.n_clusters : n.n_groups);
    oa.row_bytes[nc] = (uint16_t)rb;
    oa.cls[nc] = kObjClass[i];
    if (i == kGroupClusterCol) oa.g_cluster_idx_new = reinterpret_cast<const uint32_t *>(e->orow_d + col_off[i]);
    if (i == kHeadKeyCol) oa.h_pod_idx_new = reinterpret_cast<const uint32_t *>(e->orow_d + col_off[i]);
    first += cnt[d];
    nc++;
  }
  oa.first[nc] = first; oa.n_cols = nc;
  oa.h_pod_idx_old = reinterpret_cast<const uint32_t *>(e->d_in + e->il.off[kHeadKeyCol]);
  oa.n_heads_old = e->res_n_heads;
  // the pod -> head-aux row table follows the keys (compared here, on the host shadow)
  for (uint32_t i = 0; i < n_hd; i++)
    if (e->prev_h_pod_idx[head_rows[i]] != hb.h_pod_idx[head_rows[i]]) { e->heads_rebuild = true; e->prev_h_pod_idx[head_rows[i]] = hb.h_pod_idx[head_rows[i]]; }
  CK(cudaStreamSynchronize(e->sm));  // a pass still reading the tables must finish first
  CK(cudaEventRecord(e->ev_h2d0, e->scopy));
  CK(cudaMemcpyAsync(e->orow_d, e->orow_h, need, cudaMemcpyHostToDevice, e->scopy));
  CK(cudaEventRecord(e->ev_orow, e->scopy));
  e->orow_busy = true;
  SnapDev sd;
  bind_in(e->il, e->d_in, &sd);
  ScratchDev scd = bind_scratch(e->sl, e->d_scratch);
  Sizes zz{n.n_clusters, n.n_groups, n.n_wtd, n.n_pods, n.n_heads, n.n_jobs};
  if (first) k_inc_objects<<<(first + 255) / 256, 256, 0, e->scopy>>>(oa, sd, scd, zz);
  if (n_hd) k_inc_objects_keys<<<(n_hd + 255) / 256, 256, 0, e->scopy>>>(oa.h_pod_idx_new, const_cast<uint32_t *>(sd.h_pod_idx), n_hd, reinterpret_cast<const uint32_t *>(e->orow_d + list_off[D_HEADS]));
  if (n_cl) k_inc_refresh<<<std::min<uint32_t>((uint32_t)e->sm_count * 2, (n.n_clusters + 255) / 256 + 1), 256, 0, e->scopy>>>(sd, scd);  // (see kr_snapshot_commit_parts)
  CK(cudaGetLastError());
  CK(cudaEventRecord(e->ev_h2d1, e->scopy));
  CK(cudaEventRecord(e->ev_cols, e->scopy));
  e->h2d_timed = false;
  e->prof.h2d_bytes = (e->h2d_accum += need);
  e->committed = true;
  return KR_OK;
}

int kr_snapshot_commit_pod_rows(kr_engine *e, const uint32_t *rows, uint32_t n) { return commit_pod_patch(e, rows, nullptr, n); }

// for kr_packer.cpp: its journal holds every row once (row_dirty), so the duplicate scan is skipped
int kr_internal_commit_pod_values_distinct(kr_engine *e, const uint32_t *rows, const uint32_t *values, uint32_t n) {
  if (!values && n) return KR_E_INVALID;
  return commit_pod_patch(e, rows, values, n, true);
}

int kr_snapshot_commit_pod_values(kr_engine *e, const uint32_t *rows, const uint32_t *values, uint32_t n) {
  if (!values && n) return KR_E_INVALID;
  return commit_pod_patch(e, rows, values, n);
}

int kr_reconcile_device_only(kr_engine *e, const kr_flags *flags) {
  if (!e || !flags) return KR_E_INVALID;
  if (!e->committed) return fail(e, KR_E_STATE, "no committed snapshot");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaEventRecord(e->ev_a, e->sm));
  int rc = run_pass(e, *flags, e->ev_b);
  if (rc) return rc;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e->ev_a, e->ev_b) == cudaSuccess) e->prof.kernels_ms = ms;
  e->ran = true;
  return KR_OK;
}

int kr_reconcile_batch(kr_engine *e, const kr_flags *flags, kr_results_view *out) {
  if (!e || !flags || !out) return KR_E_INVALID;
  if (!e->committed) return fail(e, KR_E_STATE, "no committed snapshot");
  CK(cudaSetDevice(e->cfg.device));
  static const bool trace = getenv("KR_ENGINE_TRACE") != nullptr;  // development aid: host time of the two halves of a call (stderr)
  const auto t0 = std::chrono::steady_clock::now();
  CK(cudaEventRecord(e->ev_a, e->sm));
  int rc = run_pass(e, *flags, e->ev_k[KR_MAX_KERNEL_TIMES]);
  if (rc) return rc;
  e->ran = true;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e->ev_a, e->ev_k[KR_MAX_KERNEL_TIMES]) == cudaSuccess) e->prof.kernels_ms = ms;
  const auto t1 = std::chrono::steady_clock::now();
  rc = fetch_results(e, out);  // d2h_ms = ev_b..ev_c
  if (trace) {
    const auto t2 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    fprintf(stderr, "kr_reconcile_batch: pass %.0f us (device %.0f us), fetch %.0f us (device copy %.0f us, %llu bytes)%s\n", us(t0, t1), e->prof.kernels_ms * 1e3, us(t1, t2),
            e->prof.d2h_ms * 1e3, (unsigned long long)e->prof.d2h_bytes, e->ran_inc ? " [incremental]" : "");
  }
  return rc;
}

int kr_reconcile_batch_profiled(kr_engine *e, const kr_flags *flags, kr_profile *prof) {
  if (!e || !flags) return KR_E_INVALID;
  if (!e->committed) return fail(e, KR_E_STATE, "no committed snapshot");
  CK(cudaSetDevice(e->cfg.device));
  e->last_flags = *flags;
  bool inc_done = false;
  if (e->inc_valid && !e->no_incr && memcmp(&e->inc_flags, flags, sizeof *flags) == 0) {
    CK(cudaEventRecord(e->ev_a, e->sm));
    if (int rc = run_pass_inc(e, *flags, e->ev_b, true, &inc_done)) return rc;
    if (inc_done) { e->inc_n_pods = e->sizes.n_pods; e->inc_n_heads = e->sizes.n_heads; }
  }
  if (!inc_done) {
    e->inc_valid = false; e->ran_inc = false;
    if (e->inc_zero_needed) {
      CK(cudaMemsetAsync(e->d_scratch + e->sl.inc_zero, 0, e->sl.inc_zero_end - e->sl.inc_zero, e->sm));
      e->inc_zero_needed = false;
    }
  }
  for (int attempt = 0; !inc_done; attempt++) {  // same fallback ladder as run_pass
    CK(cudaEventRecord(e->ev_a, e->sm));
    int rc = launch_pass(e, *flags, true);
    if (rc) return rc;
    CK(cudaEventRecord(e->ev_b, e->sm));
    CK(cudaMemcpyAsync(e->h_totals, e->d_out + e->ol.totals, 48, cudaMemcpyDeviceToHost, e->sm));
    CK(cudaStreamSynchronize(e->sm));
    e->order_pending = false;
    if (!(e->h_totals[3] & KR_TOTALS_BIG_BUCKET)) { after_full_pass(e, *flags); break; }
    if (attempt >= 4) return fail(e, KR_E_STATE, "internal: radix pipeline flagged a big bucket");
    if (e->ran_bucket) {
      const uint32_t wider = e->bstride * 2;
      e->bstride = (wider <= 256 && (size_t)e->sizes.n_clusters * wider <= e->sl.bucket_entries) ? wider : 0;
    } else if (e->ran_fast) e->force_radix = true;
    e->gvalid = false;
  }
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e->ev_a, e->ev_b) == cudaSuccess) e->prof.kernels_ms = ms;
  uint32_t k = e->prof.n_kernels < KR_MAX_KERNEL_TIMES ? e->prof.n_kernels : KR_MAX_KERNEL_TIMES;
  for (uint32_t i = 0; i < k; i++) {
    float t = 0;
    cudaEventElapsedTime(&t, e->ev_k[i], e->ev_k[i + 1]);
    e->prof.kernel_ms[i] = t;
  }
  e->ran = true;
  if (prof) *prof = e->prof;
  return KR_OK;
}

int kr_results_fetch(kr_engine *e, kr_results_view *out) {
  if (!e || !out) return KR_E_INVALID;
  if (!e->ran) return fail(e, KR_E_STATE, "no pass has run on the committed snapshot");
  CK(cudaSetDevice(e->cfg.device));
  return fetch_results(e, out);
}

// Digests of n messages given as (pointer, length) pieces: staged 16-byte aligned in pinned memory (the copies run on `threads` host
// threads when the batch is large), one upload, one SHA-1 launch, digests back.
static int hash_pieces(kr_engine *e, const uint8_t *const *ptr, const uint64_t *len, uint32_t n, char *out32xN, uint32_t threads) {
  if (n == 0) return KR_OK;
  CK(cudaSetDevice(e->cfg.device));
  // staging layout: [aligned offsets (n+0) u64 | lens u32 | order u32 | bytes, each message 16-byte aligned | out 32n]
  size_t data = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (len[i] > 0xFFFFFFFFull) return fail(e, KR_E_CAPACITY, "message %u longer than 4 GiB", i);
    data += align_up(len[i], 16);
  }
  size_t o_off = 0, o_len = align_up(8 * (size_t)n), o_ord = align_up(o_len + 4 * (size_t)n), o_data = align_up(o_ord + 4 * (size_t)n), o_out = align_up(o_data + data + 16), total = o_out + 32 * (size_t)n;
  if (total > e->hb_cap) {
    if (e->hb_h) cudaFreeHost(e->hb_h);
    if (e->hb_d) cudaFree(e->hb_d);
    e->hb_h = nullptr; e->hb_d = nullptr; e->hb_cap = 0;
    size_t cap = total + total / 4;
    CK(cudaHostAlloc((void **)&e->hb_h, cap, cudaHostAllocDefault));
    CK(cudaMalloc((void **)&e->hb_d, cap));
    e->hb_cap = cap;
  }
  uint64_t *so = reinterpret_cast<uint64_t *>(e->hb_h + o_off);
  uint32_t *sl = reinterpret_cast<uint32_t *>(e->hb_h + o_len);
  size_t cur = 0;
  for (uint32_t i = 0; i < n; i++) { so[i] = cur; sl[i] = (uint32_t)len[i]; cur += align_up(len[i], 16); }
  auto fill = [&](uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo; i < hi; i++) {
      uint8_t *dst = e->hb_h + o_data + so[i];
      if (len[i]) memcpy(dst, ptr[i], len[i]);
      const size_t pad = align_up(len[i], 16) - len[i];
      if (pad) memset(dst + len[i], 0, pad);
    }
  };
  threads = std::min<uint32_t>(threads, (uint32_t)(data >> 20) + 1);  // a thread per MiB at most
  if (threads <= 1) fill(0, n);
  else {
    std::vector<std::thread> pool;
    const uint32_t per = (n + threads - 1) / threads;
    for (uint32_t t = 0; t < threads; t++) { const uint32_t lo = t * per, hi = std::min(n, lo + per); if (lo < hi) pool.emplace_back(fill, lo, hi); }
    for (auto &th : pool) th.join();
  }
  // message ids by descending block count, staged behind the lengths
  uint32_t *ord = reinterpret_cast<uint32_t *>(e->hb_h + o_ord);
  for (uint32_t i = 0; i < n; i++) ord[i] = i;
  std::stable_sort(ord, ord + n, [&](uint32_t a, uint32_t b) { return (sl[a] + 8) / 64 > (sl[b] + 8) / 64; });
  CK(cudaMemcpyAsync(e->hb_d, e->hb_h, o_data + data, cudaMemcpyHostToDevice, e->sh));
  const uint8_t *db = e->hb_d + o_data;
  const uint64_t *doff = reinterpret_cast<const uint64_t *>(e->hb_d + o_off);
  const uint32_t *dlen = reinterpret_cast<const uint32_t *>(e->hb_d + o_len);
  const uint32_t *dord = reinterpret_cast<const uint32_t *>(e->hb_d + o_ord);
  char *dout = reinterpret_cast<char *>(e->hb_d + o_out);
  const uint32_t ngroups = (n + 31) / 32;
  if (ngroups <= (uint32_t)e->sm_count * 4) k_hash3<1, 0><<<std::min<uint32_t>(ngroups, (uint32_t)e->sm_count * 2), 64, sizeof(H3Smem), e->sh>>>(db, doff, dlen, dord, n, dout);
  else k_hash2<4, 1><<<std::min<uint32_t>((n + 127) / 128, (uint32_t)e->sm_count * 4), 128, 0, e->sh>>>(db, doff, dlen, dord, n, dout, 1u);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(e->hb_h + o_out, dout, 32 * (size_t)n, cudaMemcpyDeviceToHost, e->sh));
  CK(cudaStreamSynchronize(e->sh));
  memcpy(out32xN, e->hb_h + o_out, 32 * (size_t)n);
  return KR_OK;
}

int kr_hash_batch(kr_engine *e, const uint8_t *bytes, const uint64_t *offsets, uint32_t n, char *out32xN) {
  if (!e || (!bytes && n) || !offsets || (!out32xN && n)) return KR_E_INVALID;
  if (n == 0) return KR_OK;
  std::vector<const uint8_t *> ptr(n);
  std::vector<uint64_t> len(n);
  for (uint32_t i = 0; i < n; i++) {
    if (offsets[i + 1] < offsets[i]) return fail(e, KR_E_INVALID, "offsets must be non-decreasing");
    ptr[i] = bytes + offsets[i]; len[i] = offsets[i + 1] - offsets[i];
  }
  return hash_pieces(e, ptr.data(), len.data(), n, out32xN, 1);
}

// strconv.Atoi: optional sign, decimal digits only, no spaces / underscores, must fit an int
static bool go_atoi(const char *t, uint32_t len, long long &v) {
  if (!t || len == 0 || len > 19) return false;
  uint32_t i = 0;
  bool neg = false;
  if (t[0] == '+' || t[0] == '-') { neg = t[0] == '-'; i = 1; }
  if (i >= len) return false;
  v = 0;
  for (; i < len; i++) { if (t[i] < '0' || t[i] > '9') return false; v = v * 10 + (t[i] - '0'); }
  if (neg) v = -v;
  return true;
}

int kr_hash_compare_batch(kr_engine *e, const kr_hash_compare_row *rows, uint32_t n, uint8_t *equal_out, char *goal_hash_out32xN) {
  if (!e || (!rows && n) || (!equal_out && n)) return KR_E_INVALID;
  if (goal_hash_out32xN) memset(goal_hash_out32xN, 0, 32 * (size_t)n);
  // 1. goal specs -> canonical muted JSON (host), rows that need no hash are settled here
  std::vector<int32_t> msg_of(n, -1);  // row -> message index, -1 = goal hash is ""
  std::vector<std::string> emitted(n);
  std::vector<uint8_t> has_msg(n, 0);
  auto emit_rows = [&](uint32_t lo, uint32_t hi) {  // rows are independent: mute + marshal on as many host threads as the batch is worth
    for (uint32_t i = lo; i < hi; i++) {
      const kr_hash_compare_row &r = rows[i];
      equal_out[i] = 2;  // undecided
      long max_groups = -1;
      if (r.partial) {
        long long ng = 0;
        if (!go_atoi(r.num_worker_groups, r.num_worker_groups_len, ng) || ng < 0 || ng > 0x7FFFFFFF) { equal_out[i] = 1; continue; }  // :1140-1142
        max_groups = (long)ng;
      }
      const int rc = kr_specjson_emit_string(r.goal_spec_json, r.goal_spec_len, true, max_groups, emitted[i], nullptr);
      if (rc == KR_E_STATE) continue;                                   // fewer goal groups than the cluster has: goal hash stays ""
      if (rc != KR_OK) { if (r.partial) equal_out[i] = 1; continue; }   // :1151-1153 / the dropped error of :1135
      has_msg[i] = 1;
    }
  };
  uint32_t emit_threads = 1;
  {
    const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t nthreads = std::min<uint32_t>(std::min<uint32_t>(hw, 32u), n / 64u);  // a thread is worth ~64 rows (~2 ms of emitting)
    emit_threads = std::max(1u, nthreads);
    if (nthreads <= 1) emit_rows(0, n);
    else {
      std::vector<std::thread> pool;
      const uint32_t per = (n + nthreads - 1) / nthreads;
      for (uint32_t t = 0; t < nthreads; t++) { const uint32_t lo = t * per, hi = std::min(n, lo + per); if (lo < hi) pool.emplace_back(emit_rows, lo, hi); }
      for (auto &th : pool) th.join();
    }
  }
  std::vector<const uint8_t *> mptr;
  std::vector<uint64_t> mlen;
  for (uint32_t i = 0; i < n; i++) {
    if (!has_msg[i]) continue;
    msg_of[i] = (int32_t)mptr.size();
    mptr.push_back(reinterpret_cast<const uint8_t *>(emitted[i].data()));
    mlen.push_back(emitted[i].size());
  }
  // 2. one GPU batch for every digest (the emitted strings go straight into the pinned staging area)
  const uint32_t nmsg = (uint32_t)mptr.size();
  std::vector<char> digests(32 * (size_t)nmsg);
  if (nmsg) {
    int rc = hash_pieces(e, mptr.data(), mlen.data(), nmsg, digests.data(), emit_threads);
    if (rc) return rc;
  }
  // 3. compare with the annotation
  for (uint32_t i = 0; i < n; i++) {
    if (equal_out[i] != 2) continue;
    const kr_hash_compare_row &r = rows[i];
    const uint32_t clen = r.cluster_hash ? r.cluster_hash_len : 0;
    if (msg_of[i] < 0) { equal_out[i] = clen == 0; continue; }
    const char *d = &digests[32 * (size_t)msg_of[i]];
    if (goal_hash_out32xN) memcpy(goal_hash_out32xN + 32 * (size_t)i, d, 32);
    equal_out[i] = clen == 32 && memcmp(r.cluster_hash, d, 32) == 0;
  }
  return KR_OK;
}

int kr_last_profile(kr_engine *e, kr_profile *prof) {
  if (!e || !prof) return KR_E_INVALID;
  *prof = e->prof;
  return KR_OK;
}

int kr_group_results_device(kr_engine *e, const void **dev_ptr, uint64_t *bytes) {
  if (!e || !dev_ptr || !bytes) return KR_E_INVALID;
  if (!e->ran) return fail(e, KR_E_STATE, "no pass has run");
  *dev_ptr = e->d_out + e->ol.groups;
  *bytes = sizeof(kr_group_result) * (uint64_t)e->sizes.n_groups;
  return KR_OK;
}

int kr_group_results_copy(kr_engine *e, void *dst_device, uint64_t dst_capacity_bytes) {
  if (!e || !dst_device) return KR_E_INVALID;
  if (!e->ran) return fail(e, KR_E_STATE, "no pass has run");
  uint64_t bytes = sizeof(kr_group_result) * (uint64_t)e->sizes.n_groups;
  if (bytes > dst_capacity_bytes) return fail(e, KR_E_CAPACITY, "destination too small (%llu < %llu)", (unsigned long long)dst_capacity_bytes, (unsigned long long)bytes);
  CK(cudaSetDevice(e->cfg.device));
  if (bytes) CK(cudaMemcpyAsync(dst_device, e->d_out + e->ol.groups, bytes, cudaMemcpyDeviceToDevice, e->sm));
  CK(cudaStreamSynchronize(e->sm));
  return KR_OK;
}

const char *kr_last_error(kr_engine *e) { return e ? e->err.c_str() : "null engine"; }

int kr_algorithmic_bytes(kr_engine *e, uint64_t *pass_bytes, uint64_t *hash_bytes, uint64_t *match_bytes) {
  if (!e || !e->begun) return KR_E_INVALID;
  const kr_sizes &n = e->sizes;
  // compulsory traffic only — every input column read once, every output written once
  uint64_t json = 0;
  if (e->committed) {
    kr_snapshot_bufs hb;
    bind_in(e->il, e->h_in, &hb);
    for (uint32_t c = 0; c < n.n_clusters; c++) json += hb.c_json_len[c];
  } else json = n.json_bytes;
  uint64_t hashb = json + 32ull * n.n_clusters;
  uint64_t matchb = 144ull * n.n_clusters - 32ull * n.n_clusters + 56ull * n.n_groups + 4ull * n.n_wtd + 33ull * n.n_pods;
  if (pass_bytes) *pass_bytes = hashb + matchb;
  if (hash_bytes) *hash_bytes = hashb;
  if (match_bytes) *match_bytes = matchb;
  return KR_OK;
}

}  // extern "C"
