/*
 * kr_oracle.h — CPU oracle for the batched reconcile engine.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the CPU arm being timed.  The product path (kuberay_b200/) never
 * falls back to it.
 *
 * It restates, one RayCluster at a time, the reference Go algorithm
 * (ray-project/kuberay: ray-operator/controllers/ray/raycluster_controller.go:619-935, 963-1125,
 * 1127-1248, 1552-1719; utils/util.go:81-175, 386-474, 584-603, 628-665; utils/consistency.go:16-34;
 * common/association.go:83-130, 179-196; rayjob_controller.go:203-216, 343, 880-905)
 * over the SAME columnar snapshot the engine consumes (include/kr_engine.h), and writes the SAME
 * result records, so parity is a byte compare.
 *
 * Pinning: the Go reference cannot be built here (no Go toolchain; k8s deps un-vendored), so this
 * restatement is pinned against the reference's own known-answer tests, transcribed to
 * tests/golden (JSON) (utils/util_test.go:408-475,555-894; raycluster_controller_unit_test.go:
 * 417-1010,1319-1513,1611-2221,2380-2503,2873-3137,3680-3814; utils/consistency_test.go:16-146)
 * and the envtest suites restated as closed loops (raycluster_controller_test.go:132-249,426-530,
 * 565-920,925-1123; tests/test_envtest_lifecycle.py, tests/test_golden.py).
 * SHA-1/base32hex are pinned by FIPS 180 / RFC 4648 vectors and python hashlib.  Literal spec-hash
 * digests vs Go are UNPINNED in the reference itself (no golden digest exists in its tree): the
 * hash input bytes are produced by Go's json.Marshal in production and passed through verbatim.
 */
#ifndef KR_ORACLE_H_
#define KR_ORACLE_H_

#include "../include/kr_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Caller-allocated output arrays, same layout as kr_results_view. */
typedef struct kr_oracle_out {
  kr_cluster_result *clusters;   /* [n_clusters] */
  char              *hash;       /* [32*n_clusters] */
  kr_group_result   *groups;     /* [n_groups] */
  int32_t           *wtd_pod_idx;/* [n_wtd] */
  uint32_t          *sorted_pod_idx; /* [n_pods] */
  uint8_t           *sorted_action;  /* [n_pods] */
  int32_t           *create_idx; /* [create_cap] */
  kr_job_result     *jobs;       /* [n_jobs] */
  uint32_t          *act_start;  /* [n_clusters + 1] */
  uint32_t          *act_cnt;    /* [n_clusters] */
  uint32_t          *act_pod_idx;/* [n_pods] capacity */
  uint8_t           *act_code;   /* [n_pods] capacity */
  uint32_t create_cap;
  uint32_t n_create_total, n_orphans, n_actions;
} kr_oracle_out;

/* list_mode: how a cached r.List is served.
 *   KR_ORACLE_INDEXED: pods pre-bucketed by (namespace, ray.io/cluster) once — the checker.
 *   KR_ORACLE_NS_SCAN: every List walks all pods of the cluster's namespace and tests the selector,
 *                      as controller-runtime's CacheReader.List does — the reference's cost structure
 *                      (SURVEY §3.2(a)); used for the CPU baseline timing. */
enum { KR_ORACLE_INDEXED = 0, KR_ORACLE_NS_SCAN = 1 };

/* Full pass over every cluster. threads<=1: serial.  Returns 0, or <0 on bad input / capacity. */
int kr_oracle_run(const kr_snapshot_bufs *s, const kr_sizes *n, const kr_flags *f,
                  kr_oracle_out *out, int list_mode, int threads);

/* Reconcile only clusters [c0, c1) against the full snapshot (bounded CPU-baseline sample).
 * Outputs for other clusters are left untouched; sorted_* / totals are not produced. */
int kr_oracle_run_range(const kr_snapshot_bufs *s, const kr_sizes *n, const kr_flags *f,
                        kr_oracle_out *out, int list_mode, int threads, uint32_t c0, uint32_t c1);

/* The shared index of one snapshot (what controller-runtime's informer cache keeps between reconciles: its namespace and
 * label indexes are maintained incrementally from watch events, never rebuilt per reconcile).  Build it once per snapshot,
 * then run any number of passes / ranges against it; `s` must stay valid and unchanged for the context's lifetime.
 * `reps` repeats the range inside the worker threads (bench.py: amortises thread start-up over a bounded sample). */
typedef struct kr_oracle_ctx kr_oracle_ctx;
int  kr_oracle_ctx_create(const kr_snapshot_bufs *s, const kr_sizes *n, kr_oracle_ctx **out);
int  kr_oracle_ctx_run(kr_oracle_ctx *cx, const kr_flags *f, kr_oracle_out *out, int list_mode, int threads);
int  kr_oracle_ctx_run_range(kr_oracle_ctx *cx, const kr_flags *f, kr_oracle_out *out, int list_mode, int threads,
                             uint32_t c0, uint32_t c1, int reps);
void kr_oracle_ctx_destroy(kr_oracle_ctx *cx);

/* 1 when this build's SHA-1 uses the x86 SHA extensions (the -march=native build of the CPU arm), 0 for portable C. */
int kr_oracle_sha1_impl(void);

/* base32hex(sha1(msg)) -> 32 chars (utils/util.go:628-640). */
void kr_oracle_hash32(const uint8_t *msg, uint64_t len, char out32[32]);
/* raw SHA-1 (FIPS 180-4) */
void kr_oracle_sha1(const uint8_t *msg, uint64_t len, uint8_t digest[20]);

/* utils.GetWorkerGroupDesiredReplicas (utils/util.go:386-404), int32 wrap semantics. */
int32_t kr_oracle_desired_replicas(int32_t replicas, int32_t min, int32_t max, int32_t num_hosts, uint32_t gflags);

/* shouldDeletePod (raycluster_controller.go:1181-1231) on a packed pod word. */
int kr_oracle_should_delete(uint32_t packed);

#ifdef __cplusplus
}
#endif
#endif
