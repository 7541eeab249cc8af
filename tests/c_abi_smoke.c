/* tests/c_abi_smoke.c — the C ABI used from plain C99, the way a cgo shim reaches it (no Python, no C++).
 * Built and run by tests/test_c_abi.py.  One RayCluster "rc" in namespace "ns" with one worker group (replicas 2) and three
 * pods: a Running+Ready head, one Running worker, one Failed worker.  Expected decisions (raycluster_controller.go:790-811):
 * the Failed worker is deleted and the pass aborts with "delete 1 unhealthy worker Pods"; nothing is created.
 * Exit code 0 = as expected (or, without a CUDA device, the documented KR_E_NO_DEVICE refusal); anything else is a failure. */
#include <stdio.h>
#include <string.h>

#include "kr_engine.h"

enum { ID_NS = 2, ID_RC = 3, ID_GROUP = 4, ID_HEAD = 5, ID_W0 = 6, ID_W1 = 7, ID_HEADGROUP = 8 }; /* interned strings; 0 absent, 1 "" */

static uint32_t pod_word(uint32_t node_type, uint32_t phase, uint32_t ready) {
  return (node_type << KR_PP_NODE_TYPE_SHIFT) | (phase << KR_PP_PHASE_SHIFT) | (ready << KR_PP_READY_SHIFT);
}

/* The host-side builders need no device: one worker Pod manifest for a RayCluster given as JSON (kr_pod_build). */
static int host_builders(void) {
  const char cluster[] =
      "{\"metadata\":{\"name\":\"rc\",\"namespace\":\"ns\",\"uid\":\"u-1\"},\"spec\":{\"headGroupSpec\":{\"rayStartParams\":{},\"template\":{\"spec\":{\"containers\":[{\"name\":\"h\",\"image\":\"ray:2\"}]}}},"
      "\"workerGroupSpecs\":[{\"groupName\":\"g\",\"rayStartParams\":{},\"template\":{\"spec\":{\"containers\":[{\"name\":\"w\",\"image\":\"ray:2\",\"resources\":{\"limits\":{\"cpu\":\"2\",\"memory\":\"1Gi\"}}}]}}}]}}";
  static uint8_t out[16384];
  kr_podbuild_env env;
  kr_podmeta_create tuple;
  uint64_t off[2], need = 0;
  memset(&env, 0, sizeof env);
  memset(&tuple, 0, sizeof tuple);
  env.kuberay_version.p = "v1.5.0"; env.kuberay_version.n = 6; env.gate_multihost_indexing = 1;
  tuple.group = 0; tuple.replica_index = 3; tuple.replica_name.p = ""; tuple.replica_name.n = 0;
  if (kr_pod_build((const uint8_t *)cluster, sizeof cluster - 1, &env, &tuple, 1, out, sizeof out - 1, off, &need) != KR_OK) {
    printf("kr_pod_build: %s\n", kr_pod_build_last_error());
    return 1;
  }
  out[need] = 0;
  if (off[0] != 0 || off[1] != need || strncmp((const char *)out, "{\"kind\":\"Pod\",\"apiVersion\":\"v1\",\"metadata\":{\"generateName\":\"rc-g-worker-\"", 72) != 0 ||
      !strstr((const char *)out, "ray start  --address=rc-head-svc.ns.svc.cluster.local:6379 ") || !strstr((const char *)out, "\"ray.io/worker-group-replica-index\":\"3\"") ||
      !strstr((const char *)out, "--num-cpus=2 ") || !strstr((const char *)out, "\"name\":\"wait-gcs-ready\"")) {
    printf("kr_pod_build wrote an unexpected manifest:\n%s\n", (const char *)out);
    return 1;
  }
  printf("host builders from C: OK (%lu-byte worker manifest)\n", (unsigned long)need);
  return 0;
}

int main(void) {
  if (host_builders()) return 1;
  int ndev = kr_device_count();
  if (ndev <= 0) { printf("no CUDA device (kr_device_count = %d): the engine refuses to run, as documented\n", ndev); return 0; }

  kr_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.device = 0; cfg.max_clusters = 4; cfg.max_groups = 4; cfg.max_wtd = 4; cfg.max_pods = 16; cfg.max_heads = 4; cfg.max_jobs = 4;
  cfg.max_creates = 16; cfg.max_json_bytes = 1024;
  kr_engine *e = NULL;
  if (kr_engine_create(&cfg, &e) != KR_OK) { printf("kr_engine_create failed\n"); return 1; }

  const char spec[] = "{\"headGroupSpec\":{},\"workerGroupSpecs\":[{\"groupName\":\"g\"}]}";
  kr_sizes n;
  memset(&n, 0, sizeof n);
  n.n_clusters = 1; n.n_groups = 1; n.n_pods = 3; n.n_heads = 1; n.json_bytes = 64;
  kr_snapshot_bufs b;
  if (kr_snapshot_begin(e, &n, &b) != KR_OK) { printf("begin: %s\n", kr_last_error(e)); return 1; }
  /* the arenas are not cleared between epochs: write every column of every live row */
  b.c_ns_id[0] = ID_NS; b.c_name_id[0] = ID_RC; b.c_uid_hash[0] = 42; b.c_flags[0] = KR_CF_HEAD_EXPECT_OK;
  b.c_suspend_status[0] = 0; b.c_ext_err_kind[0] = 0; b.c_ext_err_msg_id[0] = 0; b.c_group_off[0] = 0; b.c_group_cnt[0] = 1;
  b.c_json_off[0] = 0; b.c_json_len[0] = (uint32_t)(sizeof spec - 1);
  memset(b.json, 0, 64); memcpy(b.json, spec, sizeof spec - 1);
  b.c_old_state[0] = 0; memset(b.c_old_counts, 0, 5 * sizeof(int32_t)); memset(b.c_old_cond_status, 0, 5); memset(b.c_old_cond_variant, 0, 5);
  b.c_old_cond_reason_id[0] = 0; b.c_old_cond_msg_id[0] = b.c_old_cond_msg_id[1] = 0; memset(b.c_old_head_ids, 0, 4 * sizeof(uint32_t));
  b.c_svc_count[0] = 1; b.c_svc_ip_kind[0] = KR_SVCIP_NORMAL; b.c_svc_ip_id[0] = 9; b.c_svc_name_id[0] = 10; b.c_summary_id[0] = 0;
  b.g_cluster_idx[0] = 0; b.g_name_id[0] = ID_GROUP; b.g_replicas[0] = 2; b.g_min[0] = 0; b.g_max[0] = 4; b.g_num_hosts[0] = 1;
  b.g_flags[0] = KR_GF_EXPECT_OK; b.g_wtd_off[0] = 0; b.g_wtd_cnt[0] = 0;
  const uint32_t names[3] = {ID_HEAD, ID_W0, ID_W1}, groups[3] = {ID_HEADGROUP, ID_GROUP, ID_GROUP};
  const uint32_t words[3] = {pod_word(1, 2, 1), pod_word(2, 2, 1), pod_word(2, 4, 0)}; /* head Running Ready | worker Running | worker Failed */
  for (int p = 0; p < 3; p++) {
    b.p_ns_id[p] = ID_NS; b.p_cluster_name_id[p] = ID_RC; b.p_group_name_id[p] = groups[p]; b.p_name_id[p] = names[p];
    b.p_packed[p] = words[p]; b.p_replica_index[p] = 0; b.p_replica_name_id[p] = 0;
  }
  b.h_pod_idx[0] = 0; b.h_ready_status[0] = 1; b.h_ready_reason_id[0] = 11; b.h_ready_msg_id[0] = 1; b.h_pod_ip_id[0] = 12;
  b.h_annot_state[0] = KR_ANNOT_EMPTY; b.h_version_state[0] = KR_VER_EMPTY; memset(b.h_annot_hash, 0, 32);
  if (kr_snapshot_commit(e) != KR_OK) { printf("commit: %s\n", kr_last_error(e)); return 1; }

  kr_flags f;
  memset(&f, 0, sizeof f);
  f.gate_status_conditions = 1; f.gate_multihost_indexing = 1; f.fetch_pod_lists = 0;
  kr_results_view r;
  if (kr_reconcile_batch(e, &f, &r) != KR_OK) { printf("reconcile: %s\n", kr_last_error(e)); return 1; }

  const kr_cluster_result *c = &r.clusters[0];
  const kr_group_result *g = &r.groups[0];
  printf("path %d head_action %d err_kind %d err_arg %d n_pods %d | group expected %d unhealthy %d n_create %u | actions %u hash %.32s\n",
         c->path, c->head_action, c->err_kind, c->err_arg, c->n_pods, g->expected, g->n_unhealthy, g->n_create, r.n_actions, r.hash);
  int ok = c->err_kind == KR_ERR_UNHEALTHY_WORKERS && c->err_arg == 1 && c->n_pods == 3 && g->expected == 2 && g->n_unhealthy == 1 &&
           g->n_create == 0 && r.n_actions == 1 && r.act_cnt[0] == 1 && r.act_pod_idx[r.act_start[0]] == 2 &&
           r.act_code[r.act_start[0]] == KR_ACT_DELETE_UNHEALTHY && r.n_orphans == 0;
  /* the same digest through the RayService entry point */
  char h[32];
  const uint64_t offs[2] = {0, sizeof spec - 1};
  if (kr_hash_batch(e, (const uint8_t *)spec, offs, 1, h) != KR_OK || memcmp(h, r.hash, 32) != 0) ok = 0;
  kr_engine_destroy(e);
  printf(ok ? "C ABI smoke: OK\n" : "C ABI smoke: MISMATCH\n");
  return ok ? 0 : 2;
}
