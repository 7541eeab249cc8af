// tools/pack_bench.cpp — end-to-end epochs THROUGH the native packer (bench.py's `e2e_with_pack` leg runs this binary).
//
// Plain C++ against the C ABI only (include/kr_engine.h), the way the cgo shim drives it: a synthetic informer universe shaped
// like workload C3 (n RayClusters x p Pods, 100 RayClusters per namespace, SURVEY §8(d) distributions, objects with their
// STRINGS) is loaded through kr_packer_cluster_upsert / kr_packer_pod_upsert — interning, row placement and the muted-spec JSON
// emitter included — and flushed (full upload).  Then every timed epoch
//   1. applies informer events natively: 0.8 % of the Pods get a status update, 0.1 % are deleted, 0.1 % new ones are added,
//      2 % of the RayClusters get a status / replicas update (same generation: no spec re-emission);
//   2. kr_packer_flush: the touched pod rows + the small object tables cross PCIe (nothing is re-sent that did not change);
//   3. kr_reconcile_batch on the packer's engine (kernels + D2H of every record, fetch_pod_lists = 0).
// Prints one JSON line.  Build: see __graft_entry__.build() / tools/Makefile-free one-liner in its docstring.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/kr_engine.h"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static uint64_t rng_state = 20260921;
static uint64_t rnd() { rng_state += 0x9E3779B97F4A7C15ull; uint64_t z = rng_state; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static double unif() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static kr_str S(const std::string &s) { return kr_str{s.data(), (uint32_t)s.size()}; }
static kr_str S0() { return kr_str{nullptr, 0}; }
#define CHECK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, kr_packer_last_error(pk)); return 1; } } while (0)

struct Pod { std::string ns, name, cluster, group, ridx; uint8_t node_type, phase, ready, never, term; bool live; };

static void fill_pod(const Pod &p, kr_pod_obj &o, const std::string &ip, const std::string &ver) {
  memset(&o, 0, sizeof o);
  o.ns = S(p.ns); o.name = S(p.name); o.cluster = S(p.cluster); o.group = S(p.group); o.replica_name = S0();
  o.replica_index = p.ridx.empty() ? S0() : S(p.ridx);
  o.node_type = p.node_type; o.phase = p.phase; o.ready_cond = p.ready; o.restart_never = p.never; o.ray_terminated = p.term;
  if (p.node_type == KR_NT_HEAD) {
    static const std::string ok = "HeadPodRunningAndReady", nok = "ContainersNotReady", empty;
    o.head_ready_status = p.ready == KR_COND_TRUE ? KR_COND_TRUE : KR_COND_FALSE;
    o.head_ready_reason = S(p.ready == KR_COND_TRUE ? ok : nok); o.head_ready_msg = S(empty); o.pod_ip = S(ip);
    o.recreate_hash = S(empty); o.kuberay_version = S(ver);
  }
}

int main(int argc, char **argv) {
  const uint32_t Nc = argc > 1 ? (uint32_t)atoi(argv[1]) : 10000, P = argc > 2 ? (uint32_t)atoi(argv[2]) : 100;
  const int steps = argc > 3 ? atoi(argv[3]) : 20, warmup = argc > 4 ? atoi(argv[4]) : 3;
  const uint32_t Np = Nc * P;
  kr_config cfg{};
  cfg.device = 0; cfg.max_clusters = Nc + 16; cfg.max_groups = Nc + 16; cfg.max_wtd = Nc + 16; cfg.max_pods = Np + Np / 8 + 1024; cfg.max_heads = Nc + Nc / 8 + 16;
  cfg.max_jobs = 16; cfg.max_creates = Np; cfg.max_json_bytes = (uint64_t)Nc * 8192 + (1u << 20);
  kr_packer *pk = nullptr;
  { int rc = kr_packer_create(&cfg, &pk); if (rc) { fprintf(stderr, "kr_packer_create failed: %d\n", rc); return 1; } }
  kr_engine *eng = kr_packer_engine(pk);
  const std::string ver = "nightly", gname = "group-0", hgroup = "headgroup";
  // ---- the universe: RayClusters (spec as JSON text in API-server key order) and their Pods
  std::vector<std::string> cname(Nc), cns(Nc), cspec(Nc), csvc(Nc), csvcip(Nc), cuid(Nc);
  std::vector<int32_t> creplicas(Nc);
  const int env_n[4] = {18, 40, 72, 116};  // ~1.5 / 2.5 / 4 / 6 KB of muted JSON, like the synthetic templates
  double t0 = now_ms();
  for (uint32_t c = 0; c < Nc; c++) {
    cname[c] = "raycluster-" + std::to_string(c); cns[c] = "ns-" + std::to_string(c / 100); csvc[c] = cname[c] + "-head-svc";
    csvcip[c] = "10.0." + std::to_string(c >> 8) + "." + std::to_string(c & 255); cuid[c] = "uid-" + std::to_string(c);
    std::string env;
    for (int i = 0; i < env_n[c & 3]; i++) env += std::string(i ? "," : "") + "{\"name\":\"RAY_ENV_" + std::to_string(i) + "\",\"value\":\"v" + std::to_string(i) + "-" + std::to_string(c) + "\"}";
    const std::string tmpl = "{\"spec\":{\"containers\":[{\"env\":[" + env + "],\"image\":\"rayproject/ray:2.46.0-" + std::to_string(c) +
                             "\",\"name\":\"ray\",\"resources\":{\"limits\":{\"cpu\":\"2\",\"memory\":\"4Gi\"},\"requests\":{\"cpu\":\"2\",\"memory\":\"4Gi\"}}}]}}";
    creplicas[c] = (int32_t)(P - 1) + (int32_t)(rnd() % 7) - 3;
    cspec[c] = "{\"enableInTreeAutoscaling\":" + std::string(unif() < 0.3 ? "true" : "false") + ",\"headGroupSpec\":{\"rayStartParams\":{\"dashboard-host\":\"0.0.0.0\"},\"template\":" + tmpl +
               "},\"rayVersion\":\"2.46.0\",\"workerGroupSpecs\":[{\"groupName\":\"group-0\",\"maxReplicas\":200,\"minReplicas\":1,\"rayStartParams\":{},\"replicas\":" +
               std::to_string(creplicas[c]) + ",\"template\":" + tmpl + "}]}";
  }
  auto upsert_cluster = [&](uint32_t c, uint64_t rv) -> int {
    kr_cluster_obj o;
    memset(&o, 0, sizeof o);
    o.ns = S(cns[c]); o.name = S(cname[c]); o.uid = S(cuid[c]); o.resource_version = rv; o.generation = 1;
    o.flags = KR_CF_HEAD_EXPECT_OK | (unif() < 0.3 ? KR_CF_AUTOSCALING : 0u);
    o.old_state = KR_STATE_READY; o.old_counts[0] = o.old_counts[1] = (int32_t)P - 1; o.old_counts[2] = creplicas[c]; o.old_counts[3] = 1; o.old_counts[4] = 200;
    o.old_cond_status[KR_COND_PROVISIONED] = KR_COND_TRUE; o.old_cond_variant[KR_COND_PROVISIONED] = KR_CV_PROV_ALL_READY;
    o.old_cond_status[KR_COND_HEAD_POD_READY] = KR_COND_TRUE; o.old_cond_variant[KR_COND_HEAD_POD_READY] = KR_CV_HEAD_FROM_POD;
    static const std::string ok = "HeadPodRunningAndReady", empty;
    o.old_head_ready_reason = S(ok); o.old_head_ready_msg = S(empty);
    o.svc_count = 1; o.svc_ip_kind = KR_SVCIP_NORMAL; o.svc_ip = S(csvcip[c]); o.svc_name = S(csvc[c]); o.status_summary = S(cuid[c]);
    kr_group_obj g;
    memset(&g, 0, sizeof g);
    g.name = S(gname); g.replicas = creplicas[c]; g.min_replicas = 1; g.max_replicas = 200; g.num_hosts = 1; g.flags = KR_GF_EXPECT_OK;
    o.groups = &g; o.n_groups = 1;
    o.spec_json = (const uint8_t *)cspec[c].data(); o.spec_json_len = cspec[c].size();
    return kr_packer_cluster_upsert(pk, &o);
  };
  for (uint32_t c = 0; c < Nc; c++) CHECK(upsert_cluster(c, 1));
  const double t_clusters = now_ms() - t0;
  // Pods in a shuffled (informer-like) order
  std::vector<Pod> pods(Np);
  std::vector<uint32_t> order(Np);
  for (uint32_t i = 0; i < Np; i++) order[i] = i;
  for (uint32_t i = Np - 1; i > 0; i--) std::swap(order[i], order[rnd() % (i + 1)]);
  for (uint32_t i = 0; i < Np; i++) {
    const uint32_t c = i / P, slot = i % P;
    Pod &p = pods[i];
    p.ns = cns[c]; p.cluster = cname[c]; p.live = true;
    if (slot == 0) { p.name = cname[c] + "-head"; p.group = hgroup; p.node_type = KR_NT_HEAD; }
    else { p.name = cname[c] + "-w-" + std::to_string(slot); p.group = gname; p.node_type = KR_NT_WORKER; if (unif() < 0.9) p.ridx = std::to_string(slot - 1); }
    const double u = unif();
    p.phase = u < 0.01 ? KR_PHASE_SUCCEEDED : u < 0.02 ? KR_PHASE_FAILED : u < 0.04 ? KR_PHASE_PENDING : KR_PHASE_RUNNING;
    if (slot == 0 && unif() > 0.002) p.phase = KR_PHASE_RUNNING;
    p.ready = p.phase == KR_PHASE_RUNNING ? (unif() < 0.95 ? KR_COND_TRUE : KR_COND_FALSE) : KR_COND_ABSENT;
    p.never = unif() < 0.5; p.term = unif() < 0.005;
  }
  t0 = now_ms();
  kr_pod_obj po;
  for (uint32_t k = 0; k < Np; k++) { const Pod &p = pods[order[k]]; fill_pod(p, po, csvcip[order[k] / P], ver); CHECK(kr_packer_pod_upsert(pk, &po)); }
  const double t_pods = now_ms() - t0;
  t0 = now_ms();
  uint32_t mode = 0;
  CHECK(kr_packer_flush(pk, &mode));
  kr_flags flags;
  memset(&flags, 0, sizeof flags);
  flags.gate_status_conditions = 1; flags.gate_multihost_indexing = 1;
  { static const std::string a = "HeadPodNotFound", b = "Head Pod not found"; flags.id_head_not_found_reason = kr_packer_intern(pk, S(a)); flags.id_head_not_found_msg = kr_packer_intern(pk, S(b)); }
  kr_results_view view;
  { int rc = kr_reconcile_batch(eng, &flags, &view); if (rc) { fprintf(stderr, "kr_reconcile_batch failed: %d (%s)\n", rc, kr_last_error(eng)); return 1; } }
  const double t_first = now_ms() - t0;
  // ---- timed epochs
  uint32_t next_new = 0;
  double ev_ms = 0, flush_ms = 0, rec_ms = 0;
  uint64_t h2d = 0, d2h = 0, n_events = 0;
  std::vector<uint32_t> deleted;
  for (int step = -warmup; step < steps; step++) {
    double a = now_ms();
    const uint32_t n_upd = Np * 8 / 1000, n_del = Np / 1000;
    uint64_t ev = 0;
    for (uint32_t k = 0; k < n_upd; k++) {  // status updates
      Pod &p = pods[rnd() % Np];
      if (!p.live) continue;
      p.ready = p.ready == KR_COND_TRUE ? KR_COND_FALSE : KR_COND_TRUE;
      if (p.phase != KR_PHASE_RUNNING) p.phase = KR_PHASE_RUNNING;
      fill_pod(p, po, csvcip[0], ver); CHECK(kr_packer_pod_upsert(pk, &po)); ev++;
    }
    for (uint32_t i : deleted) {  // Pods created since the last epoch take the rows freed one epoch earlier
      Pod &p = pods[i];
      p.name = p.cluster + "-n-" + std::to_string(next_new++); p.live = true; p.phase = KR_PHASE_PENDING; p.ready = KR_COND_ABSENT;
      fill_pod(p, po, csvcip[0], ver); CHECK(kr_packer_pod_upsert(pk, &po)); ev++;
    }
    deleted.clear();
    for (uint32_t k = 0; k < n_del; k++) {
      const uint32_t i = (uint32_t)(rnd() % Np);
      Pod &p = pods[i];
      if (!p.live || p.node_type == KR_NT_HEAD) continue;
      CHECK(kr_packer_pod_delete(pk, S(p.ns), S(p.name))); p.live = false; deleted.push_back(i); ev++;
    }
    for (uint32_t k = 0; k < Nc / 50; k++) { const uint32_t c = (uint32_t)(rnd() % Nc); creplicas[c] += (int32_t)(rnd() % 3) - 1; CHECK(upsert_cluster(c, 2 + (uint64_t)(step + warmup))); ev++; }
    double b = now_ms();
    CHECK(kr_packer_flush(pk, &mode));
    kr_profile prof;
    double c0 = now_ms();
    { int rc = kr_reconcile_batch(eng, &flags, &view); if (rc) { fprintf(stderr, "kr_reconcile_batch failed: %d (%s)\n", rc, kr_last_error(eng)); return 1; } }
    double d = now_ms();
    kr_last_profile(eng, &prof);
    if (step >= 0) { ev_ms += b - a; flush_ms += c0 - b; rec_ms += d - c0; h2d += prof.h2d_bytes; d2h += prof.d2h_bytes; n_events += ev; }
  }
  const double epoch_ms = (ev_ms + flush_ms + rec_ms) / steps;
  printf("{\"workload\": \"%u RayClusters x %u pods through the native packer\", \"load_clusters_ms\": %.1f, \"load_pods_ms\": %.1f, \"pod_upserts_per_s\": %.0f, "
         "\"first_flush_and_pass_ms\": %.2f, \"steps\": %d, \"events_per_epoch\": %.0f, \"events_ms\": %.4f, \"flush_ms\": %.4f, \"reconcile_ms\": %.4f, \"epoch_ms\": %.4f, "
         "\"reconciles_per_s\": %.0f, \"h2d_bytes_per_epoch\": %.0f, \"d2h_bytes_per_epoch\": %.0f, \"last_flush_mode\": %u, \"n_actions\": %u, \"n_create_total\": %u}\n",
         Nc, P, t_clusters, t_pods, Np / (t_pods / 1e3), t_first, steps, (double)n_events / steps, ev_ms / steps, flush_ms / steps, rec_ms / steps, epoch_ms,
         Nc / (epoch_ms / 1e3), (double)h2d / steps, (double)d2h / steps, mode, view.n_actions, view.n_create_total);
  kr_packer_destroy(pk);
  return 0;
}
