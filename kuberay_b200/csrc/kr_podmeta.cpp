// Pod metadata builder (first part) — host side of the C ABI, no GPU.
//
// The engine's pass says WHICH pods to create (head_action, kr_group_result.n_create, create_idx); this file turns those tuples
// into the ObjectMeta the reference gives each new Pod, so the Go shim only has to graft the patch onto the template and run the
// container half of BuildPod.  Restated from (never copied):
//   utils.PodName / CheckName / CheckLabel / GenerateIdentifier / GenerateRayWorkerReplicaGroupName
//                                           ray-operator/controllers/ray/utils/util.go:198-265, 375-384
//   DefaultHeadPodTemplate (metadata part)  ray-operator/controllers/ray/common/pod.go:166-190
//   DefaultWorkerPodTemplate (metadata)     common/pod.go:352-357, 420-442
//   initTemplateAnnotations                 common/pod.go:62-75
//   configureGCSFaultTolerance (annotations) common/pod.go:77-87, 105-114
//   BuildPod's ray.io/serve label           common/pod.go:583-588
//   labelPod / mergeLabels                  common/pod.go:775-799, 1276-1283
//   createHeadPod's annotations             raycluster_controller.go:1313-1316
//   SetControllerReference                  sigs.k8s.io/controller-runtime v0.23.1 pkg/controller/controllerutil (absent from
//                                           /root/reference: third-party; its published behaviour is one ownerReference
//                                           {apiVersion, kind, name, uid, controller: true, blockOwnerDeletion: true})
//   rand.String alphabet                    k8s.io/apimachinery v0.36.0 pkg/util/rand (third-party, absent): "bcdfghjklmnpqrstvwxz2456789"
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/kr_engine.h"

void kr_go_string_append(std::string &out, const std::string &s);  // kr_specjson.cpp: encoding/json string encoding

namespace {

thread_local std::string g_err;

inline std::string str(kr_str s) { return (s.p && s.n) ? std::string(s.p, s.n) : std::string(); }

// unicode.IsPunct(rune(b)) for one BYTE widened to a rune (so 0x80..0xFF are the Latin-1 code points, as in the reference's
// rune(s[0])): category P in ASCII is !"#%&'()*,-./:;?@[\]_{} (the symbols $+<=>^`|~ are category S), in Latin-1 ¡ § « ¶ · » ¿.
bool go_is_punct(unsigned char b) {
  switch (b) {
    case '!': case '"': case '#': case '%': case '&': case '\'': case '(': case ')': case '*': case ',': case '-': case '.': case '/':
    case ':': case ';': case '?': case '@': case '[': case '\\': case ']': case '_': case '{': case '}':
    case 0xA1: case 0xA7: case 0xAB: case 0xB6: case 0xB7: case 0xBB: case 0xBF:
      return true;
    default:
      return false;
  }
}
inline bool go_is_digit(unsigned char b) { return b >= '0' && b <= '9'; }  // unicode.IsDigit: Nd; Latin-1 has no other Nd

// util.go:217-240.  Precondition: s not empty.
std::string check_name(std::string s) {
  const size_t max_len = 50;
  if (s.size() > max_len) s = s.substr(s.size() - max_len);
  if (go_is_digit((unsigned char)s[0])) s[0] = 'r';
  if (go_is_punct((unsigned char)s[0])) s[0] = 'r';
  return s;
}
// util.go:247-265
std::string check_label(std::string s) {
  const size_t max_len = 63;
  if (s.size() > max_len) s = s.substr(s.size() - max_len);
  if (go_is_punct((unsigned char)s[0])) s[0] = 'r';
  return s;
}
// util.go:198-215.  strings.ToLower is applied to ASCII only here: Pod name prefixes are RFC 1123 names (the RayCluster name and
// group names are validated upstream, utils/validation.go), so no other letters reach this.
std::string pod_name(const std::string &prefix, uint8_t node_type, bool is_generate) {
  std::string r = prefix.substr(0, 50);
  r += '-';
  r += node_type == KR_NT_HEAD ? "head" : "worker";
  for (char &c : r) if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
  if (is_generate) r += '-';
  return r;
}

int64_t give(const std::string &s, char *out, uint64_t cap) {
  if (out && cap) memcpy(out, s.data(), s.size() < cap ? s.size() : cap);
  return (int64_t)s.size();
}

typedef std::map<std::string, std::string> StrMap;  // ordered by bytes = the order encoding/json writes map keys in

void put_all(StrMap &m, const kr_kv *kv, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) m[str(kv[i].key)] = str(kv[i].value);
}

void emit_map(std::string &out, const StrMap &m) {
  out += '{';
  bool first = true;
  for (const auto &e : m) {
    if (!first) out += ',';
    first = false;
    kr_go_string_append(out, e.first);
    out += ':';
    kr_go_string_append(out, e.second);
  }
  out += '}';
}

struct GroupMeta {      // everything of a group that does not depend on the create tuple, built once per call
  StrMap labels, annotations;
  std::string gen_name;
  bool multi_host = false;
};

// common/pod.go:775-799 on top of :1276-1283
void label_pod(StrMap &labels, const kr_podmeta_group &g, const char *node_type, const std::string &cluster, const std::string &group) {
  StrMap merged;
  put_all(merged, g.template_labels, g.n_template_labels);
  put_all(merged, g.group_labels, g.n_group_labels);
  labels["ray.io/is-ray-node"] = "yes";
  labels["ray.io/cluster"] = cluster;
  labels["ray.io/node-type"] = node_type;
  labels["ray.io/group"] = group;
  labels["ray.io/identifier"] = check_label(cluster + "-" + node_type);
  labels["app.kubernetes.io/name"] = "kuberay";
  labels["app.kubernetes.io/created-by"] = "kuberay-operator";
  for (const auto &e : merged) {
    if (e.first == "ray.io/node-type" || e.first == "ray.io/group" || e.first == "ray.io/cluster") continue;
    labels[e.first] = e.second;
  }
}

void base_annotations(StrMap &a, const kr_podmeta_cluster &c, const kr_podmeta_group &g, bool head) {
  put_all(a, g.template_annotations, g.n_template_annotations);
  if (c.overwrite_container_cmd) a["ray.io/overwrite-container-cmd"] = "true";                 // common/pod.go:72-74
  if (head) {
    a["ray.io/ft-enabled"] = c.ft_enabled ? "true" : "false";                                    // :85-87
    if (c.ft_enabled) {                                                                          // :105-114
      std::string ns = str(c.uid);
      if (c.storage_ns_annotation.p) ns = str(c.storage_ns_annotation);
      if (c.storage_ns_option.p && c.storage_ns_option.n) ns = str(c.storage_ns_option);
      a["ray.io/external-storage-namespace"] = ns;
    }
  }
}

const char kAlphanums[] = "bcdfghjklmnpqrstvwxz2456789";
inline uint64_t splitmix64(uint64_t &x) {
  uint64_t z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace

extern "C" {

const char *kr_pod_meta_last_error(void) { return g_err.c_str(); }

int64_t kr_pod_name(kr_str prefix, uint8_t node_type, uint8_t is_generate_name, char *out, uint64_t cap) {
  if (node_type != KR_NT_HEAD && node_type != KR_NT_WORKER) { g_err = "kr_pod_name: node_type must be KR_NT_HEAD or KR_NT_WORKER"; return KR_E_INVALID; }
  return give(pod_name(str(prefix), node_type, is_generate_name != 0), out, cap);
}
int64_t kr_check_name(kr_str s, char *out, uint64_t cap) {
  if (!s.p || !s.n) { g_err = "kr_check_name: empty input (the reference reads s[0])"; return KR_E_INVALID; }
  return give(check_name(str(s)), out, cap);
}
int64_t kr_check_label(kr_str s, char *out, uint64_t cap) {
  if (!s.p || !s.n) { g_err = "kr_check_label: empty input (the reference reads s[0])"; return KR_E_INVALID; }
  return give(check_label(str(s)), out, cap);
}

int kr_pod_meta_build(const kr_podmeta_cluster *cluster, const kr_podmeta_group *head, const kr_podmeta_group *groups, uint32_t n_groups,
                      const kr_podmeta_create *creates, uint32_t n_creates, uint8_t *out, uint64_t cap, uint64_t *off, uint64_t *need) {
  if (!cluster || !off || !need || (n_creates && !creates) || (n_groups && !groups)) { g_err = "kr_pod_meta_build: null argument"; return KR_E_INVALID; }
  const kr_podmeta_cluster &c = *cluster;
  const std::string cname = str(c.name), ns = str(c.ns);
  if (cname.empty()) { g_err = "kr_pod_meta_build: the RayCluster has no name"; return KR_E_INVALID; }

  // per group: everything the tuple does not touch
  std::vector<GroupMeta> gm(n_groups);
  std::vector<char> built(n_groups, 0);
  GroupMeta hm;
  bool head_built = false;
  auto serve_label = [&](StrMap &labels, bool is_head) {                                        // common/pod.go:583-588
    if (c.crd_type == KR_CRD_RAYSERVICE) labels["ray.io/serve"] = is_head ? "false" : "true";
  };
  std::string owner;
  {
    owner = "[{\"apiVersion\":\"ray.io/v1\",\"kind\":\"RayCluster\",\"name\":";
    kr_go_string_append(owner, cname);
    owner += ",\"uid\":";
    kr_go_string_append(owner, str(c.uid));
    owner += ",\"controller\":true,\"blockOwnerDeletion\":true}]";
  }

  std::string buf;
  for (uint32_t i = 0; i < n_creates; i++) {
    off[i] = buf.size();
    const kr_podmeta_create &t = creates[i];
    const GroupMeta *g;
    bool is_head = t.group < 0;
    if (is_head) {
      if (!head) { g_err = "kr_pod_meta_build: a head create without the head group"; return KR_E_INVALID; }
      if (!head_built) {
        label_pod(hm.labels, *head, "head", cname, "headgroup");                                 // common/pod.go:187-188
        serve_label(hm.labels, true);
        base_annotations(hm.annotations, c, *head, true);
        if (c.cluster_hash.p && c.cluster_hash.n) {                                              // raycluster_controller.go:1313-1316
          hm.annotations["ray.io/upgrade-strategy-recreate-hash"] = str(c.cluster_hash);
          hm.annotations["ray.io/kuberay-version"] = str(c.kuberay_version);
        }
        hm.gen_name = pod_name(cname, KR_NT_HEAD, !c.deterministic_head_name);                   // :1389
        head_built = true;
      }
      g = &hm;
    } else {
      if ((uint32_t)t.group >= n_groups) { g_err = "kr_pod_meta_build: create " + std::to_string(i) + " names group " + std::to_string(t.group) + " of " + std::to_string(n_groups); return KR_E_INVALID; }
      GroupMeta &w = gm[t.group];
      if (!built[t.group]) {
        const kr_podmeta_group &gs = groups[t.group];
        const std::string gname = str(gs.group_name);
        label_pod(w.labels, gs, "worker", cname, gname);                                         // common/pod.go:427-428
        serve_label(w.labels, false);
        base_annotations(w.annotations, c, gs, false);
        w.gen_name = pod_name(cname + "-" + gname, KR_NT_WORKER, true);                          // raycluster_controller.go:1418
        w.multi_host = gs.num_of_hosts > 1;
        built[t.group] = 1;
      }
      g = &w;
    }
    buf += (is_head && c.deterministic_head_name) ? "{\"name\":" : "{\"generateName\":";        // common/pod.go:170-175, 354
    kr_go_string_append(buf, g->gen_name);
    buf += ",\"namespace\":";
    kr_go_string_append(buf, ns);
    buf += ",\"labels\":";
    if (!is_head && c.gate_multihost_indexing) {                                                 // common/pod.go:430-439 (the tuple's labels)
      StrMap l = g->labels;
      l["ray.io/worker-group-replica-index"] = std::to_string(t.replica_index);
      if (g->multi_host) {
        l["ray.io/worker-group-replica-name"] = str(t.replica_name);
        l["ray.io/replica-host-index"] = std::to_string(t.host_index);
      }
      emit_map(buf, l);
    } else {
      emit_map(buf, g->labels);
    }
    buf += ",\"annotations\":";
    emit_map(buf, g->annotations);
    buf += ",\"ownerReferences\":";
    buf += owner;
    buf += '}';
  }
  off[n_creates] = buf.size();
  *need = buf.size();
  if (buf.size() > cap || (!out && !buf.empty())) { g_err = "kr_pod_meta_build: output buffer too small"; return KR_E_CAPACITY; }
  if (!buf.empty()) memcpy(out, buf.data(), buf.size());
  return KR_OK;
}

int kr_pod_creates_expand(const kr_group_result *gr, const kr_podmeta_group *groups, uint32_t n_groups, const int32_t *create_idx, uint8_t head_create,
                          uint8_t gate, uint64_t seed, kr_podmeta_create *out, uint32_t cap, char *name_buf, uint64_t name_cap, uint32_t *n_out) {
  if (!n_out || (n_groups && (!gr || !groups))) { g_err = "kr_pod_creates_expand: null argument"; return KR_E_INVALID; }
  uint64_t n = head_create ? 1 : 0, name_bytes = 0;
  for (uint32_t g = 0; g < n_groups; g++) {
    const bool mh = gate && (gr[g].flags & KR_GR_MULTIHOST);
    if (gr[g].n_create && !create_idx) { g_err = "kr_pod_creates_expand: creates without create_idx"; return KR_E_INVALID; }
    if (mh && groups[g].num_of_hosts < 1) { g_err = "kr_pod_creates_expand: multi-host group with NumOfHosts < 1"; return KR_E_INVALID; }
    n += (uint64_t)gr[g].n_create * (mh ? (uint64_t)groups[g].num_of_hosts : 1);
    if (mh) name_bytes += (uint64_t)gr[g].n_create * (groups[g].group_name.n + 6);
  }
  if (n > 0xFFFFFFFFull) { g_err = "kr_pod_creates_expand: more than 2^32 creates"; return KR_E_CAPACITY; }
  *n_out = (uint32_t)n;
  if (n > cap || name_bytes > name_cap || (n && !out) || (name_bytes && !name_buf)) { g_err = "kr_pod_creates_expand: output too small"; return KR_E_CAPACITY; }
  uint32_t k = 0;
  uint64_t nb = 0, rng = seed;
  if (head_create) out[k++] = kr_podmeta_create{-1, 0, 0, kr_str{nullptr, 0}};
  for (uint32_t g = 0; g < n_groups; g++) {
    const bool mh = gate && (gr[g].flags & KR_GR_MULTIHOST);
    for (uint32_t i = 0; i < gr[g].n_create; i++) {
      const int32_t idx = gate ? create_idx[gr[g].create_off + i] : 0;     // gate off: createWorkerPod(..., "", 0, 0) (:887)
      if (!mh) { out[k++] = kr_podmeta_create{(int32_t)g, idx, 0, kr_str{nullptr, 0}}; continue; }
      char *nm = name_buf + nb;                                            // util.go:377-379: "<group>-<rand.String(5)>"
      const uint32_t gl = groups[g].group_name.n;
      if (gl) memcpy(nm, groups[g].group_name.p, gl);
      nm[gl] = '-';
      uint64_t r = splitmix64(rng);
      for (int q = 0; q < 5; q++) { nm[gl + 1 + q] = kAlphanums[r % 27]; r /= 27; }
      nb += gl + 6;
      for (int32_t j = 0; j < groups[g].num_of_hosts; j++) out[k++] = kr_podmeta_create{(int32_t)g, idx, j, kr_str{nm, gl + 6}};
    }
  }
  return KR_OK;
}

}  // extern "C"
