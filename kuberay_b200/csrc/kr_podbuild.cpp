// The whole Pod (assembled) — host side of the C ABI, no GPU.
//
// buildHeadPod / buildWorkerPod (ray-operator/controllers/ray/raycluster_controller.go:1387-1433) for every create tuple of one RayCluster:
// the group's template is copied as a JSON tree, the per-group builders of this library (kr_raytemplate.cpp, kr_raystart.cpp) answer each
// decision, their fragments are appended in the reference's order, and the tree is written back in Go's encoding (kr_json.hpp).  The
// container half depends only on the group: it is assembled once per group and call; kr_pod_meta_build supplies the ObjectMeta per tuple.
// Order followed (common/pod.go):
//   worker: wait-gcs-ready init container, a copy of the Ray container as the template has it        :359-415
//   head:   autoscaler sidecar, service account, autoscaler-v2 env + restartPolicy                   :194-220
//   GCS fault tolerance env, the head's redis-* rayStartParams                                         :222 / :443
//   the default metrics port                                                                           :224-232 / :445-453
//   worker: restartPolicy Never under autoscaler v2                                                    :455-457
//   token auth on the Ray container, the token volume, wait-gcs-ready                                  :234-236 / :459-461
//   operator-configured sidecars                                            raycluster_controller.go:1397-1399, 1424-1426
//   BuildPod: emptyDir volumes, `ray start` command, init-container env, Ray container env, probes     :577-669
// Nothing is copied from the reference; kuberay_b200/podbuilder.py is the same assembly in Python (its test cross-check).
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/kr_engine.h"
#include "kr_json.hpp"

namespace {
using namespace krjson;

thread_local std::string g_err;

inline kr_str ks(const std::string &s) { return kr_str{s.data(), (uint32_t)s.size()}; }
inline kr_str ks_opt(const Node *n) { return (n && n->t == N_STR) ? ks(n->s) : kr_str{nullptr, 0}; }
const std::string kEmpty;

const Node *child(const Node *n, const char *k) { return (n && n->t == N_OBJ) ? n->get(k) : nullptr; }
Node *child(Node *n, const char *k) { return (n && n->t == N_OBJ) ? n->get(k) : nullptr; }
const std::string &text(const Node *n) { return (n && n->t == N_STR) ? n->s : kEmpty; }
std::string lower(std::string s) { for (char &c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32); return s; }

Node &member(Node &obj, const char *k) {  // obj[k], created (as null) when missing
  if (obj.t != N_OBJ) { obj = Node(); obj.t = N_OBJ; }
  if (Node *n = obj.get(k)) return *n;
  obj.o.emplace_back(k, Node());
  return obj.o.back().second;
}
Node &array_member(Node &obj, const char *k) {
  Node &n = member(obj, k);
  if (n.t != N_ARR) { n = Node(); n.t = N_ARR; }
  return n;
}
void extend(Node &obj, const char *k, const Node *items) {  // append(obj[k], items...), a nil slice staying nil when there is nothing to add
  if (!items || items->t != N_ARR || items->a.empty()) return;
  Node &dst = array_member(obj, k);
  dst.a.insert(dst.a.end(), items->a.begin(), items->a.end());
}
Node str_node(const std::string &s) { Node n; n.t = N_STR; n.s = s; return n; }

struct Strs {  // kr_str views over strings that outlive the call
  std::vector<std::string> own;
  std::vector<kr_str> v;
  void add(const std::string &s) { own.push_back(s); }
  void seal() { v.clear(); for (auto &s : own) v.push_back(ks(s)); if (v.empty()) v.push_back(kr_str{nullptr, 0}); }
  uint32_t n() const { return (uint32_t)own.size(); }
};
Strs field_of_each(const Node *arr, const char *k) {  // arr[*][k] for the entries that are objects
  Strs s;
  if (arr && arr->t == N_ARR) for (const Node &e : arr->a) s.add(text(child(&e, k)));
  s.seal();
  return s;
}
struct Kvs {
  std::vector<kr_kv> v;
  explicit Kvs(const Node *m) {
    if (m && m->t == N_OBJ) for (auto &kv : m->o) if (kv.second.t == N_STR) v.push_back(kr_kv{ks(kv.first), ks(kv.second.s)});
    if (v.empty()) v.push_back(kr_kv{{nullptr, 0}, {nullptr, 0}});
  }
  uint32_t n(const Node *m) const { uint32_t c = 0; if (m && m->t == N_OBJ) for (auto &kv : m->o) if (kv.second.t == N_STR) c++; return c; }
};

std::string raw_text(const Node *n) {  // a tree fragment as JSON text for the builders that take raw corev1 fragments
  if (!n || n->t == N_NULL) return std::string();
  Emitter em;
  em.raw(*n);
  return em.out;
}

// Runs one of the JSON-producing builders (size, then fill) and parses its document.
template <class In, class Fn>
bool call_json(const char *who, Fn fn, const In &in, Node &out, const char *(*last_error)(void)) {
  uint64_t need = 0;
  int rc = fn(&in, nullptr, 0, &need);
  if (rc != KR_OK && rc != KR_E_CAPACITY) { g_err = std::string("kr_pod_build: ") + who + ": " + last_error(); return false; }
  std::string buf(need, '\0');
  rc = fn(&in, reinterpret_cast<uint8_t *>(&buf[0]), need, &need);
  if (rc != KR_OK) { g_err = std::string("kr_pod_build: ") + who + ": " + last_error(); return false; }
  Parser ps{buf.data(), buf.data() + buf.size(), {}};
  out = Node();
  if (!ps.value(out, 0)) { g_err = std::string("kr_pod_build: ") + who + " returned malformed JSON: " + ps.err; return false; }
  return true;
}

struct Cluster {  // what the assembly reads of the RayCluster, resolved once
  const Node *meta, *spec, *head_spec, *groups, *annots, *auth, *ft_opts, *auto_opts;
  std::string name, ns, uid, fqdn, head_port, secret_name;
  const std::string *head_redis_password;
  bool autoscaling, auto_v2, auth_on, k8s_auth, ft;
  uint8_t crd;
};

bool json_true(const Node *n) { return n && n->t == N_BOOL && n->b; }

// The PodSpec of one group (g < 0: the head), in Go's encoding.
bool build_spec(const Cluster &c, const kr_podbuild_env &env, long g, bool overwrite_cmd, std::string &out) {
  const bool head = g < 0;
  const uint8_t node = head ? KR_NT_HEAD : KR_NT_WORKER;
  const Node *grp = head ? c.head_spec : &c.groups->a[(size_t)g];
  const Node *tspec = child(child(grp, "template"), "spec");
  Node pspec;
  if (tspec && tspec->t == N_OBJ) pspec = *tspec; else pspec.t = N_OBJ;
  // every member the assembly may add exists (as null = absent) before any reference into the tree is taken, so the trees' member vectors never grow under one
  for (const char *k : {"containers", "initContainers", "volumes", "restartPolicy", "serviceAccountName"}) member(pspec, k);
  Node &containers = array_member(pspec, "containers");
  if (containers.a.empty() || containers.a[0].t != N_OBJ) { g_err = "kr_pod_build: the group's template has no Ray container (spec.containers[0])"; return false; }
  for (const char *k : {"env", "ports", "volumeMounts", "command", "args", "livenessProbe", "readinessProbe"}) member(containers.a[0], k);
  const bool login = env.login_shell != 0;
  Node got;

  // ---- DefaultWorkerPodTemplate / DefaultHeadPodTemplate
  if (!head && !env.no_init_container_injection) {
    const Node &ray = containers.a[0];
    kr_rayinit_in in{};
    in.login_shell = login;
    in.image = ks(text(child(&ray, "image"))); in.image_pull_policy = ks(text(child(&ray, "imagePullPolicy")));
    in.fqdn_ray_ip = ks(c.fqdn); in.head_port = ks(c.head_port);
    if (!call_json("kr_ray_init_container", kr_ray_init_container, in, got, kr_ray_template_last_error)) return false;
    for (const char *k : {"env", "volumeMounts", "securityContext"})  // the Ray container's, copied as the template has them (:392-397)
      if (const Node *v = child(&ray, k)) if (v->t != N_NULL) member(got, k) = *v;
    array_member(pspec, "initContainers").a.push_back(got);
  }
  if (head && c.autoscaling) {
    const std::string resources = raw_text(child(c.auto_opts, "resources")), envj = raw_text(child(c.auto_opts, "env")), env_from = raw_text(child(c.auto_opts, "envFrom")),
                      mounts = raw_text(child(c.auto_opts, "volumeMounts")), sec = raw_text(child(c.auto_opts, "securityContext"));
    kr_rayautoscaler_in in{};
    in.login_shell = login; in.autoscaler_v2 = c.auto_v2; in.auth_enabled = c.auth_on; in.k8s_token_auth = c.k8s_auth;
    in.has_options = c.auto_opts && c.auto_opts->t == N_OBJ;
    in.cluster_name = ks(c.name); in.secret_name = ks(c.secret_name);
    in.head_service_account = ks(text(child(&pspec, "serviceAccountName")));
    in.ray_image = ks(text(child(&containers.a[0], "image")));
    in.image = ks_opt(child(c.auto_opts, "image")); in.image_pull_policy = ks_opt(child(c.auto_opts, "imagePullPolicy"));
    in.resources_json = ks(resources); in.env_json = ks(envj); in.env_from_json = ks(env_from); in.volume_mounts_json = ks(mounts); in.security_context_json = ks(sec);
    if (!call_json("kr_ray_autoscaler_container", kr_ray_autoscaler_container, in, got, kr_ray_template_last_error)) return false;
    member(pspec, "serviceAccountName") = str_node(text(child(&got, "serviceAccountName")));
    Node &cs = array_member(pspec, "containers");
    if (const Node *side = child(&got, "container")) cs.a.push_back(*side);
    extend(cs.a[0], "env", child(&got, "rayContainerEnv"));
    if (!text(child(&got, "restartPolicy")).empty()) member(pspec, "restartPolicy") = str_node(text(child(&got, "restartPolicy")));
  }
  Node &ray = array_member(pspec, "containers").a[0];
  Node params;  // the group's rayStartParams as BuildPod will see them
  if (const Node *p = child(grp, "rayStartParams")) if (p->t == N_OBJ) params = *p;
  if (params.t != N_OBJ) params.t = N_OBJ;
  {
    Strs existing = field_of_each(child(&ray, "env"), "name");
    const Node *user = child(c.ft_opts, "redisUsername"), *pass = child(c.ft_opts, "redisPassword");
    const std::string user_from = raw_text(child(user, "valueFrom")), pass_from = raw_text(child(pass, "valueFrom"));
    const Node *ann_ns = child(c.annots, "ray.io/external-storage-namespace");
    kr_rayft_in in{};
    in.node_type = node; in.ft_enabled = c.ft; in.has_options = c.ft_opts && c.ft_opts->t == N_OBJ;
    in.has_redis_username = user && user->t == N_OBJ; in.has_redis_password = pass && pass->t == N_OBJ;
    in.cluster_uid = ks(c.uid); in.storage_ns_annotation = ks_opt(ann_ns);
    in.storage_ns_option = ks(text(child(c.ft_opts, "externalStorageNamespace"))); in.redis_address = ks(text(child(c.ft_opts, "redisAddress")));
    in.redis_username_value = ks(text(child(user, "value"))); in.redis_username_value_from = ks(user_from);
    in.redis_password_value = ks(text(child(pass, "value"))); in.redis_password_value_from = ks(pass_from);
    in.head_redis_password_param = c.head_redis_password ? ks(*c.head_redis_password) : kr_str{nullptr, 0};
    in.existing = existing.v.data(); in.n_existing = existing.n();
    if (!call_json("kr_ray_ft_env", kr_ray_ft_env, in, got, kr_ray_template_last_error)) return false;
    extend(ray, "env", child(&got, "env"));
    if (head) if (const Node *add = child(&got, "rayStartParams")) for (auto &kv : add->o) member(params, kv.first.c_str()) = kv.second;
  }
  {
    bool have_metrics = false;
    if (const Node *ports = child(&ray, "ports")) if (ports->t == N_ARR) for (const Node &p : ports->a) have_metrics |= text(child(&p, "name")) == "metrics";
    if (!have_metrics) {
      Node port; port.t = N_OBJ;
      member(port, "name") = str_node("metrics");
      Node num; num.t = N_NUM; num.s = "8080";
      member(port, "containerPort") = num;
      array_member(ray, "ports").a.push_back(port);
    }
  }
  if (!head && c.autoscaling && c.auto_v2) member(pspec, "restartPolicy") = str_node("Never");
  if (c.auth_on) {
    std::vector<Node *> targets{&ray};
    if (Node *inits = child(&pspec, "initContainers")) if (inits->t == N_ARR) for (Node &ic : inits->a) if (text(child(&ic, "name")) == "wait-gcs-ready") targets.push_back(&ic);
    for (Node *t : targets) {
      Strs envs = field_of_each(child(t, "env"), "name"), mounts = field_of_each(child(t, "volumeMounts"), "name"), vols = field_of_each(child(&pspec, "volumes"), "name");
      kr_rayauth_in in{};
      in.k8s_token_auth = c.k8s_auth; in.cluster_name = ks(c.name); in.secret_name = ks(c.secret_name);
      in.existing_env = envs.v.data(); in.n_existing_env = envs.n();
      in.existing_mount_names = mounts.v.data(); in.n_existing_mount_names = mounts.n();
      in.existing_volume_names = vols.v.data(); in.n_existing_volume_names = vols.n();
      if (!call_json("kr_ray_auth", kr_ray_auth, in, got, kr_ray_template_last_error)) return false;
      extend(*t, "env", child(&got, "env"));
      extend(*t, "volumeMounts", child(&got, "volumeMounts"));
      extend(pspec, "volumes", child(&got, "volumes"));
    }
  }
  {
    const kr_str side = head ? env.head_sidecars_json : env.worker_sidecars_json;
    if (side.p && side.n) {
      Parser ps{side.p, side.p + side.n, {}};
      Node extra;
      if (!ps.value(extra, 0) || (extra.t != N_ARR && extra.t != N_NULL)) { g_err = "kr_pod_build: the sidecar containers must be a JSON array"; return false; }
      extend(pspec, "containers", &extra);
    }
  }

  // ---- BuildPod
  Node &cs = array_member(pspec, "containers");
  Node &rayc = cs.a[0];
  Node *side = nullptr;
  if (head && c.autoscaling) {
    for (Node &k : cs.a) if (text(child(&k, "name")) == "autoscaler") { side = &k; break; }
    if (!side) { g_err = "kr_pod_build: the autoscaler container is missing (getAutoscalerContainerIndex panics here)"; return false; }
  }
  const Node *limits = child(child(&rayc, "resources"), "limits"), *requests = child(child(&rayc, "resources"), "requests");
  auto qty = [](const Node *m, const char *k) -> kr_str { const Node *v = child(m, k); return (v && (v->t == N_STR || v->t == N_NUM)) ? ks(v->s) : kr_str{nullptr, 0}; };
  {
    Strs vols = field_of_each(child(&pspec, "volumes"), "name"), rm = field_of_each(child(&rayc, "volumeMounts"), "mountPath"), am = field_of_each(child(side, "volumeMounts"), "mountPath");
    kr_rayvol_in in{};
    in.node_type = node; in.autoscaling_enabled = c.autoscaling; in.plasma_directory_set = params.get("plasma-directory") != nullptr;
    in.memory_limit = qty(limits, "memory"); in.memory_request = qty(requests, "memory");
    in.volume_names = vols.v.data(); in.n_volume_names = vols.n();
    in.ray_mount_paths = rm.v.data(); in.n_ray_mount_paths = rm.n();
    in.autoscaler_mount_paths = am.v.data(); in.n_autoscaler_mount_paths = am.n();
    if (!call_json("kr_ray_volumes", kr_ray_volumes, in, got, kr_ray_start_last_error)) return false;
    extend(pspec, "volumes", child(&got, "volumes"));
    extend(rayc, "volumeMounts", child(&got, "rayContainerVolumeMounts"));
    if (side) extend(*side, "volumeMounts", child(&got, "autoscalerVolumeMounts"));
  }
  Node rs;
  {
    auto quantities = [](const Node *m, std::vector<std::string> &own, std::vector<kr_kv> &kv) {  // a ResourceList: values may be JSON numbers
      if (m && m->t == N_OBJ) for (auto &e : m->o) if (e.second.t == N_STR || e.second.t == N_NUM) own.push_back(e.second.s);
      size_t i = 0;
      if (m && m->t == N_OBJ) for (auto &e : m->o) if (e.second.t == N_STR || e.second.t == N_NUM) { kv.push_back(kr_kv{ks(e.first), ks(own[i])}); i++; }
      if (kv.empty()) kv.push_back(kr_kv{{nullptr, 0}, {nullptr, 0}});
    };
    std::vector<std::string> lo, ro;
    std::vector<kr_kv> lim, req;
    quantities(limits, lo, lim); quantities(requests, ro, req);
    Kvs p(&params), gl(child(grp, "labels")), gr(child(grp, "resources"));
    auto strs = [](const Node *arr) { Strs s; if (arr && arr->t == N_ARR) for (const Node &e : arr->a) s.add(e.t == N_STR ? e.s : std::string()); s.seal(); return s; };
    Strs cmd = strs(child(&rayc, "command")), args = strs(child(&rayc, "args"));
    kr_raystart_in in{};
    in.node_type = node; in.autoscaling_enabled = c.autoscaling; in.overwrite_container_cmd = overwrite_cmd; in.login_shell = login;
    in.head_port = ks(c.head_port); in.fqdn_ray_ip = ks(c.fqdn);
    in.ray_start_params = p.v.data(); in.n_ray_start_params = p.n(&params);
    in.group_labels = gl.v.data(); in.n_group_labels = gl.n(child(grp, "labels"));
    in.group_resources = gr.v.data(); in.n_group_resources = gr.n(child(grp, "resources"));
    in.container_limits = lim.data(); in.n_container_limits = (uint32_t)lo.size();
    in.container_requests = req.data(); in.n_container_requests = (uint32_t)ro.size();
    in.command = cmd.v.data(); in.n_command = cmd.n();
    in.args = args.v.data(); in.n_args = args.n();
    if (!call_json("kr_ray_start_command", kr_ray_start_command, in, rs, kr_ray_start_last_error)) return false;
    if (json_true(child(&rs, "generated"))) {
      member(rayc, "command") = *child(&rs, "command");
      member(rayc, "args") = *child(&rs, "args");
    }
  }
  const std::string kuberay_version(env.kuberay_version.p ? env.kuberay_version.p : "", env.kuberay_version.p ? env.kuberay_version.n : 0);
  if (Node *inits = child(&pspec, "initContainers")) if (inits->t == N_ARR)
    for (Node &ic : inits->a) {
      kr_rayenv_in in{};
      in.node_type = node; in.init_container = 1; in.fqdn_ray_ip = ks(c.fqdn); in.head_port = ks(c.head_port);
      if (!call_json("kr_ray_container_env", kr_ray_container_env, in, got, kr_ray_start_last_error)) return false;
      extend(ic, "env", &got);
    }
  {
    Strs existing = field_of_each(child(&rayc, "env"), "name");
    kr_rayenv_in in{};
    in.node_type = node; in.crd_type = c.crd; in.fqdn_ray_ip = ks(c.fqdn); in.head_port = ks(c.head_port);
    in.ray_start_cmd = ks(text(child(&rs, "rayStartCommand"))); in.kuberay_version = ks(kuberay_version);
    in.existing = existing.v.data(); in.n_existing = existing.n();
    in.default_envs = env.default_envs; in.n_default_envs = env.n_default_envs;
    if (!call_json("kr_ray_container_env", kr_ray_container_env, in, got, kr_ray_start_last_error)) return false;
    extend(rayc, "env", &got);
  }
  if (!env.no_probes_injection) {
    int serve = 8000;
    if (const Node *ports = child(&rayc, "ports")) if (ports->t == N_ARR)
      for (const Node &p : ports->a) if (text(child(&p, "name")) == "serve") { const Node *cp = child(&p, "containerPort"); serve = (cp && cp->t == N_NUM) ? atoi(cp->s.c_str()) : 8000; break; }
    auto has = [&](const char *k) { const Node *v = child(&rayc, k); return v && v->t != N_NULL; };
    Kvs final_params(child(&rs, "rayStartParams"));
    kr_rayprobe_in in{};
    in.node_type = node; in.crd_type = c.crd; in.has_liveness_probe = has("livenessProbe"); in.has_readiness_probe = has("readinessProbe");
    in.serving_port = serve; in.ray_version = ks(text(child(c.spec, "rayVersion")));
    in.ray_start_params = final_params.v.data(); in.n_ray_start_params = final_params.n(child(&rs, "rayStartParams"));
    if (!call_json("kr_ray_probes", kr_ray_probes, in, got, kr_ray_start_last_error)) return false;
    for (auto &kv : got.o) member(rayc, kv.first.c_str()) = kv.second;
  }
  Emitter em;
  em.strct("PodSpec", &pspec);
  if (!em.err.empty()) { g_err = "kr_pod_build: " + em.err; return false; }
  out.swap(em.out);
  return true;
}

}  // namespace

extern "C" {

const char *kr_pod_build_last_error(void) { return g_err.c_str(); }

int kr_pod_build(const uint8_t *cluster_json, uint64_t len, const kr_podbuild_env *envp, const kr_podmeta_create *creates, uint32_t n_creates,
                 uint8_t *out, uint64_t cap, uint64_t *off, uint64_t *need) {
  if (!cluster_json || !envp || !need || !off || (!creates && n_creates)) { g_err = "kr_pod_build: null argument"; return KR_E_INVALID; }
  const kr_podbuild_env &env = *envp;
  Parser ps{reinterpret_cast<const char *>(cluster_json), reinterpret_cast<const char *>(cluster_json) + len, {}};
  Node root;
  if (!ps.value(root, 0) || root.t != N_OBJ) { g_err = "kr_pod_build: the RayCluster must be a JSON object" + (ps.err.empty() ? std::string() : ": " + ps.err); return KR_E_INVALID; }
  Cluster c{};
  c.meta = child(&root, "metadata"); c.spec = child(&root, "spec");
  c.head_spec = child(c.spec, "headGroupSpec"); c.groups = child(c.spec, "workerGroupSpecs");
  c.annots = child(c.meta, "annotations");
  c.name = text(child(c.meta, "name")); c.ns = text(child(c.meta, "namespace")); c.uid = text(child(c.meta, "uid"));
  if (c.name.empty()) { g_err = "kr_pod_build: metadata.name is empty"; return KR_E_INVALID; }
  if (c.ns.empty()) c.ns = "default";
  const size_t n_groups = (c.groups && c.groups->t == N_ARR) ? c.groups->a.size() : 0;
  const Node *head_params = child(c.head_spec, "rayStartParams");
  c.head_port = child(head_params, "port") && child(head_params, "port")->t == N_STR ? child(head_params, "port")->s : "6379";  // common.GetHeadPort
  const Node *rp = child(head_params, "redis-password");
  c.head_redis_password = (rp && rp->t == N_STR) ? &rp->s : nullptr;
  {  // utils.GenerateFQDNServiceName (utils/util.go:313-337)
    std::string svc = text(child(child(child(c.head_spec, "headService"), "metadata"), "name"));
    if (svc.empty()) svc = c.name + "-head-svc";
    const std::string domain = (env.cluster_domain.p && env.cluster_domain.n) ? std::string(env.cluster_domain.p, env.cluster_domain.n) : "cluster.local";
    c.fqdn = svc + "." + c.ns + ".svc." + domain;
  }
  c.autoscaling = json_true(child(c.spec, "enableInTreeAutoscaling"));
  c.auto_opts = child(c.spec, "autoscalerOptions");
  if (c.auto_opts && c.auto_opts->t != N_OBJ) c.auto_opts = nullptr;
  c.auto_v2 = text(child(c.auto_opts, "version")) == "v2";
  c.auth = child(c.spec, "authOptions");
  if (c.auth && c.auth->t != N_OBJ) c.auth = nullptr;
  c.auth_on = text(child(c.auth, "mode")) == "token";
  c.k8s_auth = json_true(child(c.auth, "enableK8sTokenAuth"));
  c.secret_name = text(child(c.auth, "secretName"));
  c.ft_opts = child(c.spec, "gcsFaultToleranceOptions");
  if (c.ft_opts && c.ft_opts->t != N_OBJ) c.ft_opts = nullptr;
  const Node *ft_ann = child(c.annots, "ray.io/ft-enabled");
  c.ft = (ft_ann && lower(text(ft_ann)) == "true") || c.ft_opts != nullptr;
  const std::string &origin = text(child(child(c.meta, "labels"), "ray.io/originated-from-crd"));
  c.crd = origin == "RayService" ? KR_CRD_RAYSERVICE : origin == "RayJob" ? KR_CRD_RAYJOB : KR_CRD_RAYCLUSTER;
  const Node *ow = child(c.annots, "ray.io/overwrite-container-cmd");
  const bool cluster_overwrite = ow && lower(text(ow)) == "true";

  // ---- ObjectMeta of every tuple (kr_pod_meta_build)
  kr_podmeta_cluster pc{};
  pc.name = ks(c.name); pc.ns = ks(c.ns); pc.uid = ks(c.uid);
  pc.cluster_hash = env.cluster_hash; pc.kuberay_version = env.kuberay_version;
  pc.storage_ns_annotation = ks_opt(child(c.annots, "ray.io/external-storage-namespace"));
  pc.storage_ns_option = ks(text(child(c.ft_opts, "externalStorageNamespace")));
  pc.overwrite_container_cmd = cluster_overwrite; pc.ft_enabled = c.ft; pc.crd_type = c.crd;
  pc.deterministic_head_name = env.deterministic_head_name; pc.gate_multihost_indexing = env.gate_multihost_indexing;
  std::vector<Kvs> keep;
  keep.reserve(3 * (n_groups + 1));
  auto group_struct = [&](const Node *grp, bool head) {
    kr_podmeta_group g{};
    const Node *tmeta = child(child(grp, "template"), "metadata");
    static const std::string kHeadGroup = "headgroup";
    g.group_name = head ? ks(kHeadGroup) : ks(text(child(grp, "groupName")));
    const Node *hosts = child(grp, "numOfHosts");
    g.num_of_hosts = (!head && hosts && hosts->t == N_NUM) ? atoi(hosts->s.c_str()) : 1;
    const Node *tl = child(tmeta, "labels"), *gl = child(grp, "labels"), *ta = child(tmeta, "annotations");
    keep.emplace_back(tl); g.template_labels = keep.back().v.data(); g.n_template_labels = keep.back().n(tl);
    keep.emplace_back(gl); g.group_labels = keep.back().v.data(); g.n_group_labels = keep.back().n(gl);
    keep.emplace_back(ta); g.template_annotations = keep.back().v.data(); g.n_template_annotations = keep.back().n(ta);
    return g;
  };
  const kr_podmeta_group head_group = group_struct(c.head_spec, true);
  std::vector<kr_podmeta_group> groups;
  for (size_t i = 0; i < n_groups; i++) groups.push_back(group_struct(&c.groups->a[i], false));
  if (groups.empty()) groups.push_back(kr_podmeta_group{});
  for (uint32_t i = 0; i < n_creates; i++)
    if (creates[i].group >= (int32_t)n_groups) { g_err = "kr_pod_build: a create tuple names a worker group the RayCluster does not have"; return KR_E_INVALID; }
  std::vector<uint64_t> moff(n_creates + 1, 0);
  uint64_t mneed = 0;
  int rc = kr_pod_meta_build(&pc, &head_group, groups.data(), (uint32_t)n_groups, creates, n_creates, nullptr, 0, moff.data(), &mneed);
  if (rc != KR_OK && rc != KR_E_CAPACITY) { g_err = std::string("kr_pod_build: kr_pod_meta_build: ") + kr_pod_meta_last_error(); return rc; }
  std::string metas(mneed, '\0');
  if (n_creates) {
    rc = kr_pod_meta_build(&pc, &head_group, groups.data(), (uint32_t)n_groups, creates, n_creates, reinterpret_cast<uint8_t *>(&metas[0]), mneed, moff.data(), &mneed);
    if (rc != KR_OK) { g_err = std::string("kr_pod_build: kr_pod_meta_build: ") + kr_pod_meta_last_error(); return rc; }
  }

  // ---- the PodSpec of each group that has a tuple, once
  std::vector<std::string> specs(n_groups + 1);
  std::vector<char> built(n_groups + 1, 0);
  std::string doc;
  for (uint32_t i = 0; i < n_creates; i++) {
    const long g = creates[i].group < 0 ? -1 : creates[i].group;
    const size_t slot = (size_t)(g + 1);
    if (!built[slot]) {
      const Node *grp = g < 0 ? c.head_spec : &c.groups->a[(size_t)g];
      const Node *tann = child(child(child(grp, "template"), "metadata"), "annotations");
      const Node *tow = child(tann, "ray.io/overwrite-container-cmd");
      // BuildPod reads the annotation off the template's metadata (common/pod.go:631-634): the user's own value, unless the RayCluster's "true" replaced it (:72-74)
      const bool overwrite = cluster_overwrite || (tow && lower(text(tow)) == "true");
      if (!build_spec(c, env, g, overwrite, specs[slot])) return KR_E_INVALID;
      built[slot] = 1;
    }
    off[i] = doc.size();
    doc += "{\"kind\":\"Pod\",\"apiVersion\":\"v1\",\"metadata\":";
    {
      // ObjectMeta: podTemplateSpec.ObjectMeta (common/pod.go:598) — the template's own metadata with what the builders decided laid over it.
      // The worker's name is cleared (:418); the head keeps a name its template brought when only generateName is set (:171-175).
      const Node *grp = g < 0 ? c.head_spec : &c.groups->a[(size_t)g];
      const Node *tmeta = child(child(grp, "template"), "metadata");
      bool plain = true;  // nothing but what the patch replaces: the patch is the metadata
      if (tmeta && tmeta->t == N_OBJ)
        for (auto &kv : tmeta->o) {
          const std::string &k = kv.first;
          if (k == "labels" || k == "annotations" || k == "namespace" || k == "ownerReferences" || kv.second.t == N_NULL) continue;
          if (k == "generateName" && !(g < 0 && env.deterministic_head_name && kv.second.t == N_STR && !kv.second.s.empty())) continue;
          if (k == "name" && (g >= 0 || env.deterministic_head_name || kv.second.t != N_STR || kv.second.s.empty())) continue;
          plain = false;
        }
      if (plain) doc.append(metas, moff[i], moff[i + 1] - moff[i]);
      else {
        Node base = *tmeta, patch;
        if (g >= 0) base.erase("name");
        Parser mp{metas.data() + moff[i], metas.data() + moff[i + 1], {}};
        if (!mp.value(patch, 0) || patch.t != N_OBJ) { g_err = "kr_pod_build: kr_pod_meta_build returned malformed JSON"; return KR_E_INVALID; }
        for (auto &kv : patch.o) member(base, kv.first.c_str()) = kv.second;
        Emitter em;
        em.strct("ObjectMeta", &base);
        if (!em.err.empty()) { g_err = "kr_pod_build: " + em.err; return KR_E_INVALID; }
        doc += em.out;
      }
    }
    doc += ",\"spec\":";
    doc += specs[slot];
    doc += ",\"status\":{}}";
  }
  off[n_creates] = doc.size();
  *need = doc.size();
  if (doc.size() > cap || (!out && !doc.empty())) { g_err = "kr_pod_build: output buffer too small"; return KR_E_CAPACITY; }
  if (!doc.empty()) memcpy(out, doc.data(), doc.size());
  return KR_OK;
}

}  // extern "C"
