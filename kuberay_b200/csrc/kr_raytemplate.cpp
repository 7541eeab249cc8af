// Pod template surgery (last part) — host side of the C ABI, no GPU.
//
// The pieces DefaultHeadPodTemplate / DefaultWorkerPodTemplate bolt onto a group's template before BuildPod runs, once per group and
// reconcile.  Restated from (never copied), paths relative to ray-operator/controllers/ray/:
//   configureGCSFaultTolerance      common/pod.go:77-163     kr_ray_ft_env
//   configureTokenAuth, AddRayTokenVolume, SetContainerTokenAuthEnvVars   common/pod.go:254-335     kr_ray_auth
//   BuildAutoscalerContainer, mergeAutoscalerOverrides, setAutoscalerV2EnvVars, the head's service account
//                                   common/pod.go:673-751, 242-251, 194-220; utils/util.go:575-581   kr_ray_autoscaler_container
//   the wait-gcs-ready init container of DefaultWorkerPodTemplate          common/pod.go:359-415     kr_ray_init_container
//   utils.GetContainerCommand       utils/util.go:884-892
// corev1 fragments the caller already holds travel as raw JSON (Go's encoding) and are spliced in unchanged; everything this file
// produces itself follows encoding/json: struct fields in declaration order, omitempty honoured, strings escaped as Go does.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kr_engine.h"

void kr_go_string_append(std::string &out, const std::string &s);  // kr_specjson.cpp

namespace {

thread_local std::string g_err;

inline std::string str(kr_str s) { return (s.p && s.n) ? std::string(s.p, s.n) : std::string(); }
inline bool present(kr_str s) { return s.p != nullptr; }

// A raw JSON fragment: trimmed; "" / "null" -> absent.
std::string raw(kr_str s) {
  std::string t = str(s);
  size_t a = 0, b = t.size();
  auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; };
  while (a < b && ws(t[a])) a++;
  while (b > a && ws(t[b - 1])) b--;
  t = t.substr(a, b - a);
  return t == "null" ? std::string() : t;
}
// The elements of a raw JSON array without the brackets ("" for an absent or empty list); false when the text is not an array.
bool raw_elements(kr_str s, std::string &out) {
  const std::string t = raw(s);
  out.clear();
  if (t.empty()) return true;
  if (t.size() < 2 || t.front() != '[' || t.back() != ']') return false;
  out = t.substr(1, t.size() - 2);
  if (out.find_first_not_of(" \t\r\n") == std::string::npos) out.clear();
  return true;
}
bool raw_object(kr_str s, std::string &out) {
  out = raw(s);
  return out.empty() || (out.size() >= 2 && out.front() == '{' && out.back() == '}');
}

struct Names {
  std::vector<std::string> v;
  Names(const kr_str *p, uint32_t n) { for (uint32_t i = 0; i < n; i++) v.push_back(str(p[i])); }
  bool has(const char *x) const { return std::find(v.begin(), v.end(), x) != v.end(); }
  void add(const char *x) { v.push_back(x); }
};

struct List {  // a JSON array under construction
  std::string js;
  void item(const std::string &x) { if (!js.empty()) js += ','; js += x; }
  void splice(const std::string &elements) { if (!elements.empty()) item(elements); }
  std::string done() const { return "[" + js + "]"; }
};

std::string env_value(const char *name, const std::string &value, const std::string &value_from = std::string()) {  // corev1.EnvVar
  std::string js = "{\"name\":";
  kr_go_string_append(js, name);
  if (!value.empty()) { js += ",\"value\":"; kr_go_string_append(js, value); }
  if (!value_from.empty()) js += ",\"valueFrom\":" + value_from;
  return js + "}";
}
std::string env_field(const char *name, const char *path) {
  std::string js = "{\"fieldRef\":{\"fieldPath\":";
  kr_go_string_append(js, path);
  return env_value(name, "", js + "}}");
}

std::string container_command(bool login_shell) {  // utils.GetContainerCommand([]string{})
  return login_shell ? "[\"/bin/bash\",\"-cl\",\"--\"]" : "[\"/bin/bash\",\"-c\",\"--\"]";
}

std::string check_name(const std::string &s) {  // utils.CheckName through the podmeta entry point
  if (s.empty()) return s;
  char buf[128];
  const int64_t n = kr_check_name(kr_str{s.data(), (uint32_t)s.size()}, buf, sizeof buf);
  return n > 0 ? std::string(buf, (size_t)std::min<int64_t>(n, sizeof buf)) : s;
}

// SetContainerTokenAuthEnvVars (common/pod.go:296-335) on one container.
void token_auth(bool k8s, const std::string &cluster, const std::string &secret_opt, Names &env, Names &mounts, List &jenv, List &jmounts) {
  if (!env.has("RAY_AUTH_MODE")) { jenv.item(env_value("RAY_AUTH_MODE", "token")); env.add("RAY_AUTH_MODE"); }
  if (k8s) {
    if (!env.has("RAY_ENABLE_K8S_TOKEN_AUTH")) { jenv.item(env_value("RAY_ENABLE_K8S_TOKEN_AUTH", "true")); env.add("RAY_ENABLE_K8S_TOKEN_AUTH"); }
    if (!mounts.has("ray-token")) {
      jmounts.item("{\"name\":\"ray-token\",\"readOnly\":true,\"mountPath\":\"/var/run/secrets/ray.io/serviceaccount\"}");
      mounts.add("ray-token");
    }
  } else if (!env.has("RAY_AUTH_TOKEN")) {
    const std::string secret = secret_opt.empty() ? check_name(cluster) : secret_opt;
    std::string ref = "{\"secretKeyRef\":{";
    if (!secret.empty()) { ref += "\"name\":"; kr_go_string_append(ref, secret); ref += ','; }  // LocalObjectReference.Name is omitempty
    ref += "\"key\":\"auth_token\"}}";
    jenv.item(env_value("RAY_AUTH_TOKEN", "", ref));
    env.add("RAY_AUTH_TOKEN");
  }
}

const char kTokenVolume[] = "{\"name\":\"ray-token\",\"projected\":{\"sources\":[{\"serviceAccountToken\":{\"path\":\"token\"}}]}}";

int finish(const char *who, const std::string &js, uint8_t *out, uint64_t cap, uint64_t *need) {
  *need = js.size();
  if (js.size() > cap || !out) { g_err = std::string(who) + ": output buffer too small"; return KR_E_CAPACITY; }
  memcpy(out, js.data(), js.size());
  return KR_OK;
}

}  // namespace

extern "C" {

const char *kr_ray_template_last_error(void) { return g_err.c_str(); }

int kr_ray_ft_env(const kr_rayft_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_ft_env: null argument"; return KR_E_INVALID; }
  if (in->node_type != KR_NT_HEAD && in->node_type != KR_NT_WORKER) { g_err = "kr_ray_ft_env: node_type must be KR_NT_HEAD or KR_NT_WORKER"; return KR_E_INVALID; }
  Names env(in->existing, in->n_existing);
  List jenv;
  std::string params;
  auto param = [&](const char *k, const char *v) { if (!params.empty()) params += ','; params += std::string("\"") + k + "\":\"" + v + "\""; };
  if (in->ft_enabled) {
    const bool head = in->node_type == KR_NT_HEAD;
    // workers wait 600 s for a restarted GCS instead of Ray's 60 s (:94-102)
    if (!head && !env.has("RAY_gcs_rpc_server_reconnect_timeout_s")) { jenv.item(env_value("RAY_gcs_rpc_server_reconnect_timeout_s", "600")); env.add("RAY_gcs_rpc_server_reconnect_timeout_s"); }
    if (head) {
      std::string ns = str(in->cluster_uid);                       // :107-113: UID, then the annotation, then the option
      if (present(in->storage_ns_annotation)) ns = str(in->storage_ns_annotation);
      if (in->has_options && in->storage_ns_option.p && in->storage_ns_option.n) ns = str(in->storage_ns_option);
      if (!env.has("RAY_external_storage_namespace")) { jenv.item(env_value("RAY_external_storage_namespace", ns)); env.add("RAY_external_storage_namespace"); }
      if (in->has_options) {                                       // :120-147: appended without "exists" checks
        std::string from;
        jenv.item(env_value("RAY_REDIS_ADDRESS", str(in->redis_address)));
        if (in->has_redis_username) {
          if (!raw_object(in->redis_username_value_from, from)) { g_err = "kr_ray_ft_env: redis_username_value_from is not a JSON object"; return KR_E_INVALID; }
          param("redis-username", "$REDIS_USERNAME");
          jenv.item(env_value("REDIS_USERNAME", str(in->redis_username_value), from));
        }
        if (in->has_redis_password) {
          if (!raw_object(in->redis_password_value_from, from)) { g_err = "kr_ray_ft_env: redis_password_value_from is not a JSON object"; return KR_E_INVALID; }
          param("redis-password", "$REDIS_PASSWORD");
          jenv.item(env_value("REDIS_PASSWORD", str(in->redis_password_value), from));
        }
      } else if (!env.has("REDIS_PASSWORD") && present(in->head_redis_password_param)) {
        // a password written straight into rayStartParams is mirrored for the Redis cleanup job (:148-159)
        jenv.item(env_value("REDIS_PASSWORD", str(in->head_redis_password_param)));
      }
    }
  }
  return finish("kr_ray_ft_env", "{\"env\":" + jenv.done() + ",\"rayStartParams\":{" + params + "}}", out, cap, need);
}

int kr_ray_auth(const kr_rayauth_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_auth: null argument"; return KR_E_INVALID; }
  Names env(in->existing_env, in->n_existing_env), mounts(in->existing_mount_names, in->n_existing_mount_names), vols(in->existing_volume_names, in->n_existing_volume_names);
  List jenv, jmounts, jvols;
  token_auth(in->k8s_token_auth != 0, str(in->cluster_name), str(in->secret_name), env, mounts, jenv, jmounts);
  if (in->k8s_token_auth && !vols.has("ray-token")) jvols.item(kTokenVolume);  // AddRayTokenVolume (:274-293)
  return finish("kr_ray_auth", "{\"env\":" + jenv.done() + ",\"volumeMounts\":" + jmounts.done() + ",\"volumes\":" + jvols.done() + "}", out, cap, need);
}

int kr_ray_autoscaler_container(const kr_rayautoscaler_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_autoscaler_container: null argument"; return KR_E_INVALID; }
  // BuildAutoscalerContainer (:673-724)
  std::string image = str(in->ray_image), pull = "IfNotPresent";
  std::string resources = "{\"limits\":{\"cpu\":\"500m\",\"memory\":\"512Mi\"},\"requests\":{\"cpu\":\"500m\",\"memory\":\"512Mi\"}}";
  std::string env_from, security;
  List jenv, jmounts;
  jenv.item(env_field("RAY_CLUSTER_NAME", "metadata.labels['ray.io/cluster']"));
  jenv.item(env_field("RAY_CLUSTER_NAMESPACE", "metadata.namespace"));
  jenv.item(env_field("RAY_HEAD_POD_NAME", "metadata.name"));
  jenv.item(env_value("KUBERAY_CRD_VER", "v1"));
  const kr_str base[] = {{"RAY_CLUSTER_NAME", 16}, {"RAY_CLUSTER_NAMESPACE", 21}, {"RAY_HEAD_POD_NAME", 17}, {"KUBERAY_CRD_VER", 15}};
  Names env(base, 4), mounts(nullptr, 0);
  if (in->auth_enabled) token_auth(in->k8s_token_auth != 0, str(in->cluster_name), str(in->secret_name), env, mounts, jenv, jmounts);  // :207-210
  if (in->has_options) {  // mergeAutoscalerOverrides (:727-751)
    std::string el;
    if (!raw_object(in->resources_json, el)) { g_err = "kr_ray_autoscaler_container: resources_json is not a JSON object"; return KR_E_INVALID; }
    if (!el.empty()) resources = el;
    if (present(in->image)) image = str(in->image);
    if (present(in->image_pull_policy)) pull = str(in->image_pull_policy);
    if (!raw_elements(in->env_json, el)) { g_err = "kr_ray_autoscaler_container: env_json is not a JSON array"; return KR_E_INVALID; }
    jenv.splice(el);
    if (!raw_elements(in->env_from_json, el)) { g_err = "kr_ray_autoscaler_container: env_from_json is not a JSON array"; return KR_E_INVALID; }
    env_from = el;
    if (!raw_elements(in->volume_mounts_json, el)) { g_err = "kr_ray_autoscaler_container: volume_mounts_json is not a JSON array"; return KR_E_INVALID; }
    jmounts.splice(el);
    if (!raw_object(in->security_context_json, security)) { g_err = "kr_ray_autoscaler_container: security_context_json is not a JSON object"; return KR_E_INVALID; }
  }
  std::string c = "{\"name\":\"autoscaler\"";  // corev1.Container, fields in declaration order
  if (!image.empty()) { c += ",\"image\":"; kr_go_string_append(c, image); }
  c += ",\"command\":" + container_command(in->login_shell != 0);
  c += ",\"args\":[\"ray kuberay-autoscaler --cluster-name $(RAY_CLUSTER_NAME) --cluster-namespace $(RAY_CLUSTER_NAMESPACE)\"]";
  if (!env_from.empty()) c += ",\"envFrom\":[" + env_from + "]";
  c += ",\"env\":" + jenv.done();
  c += ",\"resources\":" + resources;
  if (!jmounts.js.empty()) c += ",\"volumeMounts\":" + jmounts.done();
  if (!pull.empty()) { c += ",\"imagePullPolicy\":"; kr_go_string_append(c, pull); }
  if (!security.empty()) c += ",\"securityContext\":" + security;
  c += "}";
  // the head runs under the autoscaler's service account (:199-201); autoscaler v2 (:216-219)
  std::string sa = (in->head_service_account.p && in->head_service_account.n) ? str(in->head_service_account) : str(in->cluster_name);
  std::string js = "{\"container\":" + c + ",\"serviceAccountName\":";
  kr_go_string_append(js, check_name(sa));
  js += ",\"rayContainerEnv\":[";
  if (in->autoscaler_v2) js += env_value("RAY_enable_autoscaler_v2", "true");
  js += std::string("],\"restartPolicy\":") + (in->autoscaler_v2 ? "\"Never\"" : "\"\"") + "}";
  return finish("kr_ray_autoscaler_container", js, out, cap, need);
}

int kr_ray_init_container(const kr_rayinit_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_init_container: null argument"; return KR_E_INVALID; }
  const std::string addr = str(in->fqdn_ray_ip) + ":" + str(in->head_port);
  // the polling script (:372-390): quiet for two minutes, then it lets ray health-check print why it fails; the reference's raw string
  // keeps its source indentation, so does this
  const std::string t5(5, '\t'), t6(6, '\t'), t7(7, '\t'), t8(8, '\t');
  std::string sh = "\n";
  sh += t5 + "SECONDS=0\n";
  sh += t5 + "while true; do\n";
  sh += t6 + "if (( SECONDS <= 120 )); then\n";
  sh += t7 + "if ray health-check --address " + addr + " > /dev/null 2>&1; then\n";
  sh += t8 + "echo \"GCS is ready.\"\n";
  sh += t8 + "break\n";
  sh += t7 + "fi\n";
  sh += t7 + "echo \"$SECONDS seconds elapsed: Waiting for GCS to be ready.\"\n";
  sh += t6 + "else\n";
  sh += t7 + "if ray health-check --address " + addr + "; then\n";
  sh += t8 + "echo \"GCS is ready. Any error messages above can be safely ignored.\"\n";
  sh += t8 + "break\n";
  sh += t7 + "fi\n";
  sh += t7 + "echo \"$SECONDS seconds elapsed: Still waiting for GCS to be ready. For troubleshooting, refer to the FAQ at https://docs.ray.io/en/master/cluster/kubernetes/troubleshooting.html.\"\n";
  sh += t6 + "fi\n";
  sh += t6 + "sleep 5\n";
  sh += t5 + "done\n";
  sh += std::string(4, '\t');
  std::string env, mounts, security;
  if (!raw_elements(in->env_json, env)) { g_err = "kr_ray_init_container: env_json is not a JSON array"; return KR_E_INVALID; }
  if (!raw_elements(in->volume_mounts_json, mounts)) { g_err = "kr_ray_init_container: volume_mounts_json is not a JSON array"; return KR_E_INVALID; }
  if (!raw_object(in->security_context_json, security)) { g_err = "kr_ray_init_container: security_context_json is not a JSON object"; return KR_E_INVALID; }
  std::string c = "{\"name\":\"wait-gcs-ready\"";
  if (in->image.p && in->image.n) { c += ",\"image\":"; kr_go_string_append(c, str(in->image)); }
  c += ",\"command\":" + container_command(in->login_shell != 0) + ",\"args\":[";
  kr_go_string_append(c, sh);
  c += "]";
  if (!env.empty()) c += ",\"env\":[" + env + "]";
  // fixed and small: a ResourceQuota needs explicit numbers and GKE Autopilot rejects GPU init containers (:398-412)
  c += ",\"resources\":{\"limits\":{\"cpu\":\"200m\",\"memory\":\"256Mi\"},\"requests\":{\"cpu\":\"200m\",\"memory\":\"256Mi\"}}";
  if (!mounts.empty()) c += ",\"volumeMounts\":[" + mounts + "]";
  if (in->image_pull_policy.p && in->image_pull_policy.n) { c += ",\"imagePullPolicy\":"; kr_go_string_append(c, str(in->image_pull_policy)); }
  if (!security.empty()) c += ",\"securityContext\":" + security;
  c += "}";
  return finish("kr_ray_init_container", c, out, cap, need);
}

}  // extern "C"
