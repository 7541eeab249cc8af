// kr_json.hpp — the JSON tree, parser and Go-encoding emitter shared by the host-side builders (kr_specjson.cpp: the muted-spec emitter;
// kr_podbuild.cpp: the whole-Pod builder).  Internal to libkrengine.so; nothing here is part of the C ABI.
//
// The struct tables for rayv1 follow ray-operator/apis/ray/v1/raycluster_types.go:13-225 (in the reference tree); the corev1 tables come from
// k8s.io/api v0.36.0 (ray-operator/go.mod:23), which is not vendored there: they are restated from the published API (see kr_specjson.cpp).
#ifndef KR_JSON_HPP_
#define KR_JSON_HPP_
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace krjson {

// ------------------------------------------------------------------------------------------------ JSON tree
enum NodeType : uint8_t { N_NULL, N_BOOL, N_NUM, N_STR, N_ARR, N_OBJ };
struct Node {
  NodeType t = N_NULL;
  bool b = false;
  std::string s;                                        // N_STR: decoded UTF-8; N_NUM: the number's text
  std::vector<Node> a;                                  // N_ARR
  std::vector<std::pair<std::string, Node>> o;          // N_OBJ, input order
  const Node *get(const char *k) const {
    for (auto &kv : o) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  Node *get(const char *k) {
    for (auto &kv : o) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Node *get(const std::string &k) const {
    for (auto &kv : o) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  void erase(const char *k) {
    o.erase(std::remove_if(o.begin(), o.end(), [&](const std::pair<std::string, Node> &kv) { return kv.first == k; }), o.end());
  }
};

struct Parser {
  const char *p, *end;
  std::string err;
  bool fail(const char *m) { if (err.empty()) err = m; return false; }
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  static void utf8(std::string &out, uint32_t cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 63)); out += (char)(0x80 | (cp & 63)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 63)); out += (char)(0x80 | ((cp >> 6) & 63)); out += (char)(0x80 | (cp & 63)); }
  }
  bool hex4(uint32_t &v) {
    if (end - p < 4) return fail("short \\u escape");
    v = 0;
    for (int i = 0; i < 4; i++) {
      char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return fail("bad \\u escape");
    }
    return true;
  }
  bool str(std::string &out) {
    if (p >= end || *p != '"') return fail("expected string");
    p++;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) return fail("short escape");
        char c = *p++;
        switch (c) {
          case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
          case 'r': out += '\r'; break; case 't': out += '\t'; break;
          case 'u': {
            uint32_t cp;
            if (!hex4(cp)) return false;
            if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {  // surrogate pair
              const char *save = p;
              p += 2;
              uint32_t lo;
              if (!hex4(lo)) return false;
              if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              else { p = save; cp = 0xFFFD; }
            } else if (cp >= 0xD800 && cp < 0xE000) cp = 0xFFFD;  // lone surrogate: encoding/json decodes it to U+FFFD
            utf8(out, cp);
            break;
          }
          default: return fail("bad escape");
        }
      } else {
        const char *q = p + 1;
        while (q < end && *q != '"' && *q != '\\') q++;
        out.append(p, q);
        p = q;
      }
    }
    if (p >= end) return fail("unterminated string");
    p++;
    return true;
  }
  bool value(Node &n, int depth) {
    if (depth > 200) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("unexpected end");
    char c = *p;
    if (c == '{') {
      n.t = N_OBJ; p++; ws();
      if (p < end && *p == '}') { p++; return true; }
      while (true) {
        ws();
        std::string k;
        if (!str(k)) return false;
        ws();
        if (p >= end || *p != ':') return fail("expected ':'");
        p++;
        // a repeated key: the last one wins (encoding/json); the value is parsed in place
        Node *slot = nullptr;
        for (auto &kv : n.o) if (kv.first == k) { kv.second = Node(); slot = &kv.second; break; }
        if (!slot) {
          if (n.o.empty()) n.o.reserve(6);
          n.o.emplace_back(std::move(k), Node());
          slot = &n.o.back().second;
        }
        if (!value(*slot, depth + 1)) return false;
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return true; }
        return fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      n.t = N_ARR; p++; ws();
      if (p < end && *p == ']') { p++; return true; }
      while (true) {
        if (n.a.empty()) n.a.reserve(4);
        n.a.emplace_back();
        if (!value(n.a.back(), depth + 1)) return false;
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return true; }
        return fail("expected ',' or ']'");
      }
    }
    if (c == '"') { n.t = N_STR; return str(n.s); }
    if (end - p >= 4 && !memcmp(p, "true", 4)) { n.t = N_BOOL; n.b = true; p += 4; return true; }
    if (end - p >= 5 && !memcmp(p, "false", 5)) { n.t = N_BOOL; n.b = false; p += 5; return true; }
    if (end - p >= 4 && !memcmp(p, "null", 4)) { n.t = N_NULL; p += 4; return true; }
    if (c == '-' || (c >= '0' && c <= '9')) {
      const char *q = p;
      if (*q == '-') q++;
      while (q < end && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-')) q++;
      n.t = N_NUM; n.s.assign(p, q); p = q;
      return true;
    }
    return fail("unexpected character");
  }
};

// ------------------------------------------------------------------------------------------------ Go string encoding
inline void go_string(std::string &out, const std::string &s) {  // encoding/json encodeState.string with escapeHTML = true
  static const char *hex = "0123456789abcdef";
  out += '"';
  size_t i = 0, n = s.size();
  while (i < n) {
    unsigned char c = (unsigned char)s[i];
    if (c >= 0x20 && c < 0x80 && c != '"' && c != '\\' && c != '<' && c != '>' && c != '&') {  // the common case: a run of bytes written as they are
      size_t j = i + 1;
      while (j < n) { const unsigned char d = (unsigned char)s[j]; if (d < 0x20 || d >= 0x80 || d == '"' || d == '\\' || d == '<' || d == '>' || d == '&') break; j++; }
      out.append(s, i, j - i);
      i = j;
      continue;
    }
    if (c < 0x80) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\b': out += "\\b"; break;
        case '\f': out += "\\f"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
          if (c < 0x20 || c == '<' || c == '>' || c == '&') { out += "\\u00"; out += hex[c >> 4]; out += hex[c & 15]; }
          else out += (char)c;
      }
      i++;
      continue;
    }
    // multi-byte: validate; invalid bytes become U+FFFD, U+2028 / U+2029 are escaped
    int len = (c >= 0xF0 && c <= 0xF4) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC2 && c < 0xE0) ? 2 : 0;
    bool ok = len && i + len <= n;
    uint32_t cp = 0;
    if (ok) {
      cp = c & (0xFF >> (len + 1));
      for (int k = 1; k < len; k++) { unsigned char d = (unsigned char)s[i + k]; if ((d & 0xC0) != 0x80) { ok = false; break; } cp = (cp << 6) | (d & 63); }
      if (ok && ((len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp < 0xE000))) || (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)))) ok = false;
    }
    if (!ok) { out += "\\ufffd"; i++; continue; }
    if (cp == 0x2028 || cp == 0x2029) { out += "\\u202"; out += hex[cp & 15]; }
    else out.append(s, i, len);
    i += len;
  }
  out += '"';
}

// ------------------------------------------------------------------------------------------------ resource.Quantity
// Canonical string of a quantity (apimachinery resource.Quantity.String()).  Returns false when the text is not a quantity (the
// caller then emits the text unchanged).
inline bool canon_quantity(const std::string &in, std::string &out) {
  size_t i = 0, n = in.size();
  bool neg = false;
  if (i < n && (in[i] == '+' || in[i] == '-')) { neg = in[i] == '-'; i++; }
  __int128 m = 0;
  int e10 = 0, digits = 0;
  bool seen_dot = false;
  for (; i < n; i++) {
    char c = in[i];
    if (c >= '0' && c <= '9') {
      if (m > ((__int128)1 << 100)) return false;
      m = m * 10 + (c - '0'); digits++;
      if (seen_dot) e10--;
    } else if (c == '.' && !seen_dot) seen_dot = true;
    else break;
  }
  if (!digits) return false;
  std::string suf = in.substr(i);
  enum { DEC_SI, BIN_SI, DEC_EXP } fmt = DEC_SI;
  int bin_pow = 0;
  if (suf.empty()) {}
  else if (suf == "Ki") { fmt = BIN_SI; bin_pow = 1; } else if (suf == "Mi") { fmt = BIN_SI; bin_pow = 2; }
  else if (suf == "Gi") { fmt = BIN_SI; bin_pow = 3; } else if (suf == "Ti") { fmt = BIN_SI; bin_pow = 4; }
  else if (suf == "Pi") { fmt = BIN_SI; bin_pow = 5; } else if (suf == "Ei") { fmt = BIN_SI; bin_pow = 6; }
  else if (suf == "n") e10 -= 9; else if (suf == "u") e10 -= 6; else if (suf == "m") e10 -= 3;
  else if (suf == "k") e10 += 3; else if (suf == "M") e10 += 6; else if (suf == "G") e10 += 9;
  else if (suf == "T") e10 += 12; else if (suf == "P") e10 += 15; else if (suf == "E") e10 += 18;
  else if (suf[0] == 'e' || suf[0] == 'E') {
    size_t k = 1;
    bool eneg = false;
    if (k < suf.size() && (suf[k] == '+' || suf[k] == '-')) { eneg = suf[k] == '-'; k++; }
    if (k >= suf.size()) return false;
    int ev = 0;
    for (; k < suf.size(); k++) { if (suf[k] < '0' || suf[k] > '9' || ev > 100) return false; ev = ev * 10 + (suf[k] - '0'); }
    e10 += eneg ? -ev : ev;
    fmt = DEC_EXP;
  } else return false;
  if (m == 0) { out = "0"; return true; }
  for (int k = 0; k < bin_pow; k++) { if (m > ((__int128)1 << 110)) return false; m *= 1024; }
  auto dec_to_str = [](__int128 v) { std::string s; if (v == 0) s = "0"; while (v > 0) { s += (char)('0' + (int)(v % 10)); v /= 10; } std::reverse(s.begin(), s.end()); return s; };
  if (fmt == BIN_SI) {
    // exact integer at least 1024 in magnitude -> largest power of 1024 that divides it; otherwise shown as DecimalSI
    __int128 v = m;
    int e = e10;
    bool exact = true;
    while (e < 0) { if (v % 10) { exact = false; break; } v /= 10; e++; }
    if (exact) { while (e > 0) { if (v > ((__int128)1 << 120)) return false; v *= 10; e--; } }
    if (exact && v >= 1024) {
      static const char *bs[] = {"", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei"};
      int p = 0;
      while (p < 6 && v % 1024 == 0) { v /= 1024; p++; }
      out = (neg ? "-" : "") + dec_to_str(v) + bs[p];
      return true;
    }
    fmt = DEC_SI;
  }
  // mantissa without trailing zeros, exponent a multiple of 3 (never below nano: the API rounds up there; not restated)
  while (m % 10 == 0) { m /= 10; e10++; }
  while (e10 % 3 != 0) { m *= 10; e10--; }
  while (e10 > 18) { m *= 1000; e10 -= 3; }
  if (e10 < -9) return false;
  if (fmt == DEC_EXP) {
    out = (neg ? "-" : "") + dec_to_str(m);
    if (e10) out += "e" + std::to_string(e10);
    return true;
  }
  static const char *ds[] = {"n", "u", "m", "", "k", "M", "G", "T", "P", "E"};
  out = (neg ? "-" : "") + dec_to_str(m) + ds[(e10 + 9) / 3];
  return true;
}

// ------------------------------------------------------------------------------------------------ struct tables
// "Type = field<sep>kind ..." — <sep> '?' = omitempty, ':' = always.  kinds: bool int string (values); *bool *int *string
// (pointers); map (map[string]string); mapq (ResourceList); quantity; intstr; []string; []int; raw (caller's order);
// T / *T / []T for a struct type T; "+T" inlines T's fields at this position (embedded struct).
static const char *const kSchema[] = {
    // ---- ray-operator/apis/ray/v1/raycluster_types.go:13-225
    "RayClusterSpec = upgradeStrategy?*RayClusterUpgradeStrategy authOptions?*AuthOptions suspend?*bool managedBy?*string autoscalerOptions?*AutoscalerOptions "
    "headServiceAnnotations?map enableInTreeAutoscaling?*bool gcsFaultToleranceOptions?*GcsFaultToleranceOptions headGroupSpec:HeadGroupSpec rayVersion?string "
    "workerGroupSpecs?[]WorkerGroupSpec",
    "RayClusterUpgradeStrategy = type?*string",
    "AuthOptions = enableK8sTokenAuth?*bool secretName?*string mode?string",
    "GcsFaultToleranceOptions = redisUsername?*RedisCredential redisPassword?*RedisCredential externalStorageNamespace?string redisAddress:string",
    "RedisCredential = valueFrom?*EnvVarSource value?string",
    "HeadGroupSpec = template:PodTemplateSpec headService?*Service enableIngress?*bool resources?map labels?map rayStartParams:map serviceType?string",
    "WorkerGroupSpec = suspend?*bool groupName:string replicas?*int minReplicas:*int maxReplicas:*int idleTimeoutSeconds?*int resources?map labels?map "
    "rayStartParams:map template:PodTemplateSpec scaleStrategy?ScaleStrategy numOfHosts?int",
    "ScaleStrategy = workersToDelete?[]string",
    "AutoscalerOptions = resources?*ResourceRequirements image?*string imagePullPolicy?*string securityContext?*SecurityContext idleTimeoutSeconds?*int "
    "upscalingMode?*string version?*string env?[]EnvVar envFrom?[]EnvFromSource volumeMounts?[]VolumeMount",
    // ---- k8s.io/api core/v1 + apimachinery meta/v1, v0.36.0 (not vendored in the reference: restated, byte-level unverified)
    "PodTemplateSpec = metadata?ObjectMeta spec?PodSpec",
    "ObjectMeta = name?string generateName?string namespace?string selfLink?string uid?string resourceVersion?string generation?int creationTimestamp?time "
    "deletionTimestamp?*string deletionGracePeriodSeconds?*int labels?map annotations?map ownerReferences?[]OwnerReference finalizers?[]string managedFields?raw",
    "OwnerReference = apiVersion:string kind:string name:string uid:string controller?*bool blockOwnerDeletion?*bool",
    "PodSpec = volumes?[]Volume initContainers?[]Container containers:[]Container ephemeralContainers?raw restartPolicy?string terminationGracePeriodSeconds?*int "
    "activeDeadlineSeconds?*int dnsPolicy?string nodeSelector?map serviceAccountName?string serviceAccount?string automountServiceAccountToken?*bool nodeName?string "
    "hostNetwork?bool hostPID?bool hostIPC?bool shareProcessNamespace?*bool securityContext?*PodSecurityContext imagePullSecrets?[]LocalObjectReference hostname?string "
    "subdomain?string affinity?*Affinity schedulerName?string tolerations?[]Toleration hostAliases?[]HostAlias priorityClassName?string priority?*int dnsConfig?*PodDNSConfig "
    "readinessGates?[]PodReadinessGate runtimeClassName?*string enableServiceLinks?*bool preemptionPolicy?*string overhead?mapq "
    "topologySpreadConstraints?[]TopologySpreadConstraint setHostnameAsFQDN?*bool os?*PodOS hostUsers?*bool schedulingGates?[]PodSchedulingGate "
    "resourceClaims?[]PodResourceClaim resources?*ResourceRequirements hostnameOverride?*string",
    "Container = name:string image?string command?[]string args?[]string workingDir?string ports?[]ContainerPort envFrom?[]EnvFromSource env?[]EnvVar "
    "resources?ResourceRequirements resizePolicy?[]ContainerResizePolicy restartPolicy?*string restartPolicyRules?raw volumeMounts?[]VolumeMount "
    "volumeDevices?[]VolumeDevice livenessProbe?*Probe readinessProbe?*Probe startupProbe?*Probe lifecycle?*Lifecycle terminationMessagePath?string "
    "terminationMessagePolicy?string imagePullPolicy?string securityContext?*SecurityContext stdin?bool stdinOnce?bool tty?bool",
    "ContainerResizePolicy = resourceName:string restartPolicy:string",
    "ContainerPort = name?string hostPort?int containerPort:int protocol?string hostIP?string",
    "EnvVar = name:string value?string valueFrom?*EnvVarSource",
    "EnvVarSource = fieldRef?*ObjectFieldSelector resourceFieldRef?*ResourceFieldSelector configMapKeyRef?*ConfigMapKeySelector secretKeyRef?*SecretKeySelector fileKeyRef?raw",
    "ObjectFieldSelector = apiVersion?string fieldPath:string",
    "ResourceFieldSelector = containerName?string resource:string divisor:quantity",
    "ConfigMapKeySelector = name?string key:string optional?*bool",
    "SecretKeySelector = name?string key:string optional?*bool",
    "EnvFromSource = prefix?string configMapRef?*ConfigMapEnvSource secretRef?*SecretEnvSource",
    "ConfigMapEnvSource = name?string optional?*bool",
    "SecretEnvSource = name?string optional?*bool",
    "ResourceRequirements = limits?mapq requests?mapq claims?[]ResourceClaim",
    "ResourceClaim = name:string request?string",
    "VolumeMount = name:string readOnly?bool recursiveReadOnly?*string mountPath:string subPath?string mountPropagation?*string subPathExpr?string",
    "VolumeDevice = name:string devicePath:string",
    "Probe = +ProbeHandler initialDelaySeconds?int timeoutSeconds?int periodSeconds?int successThreshold?int failureThreshold?int terminationGracePeriodSeconds?*int",
    "ProbeHandler = exec?*ExecAction httpGet?*HTTPGetAction tcpSocket?*TCPSocketAction grpc?*GRPCAction",
    "ExecAction = command?[]string",
    "HTTPGetAction = path?string port:intstr host?string scheme?string httpHeaders?[]HTTPHeader",
    "HTTPHeader = name:string value:string",
    "TCPSocketAction = port:intstr host?string",
    "GRPCAction = port:int service:*string",
    "Lifecycle = postStart?*LifecycleHandler preStop?*LifecycleHandler stopSignal?*string",
    "LifecycleHandler = exec?*ExecAction httpGet?*HTTPGetAction tcpSocket?*TCPSocketAction sleep?*SleepAction",
    "SleepAction = seconds:int",
    "SecurityContext = capabilities?*Capabilities privileged?*bool seLinuxOptions?*SELinuxOptions windowsOptions?raw runAsUser?*int runAsGroup?*int runAsNonRoot?*bool "
    "readOnlyRootFilesystem?*bool allowPrivilegeEscalation?*bool procMount?*string seccompProfile?*SeccompProfile appArmorProfile?*AppArmorProfile",
    "Capabilities = add?[]string drop?[]string",
    "SELinuxOptions = user?string role?string type?string level?string",
    "SeccompProfile = type:string localhostProfile?*string",
    "AppArmorProfile = type:string localhostProfile?*string",
    "PodSecurityContext = seLinuxOptions?*SELinuxOptions windowsOptions?raw runAsUser?*int runAsGroup?*int runAsNonRoot?*bool supplementalGroups?[]int "
    "supplementalGroupsPolicy?*string fsGroup?*int sysctls?[]Sysctl fsGroupChangePolicy?*string seccompProfile?*SeccompProfile appArmorProfile?*AppArmorProfile "
    "seLinuxChangePolicy?*string",
    "Sysctl = name:string value:string",
    "LocalObjectReference = name?string",
    "Volume = name:string hostPath?*HostPathVolumeSource emptyDir?*EmptyDirVolumeSource gcePersistentDisk?raw awsElasticBlockStore?raw gitRepo?raw "
    "secret?*SecretVolumeSource nfs?*NFSVolumeSource iscsi?raw glusterfs?raw persistentVolumeClaim?*PersistentVolumeClaimVolumeSource rbd?raw flexVolume?raw cinder?raw "
    "cephfs?raw flocker?raw downwardAPI?raw fc?raw azureFile?raw configMap?*ConfigMapVolumeSource vsphereVolume?raw quobyte?raw azureDisk?raw photonPersistentDisk?raw "
    "projected?raw portworxVolume?raw scaleIO?raw storageos?raw csi?*CSIVolumeSource ephemeral?raw image?raw",
    "HostPathVolumeSource = path:string type?*string",
    "EmptyDirVolumeSource = medium?string sizeLimit?*quantity",
    "SecretVolumeSource = secretName?string items?[]KeyToPath defaultMode?*int optional?*bool",
    "KeyToPath = key:string path:string mode?*int",
    "NFSVolumeSource = server:string path:string readOnly?bool",
    "PersistentVolumeClaimVolumeSource = claimName:string readOnly?bool",
    "ConfigMapVolumeSource = name?string items?[]KeyToPath defaultMode?*int optional?*bool",
    "CSIVolumeSource = driver:string readOnly?*bool fsType?*string volumeAttributes?map nodePublishSecretRef?*LocalObjectReference",
    "Toleration = key?string operator?string value?string effect?string tolerationSeconds?*int",
    "HostAlias = ip:string hostnames?[]string",
    "PodDNSConfig = nameservers?[]string searches?[]string options?[]PodDNSConfigOption",
    "PodDNSConfigOption = name?string value?*string",
    "PodReadinessGate = conditionType:string",
    "PodOS = name:string",
    "PodSchedulingGate = name:string",
    "PodResourceClaim = name:string resourceClaimName?*string resourceClaimTemplateName?*string",
    "Affinity = nodeAffinity?*NodeAffinity podAffinity?*PodAffinity podAntiAffinity?*PodAffinity",
    "NodeAffinity = requiredDuringSchedulingIgnoredDuringExecution?*NodeSelector preferredDuringSchedulingIgnoredDuringExecution?[]PreferredSchedulingTerm",
    "NodeSelector = nodeSelectorTerms:[]NodeSelectorTerm",
    "NodeSelectorTerm = matchExpressions?[]SelectorRequirement matchFields?[]SelectorRequirement",
    "SelectorRequirement = key:string operator:string values?[]string",
    "PreferredSchedulingTerm = weight:int preference:NodeSelectorTerm",
    "PodAffinity = requiredDuringSchedulingIgnoredDuringExecution?[]PodAffinityTerm preferredDuringSchedulingIgnoredDuringExecution?[]WeightedPodAffinityTerm",
    "PodAffinityTerm = labelSelector?*LabelSelector namespaces?[]string topologyKey:string namespaceSelector?*LabelSelector matchLabelKeys?[]string mismatchLabelKeys?[]string",
    "WeightedPodAffinityTerm = weight:int podAffinityTerm:PodAffinityTerm",
    "LabelSelector = matchLabels?map matchExpressions?[]SelectorRequirement",
    "TopologySpreadConstraint = maxSkew:int topologyKey:string whenUnsatisfiable:string labelSelector?*LabelSelector minDomains?*int nodeAffinityPolicy?*string "
    "nodeTaintsPolicy?*string matchLabelKeys?[]string",
    "Service = kind?string apiVersion?string metadata?ObjectMeta spec?ServiceSpec status?ServiceStatus",
    "ServiceSpec = ports?[]ServicePort selector?map clusterIP?string clusterIPs?[]string type?string externalIPs?[]string sessionAffinity?string loadBalancerIP?string "
    "loadBalancerSourceRanges?[]string externalName?string externalTrafficPolicy?string healthCheckNodePort?int publishNotReadyAddresses?bool sessionAffinityConfig?raw "
    "ipFamilies?[]string ipFamilyPolicy?*string allocateLoadBalancerNodePorts?*bool loadBalancerClass?*string internalTrafficPolicy?*string trafficDistribution?*string",
    "ServicePort = name?string protocol?string appProtocol?*string port:int targetPort:intstr nodePort?int",
    "ServiceStatus = loadBalancer?LoadBalancerStatus conditions?raw",
    "LoadBalancerStatus = ingress?raw",
};

// The tables above, compiled once: every kind string becomes a small tree of KindNodes, every struct a vector of Fields.
enum KindOp : uint8_t { K_RAW, K_MAP, K_MAPQ, K_BOOL, K_INT, K_STRING, K_QUANTITY, K_INTSTR, K_TIME, K_STRUCT, K_PTR, K_ARR };
struct KindNode { KindOp op; int sub; };  // K_PTR / K_ARR: sub = kind id of the pointee / element; K_STRUCT: sub = type id (-1: not in the tables)
struct Field { std::string name; int kind; bool omitempty; int inline_type; };  // inline_type >= 0: an embedded struct ("+Type")
struct Schema {
  std::map<std::string, int> type_id;
  std::vector<std::string> type_name;
  std::vector<std::vector<Field>> types;
  std::vector<KindNode> kinds;
  std::map<std::string, int> kind_id;
  std::vector<std::string> unknown;  // struct names a kind refers to that no table defines (reported when reached)

  int type_of(const std::string &name) const { auto it = type_id.find(name); return it == type_id.end() ? -1 : it->second; }
  int compile(const std::string &k) {
    auto it = kind_id.find(k);
    if (it != kind_id.end()) return it->second;
    KindNode kn{K_STRUCT, -1};
    if (k.compare(0, 2, "[]") == 0) kn = {K_ARR, compile(k.substr(2))};
    else if (k[0] == '*') kn = {K_PTR, compile(k.substr(1))};
    else if (k == "raw") kn.op = K_RAW;
    else if (k == "map") kn.op = K_MAP;
    else if (k == "mapq") kn.op = K_MAPQ;
    else if (k == "bool") kn.op = K_BOOL;
    else if (k == "int") kn.op = K_INT;
    else if (k == "string") kn.op = K_STRING;
    else if (k == "quantity") kn.op = K_QUANTITY;
    else if (k == "intstr") kn.op = K_INTSTR;
    else if (k == "time") kn.op = K_TIME;
    else {
      kn.sub = type_of(k);
      if (kn.sub < 0) { unknown.push_back(k); kn.sub = -(int)unknown.size() - 1; }  // -(index + 2)
    }
    kinds.push_back(kn);
    kind_id[k] = (int)kinds.size() - 1;
    return (int)kinds.size() - 1;
  }
  Schema() {
    std::vector<std::pair<std::string, std::string>> lines;
    for (const char *line : kSchema) {
      std::string l(line);
      const size_t eq = l.find(" = ");
      lines.emplace_back(l.substr(0, eq), l.substr(eq + 3));
      type_id[lines.back().first] = (int)type_name.size();
      type_name.push_back(lines.back().first);
    }
    types.resize(lines.size());
    for (size_t t = 0; t < lines.size(); t++) {
      const std::string &l = lines[t].second;
      size_t i = 0;
      while (i < l.size()) {
        size_t j = l.find(' ', i);
        if (j == std::string::npos) j = l.size();
        const std::string tok = l.substr(i, j - i);
        i = j + 1;
        if (tok.empty()) continue;
        if (tok[0] == '+') { types[t].push_back({"", -1, false, type_of(tok.substr(1))}); continue; }
        const size_t sep = tok.find_first_of("?:");
        types[t].push_back({tok.substr(0, sep), compile(tok.substr(sep + 1)), tok[sep] == '?', -1});
      }
    }
  }
};
inline const Schema &schema() { static Schema s; return s; }

struct Emitter {
  std::string out, err;

  void raw(const Node &n) {  // a value of a type the tables do not describe: the caller's order, Go's scalars
    switch (n.t) {
      case N_NULL: out += "null"; break;
      case N_BOOL: out += n.b ? "true" : "false"; break;
      case N_NUM: out += n.s; break;
      case N_STR: go_string(out, n.s); break;
      case N_ARR: out += '['; for (size_t i = 0; i < n.a.size(); i++) { if (i) out += ','; raw(n.a[i]); } out += ']'; break;
      case N_OBJ:
        out += '{';
        for (size_t i = 0; i < n.o.size(); i++) { if (i) out += ','; go_string(out, n.o[i].first); out += ':'; raw(n.o[i].second); }
        out += '}';
        break;
    }
  }
  static bool is_scalar(KindOp op) { return op >= K_BOOL && op <= K_TIME; }
  void scalar(KindOp op, const Node *n) {
    static const std::string kNone;
    switch (op) {
      case K_BOOL: out += (n && n->t == N_BOOL && n->b) ? "true" : "false"; break;
      case K_INT: if (n && n->t == N_NUM) out += n->s; else out += '0'; break;
      case K_STRING: go_string(out, (n && n->t == N_STR) ? n->s : kNone); break;
      case K_QUANTITY: {
        std::string text = !n ? "0" : (n->t == N_STR || n->t == N_NUM) ? n->s : "0", canon;
        go_string(out, canon_quantity(text, canon) ? canon : text);
        break;
      }
      case K_INTSTR:
        if (n && n->t == N_STR) go_string(out, n->s);
        else if (n && n->t == N_NUM) out += n->s; else out += '0';
        break;
      case K_TIME: if (n && n->t == N_STR) go_string(out, n->s); else out += "null"; break;
      default: break;
    }
  }
  static bool scalar_zero(KindOp op, const Node *n) {
    if (!n || n->t == N_NULL) return true;
    switch (op) {
      case K_BOOL: return !(n->t == N_BOOL && n->b);
      case K_INT: return n->t != N_NUM || n->s == "0" || n->s == "-0";
      case K_STRING: return n->t != N_STR || n->s.empty();
      case K_TIME: return n->t != N_STR;
      default: return false;  // quantity / intstr are struct values: omitempty never drops them
    }
  }
  void string_map(const Node &n, bool quantities) {
    std::vector<const std::pair<std::string, Node> *> items;
    items.reserve(n.o.size());
    for (auto &kv : n.o) items.push_back(&kv);
    std::sort(items.begin(), items.end(), [](auto *a, auto *b) { return a->first < b->first; });  // bytewise, like encoding/json
    out += '{';
    for (size_t i = 0; i < items.size(); i++) {
      if (i) out += ',';
      go_string(out, items[i]->first);
      out += ':';
      const Node &v = items[i]->second;
      if (quantities) scalar(K_QUANTITY, &v);
      else if (v.t == N_STR) go_string(out, v.s);
      else raw(v);
    }
    out += '}';
  }
  // one value of kind `k`; `n` may be null (absent)
  void value(int k, const Node *n) {
    const KindNode kn = schema().kinds[(size_t)k];
    const bool absent = !n || n->t == N_NULL;
    switch (kn.op) {
      case K_RAW: if (absent) out += "null"; else raw(*n); return;
      case K_MAP: case K_MAPQ:
        if (absent || n->t != N_OBJ) out += "null"; else string_map(*n, kn.op == K_MAPQ);
        return;
      case K_ARR:
        if (absent || n->t != N_ARR) { out += "null"; return; }
        out += '[';
        for (size_t i = 0; i < n->a.size(); i++) { if (i) out += ','; value(kn.sub, &n->a[i]); }
        out += ']';
        return;
      case K_PTR: if (absent) out += "null"; else value(kn.sub, n); return;
      case K_STRUCT: strct_id(kn.sub, absent || n->t != N_OBJ ? nullptr : n); return;
      default: scalar(kn.op, absent ? nullptr : n); return;
    }
  }
  // does `omitempty` drop this field?
  bool omitted(int k, const Node *n) {
    const KindNode kn = schema().kinds[(size_t)k];
    const bool absent = !n || n->t == N_NULL;
    switch (kn.op) {
      case K_RAW: return absent || (n->t == N_ARR && n->a.empty()) || (n->t == N_OBJ && n->o.empty()) || (n->t == N_STR && n->s.empty()) ||
                         (n->t == N_BOOL && !n->b) || (n->t == N_NUM && n->s == "0");
      case K_PTR: return absent;
      case K_MAP: case K_MAPQ: return absent || n->t != N_OBJ || n->o.empty();
      case K_ARR: return absent || n->t != N_ARR || n->a.empty();
      case K_STRUCT: return false;  // a struct value is never empty for encoding/json
      default: return scalar_zero(kn.op, n);
    }
  }
  void fields_of(int type, const Node *n, bool &first) {
    if (type < 0) { err = "unknown type " + (type <= -2 ? schema().unknown[(size_t)(-type - 2)] : std::string("?")); return; }
    for (const Field &f : schema().types[(size_t)type]) {
      if (f.kind < 0) { fields_of(f.inline_type, n, first); continue; }  // an embedded struct
      const Node *v = n ? n->get(f.name) : nullptr;
      if (f.omitempty && omitted(f.kind, v)) continue;
      if (!first) out += ',';
      first = false;
      out += '"'; out += f.name; out += "\":";  // field names are plain identifiers: nothing to escape
      value(f.kind, v);
    }
  }
  void strct_id(int type, const Node *n) {
    out += '{';
    bool first = true;
    fields_of(type, n, first);
    out += '}';
  }
  void strct(const std::string &tname, const Node *n) {
    const int t = schema().type_of(tname);
    if (t < 0) { out += "{}"; err = "unknown type " + tname; return; }
    strct_id(t, n);
  }
};

}  // namespace krjson
#endif  // KR_JSON_HPP_
