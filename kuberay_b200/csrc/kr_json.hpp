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
      } else out += *p++;
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
        Node v;
        if (!value(v, depth + 1)) return false;
        // a repeated key: the last one wins (encoding/json)
        bool replaced = false;
        for (auto &kv : n.o) if (kv.first == k) { kv.second = std::move(v); replaced = true; break; }
        if (!replaced) n.o.emplace_back(std::move(k), std::move(v));
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
        Node v;
        if (!value(v, depth + 1)) return false;
        n.a.push_back(std::move(v));
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

struct Field { std::string name, kind; bool omitempty; };
struct Schema {
  std::map<std::string, std::vector<Field>> types;
  Schema() {
    for (const char *line : kSchema) {
      std::string l(line);
      size_t eq = l.find(" = ");
      std::string tname = l.substr(0, eq);
      std::vector<Field> fields;
      size_t i = eq + 3;
      while (i < l.size()) {
        size_t j = l.find(' ', i);
        if (j == std::string::npos) j = l.size();
        std::string tok = l.substr(i, j - i);
        i = j + 1;
        if (tok.empty()) continue;
        if (tok[0] == '+') { fields.push_back({"", tok, false}); continue; }
        size_t sep = tok.find_first_of("?:");
        fields.push_back({tok.substr(0, sep), tok.substr(sep + 1), tok[sep] == '?'});
      }
      types[tname] = std::move(fields);
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
  static bool is_scalar_kind(const std::string &k) { return k == "bool" || k == "int" || k == "string" || k == "quantity" || k == "intstr" || k == "time"; }
  void scalar(const std::string &kind, const Node *n) {
    if (kind == "bool") out += (n && n->t == N_BOOL && n->b) ? "true" : "false";
    else if (kind == "int") out += (n && n->t == N_NUM) ? n->s : "0";
    else if (kind == "string") go_string(out, (n && n->t == N_STR) ? n->s : std::string());
    else if (kind == "quantity") {
      std::string text = !n ? "0" : (n->t == N_STR || n->t == N_NUM) ? n->s : "0", canon;
      go_string(out, canon_quantity(text, canon) ? canon : text);
    } else if (kind == "intstr") {
      if (n && n->t == N_STR) go_string(out, n->s);
      else out += (n && n->t == N_NUM) ? n->s : "0";
    } else if (kind == "time") {
      if (n && n->t == N_STR) go_string(out, n->s); else out += "null";
    }
  }
  static bool scalar_zero(const std::string &kind, const Node *n) {
    if (!n || n->t == N_NULL) return true;
    if (kind == "bool") return !(n->t == N_BOOL && n->b);
    if (kind == "int") return n->t != N_NUM || n->s == "0" || n->s == "-0";
    if (kind == "string") return n->t != N_STR || n->s.empty();
    if (kind == "time") return n->t != N_STR;
    return false;  // quantity / intstr are struct values: omitempty never drops them
  }
  void string_map(const Node &n, bool quantities) {
    std::vector<const std::pair<std::string, Node> *> items;
    for (auto &kv : n.o) items.push_back(&kv);
    std::sort(items.begin(), items.end(), [](auto *a, auto *b) { return a->first < b->first; });  // bytewise, like encoding/json
    out += '{';
    for (size_t i = 0; i < items.size(); i++) {
      if (i) out += ',';
      go_string(out, items[i]->first);
      out += ':';
      const Node &v = items[i]->second;
      if (quantities) scalar("quantity", &v);
      else if (v.t == N_STR) go_string(out, v.s);
      else raw(v);
    }
    out += '}';
  }
  // one value of `kind`; `n` may be null (absent)
  void value(const std::string &kind, const Node *n) {
    const bool absent = !n || n->t == N_NULL;
    if (kind == "raw") { if (absent) out += "null"; else raw(*n); return; }
    if (kind == "map" || kind == "mapq") {
      if (absent || n->t != N_OBJ) out += "null"; else string_map(*n, kind == "mapq");
      return;
    }
    if (kind.compare(0, 2, "[]") == 0) {
      if (absent || n->t != N_ARR) { out += "null"; return; }
      const std::string elem = kind.substr(2);
      out += '[';
      for (size_t i = 0; i < n->a.size(); i++) { if (i) out += ','; value(elem, &n->a[i]); }
      out += ']';
      return;
    }
    if (kind[0] == '*') { if (absent) out += "null"; else value(kind.substr(1), n); return; }
    if (is_scalar_kind(kind)) { scalar(kind, absent ? nullptr : n); return; }
    strct(kind, absent || n->t != N_OBJ ? nullptr : n);
  }
  // does `omitempty` drop this field?
  bool omitted(const std::string &kind, const Node *n) {
    const bool absent = !n || n->t == N_NULL;
    if (kind == "raw") return absent || (n->t == N_ARR && n->a.empty()) || (n->t == N_OBJ && n->o.empty()) || (n->t == N_STR && n->s.empty()) ||
                              (n->t == N_BOOL && !n->b) || (n->t == N_NUM && n->s == "0");
    if (kind[0] == '*') return absent;
    if (kind == "map" || kind == "mapq") return absent || n->t != N_OBJ || n->o.empty();
    if (kind.compare(0, 2, "[]") == 0) return absent || n->t != N_ARR || n->a.empty();
    if (is_scalar_kind(kind)) return scalar_zero(kind, n);
    return false;  // a struct value is never empty for encoding/json
  }
  void fields_of(const std::string &tname, const Node *n, bool &first) {
    auto it = schema().types.find(tname);
    if (it == schema().types.end()) { err = "unknown type " + tname; return; }
    for (const Field &f : it->second) {
      if (f.kind[0] == '+') { fields_of(f.kind.substr(1), n, first); continue; }
      const Node *v = n ? n->get(f.name.c_str()) : nullptr;
      if (f.omitempty && omitted(f.kind, v)) continue;
      if (!first) out += ',';
      first = false;
      go_string(out, f.name);
      out += ':';
      value(f.kind, v);
    }
  }
  void strct(const std::string &tname, const Node *n) {
    out += '{';
    bool first = true;
    fields_of(tname, n, first);
    out += '}';
  }
};

}  // namespace krjson
#endif  // KR_JSON_HPP_
