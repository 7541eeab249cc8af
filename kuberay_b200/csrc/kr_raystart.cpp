// `ray start` command builder (second part) — host side of the C ABI, no GPU.
//
// What DefaultHeadPodTemplate / DefaultWorkerPodTemplate and BuildPod do to a group's rayStartParams and to the Ray container's
// command line, per worker group (it does not depend on the create tuple: the shim calls it once per group and reconcile).
// Restated from (never copied), paths relative to ray-operator/controllers/ray/:
//   GetHeadPort                           common/pod.go:54-59
//   updateRayStartParamsResources         common/pod.go:1238-1276
//   updateRayStartParamsLabels            common/pod.go:1219-1235
//   setMissingRayStartParams              common/pod.go:935-978
//   head: no-monitor with the autoscaler  common/pod.go:196-200
//   generateRayStartCommand               common/pod.go:980-1020
//   addWellKnownAcceleratorResources      common/pod.go:1022-1097
//   convertParamMap                       common/pod.go:1108-1135
//   the container command / args          common/pod.go:617-650 (BuildPod), utils/util.go:884-892 (GetContainerCommand)
//   utils.IsGPUResourceKey                utils/resources.go:8-17
//   resource.Quantity (ParseQuantity, Value, IsZero, AsApproximateFloat64): k8s.io/apimachinery v0.36.0 pkg/api/resource —
//     third-party, absent from /root/reference; restated from its published behaviour: <sign><digits>[.<digits>][<suffix>] with
//     suffixes n u m "" k M G T P E (powers of 1000), Ki Mi Gi Ti Pi Ei (powers of 1024) or a decimal exponent e<N> / E<N>;
//     Value() rounds up to the next integer.
//   encoding/json float64 formatting (the resources map): shortest representation, 'f' form for 1e-6 <= |x| < 1e21, else 'e' form
//     with a two-digit exponent reduced to one when it has a leading zero.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/kr_engine.h"

void kr_go_string_append(std::string &out, const std::string &s);  // kr_specjson.cpp

namespace {

thread_local std::string g_err;
typedef std::map<std::string, std::string> StrMap;

inline std::string str(kr_str s) { return (s.p && s.n) ? std::string(s.p, s.n) : std::string(); }
std::string lower(std::string s) { for (char &c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32); return s; }

// ---- resource.Quantity: value = mant * 10^e10 * 2^e2 (mant >= 0, sign apart) ------------------------------------------------
struct Quantity { bool ok = false, neg = false; unsigned __int128 mant = 0; int e10 = 0, e2 = 0; };

Quantity parse_quantity(const std::string &in) {
  Quantity q;
  size_t i = 0, n = in.size();
  if (n == 0) return q;
  if (in[i] == '+' || in[i] == '-') { q.neg = in[i] == '-'; i++; }
  int digits = 0, frac = 0;
  bool dot = false;
  for (; i < n; i++) {
    const char c = in[i];
    if (c >= '0' && c <= '9') {
      if (q.mant > ((unsigned __int128)1 << 120)) return q;  // far beyond anything a resource field carries
      q.mant = q.mant * 10 + (unsigned)(c - '0'); digits++;
      if (dot) frac++;
    } else if (c == '.' && !dot) dot = true;
    else break;
  }
  if (digits == 0) return q;
  q.e10 = -frac;
  const std::string suf = in.substr(i);
  if (suf.empty()) { q.ok = true; return q; }
  if (suf == "Ki" || suf == "Mi" || suf == "Gi" || suf == "Ti" || suf == "Pi" || suf == "Ei") {
    q.e2 = 10 * (int)(std::string("KMGTPE").find(suf[0]) + 1); q.ok = true; return q;
  }
  if (suf.size() == 1) {
    static const char *dec = "numkMGTPE";
    static const int exp10[] = {-9, -6, -3, 3, 6, 9, 12, 15, 18};
    const char *p = strchr(dec, suf[0]);
    if (p && *p) { q.e10 += exp10[p - dec]; q.ok = true; return q; }
  }
  if (suf[0] == 'e' || suf[0] == 'E') {
    size_t k = 1;
    bool eneg = false;
    if (k < suf.size() && (suf[k] == '+' || suf[k] == '-')) { eneg = suf[k] == '-'; k++; }
    if (k >= suf.size()) return q;
    int ex = 0;
    for (; k < suf.size(); k++) { if (suf[k] < '0' || suf[k] > '9' || ex > 1000) return q; ex = ex * 10 + (suf[k] - '0'); }
    q.e10 += eneg ? -ex : ex; q.ok = true; return q;
  }
  return q;
}
bool q_is_zero(const Quantity &q) { return q.mant == 0; }
// Value(): the quantity rounded up to an integer, as int64 (saturating)
long long q_value(const Quantity &q) {
  unsigned __int128 m = q.mant;
  bool rem = false;
  int e10 = q.e10;
  for (int k = 0; k < q.e2; k++) { if (m > ((unsigned __int128)1 << 125)) return q.neg ? INT64_MIN : INT64_MAX; m <<= 1; }
  while (e10 > 0) { if (m > ((unsigned __int128)1 << 120)) return q.neg ? INT64_MIN : INT64_MAX; m *= 10; e10--; }
  while (e10 < 0) { if (m % 10) rem = true; m /= 10; e10++; }
  if (rem && !q.neg) m += 1;  // round up (towards +inf; a negative value truncates towards zero, which is up)
  if (m > (unsigned __int128)INT64_MAX) return q.neg ? INT64_MIN : INT64_MAX;
  return q.neg ? -(long long)m : (long long)m;
}
// math.Pow10 as Go computes it (tables of literals: exact up to 1e22, correctly rounded beyond; negative powers by one division)
double go_pow10(int n) {
  static const double tab[32] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20,
                                 1e21, 1e22, 1e23, 1e24, 1e25, 1e26, 1e27, 1e28, 1e29, 1e30, 1e31};
  static const double pos32[10] = {1e0, 1e32, 1e64, 1e96, 1e128, 1e160, 1e192, 1e224, 1e256, 1e288};
  static const double neg32[11] = {1e-0, 1e-32, 1e-64, 1e-96, 1e-128, 1e-160, 1e-192, 1e-224, 1e-256, 1e-288, 1e-320};
  if (n >= 0 && n <= 308) return pos32[n / 32] * tab[n % 32];
  if (n >= -323 && n < 0) return neg32[(-n) / 32] / tab[(-n) % 32];
  return n > 0 ? INFINITY : 0.0;
}
// AsApproximateFloat64(): float64(unscaled) * math.Pow10(-scale), after ParseQuantity rounded anything finer than nano up to nano
double q_float(const Quantity &q) {
  unsigned __int128 m = q.mant;
  int e10 = q.e10;
  for (int k = 0; k < q.e2; k++) { if (m > ((unsigned __int128)1 << 125)) return q.neg ? -INFINITY : INFINITY; m <<= 1; }
  if (e10 < -9 && m != 0) {
    bool rem = false;
    while (e10 < -9) { if (m % 10) rem = true; m /= 10; e10++; }
    if (rem) m += 1;
  }
  const double base = (double)m;
  const double v = e10 == 0 ? base : base * go_pow10(e10);
  return q.neg ? -v : v;
}

// encoding/json's float64 encoder
std::string go_float(double f) {
  if (f == 0) return std::signbit(f) ? "-0" : "0";
  // shortest round-trip digits and decimal exponent (strconv's 'e' / -1), then encoding/json's layout
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);
  std::string sci(buf, r.ptr), digits;
  const size_t epos = sci.find('e');
  for (size_t i = 0; i < epos; i++) if (sci[i] >= '0' && sci[i] <= '9') digits += sci[i];
  const int e10 = atoi(sci.c_str() + epos + 1);
  const std::string sign = f < 0 ? "-" : "";
  const double a = std::fabs(f);
  const int nd = (int)digits.size();
  if (a >= 1e-6 && a < 1e21) {  // 'f' form: the digits, padded with zeros up to the decimal point
    if (e10 >= nd - 1) return sign + digits + std::string((size_t)(e10 - nd + 1), '0');
    if (e10 >= 0) return sign + digits.substr(0, (size_t)e10 + 1) + "." + digits.substr((size_t)e10 + 1);
    return sign + "0." + std::string((size_t)(-e10 - 1), '0') + digits;
  }
  std::string out = sign + digits.substr(0, 1) + (nd > 1 ? "." + digits.substr(1) : "");
  char eb[16];
  snprintf(eb, sizeof eb, "e%c%02d", e10 < 0 ? '-' : '+', e10 < 0 ? -e10 : e10);
  std::string es(eb);  // e-07 -> e-7 (Go cleans a two-digit exponent with a leading zero)
  if (es.size() == 4 && es[2] == '0') es = es.substr(0, 2) + es.substr(3);
  return out + es;
}

// json.Unmarshal into map[string]float64: an object whose values are all numbers (null leaves the key out); anything else fails
bool parse_float_map(const std::string &text, std::map<std::string, double> &out) {
  size_t i = 0, n = text.size();
  auto ws = [&] { while (i < n && (text[i] == ' ' || text[i] == '\t' || text[i] == '\n' || text[i] == '\r')) i++; };
  ws();
  if (text.compare(i, 4, "null") == 0) { i += 4; ws(); return i == n; }
  if (i >= n || text[i] != '{') return false;
  i++; ws();
  if (i < n && text[i] == '}') { i++; ws(); return i == n; }
  while (true) {
    ws();
    if (i >= n || text[i] != '"') return false;
    i++;
    std::string key;
    while (i < n && text[i] != '"') {
      if (text[i] == '\\') {
        if (i + 1 >= n) return false;
        const char e = text[i + 1];
        if (e == 'u') {  // \uXXXX (json.Marshal writes <, >, & of a resource name this way), surrogate pairs included
          auto hex4 = [&](size_t at, uint32_t &v) {
            if (at + 4 > n) return false;
            v = 0;
            for (size_t k = 0; k < 4; k++) {
              const char c = text[at + k];
              const int d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
              if (d < 0) return false;
              v = v * 16 + (uint32_t)d;
            }
            return true;
          };
          uint32_t cp = 0, lo = 0;
          if (!hex4(i + 2, cp)) return false;
          i += 6;
          if (cp >= 0xD800 && cp < 0xDC00 && i + 1 < n && text[i] == '\\' && text[i + 1] == 'u' && hex4(i + 2, lo) && lo >= 0xDC00 && lo < 0xE000) {
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); i += 6;
          } else if (cp >= 0xD800 && cp < 0xE000) cp = 0xFFFD;  // lone surrogate
          if (cp < 0x80) key += (char)cp;
          else if (cp < 0x800) { key += (char)(0xC0 | (cp >> 6)); key += (char)(0x80 | (cp & 63)); }
          else if (cp < 0x10000) { key += (char)(0xE0 | (cp >> 12)); key += (char)(0x80 | ((cp >> 6) & 63)); key += (char)(0x80 | (cp & 63)); }
          else { key += (char)(0xF0 | (cp >> 18)); key += (char)(0x80 | ((cp >> 12) & 63)); key += (char)(0x80 | ((cp >> 6) & 63)); key += (char)(0x80 | (cp & 63)); }
          continue;
        }
        if (e != '"' && e != '\\' && e != '/' && e != 'n' && e != 't' && e != 'r' && e != 'b' && e != 'f') return false;
        key += (e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e == 'b' ? '\b' : e == 'f' ? '\f' : e);
        i += 2;
      } else if ((unsigned char)text[i] < 0x20) return false;  // raw control characters are not JSON
      else key += text[i++];
    }
    if (i >= n) return false;
    i++; ws();
    if (i >= n || text[i] != ':') return false;
    i++; ws();
    if (text.compare(i, 4, "null") == 0) { i += 4; }
    else {
      const size_t s0 = i;
      if (i < n && text[i] == '-') i++;
      if (i >= n || text[i] < '0' || text[i] > '9') return false;
      if (text[i] == '0') i++; else while (i < n && text[i] >= '0' && text[i] <= '9') i++;
      if (i < n && text[i] == '.') { i++; if (i >= n || text[i] < '0' || text[i] > '9') return false; while (i < n && text[i] >= '0' && text[i] <= '9') i++; }
      if (i < n && (text[i] == 'e' || text[i] == 'E')) {
        i++;
        if (i < n && (text[i] == '+' || text[i] == '-')) i++;
        if (i >= n || text[i] < '0' || text[i] > '9') return false;
        while (i < n && text[i] >= '0' && text[i] <= '9') i++;
      }
      double v = 0;
      auto r = std::from_chars(text.data() + s0, text.data() + i, v);
      if (r.ec != std::errc()) return false;
      out[key] = v;
    }
    ws();
    if (i < n && text[i] == ',') { i++; continue; }
    if (i < n && text[i] == '}') { i++; ws(); return i == n; }
    return false;
  }
}

std::string marshal_float_map(const std::map<std::string, double> &m) {  // json.Marshal(map[string]float64): keys sorted
  std::string out = "{";
  bool first = true;
  for (const auto &e : m) {
    if (!first) out += ',';
    first = false;
    kr_go_string_append(out, e.first);
    out += ':';
    out += go_float(e.second);
  }
  return out + "}";
}

bool is_gpu_key(const std::string &k) {  // utils/resources.go:8-17
  if (k.size() >= 3 && k.compare(k.size() - 3, 3, "gpu") == 0) return true;
  // nvidia\.com/mig-\d+g\.\d+gb$ (unanchored at the front)
  static const std::string pre = "nvidia.com/mig-";
  for (size_t s0 = k.find(pre); s0 != std::string::npos; s0 = k.find(pre, s0 + 1)) {
    size_t i = s0 + pre.size(), d = 0;
    while (i < k.size() && k[i] >= '0' && k[i] <= '9') { i++; d++; }
    if (!d || i + 1 >= k.size() || k[i] != 'g' || k[i + 1] != '.') continue;
    i += 2; d = 0;
    while (i < k.size() && k[i] >= '0' && k[i] <= '9') { i++; d++; }
    if (d && k.compare(i, std::string::npos, "gb") == 0) return true;
  }
  return false;
}

const char *custom_accelerator(const std::string &k) {  // common/pod.go:40-49
  if (k == "aws.amazon.com/neuroncore") return "neuron_cores";
  if (k == "google.com/tpu") return "TPU";
  return nullptr;
}

std::string trim_set(const std::string &s, const char *set) {  // strings.Trim
  size_t a = 0, b = s.size();
  while (a < b && strchr(set, s[a])) a++;
  while (b > a && strchr(set, s[b - 1])) b--;
  return s.substr(a, b - a);
}

void put_all(StrMap &m, const kr_kv *kv, uint32_t n) { for (uint32_t i = 0; i < n; i++) m[str(kv[i].key)] = str(kv[i].value); }

// common/pod.go:1238-1276.  The reference ranges over a Go map; two names that normalise to the same key ("CPU" and "cpu") would
// make its result depend on iteration order — here: byte order of the names, the later one wins.
void update_resources(StrMap &p, const StrMap &group_resources) {
  if (group_resources.empty()) return;
  std::map<std::string, double> custom;
  for (const auto &e : group_resources) {
    const Quantity q = parse_quantity(e.second);
    if (!q.ok) continue;
    const std::string nm = lower(e.first);
    if (nm == "cpu") p["num-cpus"] = std::to_string(q_value(q));
    else if (nm == "memory") p["memory"] = std::to_string(q_value(q));
    else if (is_gpu_key(nm)) p["num-gpus"] = std::to_string(q_value(q));
    else custom[e.first] = q_float(q);
  }
  if (!custom.empty()) p["resources"] = "'" + marshal_float_map(custom) + "'";
}

void update_labels(StrMap &p, const StrMap &group_labels) {  // common/pod.go:1219-1235
  if (group_labels.empty()) return;
  std::string joined;
  for (const auto &e : group_labels) { if (!joined.empty()) joined += ','; joined += e.first + "=" + e.second; }
  p["labels"] = joined;
}

void add_accelerators(StrMap &p, const StrMap &limits) {  // common/pod.go:1022-1069
  if (limits.empty()) return;
  std::map<std::string, double> res;
  auto it = p.find("resources");
  if (it != p.end() && !parse_float_map(trim_set(it->second, "'\"` "), res)) return;  // the error is logged and nothing is added
  bool have_custom = res.count("neuron_cores") || res.count("TPU");
  for (const auto &e : limits) {  // sorted resource keys
    const Quantity q = parse_quantity(e.second);
    const bool zero = !q.ok || q_is_zero(q);
    if (!p.count("num-gpus") && is_gpu_key(e.first) && !zero) p["num-gpus"] = std::to_string(q_value(q));
    if (!have_custom) {
      const char *ray_name = custom_accelerator(e.first);
      if (ray_name && !zero) {
        if (!res.count(ray_name)) { res[ray_name] = q_float(q); p["resources"] = "'" + marshal_float_map(res) + "'"; }
        have_custom = true;
      }
    }
  }
}

std::string convert_param_map(const StrMap &p) {  // common/pod.go:1108-1135
  std::string out;
  for (const auto &e : p) {
    const std::string lv = lower(e.second);
    const bool boolean = (lv == "true" || lv == "false") && e.first != "log-color" && e.first != "include-dashboard";
    if (boolean) { if (lv == "true") out += " --" + e.first + " "; }
    else out += " --" + e.first + "=" + e.second + " ";
  }
  return out;
}

void json_str_array(std::string &out, const std::vector<std::string> &v) {
  out += '[';
  for (size_t i = 0; i < v.size(); i++) { if (i) out += ','; kr_go_string_append(out, v[i]); }
  out += ']';
}

}  // namespace

extern "C" {

const char *kr_ray_start_last_error(void) { return g_err.c_str(); }

int64_t kr_quantity_value(kr_str text, int64_t *value_out, double *approx_out, uint8_t *is_zero_out) {
  const Quantity q = parse_quantity(str(text));
  if (!q.ok) { g_err = "kr_quantity_value: not a quantity"; return KR_E_INVALID; }
  if (value_out) *value_out = q_value(q);
  if (approx_out) *approx_out = q_float(q);
  if (is_zero_out) *is_zero_out = q_is_zero(q) ? 1 : 0;
  return KR_OK;
}

// setContainerEnvVars (common/pod.go:815-933) / setInitContainerEnvVars (:801-813): the EnvVars BuildPod APPENDS to a container, in
// its order, as a JSON array in Go's encoding ({"name":..,"value":..} — value omitted when empty — or
// {"name":..,"valueFrom":{"fieldRef":{"fieldPath":..}}}).  "Exists" checks see the template's names and everything appended so far.
int kr_ray_container_env(const kr_rayenv_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_container_env: null argument"; return KR_E_INVALID; }
  if (!in->init_container && in->node_type != KR_NT_HEAD && in->node_type != KR_NT_WORKER) { g_err = "kr_ray_container_env: node_type must be KR_NT_HEAD or KR_NT_WORKER"; return KR_E_INVALID; }
  std::vector<std::string> names;
  for (uint32_t i = 0; i < in->n_existing; i++) names.push_back(str(in->existing[i]));
  auto exists = [&](const char *n) { return std::find(names.begin(), names.end(), n) != names.end(); };
  std::string js = "[";
  auto sep = [&] { if (js.size() > 1) js += ','; };
  auto add_value = [&](const std::string &name, const std::string &value) {
    sep(); js += "{\"name\":"; kr_go_string_append(js, name);
    if (!value.empty()) { js += ",\"value\":"; kr_go_string_append(js, value); }
    js += '}'; names.push_back(name);
  };
  auto add_field = [&](const std::string &name, const std::string &path) {
    sep(); js += "{\"name\":"; kr_go_string_append(js, name); js += ",\"valueFrom\":{\"fieldRef\":{\"fieldPath\":"; kr_go_string_append(js, path); js += "}}}";
    names.push_back(name);
  };
  const std::string fqdn = str(in->fqdn_ray_ip), head_port = str(in->head_port);
  const std::string short_ip = fqdn.substr(0, fqdn.find('.'));  // utils.ExtractRayIPFromFQDN (util.go:341-343)
  if (in->init_container) {  // common/pod.go:801-813
    add_value("FQ_RAY_IP", fqdn); add_value("RAY_IP", short_ip);
  } else {
    const bool head = in->node_type == KR_NT_HEAD;
    for (uint32_t i = 0; i < in->n_default_envs; i++) {  // configuration defaults, unless the template already sets the name (:822-827)
      const std::string n = str(in->default_envs[i].key);
      if (!exists(n.c_str())) add_value(n, str(in->default_envs[i].value));
    }
    std::string ip = "127.0.0.1";
    if (!head) { ip = fqdn; add_value("FQ_RAY_IP", ip); add_value("RAY_IP", short_ip); }
    add_field("RAY_CLUSTER_NAME", "metadata.labels['ray.io/cluster']");
    add_field("RAY_CLUSTER_NAMESPACE", "metadata.namespace");
    add_field("RAY_CLOUD_INSTANCE_ID", "metadata.name");
    add_field("RAY_NODE_TYPE_NAME", "metadata.labels['ray.io/group']");
    add_value("KUBERAY_GEN_RAY_START_CMD", str(in->ray_start_cmd));
    if (!exists("RAY_PORT")) add_value("RAY_PORT", head_port);
    if (in->crd_type == KR_CRD_RAYSERVICE) {  // :895-909
      if (!exists("RAY_timeout_ms_task_wait_for_death_info")) add_value("RAY_timeout_ms_task_wait_for_death_info", "0");
      if (!exists("RAY_gcs_server_request_timeout_seconds")) add_value("RAY_gcs_server_request_timeout_seconds", "5");
      if (!exists("RAY_SERVE_KV_TIMEOUT_S")) add_value("RAY_SERVE_KV_TIMEOUT_S", "5");
    }
    if (!exists("RAY_ADDRESS")) add_value("RAY_ADDRESS", ip + ":" + head_port);
    if (!exists("RAY_USAGE_STATS_KUBERAY_IN_USE")) add_value("RAY_USAGE_STATS_KUBERAY_IN_USE", "1");
    if (head) {
      static const char *crd[] = {"RayCluster", "RayJob", "RayService"};
      add_value("RAY_USAGE_STATS_EXTRA_TAGS", "kuberay_version=" + str(in->kuberay_version) + ";kuberay_crd=" + crd[in->crd_type <= KR_CRD_RAYSERVICE ? in->crd_type : 0]);
    }
    if (!exists("RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE")) add_value("RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE", "1");
  }
  js += ']';
  *need = js.size();
  if (js.size() > cap || (!out && !js.empty())) { g_err = "kr_ray_container_env: output buffer too small"; return KR_E_CAPACITY; }
  memcpy(out, js.data(), js.size());
  return KR_OK;
}

// supportsUnifiedHealthCheck (common/pod.go:466-475): version.ParseGeneric(rayVersion).AtLeast(2.53.0) — optional "v", two or more
// dot-separated numeric fields (the first without a leading zero), anything after them ignored (k8s.io/apimachinery pkg/util/version).
bool ray_version_at_least(const std::string &text, const unsigned long long (&min)[3]) {
  size_t i = 0, n = text.size();
  while (i < n && (text[i] == ' ' || text[i] == '\t' || text[i] == '\n' || text[i] == '\r')) i++;
  if (i < n && text[i] == 'v') i++;
  std::vector<unsigned long long> comp;
  while (i < n && text[i] >= '0' && text[i] <= '9') {
    const size_t s0 = i;
    unsigned long long v = 0;
    while (i < n && text[i] >= '0' && text[i] <= '9') { if (v > (~0ull - 9) / 10) return false; v = v * 10 + (unsigned)(text[i] - '0'); i++; }
    if (comp.empty() && i - s0 > 1 && text[s0] == '0') return false;  // zero-prefixed first component
    comp.push_back(v);
    if (i + 1 < n && text[i] == '.' && text[i + 1] >= '0' && text[i + 1] <= '9') i++; else break;
  }
  if (comp.size() < 2) return false;
  for (size_t k = 0; k < 3 || k < comp.size(); k++) {
    const unsigned long long a = k < comp.size() ? comp[k] : 0, b = k < 3 ? min[k] : 0;
    if (a != b) return a > b;
  }
  return true;
}

// initLivenessAndReadinessProbe (common/pod.go:477-573): the probes KubeRay injects when the template's Ray container has none.
int kr_ray_probes(const kr_rayprobe_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_probes: null argument"; return KR_E_INVALID; }
  if (in->node_type != KR_NT_HEAD && in->node_type != KR_NT_WORKER) { g_err = "kr_ray_probes: node_type must be KR_NT_HEAD or KR_NT_WORKER"; return KR_E_INVALID; }
  const bool head = in->node_type == KR_NT_HEAD;
  StrMap p;
  put_all(p, in->ray_start_params, in->n_ray_start_params);
  auto port = [&](const char *key, long long dflt) -> long long {  // strconv.ParseInt(s, 10, 32)
    auto it = p.find(key);
    if (it == p.end() || it->second.empty()) return dflt;
    const std::string &t = it->second;
    size_t i = 0;
    bool neg = false;
    if (t[0] == '+' || t[0] == '-') { neg = t[0] == '-'; i = 1; }
    if (i >= t.size()) return dflt;
    long long v = 0;
    for (; i < t.size(); i++) { if (t[i] < '0' || t[i] > '9') return dflt; v = v * 10 + (t[i] - '0'); if (v > 2147483648LL) return dflt; }
    if (neg) v = -v;
    return (v < -2147483648LL || v > 2147483647LL) ? dflt : v;
  };
  static const unsigned long long kMin[3] = {2, 53, 0};
  const bool http = ray_version_at_least(str(in->ray_version), kMin);
  const long long agent = port("dashboard-agent-listen-port", 52365), dash = port("dashboard-port", 8265);
  auto wget = [](long long timeout, long long prt, const char *path) {
    return "wget --tries 1 -T " + std::to_string(timeout) + " -q -O- http://localhost:" + std::to_string(prt) + "/" + path + " | grep success";
  };
  std::vector<std::string> commands = {wget(2, agent, "api/local_raylet_healthz")};
  if (head) commands.push_back(wget(10, dash, "api/gcs_healthz"));  // (the second argument really is DefaultReadinessProbeFailureThreshold, :505)
  auto join = [](const std::vector<std::string> &v) { std::string o; for (size_t i = 0; i < v.size(); i++) { if (i) o += " && "; o += v[i]; } return o; };
  auto probe = [&](bool use_http, const std::vector<std::string> &cmds, int delay, int timeout, int period, int success, int failure) {
    std::string js = "{";
    if (use_http) js += "\"httpGet\":{\"path\":\"/api/healthz\",\"port\":" + std::to_string(agent) + "}";
    else { js += "\"exec\":{\"command\":[\"bash\",\"-c\","; kr_go_string_append(js, join(cmds)); js += "]}"; }
    js += ",\"initialDelaySeconds\":" + std::to_string(delay) + ",\"timeoutSeconds\":" + std::to_string(timeout) + ",\"periodSeconds\":" + std::to_string(period) +
          ",\"successThreshold\":" + std::to_string(success) + ",\"failureThreshold\":" + std::to_string(failure) + "}";
    return js;
  };
  std::string js = "{";
  if (!in->has_liveness_probe) js += "\"livenessProbe\":" + probe(http, commands, 30, head ? 5 : 2, 5, 1, 120);
  if (!in->has_readiness_probe) {
    if (js.size() > 1) js += ',';
    bool use_http = http;
    int failure = 10;
    if (in->crd_type == KR_CRD_RAYSERVICE && !head) {  // a worker that serves traffic also checks the Serve proxy, always by exec (:557-571)
      failure = 1;
      commands.push_back(wget(10, in->serving_port > 0 ? in->serving_port : 8000, "-/healthz"));
      use_http = false;
    }
    js += "\"readinessProbe\":" + probe(use_http, commands, 10, head ? 5 : 2, 5, 1, failure);
  }
  js += '}';
  *need = js.size();
  if (js.size() > cap || (!out && !js.empty())) { g_err = "kr_ray_probes: output buffer too small"; return KR_E_CAPACITY; }
  memcpy(out, js.data(), js.size());
  return KR_OK;
}

// The emptyDir volumes BuildPod adds (common/pod.go:600-615 through addEmptyDir :1137-1161, makeEmptyDirVolume :1163-1183,
// findMemoryReqOrLimit :1204-1217): /dev/shm for the object store (memory medium, sized by the Ray container's memory limit, else
// request) unless the user set plasma-directory, and the shared Ray log directory when the head runs the autoscaler sidecar.
int kr_ray_volumes(const kr_rayvol_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_volumes: null argument"; return KR_E_INVALID; }
  std::vector<std::string> vols, ray_mounts, as_mounts;
  for (uint32_t i = 0; i < in->n_volume_names; i++) vols.push_back(str(in->volume_names[i]));
  for (uint32_t i = 0; i < in->n_ray_mount_paths; i++) ray_mounts.push_back(str(in->ray_mount_paths[i]));
  for (uint32_t i = 0; i < in->n_autoscaler_mount_paths; i++) as_mounts.push_back(str(in->autoscaler_mount_paths[i]));
  std::string jv, jr, ja;
  auto has = [](const std::vector<std::string> &v, const char *x) { return std::find(v.begin(), v.end(), x) != v.end(); };
  auto add = [&](std::vector<std::string> &mounts, std::string &jm, const char *name, const char *path, bool memory) -> int {
    if (has(mounts, path)) return KR_OK;  // already mounted (checkIfVolumeMounted compares the PATH)
    if (!has(vols, name)) {
      if (!jv.empty()) jv += ',';
      jv += "{\"name\":"; kr_go_string_append(jv, name); jv += ",\"emptyDir\":{";
      if (memory) {
        jv += "\"medium\":\"Memory\"";
        const kr_str q = (in->memory_limit.p && in->memory_limit.n) ? in->memory_limit : in->memory_request;
        if (q.p && q.n) {
          char canon[128];
          const std::string text = str(q);
          if (kr_quantity_canonical(text.c_str(), canon, sizeof canon) != KR_OK) { g_err = "kr_ray_volumes: the memory quantity does not parse"; return KR_E_INVALID; }
          jv += ",\"sizeLimit\":"; kr_go_string_append(jv, canon);
        }
      }
      jv += "}}";
      vols.push_back(name);
    }
    if (!jm.empty()) jm += ',';
    jm += "{\"name\":"; kr_go_string_append(jm, name); jm += ",\"mountPath\":"; kr_go_string_append(jm, path); jm += '}';
    mounts.push_back(path);
    return KR_OK;
  };
  if (!in->plasma_directory_set) { if (int rc = add(ray_mounts, jr, "shared-mem", "/dev/shm", true)) return rc; }
  if (in->node_type == KR_NT_HEAD && in->autoscaling_enabled) {
    if (int rc = add(ray_mounts, jr, "ray-logs", "/tmp/ray", false)) return rc;
    if (int rc = add(as_mounts, ja, "ray-logs", "/tmp/ray", false)) return rc;
  }
  const std::string js = "{\"volumes\":[" + jv + "],\"rayContainerVolumeMounts\":[" + jr + "],\"autoscalerVolumeMounts\":[" + ja + "]}";
  *need = js.size();
  if (js.size() > cap || (!out && !js.empty())) { g_err = "kr_ray_volumes: output buffer too small"; return KR_E_CAPACITY; }
  memcpy(out, js.data(), js.size());
  return KR_OK;
}

int kr_ray_start_command(const kr_raystart_in *in, uint8_t *out, uint64_t cap, uint64_t *need) {
  if (!in || !need) { g_err = "kr_ray_start_command: null argument"; return KR_E_INVALID; }
  if (in->node_type != KR_NT_HEAD && in->node_type != KR_NT_WORKER) { g_err = "kr_ray_start_command: node_type must be KR_NT_HEAD or KR_NT_WORKER"; return KR_E_INVALID; }
  const bool head = in->node_type == KR_NT_HEAD;
  StrMap p, labels, resources, limits, requests;
  put_all(p, in->ray_start_params, in->n_ray_start_params);
  put_all(labels, in->group_labels, in->n_group_labels);
  put_all(resources, in->group_resources, in->n_group_resources);
  put_all(limits, in->container_limits, in->n_container_limits);
  put_all(requests, in->container_requests, in->n_container_requests);
  // head port as the workers dial it: GetHeadPort(instance.Spec.HeadGroupSpec.RayStartParams) (raycluster_controller.go:1393,1422)
  const std::string head_port = (in->head_port.p && in->head_port.n) ? str(in->head_port) : std::string("6379");

  const uint32_t steps = in->steps ? in->steps : 0xFFFFFFFFu;  // (single steps: the reference's unit tests call them one by one)
  // DefaultHeadPodTemplate / DefaultWorkerPodTemplate (common/pod.go:179-190, 420-440)
  if (steps & KR_RS_UPDATE_RESOURCES) update_resources(p, resources);
  if (steps & KR_RS_UPDATE_LABELS) update_labels(p, labels);
  if (steps & KR_RS_SET_MISSING) {  // setMissingRayStartParams (common/pod.go:935-978)
    if (!head && !p.count("address")) p["address"] = str(in->fqdn_ray_ip) + ":" + head_port;
    if (head && !p.count("dashboard-host")) p["dashboard-host"] = "0.0.0.0";
    if (!p.count("metrics-export-port")) p["metrics-export-port"] = "8080";
    p["block"] = "true";
    if (!p.count("dashboard-agent-listen-port")) p["dashboard-agent-listen-port"] = "52365";
    if (head && in->autoscaling_enabled) p["no-monitor"] = "true";  // common/pod.go:196-200
  }

  // generateRayStartCommand (common/pod.go:980-1020)
  auto quantity_of = [](const StrMap &m, const char *k, Quantity &q) { auto it = m.find(k); if (it == m.end()) return false; q = parse_quantity(it->second); return q.ok && !q_is_zero(q); };
  Quantity q;
  if (steps & KR_RS_GENERATE) {
    if (!p.count("num-cpus")) {
      if (quantity_of(limits, "cpu", q) || quantity_of(requests, "cpu", q)) p["num-cpus"] = std::to_string(q_value(q));
    }
    if (!p.count("memory") && quantity_of(limits, "memory", q)) p["memory"] = std::to_string(q_value(q));
    add_accelerators(p, limits);
  }
  const std::string ray_start = std::string(head ? "ray start --head " : "ray start ") + convert_param_map(p);

  // the Ray container's command line (common/pod.go:617-650)
  std::vector<std::string> command, args;
  std::string cmd;
  for (uint32_t i = 0; i < in->n_command; i++) { command.push_back(str(in->command[i])); cmd += " " + command.back() + " "; }
  for (uint32_t i = 0; i < in->n_args; i++) { args.push_back(str(in->args[i])); cmd += " " + args.back() + " "; }
  const bool generated = !in->overwrite_container_cmd && cmd.find("ray start") == std::string::npos;
  if (generated) {
    command = {"/bin/bash", std::string("-c") + (in->login_shell ? "l" : ""), "--"};  // utils.GetContainerCommand
    const std::string gen = "ulimit -n 65536; " + ray_start;
    args = {cmd.empty() ? gen : cmd + " && " + gen};
  }

  std::string js = "{\"rayStartParams\":{";
  bool first = true;
  for (const auto &e : p) { if (!first) js += ','; first = false; kr_go_string_append(js, e.first); js += ':'; kr_go_string_append(js, e.second); }
  js += "},\"rayStartCommand\":";
  kr_go_string_append(js, ray_start);
  js += ",\"generated\":";
  js += generated ? "true" : "false";
  js += ",\"command\":";
  json_str_array(js, command);
  js += ",\"args\":";
  json_str_array(js, args);
  js += '}';
  *need = js.size();
  if (js.size() > cap || (!out && !js.empty())) { g_err = "kr_ray_start_command: output buffer too small"; return KR_E_CAPACITY; }
  memcpy(out, js.data(), js.size());
  return KR_OK;
}

}  // extern "C"
