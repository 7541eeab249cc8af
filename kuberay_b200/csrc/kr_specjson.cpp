// kr_specjson.cpp — native canonical muted-spec JSON emitter (host code; SURVEY §8(f) rank 2, include/kr_engine.h kr_spec_json_*).
//
// Produces the bytes the reference hashes: json.Marshal(mute(RayClusterSpec)) as in
// utils.GenerateHashWithoutReplicasAndWorkersToDelete (ray-operator/controllers/ray/utils/util.go:642-665), from the spec as
// JSON text in ANY key order (the API server serves custom resources with alphabetically sorted keys, not in Go struct order),
// so the operator no longer DeepCopies + reflect-marshals the spec on every reconcile: the shim emits once per
// metadata.generation straight into the engine's 16-byte aligned JSON arena.
//
// What it restates:
//   * the muting (util.go:645-661): upgradeStrategy, every worker group's replicas / minReplicas / maxReplicas /
//     scaleStrategy.workersToDelete, and tolerations + schedulingGates of every pod template;
//   * Go's encoding/json: fields in struct declaration order, `omitempty` dropping nil pointers / false / 0 /
//     "" / empty slices and maps but never a struct value and never a non-nil pointer (so `"suspend":false` stays), `null` for
//     nil maps / slices / pointers WITHOUT omitempty (minReplicas, maxReplicas, rayStartParams, containers), map keys sorted
//     bytewise, strings escaped like encoding/json with HTML escaping on, no whitespace;
//   * resource.Quantity's canonical string (k8s.io/apimachinery pkg/api/resource: Quantity.CanonicalizeBytes) for
//     ResourceList values and sizeLimit / divisor.
// The struct tables for rayv1 follow ray-operator/apis/ray/v1/raycluster_types.go:13-225 (in the reference tree).  The tables
// for the embedded corev1.PodTemplateSpec come from k8s.io/api v0.36.0 (ray-operator/go.mod:23), which is NOT vendored in the
// reference: they are restated from the published API and are byte-level UNVERIFIED here; object types the
// tables do not know keep the caller's key order.  Unknown keys inside a known struct are dropped, as a typed decode drops them.
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/kr_engine.h"

#include "kr_json.hpp"

namespace {
using namespace krjson;

// util.go:645-661 on the tree
void mute_template(Node *tmpl) {
  if (!tmpl || tmpl->t != N_OBJ) return;
  Node *spec = tmpl->get("spec");
  if (spec && spec->t == N_OBJ) { spec->erase("tolerations"); spec->erase("schedulingGates"); }
}
void mute(Node &spec) {
  if (spec.t != N_OBJ) return;
  spec.erase("upgradeStrategy");
  if (Node *h = spec.get("headGroupSpec")) if (h->t == N_OBJ) mute_template(h->get("template"));
  if (Node *ws = spec.get("workerGroupSpecs"))
    if (ws->t == N_ARR)
      for (Node &w : ws->a) {
        if (w.t != N_OBJ) continue;
        w.erase("replicas"); w.erase("minReplicas"); w.erase("maxReplicas");
        if (Node *ss = w.get("scaleStrategy")) if (ss->t == N_OBJ) ss->erase("workersToDelete");
        mute_template(w.get("template"));
      }
}

thread_local std::string g_err;

// max_groups >= 0: keep only the first max_groups worker groups (rayservice_controller.go:1146-1147) — when the spec has fewer, nothing
// is emitted and KR_E_STATE comes back with *n_groups set.
int emit_impl(const uint8_t *spec_json, uint64_t len, bool muted, std::string &out, long max_groups = -1, long *n_groups = nullptr) {
  Parser ps{reinterpret_cast<const char *>(spec_json), reinterpret_cast<const char *>(spec_json) + len, {}};
  Node root;
  if (!ps.value(root, 0)) { g_err = "kr_spec_json: parse error: " + ps.err; return KR_E_INVALID; }
  ps.ws();
  if (ps.p != ps.end) { g_err = "kr_spec_json: trailing characters after the JSON value"; return KR_E_INVALID; }
  if (root.t != N_OBJ) { g_err = "kr_spec_json: the spec must be a JSON object"; return KR_E_INVALID; }
  {
    Node *ws = root.get("workerGroupSpecs");
    const long have = (ws && ws->t == N_ARR) ? (long)ws->a.size() : 0;
    if (n_groups) *n_groups = have;
    if (max_groups >= 0) {
      if (have < max_groups) return KR_E_STATE;
      if (ws && ws->t == N_ARR) ws->a.resize((size_t)max_groups);
    }
  }
  if (muted) mute(root);
  Emitter em;
  em.out.reserve((size_t)len + 64);
  em.strct("RayClusterSpec", &root);
  if (!em.err.empty()) { g_err = "kr_spec_json: " + em.err; return KR_E_INVALID; }
  out.swap(em.out);
  return KR_OK;
}

}  // namespace

// for kr_podmeta.cpp: the same Go string encoder
void kr_go_string_append(std::string &out, const std::string &s) { go_string(out, s); }

// for kr_engine.cu (kr_hash_compare_batch): same emitter, optional truncation of the worker groups
int kr_specjson_emit_string(const uint8_t *spec_json, uint64_t len, bool muted, long max_groups, std::string &out, long *n_groups) {
  return emit_impl(spec_json, len, muted, out, max_groups, n_groups);
}

extern "C" {

int kr_spec_json_emit(const uint8_t *spec_json, uint64_t len, uint32_t flags, uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
  if ((!spec_json && len) || !out_len) return KR_E_INVALID;
  std::string s;
  int rc = emit_impl(spec_json, len, !(flags & KR_SPEC_JSON_UNMUTED), s);
  if (rc) return rc;
  *out_len = s.size();
  if (s.size() > out_cap || (!out && s.size())) { g_err = "kr_spec_json: output buffer too small"; return KR_E_CAPACITY; }
  memcpy(out, s.data(), s.size());
  return KR_OK;
}

int kr_spec_json_emit_arena(const uint8_t *spec_json, uint64_t len, uint8_t *arena, uint64_t arena_cap, uint64_t *cursor, uint64_t *off_out, uint32_t *len_out) {
  if ((!spec_json && len) || !arena || !cursor || !off_out || !len_out) return KR_E_INVALID;
  std::string s;
  int rc = emit_impl(spec_json, len, true, s);
  if (rc) return rc;
  const uint64_t off = (*cursor + 15) & ~15ull, padded = (s.size() + 15) & ~15ull;
  if (s.size() > 0xFFFFFFFFull || off + padded > arena_cap) { g_err = "kr_spec_json: JSON arena full"; return KR_E_CAPACITY; }
  memcpy(arena + off, s.data(), s.size());
  memset(arena + off + s.size(), 0, padded - s.size());  // the hash kernels fetch whole 16-byte pieces
  *off_out = off; *len_out = (uint32_t)s.size(); *cursor = off + padded;
  return KR_OK;
}

int kr_quantity_canonical(const char *text, char *out, uint64_t out_cap) {
  if (!text || !out) return KR_E_INVALID;
  std::string c;
  if (!canon_quantity(text, c)) { g_err = "kr_quantity_canonical: not a quantity"; return KR_E_INVALID; }
  if (c.size() + 1 > out_cap) return KR_E_CAPACITY;
  memcpy(out, c.c_str(), c.size() + 1);
  return KR_OK;
}

const char *kr_spec_json_last_error(void) { return g_err.c_str(); }

}  // extern "C"
