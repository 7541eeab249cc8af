"""Muted-spec JSON for the RayCluster spec hash — TEST canonicalizer (an independent restatement used to cross-check the
native emitter kuberay_b200/csrc/kr_specjson.cpp, which is what the product path calls: include/kr_engine.h kr_spec_json_emit).

The hash input is json.Marshal of the muted RayClusterSpec (ray-operator/controllers/ray/utils/util.go:629,642-661); the GPU
does SHA-1 + base32hex.  This module restates the muting (util.go:645-661) and the Go encoding/json rules of
SURVEY.md Appendix B for the subset of fields the fixtures / synthetic generator emit, so that the reference's
*relational* hash tests (rayservice_controller_unit_test.go:39-97) can be replayed.  Literal digest parity with Go
is unpinned in the reference (no golden digest exists in its tree); nested corev1 structs are emitted in the
caller's dict order (the caller is responsible for Go declaration order).
"""
from __future__ import annotations

import copy

# RayClusterSpec field order (apis/ray/v1/raycluster_types.go:13-53); True = omitempty
_SPEC_FIELDS = [("upgradeStrategy", True), ("authOptions", True), ("suspend", True), ("managedBy", True),
                ("autoscalerOptions", True), ("headServiceAnnotations", True), ("enableInTreeAutoscaling", True),
                ("gcsFaultToleranceOptions", True), ("headGroupSpec", False), ("rayVersion", True), ("workerGroupSpecs", True)]
# HeadGroupSpec (:129-154)
_HEAD_FIELDS = [("template", False), ("headService", True), ("enableIngress", True), ("resources", True), ("labels", True),
                ("rayStartParams", False), ("serviceType", True)]
# WorkerGroupSpec (:157-207)
_WORKER_FIELDS = [("suspend", True), ("groupName", False), ("replicas", True), ("minReplicas", False), ("maxReplicas", False),
                  ("idleTimeoutSeconds", True), ("resources", True), ("labels", True), ("rayStartParams", False),
                  ("template", False), ("scaleStrategy", False), ("numOfHosts", True)]
_MAP_FIELDS = {"headServiceAnnotations", "resources", "labels", "rayStartParams", "annotations", "nodeSelector", "limits", "requests"}


def _is_empty(v) -> bool:
    return v is None or v is False or v == 0 or v == "" or v == [] or v == {}


def _enc_str(s: str) -> str:
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif o < 0x20:
            out.append(f"\\u{o:04x}")
        elif ch in "<>&":
            out.append(f"\\u{o:04x}")  # encoding/json HTML escaping is on by default
        elif o in (0x2028, 0x2029):
            out.append(f"\\u{o:04x}")
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def _enc(v, key: str | None = None) -> str:
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return repr(v)
    if isinstance(v, str):
        return _enc_str(v)
    if isinstance(v, (list, tuple)):
        return "[" + ",".join(_enc(x) for x in v) + "]"
    if isinstance(v, dict):
        items = sorted(v.items()) if key in _MAP_FIELDS else v.items()  # Go maps: keys sorted bytewise
        return "{" + ",".join(_enc_str(k) + ":" + _enc(x, k) for k, x in items) + "}"
    raise TypeError(type(v))


# pointer-typed fields (*bool / *int32 / *string): omitempty drops them only when nil — `"suspend":false` stays
_POINTER_FIELDS = {"suspend", "managedBy", "enableInTreeAutoscaling", "enableIngress", "idleTimeoutSeconds", "replicas"}


def _struct(obj: dict, fields, enc_field) -> str:
    parts = []
    for name, omitempty in fields:
        v = obj.get(name)
        if omitempty and (v is None if name in _POINTER_FIELDS else _is_empty(v)):
            continue
        parts.append(_enc_str(name) + ":" + enc_field(name, v))
    return "{" + ",".join(parts) + "}"


def _mute_template(t):
    """util.go:650-651,658-659: Tolerations and SchedulingGates set to nil (omitempty => dropped)."""
    if not isinstance(t, dict):
        return t
    t = copy.deepcopy(t)
    spec = t.get("spec")
    if isinstance(spec, dict):
        spec.pop("tolerations", None)
        spec.pop("schedulingGates", None)
    return t


def muted_spec_json(spec: dict) -> bytes:
    """json.Marshal(mute(spec)) — GenerateHashWithoutReplicasAndWorkersToDelete's hash input (util.go:642-665)."""
    spec = copy.deepcopy(spec or {})
    spec["upgradeStrategy"] = None  # util.go:661
    if isinstance(spec.get("upgradeStrategy"), dict):
        spec["upgradeStrategy"] = None

    def enc_head(name, v):
        if name == "template":
            return _enc(_mute_template(v) if v is not None else {"metadata": {}, "spec": {"containers": None}})
        return _enc(v, name)

    def enc_worker_field(name, v):
        if name == "template":
            return _enc(_mute_template(v) if v is not None else {"metadata": {}, "spec": {"containers": None}})
        if name == "scaleStrategy":
            return "{}"  # workersToDelete nil'ed (util.go:657); a struct value is always emitted
        if name in ("minReplicas", "maxReplicas"):
            return "null"  # nil'ed, no omitempty (util.go:655-656)
        return _enc(v, name)

    def enc_spec_field(name, v):
        if name == "headGroupSpec":
            return _struct(v or {}, _HEAD_FIELDS, enc_head)
        if name == "workerGroupSpecs":
            out = []
            for w in v:
                w = dict(w)
                w["replicas"] = None  # util.go:654 (omitempty => dropped)
                w.setdefault("groupName", "")
                out.append(_struct(w, _WORKER_FIELDS, enc_worker_field))
            return "[" + ",".join(out) + "]"
        return _enc(v, name)

    return _struct(spec, _SPEC_FIELDS, enc_spec_field).encode("utf-8")
