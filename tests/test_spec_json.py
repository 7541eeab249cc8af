"""The native muted-spec JSON emitter (kr_spec_json_emit, kuberay_b200/csrc/kr_specjson.cpp — SURVEY §8(f) rank 2) against
  * hand-written expectations of Go's encoding/json rules and of the muting (utils/util.go:645-661),
  * the independent Python restatement kuberay_b200/specjson.py on Go-complete specs,
  * the fixed-point property on the synthetic generator's json.Marshal-shaped templates (any key order in -> the same bytes out),
  * the reference's relational hash test (rayservice_controller_unit_test.go:39-97), on hashlib here and through the GPU
    hash kernel in the `-m gpu` variant.
The emitter is host code inside libkrengine.so: these tests need no device."""
import base64
import copy
import ctypes as C
import hashlib
import json
import random

import numpy as np
import pytest

from kuberay_b200 import abi, engine, specjson, synthetic


def emit(spec, muted=True, shuffle_seed=None) -> str:
    text = json.dumps(spec, sort_keys=shuffle_seed is None)
    if shuffle_seed is not None:
        rnd = random.Random(shuffle_seed)

        def shuf(v):
            if isinstance(v, dict):
                items = [(k, shuf(x)) for k, x in v.items()]
                rnd.shuffle(items)
                return dict(items)
            if isinstance(v, list):
                return [shuf(x) for x in v]
            return v
        text = json.dumps(shuf(spec))
    return engine.spec_json_emit(text.encode(), muted).decode()


def test_go_encoding_rules_by_hand():
    spec = {"rayVersion": "2.9<&>", "suspend": False, "enableInTreeAutoscaling": True, "managedBy": "",
            "headGroupSpec": {"rayStartParams": {"b": "2", "a": "1"}, "enableIngress": False, "serviceType": "",
                              "template": {"spec": {"containers": [{"name": "h"}]}}},
            "workerGroupSpecs": [{"groupName": "g", "replicas": 3, "minReplicas": 0, "maxReplicas": 5, "numOfHosts": 0, "suspend": None,
                                  "idleTimeoutSeconds": 0, "labels": {}, "template": {"spec": {"containers": [{"name": "w"}]}}}],
            "upgradeStrategy": {"type": "Recreate"}, "unknownField": 1}
    got = emit(spec)
    assert got == ('{"suspend":false,"managedBy":"","enableInTreeAutoscaling":true,'
                   '"headGroupSpec":{"template":{"metadata":{},"spec":{"containers":[{"name":"h","resources":{}}]}},"enableIngress":false,'
                   '"rayStartParams":{"a":"1","b":"2"}},'
                   '"rayVersion":"2.9\\u003c\\u0026\\u003e",'
                   '"workerGroupSpecs":[{"groupName":"g","minReplicas":null,"maxReplicas":null,"idleTimeoutSeconds":0,"rayStartParams":null,'
                   '"template":{"metadata":{},"spec":{"containers":[{"name":"w","resources":{}}]}},"scaleStrategy":{}}]}')
    # pointers: nil is dropped, a pointer to the zero value is not; values: the zero value is dropped; nil map without omitempty: null
    assert '"suspend":false' in got and '"numOfHosts"' not in got and '"labels"' not in got and '"unknownField"' not in got
    # without the muting (utils.GenerateJsonHash callers): replicas, min/max, workersToDelete and the strategy are there
    plain = emit(spec, muted=False)
    assert '"upgradeStrategy":{"type":"Recreate"}' in plain and '"replicas":3,"minReplicas":0,"maxReplicas":5' in plain


def test_string_escaping_like_encoding_json():
    L = engine.lib()

    def esc(raw: bytes) -> str:
        text = b'{"rayVersion":' + raw + b',"headGroupSpec":{"rayStartParams":{}}}'
        out = engine.spec_json_emit(text).decode()
        return out[out.index('"rayVersion":') + len('"rayVersion":'):out.rindex("}")]

    assert esc(b'"a\\"b\\\\c/d"') == '"a\\"b\\\\c/d"'
    assert esc(b'"\\n\\r\\t\\b\\f\\u0001\\u001f"') == '"\\n\\r\\t\\b\\f\\u0001\\u001f"'
    assert esc(b'"<script>&amp;"') == '"\\u003cscript\\u003e\\u0026amp;"'
    assert esc('" x é€😀"'.encode()) == '"\\u2028x\\u2029é€😀"'
    assert esc(b'"\\ud83d\\ude00"') == '"😀"'                       # surrogate pair in the input
    assert esc(b'"a\xffb\xc3"') == '"a\\ufffdb\\ufffd"'            # invalid UTF-8 bytes -> U+FFFD
    assert L.kr_spec_json_emit(b"{", 1, 0, None, 0, C.byref(C.c_uint64())) == abi.KR_E_INVALID
    assert b"parse error" in L.kr_spec_json_last_error()


QUANTITIES = [("100m", "100m"), ("1000m", "1"), ("0.5", "500m"), ("1.5", "1500m"), ("2", "2"), ("1Gi", "1Gi"), ("1024Mi", "1Gi"),
              ("1.5Gi", "1536Mi"), ("4G", "4G"), ("1000k", "1M"), ("0.1", "100m"), ("1e3", "1e3"), ("100Mi", "100Mi"), ("1000Mi", "1000Mi"),
              ("2048Ki", "2Mi"), ("0.5Gi", "512Mi"), ("0.5Ki", "512"), ("0", "0"), ("1000", "1k"), ("1500", "1500"), ("+2", "2"), ("0.001", "1m")]


@pytest.mark.parametrize("text,want", QUANTITIES)
def test_quantity_canonical_form(text, want):
    """resource.Quantity.String(): DecimalSI mantissa without trailing zeros at an exponent that is a multiple of 3; BinarySI
    the largest power of 1024 that divides an integer value >= 1024, DecimalSI otherwise; DecimalExponent keeps its e-notation."""
    assert engine.quantity_canonical(text) == want


def test_quantities_are_canonicalised_inside_resource_lists():
    spec = {"headGroupSpec": {"rayStartParams": {}, "template": {"spec": {"containers": [
        {"name": "h", "resources": {"requests": {"memory": "2048Mi", "cpu": 0.5}, "limits": {"nvidia.com/gpu": 1, "cpu": "2000m"}}}],
        "volumes": [{"name": "v", "emptyDir": {"sizeLimit": "1024Mi"}}]}}}}
    got = emit(spec)
    assert '"resources":{"limits":{"cpu":"2","nvidia.com/gpu":"1"},"requests":{"cpu":"500m","memory":"2Gi"}}' in got
    assert '"volumes":[{"name":"v","emptyDir":{"sizeLimit":"1Gi"}}]' in got
    assert got.index('"volumes"') < got.index('"containers"')  # PodSpec declaration order, not the caller's


@pytest.mark.parametrize("t", range(4))
def test_canonical_bytes_are_a_fixed_point_for_any_key_order(t):
    """The synthetic generator's spec bodies are shaped like json.Marshal output (Go struct order, omitempty applied, metadata
    and resources present): emitting them again — from alphabetically sorted keys, as the API server serves custom resources,
    or from any shuffled order — must reproduce them byte for byte."""
    body = synthetic._json_templates()[t].decode().replace("XXXXXXXX", "00012345")
    spec = json.loads(body)
    assert emit(spec) == body
    for seed in range(5):
        assert emit(spec, shuffle_seed=seed) == body


def _go_complete(spec):
    """What the Python restatement needs spelled out (it keeps the caller's order inside pod templates): metadata and the
    containers' resources present, keys in Go declaration order."""
    spec = copy.deepcopy(spec)
    groups = [spec["headGroupSpec"]] + list(spec.get("workerGroupSpecs") or [])
    for g in groups:
        g.setdefault("rayStartParams", {})
        g.pop("workersToDelete", None)  # (fixture shorthand of scaleStrategy.workersToDelete: no such field in the CRD)
        t = g["template"]
        g["template"] = {"metadata": t.get("metadata", {}), "spec": t["spec"]}
        for c in t["spec"]["containers"]:
            c.setdefault("resources", {})
    return spec


def test_native_emitter_agrees_with_the_python_restatement():
    sc = json.load(open(__file__.rsplit("/", 1)[0] + "/golden/reconcile_scenarios.json"))["base"]["cluster"]["spec"]
    variants = [sc]
    v = copy.deepcopy(sc); v["enableInTreeAutoscaling"] = False; v["suspend"] = False; v["headServiceAnnotations"] = {"z": "1", "a": "<>"}
    variants.append(v)
    v = copy.deepcopy(sc); v["workerGroupSpecs"].append(copy.deepcopy(v["workerGroupSpecs"][0])); v["workerGroupSpecs"][1].update(groupName="g2", numOfHosts=4, suspend=True)
    variants.append(v)
    v = copy.deepcopy(sc); v["headGroupSpec"]["template"]["spec"]["tolerations"] = [{"key": "k"}]; v["upgradeStrategy"] = {"type": "Recreate"}
    variants.append(v)
    for spec in variants:
        full = _go_complete(spec)
        assert emit(full) == specjson.muted_spec_json(full).decode()
        assert emit(spec) == emit(full)  # ... and the defaults spelled out by _go_complete are what the emitter adds by itself


def _relations(h):
    """TestGenerateHashWithoutReplicasAndWorkersToDelete rayservice_controller_unit_test.go:39-97 (+ the other muted fields)."""
    sc = json.load(open(__file__.rsplit("/", 1)[0] + "/golden/reconcile_scenarios.json"))["base"]["cluster"]["spec"]
    sc = copy.deepcopy(sc)
    sc["workerGroupSpecs"][0].pop("workersToDelete", None)
    variants = {"base": sc}
    s = copy.deepcopy(sc); s["workerGroupSpecs"][0]["replicas"] += 1; variants["replicas+1"] = s
    s = copy.deepcopy(sc); s["rayVersion"] = "2.100.0"; variants["rayVersion"] = s
    s = copy.deepcopy(sc)
    s["headGroupSpec"]["template"]["spec"]["tolerations"] = [{"key": "k", "operator": "Exists"}]
    s["workerGroupSpecs"][0]["template"]["spec"]["tolerations"] = [{"key": "k", "operator": "Exists"}]
    variants["tolerations"] = s
    s = copy.deepcopy(sc); s["headGroupSpec"]["template"]["spec"]["schedulingGates"] = [{"name": "kueue.x-k8s.io/admission"}]; variants["gates"] = s
    s = copy.deepcopy(sc)
    s["workerGroupSpecs"][0]["scaleStrategy"] = {"workersToDelete": ["a", "b"]}; s["workerGroupSpecs"][0]["minReplicas"] = 7
    s["workerGroupSpecs"][0]["maxReplicas"] = 9; s["upgradeStrategy"] = {"type": "Recreate"}
    variants["scale+strategy"] = s
    s = copy.deepcopy(sc); s["workerGroupSpecs"][0]["template"]["spec"]["containers"][0]["image"] = "other"; variants["image"] = s
    names = list(variants)
    digests = dict(zip(names, h([engine.spec_json_emit(json.dumps(variants[n]).encode()) for n in names])))
    base = digests["base"]
    assert len(base) == 32 and set(base) <= set("0123456789ABCDEFGHIJKLMNOPQRSTUV")
    for same in ("replicas+1", "tolerations", "gates", "scale+strategy"):
        assert digests[same] == base, same
    for other in ("rayVersion", "image"):
        assert digests[other] != base, other


def test_hash_relations_of_the_emitted_bytes():
    _relations(lambda msgs: [base64.b32hexencode(hashlib.sha1(m).digest()).decode() for m in msgs])


@pytest.mark.gpu
def test_hash_relations_through_the_gpu_hash_kernel():
    from kuberay_b200.engine import Engine
    eng = Engine(0, max_clusters=1)
    try:
        _relations(eng.hash_batch)
    finally:
        eng.close()


def test_emit_into_the_json_arena():
    L = engine.lib()
    arena = np.full(4096, 0xAB, dtype=np.uint8)
    cursor, off, ln = C.c_uint64(5), C.c_uint64(), C.c_uint32()
    spec = b'{"headGroupSpec":{"rayStartParams":{}},"rayVersion":"2.46.0"}'
    want = engine.spec_json_emit(spec)
    for k in range(3):
        assert L.kr_spec_json_emit_arena(spec, len(spec), arena.ctypes.data, arena.size, C.byref(cursor), C.byref(off), C.byref(ln)) == 0
        assert off.value % 16 == 0 and cursor.value % 16 == 0 and ln.value == len(want)
        assert bytes(arena[off.value:off.value + ln.value]) == want
        assert not arena[off.value + ln.value:cursor.value].any()      # zero padding up to the next 16-byte piece
    small = np.zeros(64, dtype=np.uint8)
    cursor = C.c_uint64(0)
    assert L.kr_spec_json_emit_arena(spec, len(spec), small.ctypes.data, small.size, C.byref(cursor), C.byref(off), C.byref(ln)) == abi.KR_E_CAPACITY


# ---------------------------------------------------------------------------------------------- f4: RayService hash comparison
# TestIsClusterSpecHashEqual, rayservice_controller_unit_test.go:1057-1148 (table transcribed: partial, diffReplicas,
# addNewWorkerGroup, updateClusterSpec -> expected) + the branches of isClusterSpecHashEqual the table does not reach
# (rayservice_controller.go:1139-1153: Atoi failure => true, fewer goal groups => goal hash "").
IS_EQUAL_TABLE = [
    ("[full] diff replicas", False, True, False, False, True),
    ("[full] completely identical", False, False, False, False, True),
    ("[full] update cluster spec", False, False, False, True, False),
    ("[partial] new worker group", True, False, True, False, True),
    ("[partial] diff replicas + new worker group", True, True, True, False, True),
    ("[partial] diff replicas", True, True, False, False, True),
    ("[partial] update cluster spec", True, False, False, True, False),
]


@pytest.mark.gpu
def test_hash_compare_batch_is_cluster_spec_hash_equal():
    from kuberay_b200.engine import Engine
    sc = json.load(open(__file__.rsplit("/", 1)[0] + "/golden/reconcile_scenarios.json"))["base"]["cluster"]["spec"]
    sc = copy.deepcopy(sc)
    sc["workerGroupSpecs"][0].pop("workersToDelete", None)
    base_hash = base64.b32hexencode(hashlib.sha1(engine.spec_json_emit(json.dumps(sc).encode())).digest()).decode()
    nwg = str(len(sc["workerGroupSpecs"]))
    rows, want = [], []
    for _name, partial, diff_replicas, add_group, update_spec, expected in IS_EQUAL_TABLE:
        svc = copy.deepcopy(sc)
        if diff_replicas:
            svc["workerGroupSpecs"][0]["replicas"] += 1
        if add_group:
            svc["workerGroupSpecs"].append({"groupName": "worker-group-2", "replicas": 1})
        if update_spec:
            svc["rayVersion"] = "new-version"
        rows.append((json.dumps(svc).encode(), base_hash, nwg, partial))
        want.append(expected)
    two = copy.deepcopy(sc); two["workerGroupSpecs"].append({"groupName": "g2"})
    extra = [
        ((json.dumps(sc).encode(), base_hash, "one", True), True),        # Atoi fails: the reference returns true (:1140-1142)
        ((json.dumps(sc).encode(), base_hash, " 1", True), True),         # strconv.Atoi takes no spaces
        ((json.dumps(sc).encode(), base_hash, "+1", True), True),         # ... but a sign
        ((json.dumps(sc).encode(), base_hash, "2", True), False),         # fewer goal groups than the cluster: goal hash "" != annotation
        ((json.dumps(sc).encode(), None, "2", True), True),               # ... "" == "" (no annotation)
        ((json.dumps(sc).encode(), None, nwg, False), False),             # full compare against a missing annotation
        ((b"{not json", "", nwg, False), True),                           # the error of :1135 is dropped: "" == ""
        ((b"{not json", base_hash, nwg, True), True),                     # :1151-1153
        ((json.dumps(two).encode(), base_hash, "0", True), False),        # the first 0 groups: a different spec
        ((json.dumps(sc).encode(), base_hash.lower(), nwg, False), False),
    ]
    rows += [r for r, _ in extra]; want += [w for _, w in extra]
    eng = Engine(0, max_clusters=1)
    try:
        got, hashes = eng.hash_compare_batch(rows)
        assert eng.hash_compare_batch([]) == ([], [])
    finally:
        eng.close()
    assert got == want, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert hashes[1] == base_hash and hashes[3] == base_hash and hashes[2] != base_hash and hashes[len(IS_EQUAL_TABLE) + 3] == ""


@pytest.mark.gpu
def test_hash_compare_batch_in_bulk():
    """Enough rows for the emit thread pool and the multi-threaded staging copy (>= 64 rows per thread), every decision and digest against
    the row-by-row form (native emitter + hashlib) — including rows the host settles without a digest, interleaved."""
    from kuberay_b200.engine import Engine
    rng = random.Random(7)
    rows, want_eq, want_hash = [], [], []
    for i in range(700):
        groups = [{"groupName": f"g{k}", "replicas": rng.randint(0, 9), "rayStartParams": {}, "template": {"spec": {"containers": [
            {"name": "w", "image": f"ray:{i % 13}", "env": [{"name": f"E{j}", "value": "x" * rng.randint(0, 40)} for j in range(rng.randint(0, 30))]}]}}} for k in range(rng.randint(0, 3))]
        spec = {"rayVersion": f"2.{i % 7}", "headGroupSpec": {"rayStartParams": {}, "template": {"spec": {"containers": [{"name": "h", "image": "ray"}]}}}, "workerGroupSpecs": groups}
        text = json.dumps(spec).encode()
        kind = rng.choice(["equal", "stale", "partial", "partial-short", "bad-atoi", "bad-json"])
        full = base64.b32hexencode(hashlib.sha1(engine.spec_json_emit(text)).digest()).decode()
        if kind == "equal":
            rows.append((text, full, None, False)); want_eq.append(True); want_hash.append(full)
        elif kind == "stale":
            rows.append((text, "0" * 32, None, False)); want_eq.append(False); want_hash.append(full)
        elif kind == "partial":          # the cluster was created from the first k groups
            k = rng.randint(0, len(groups))
            cut = dict(spec, workerGroupSpecs=groups[:k])
            h = base64.b32hexencode(hashlib.sha1(engine.spec_json_emit(json.dumps(cut).encode())).digest()).decode()
            rows.append((text, h, str(k), True)); want_eq.append(True); want_hash.append(h)
        elif kind == "partial-short":    # the cluster has MORE groups than the goal: goal hash stays ""
            rows.append((text, full, str(len(groups) + 1), True)); want_eq.append(False); want_hash.append("")
        elif kind == "bad-atoi":
            rows.append((text, full, "x1", True)); want_eq.append(True); want_hash.append("")
        else:
            rows.append((b"{" + text, "", None, False)); want_eq.append(True); want_hash.append("")
    eng = Engine(0, max_clusters=1)
    try:
        got, hashes = eng.hash_compare_batch(rows)
    finally:
        eng.close()
    assert got == want_eq and hashes == want_hash


# ---- property: arbitrary JSON through the parser and Go's string encoder ------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_json_leaf = st.one_of(st.none(), st.booleans(), st.integers(min_value=-2**53, max_value=2**53), st.text(max_size=40))
_json_value = st.recursive(_json_leaf, lambda inner: st.one_of(st.lists(inner, max_size=5), st.dictionaries(st.text(max_size=12), inner, max_size=5)), max_leaves=40)


@settings(max_examples=300, deadline=None)
@given(payload=st.lists(_json_value, min_size=1, max_size=4), ascii_only=st.booleans())
def test_any_json_survives_the_parser_and_the_go_string_encoder(payload, ascii_only):
    """A field the tables leave untyped (`ephemeralContainers`, kind raw) carries arbitrary JSON: whatever the parser read must come back, value for
    value, through the emitter — escapes (\\uXXXX input, HTML-safe output), multi-byte UTF-8, astral code points, control characters, nesting,
    repeated structure — and the output must stay valid JSON that encoding/json would have produced (no raw <, >, &, U+2028/9)."""
    spec = {"headGroupSpec": {"rayStartParams": {}, "template": {"spec": {"containers": [{"name": "c"}], "ephemeralContainers": payload}}}}
    text = json.dumps(spec, ensure_ascii=ascii_only).encode("utf-8", "surrogatepass") if ascii_only else json.dumps(spec, ensure_ascii=False).encode("utf-8", "surrogatepass")
    try:
        text.decode("utf-8")
    except UnicodeDecodeError:
        return  # a lone surrogate written raw is not UTF-8: not a case the API server can produce
    out = engine.spec_json_emit(text)
    back = json.loads(out)["headGroupSpec"]["template"]["spec"]["ephemeralContainers"]
    want = json.loads(text)["headGroupSpec"]["template"]["spec"]["ephemeralContainers"]
    # a lone surrogate escape decodes to U+FFFD in Go
    fix = lambda v: (v.encode("utf-16", "surrogatepass").decode("utf-16", "replace") if isinstance(v, str) else  # noqa: E731
                     [fix(x) for x in v] if isinstance(v, list) else {fix(k): fix(x) for k, x in v.items()} if isinstance(v, dict) else v)
    assert back == fix(want)
    assert not any(ch in out for ch in (b"<", b">", b"&", " ".encode(), " ".encode()))
    assert engine.spec_json_emit(out) == out                  # and the canonical bytes are a fixed point
