"""CPU restatement of the reference's Pod-metadata step — TEST INFRASTRUCTURE ONLY.

Only tests/ may import this; the product path is kuberay_b200/csrc/kr_podmeta.cpp behind the C ABI.  Each function follows the
reference lines it cites (ray-operator/controllers/ray/...); nothing is copied.  Pinned by tests/test_podmeta.py against the
reference's own unit-test vectors (utils/util_test.go:110-248, common/pod_test.go:1328-1350, 2188-2227)."""
from __future__ import annotations

import json

# unicode.IsPunct(rune(b)) for a byte widened to a rune: ASCII category P, plus the Latin-1 members
_PUNCT = set(b"!\"#%&'()*,-./:;?@[\\]_{}") | {0xA1, 0xA7, 0xAB, 0xB6, 0xB7, 0xBB, 0xBF}


def _enc(s: str) -> bytes:
    return s.encode("utf-8", "surrogateescape")


def _dec(b: bytes) -> str:
    return b.decode("utf-8", "surrogateescape")


def check_name(s: bytes) -> bytes:
    """utils/util.go:217-240."""
    if len(s) > 50:
        s = s[len(s) - 50:]
    if 0x30 <= s[0] <= 0x39:
        s = b"r" + s[1:]
    if s[0] in _PUNCT:
        s = b"r" + s[1:]
    return s


def check_label(s: bytes) -> bytes:
    """utils/util.go:247-265."""
    if len(s) > 63:
        s = s[len(s) - 63:]
    if s[0] in _PUNCT:
        s = b"r" + s[1:]
    return s


def pod_name(prefix: bytes, node_type: str, is_generate_name: bool) -> bytes:
    """utils/util.go:198-215 (ToLower on ASCII letters: prefixes are RFC 1123 names)."""
    r = bytes((c + 32 if 0x41 <= c <= 0x5A else c) for c in prefix[:50] + b"-" + node_type.encode())
    return r + (b"-" if is_generate_name else b"")


def label_pod(node_type: str, cluster: str, group: str, override: dict) -> dict:
    """common/pod.go:775-799."""
    labels = {
        "ray.io/is-ray-node": "yes", "ray.io/cluster": cluster, "ray.io/node-type": node_type, "ray.io/group": group,
        "ray.io/identifier": _dec(check_label(_enc(f"{cluster}-{node_type}"))),
        "app.kubernetes.io/name": "kuberay", "app.kubernetes.io/created-by": "kuberay-operator",
    }
    for k, v in override.items():
        if k in ("ray.io/node-type", "ray.io/group", "ray.io/cluster"):
            continue
        labels[k] = v
    return labels


def merge_labels(template_labels: dict | None, group_labels: dict | None) -> dict:
    """common/pod.go:1276-1283."""
    m = dict(template_labels or {})
    m.update(group_labels or {})
    return m


def pod_meta(cluster: dict, create: tuple, *, kuberay_version="v1.5.0", deterministic_head_name=False,
             multihost_indexing_gate=True, cluster_hash: str | None = None) -> dict:
    """ObjectMeta of one Pod: DefaultHeadPodTemplate / DefaultWorkerPodTemplate metadata (common/pod.go:166-190, 352-357, 420-443),
    BuildPod's serve label (:583-588), createHeadPod's annotations (raycluster_controller.go:1313-1316), SetControllerReference."""
    spec = cluster.get("spec") or {}
    annots = cluster.get("annotations") or {}
    name, ns = cluster["name"], cluster.get("namespace", "default")
    g, replica_index, host_index, replica_name = create
    head = g < 0
    grp = (spec.get("headGroupSpec") or {}) if head else spec["workerGroupSpecs"][g]
    tmeta = (grp.get("template") or {}).get("metadata") or {}
    node_type = "head" if head else "worker"
    gname = "headgroup" if head else grp.get("groupName", "")
    labels = label_pod(node_type, name, gname, merge_labels(tmeta.get("labels"), grp.get("labels")))
    if not head and multihost_indexing_gate:                                              # common/pod.go:430-439
        labels["ray.io/worker-group-replica-index"] = str(replica_index)
        if int(grp.get("numOfHosts", 1)) > 1:
            labels["ray.io/worker-group-replica-name"] = replica_name
            labels["ray.io/replica-host-index"] = str(host_index)
    crd = (cluster.get("labels") or {}).get("ray.io/originated-from-crd")
    if crd == "RayService":                                                               # common/pod.go:583-588
        labels["ray.io/serve"] = "false" if head else "true"
    a = dict(tmeta.get("annotations") or {})
    if "ray.io/overwrite-container-cmd" in annots and annots["ray.io/overwrite-container-cmd"].lower() == "true":   # :62-75
        a["ray.io/overwrite-container-cmd"] = "true"
    ft_opts = spec.get("gcsFaultToleranceOptions")
    ft = ("ray.io/ft-enabled" in annots and annots["ray.io/ft-enabled"].lower() == "true") or ft_opts is not None  # util.go:753-756
    if head:
        a["ray.io/ft-enabled"] = "true" if ft else "false"                                # :85-87
        if ft:                                                                            # :105-114
            sns = cluster.get("uid", "")
            if "ray.io/external-storage-namespace" in annots:
                sns = annots["ray.io/external-storage-namespace"]
            if ft_opts and ft_opts.get("externalStorageNamespace"):
                sns = ft_opts["externalStorageNamespace"]
            a["ray.io/external-storage-namespace"] = sns
        if cluster_hash:                                                                  # raycluster_controller.go:1313-1316
            a["ray.io/upgrade-strategy-recreate-hash"] = cluster_hash
            a["ray.io/kuberay-version"] = kuberay_version
    if head:
        nm = _dec(pod_name(_enc(name), "head", not deterministic_head_name))
        key = "name" if deterministic_head_name else "generateName"
    else:
        nm = _dec(pod_name(_enc(f"{name}-{gname}"), "worker", True))
        key = "generateName"
    return {key: nm, "namespace": ns, "labels": labels, "annotations": a,
            "ownerReferences": [{"apiVersion": "ray.io/v1", "kind": "RayCluster", "name": name, "uid": cluster.get("uid", ""),
                                 "controller": True, "blockOwnerDeletion": True}]}


def go_string(s: str) -> str:
    """encoding/json's string encoding with HTML escaping, for valid UTF-8 input."""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o < 0x20 or ch in "<>&":
            out.append("\\u%04x" % o)
        elif o in (0x2028, 0x2029):
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def go_marshal(meta: dict) -> bytes:
    """The bytes encoding/json writes for the patch: fields in the order ObjectMeta declares them, map keys sorted by bytes."""
    def m(d: dict) -> str:
        return "{" + ",".join(f"{go_string(k)}:{go_string(v)}" for k, v in sorted(d.items(), key=lambda kv: _enc(kv[0]))) + "}"
    key = "name" if "name" in meta else "generateName"
    o = meta["ownerReferences"][0]
    return _enc(f'{{"{key}":{go_string(meta[key])},"namespace":{go_string(meta["namespace"])},"labels":{m(meta["labels"])},'
                f'"annotations":{m(meta["annotations"])},"ownerReferences":[{{"apiVersion":"ray.io/v1","kind":"RayCluster","name":{go_string(o["name"])},'
                f'"uid":{go_string(o["uid"])},"controller":true,"blockOwnerDeletion":true}}]}}')


def expand_creates(group_results, create_idx, groups: list[dict], head_create: bool, multihost_indexing_gate=True) -> list[tuple]:
    """Create order of reconcilePods (:692-735 head; :869-889 workers; :1081-1094 multi-host) as (group, replicaIndex, hostIndex);
    the generated replica name is random in the reference, so it is not part of the restated tuple."""
    MULTIHOST = 1 << 0
    out = [(-1, 0, 0)] if head_create else []
    for g, (gr, grp) in enumerate(zip(group_results, groups)):
        mh = multihost_indexing_gate and int(grp.get("numOfHosts", 1)) > 1
        for k in range(int(gr["n_create"])):
            idx = int(create_idx[int(gr["create_off"]) + k]) if multihost_indexing_gate else 0
            if mh:
                out += [(g, idx, j) for j in range(int(grp["numOfHosts"]))]
            else:
                out.append((g, idx, 0))
    return out


# ------------------------------------------------------------------------------------------------ `ray start` command (second part of f3)
# Restated from common/pod.go:935-1135, 1219-1276, 617-650 and utils/resources.go:8-17; resource.Quantity and encoding/json float
# formatting from their published behaviour (k8s.io/apimachinery v0.36.0 is not vendored under /root/reference).
import math
import re
from fractions import Fraction

_QTY = re.compile(r"^([+-]?)(\d+\.?\d*|\.\d+)(Ki|Mi|Gi|Ti|Pi|Ei|[numkMGTPE]|[eE][+-]?\d+)?$")
_DEC = {"n": -9, "u": -6, "m": -3, "k": 3, "M": 6, "G": 9, "T": 12, "P": 15, "E": 18}


def parse_quantity(text: str):
    """-> exact Fraction, or None when the text is not a quantity."""
    m = _QTY.match(text)
    if not m:
        return None
    sign, num, suf = m.groups()
    v = Fraction(num if not num.endswith(".") else num + "0") if not num.startswith(".") else Fraction("0" + num)
    if suf:
        if suf in ("Ki", "Mi", "Gi", "Ti", "Pi", "Ei"):
            v *= 2 ** (10 * ("KMGTPE".index(suf[0]) + 1))
        elif suf in _DEC:
            v *= Fraction(10) ** _DEC[suf]
        else:
            v *= Fraction(10) ** int(suf[1:])
    return -v if sign == "-" else v


def quantity_value(v: Fraction) -> int:
    return math.ceil(v)


def go_pow10(n: int) -> float:
    """math.Pow10: literal tables, negative powers by one division."""
    if 0 <= n <= 308:
        return float(f"1e{32 * (n // 32)}") * float(f"1e{n % 32}")
    if -323 <= n < 0:
        return float(f"1e-{32 * ((-n) // 32)}") / float(f"1e{(-n) % 32}")
    return math.inf if n > 0 else 0.0


def quantity_float(text: str) -> float:
    """AsApproximateFloat64 on the parsed text: float64(unscaled) * math.Pow10(-scale); finer than nano is rounded up to nano first."""
    m = _QTY.match(text)
    sign, num, suf = m.groups()
    ip, _, fp = num.partition(".")
    mant, e10, e2 = int((ip + fp) or "0"), -len(fp), 0
    if suf in ("Ki", "Mi", "Gi", "Ti", "Pi", "Ei"):
        e2 = 10 * ("KMGTPE".index(suf[0]) + 1)
    elif suf in _DEC:
        e10 += _DEC[suf]
    elif suf:
        e10 += int(suf[1:])
    mant <<= e2
    if e10 < -9 and mant:
        mant = -(-mant // 10 ** (-9 - e10))
        e10 = -9
    v = float(mant) if e10 == 0 else float(mant) * go_pow10(e10)
    return -v if sign == "-" else v


def go_float(f: float) -> str:
    """encoding/json float64: shortest repr, 'f' form in [1e-6, 1e21), else 'e' form with Go's exponent clean-up."""
    if f == 0:
        return "-0" if math.copysign(1, f) < 0 else "0"
    r = repr(float(f))
    mant, _, exp = r.partition("e")
    digits = mant.replace("-", "").replace(".", "").lstrip("0") or "0"
    # decimal exponent of the first significant digit
    from decimal import Decimal
    d = Decimal(r)
    sign = "-" if d < 0 else ""
    t = d.as_tuple()
    ds = "".join(map(str, t.digits)).rstrip("0") or "0"
    e10 = len(t.digits) + t.exponent - 1  # exponent of the leading digit
    a = abs(f)
    if 1e-6 <= a < 1e21:
        if e10 >= len(ds) - 1:
            return sign + ds + "0" * (e10 - len(ds) + 1)
        if e10 >= 0:
            return sign + ds[:e10 + 1] + "." + ds[e10 + 1:]
        return sign + "0." + "0" * (-e10 - 1) + ds
    body = ds[0] + ("." + ds[1:] if len(ds) > 1 else "")
    es = f"{abs(e10):02d}"
    if es[0] == "0" and len(es) == 2:
        es = es[1]
    return f"{sign}{body}e{'-' if e10 < 0 else '+'}{es}"


def marshal_float_map(m: dict) -> str:
    return "{" + ",".join(f"{go_string(k)}:{go_float(v)}" for k, v in sorted(m.items(), key=lambda kv: _enc(kv[0]))) + "}"


def is_gpu_resource_key(key: str) -> bool:
    return key.endswith("gpu") or re.search(r"nvidia\.com/mig-\d+g\.\d+gb$", key) is not None


CUSTOM_ACCELERATORS = {"aws.amazon.com/neuroncore": "neuron_cores", "google.com/tpu": "TPU"}


def update_ray_start_params_resources(params: dict, group_resources: dict | None):
    if not group_resources:
        return
    custom = {}
    for name in sorted(group_resources, key=_enc):
        q = parse_quantity(group_resources[name])
        if q is None:
            continue
        nm = name.lower()
        if nm == "cpu":
            params["num-cpus"] = str(quantity_value(q))
        elif nm == "memory":
            params["memory"] = str(quantity_value(q))
        elif is_gpu_resource_key(nm):
            params["num-gpus"] = str(quantity_value(q))
        else:
            custom[name] = quantity_float(group_resources[name])
    if custom:
        params["resources"] = "'" + marshal_float_map(custom) + "'"


def update_ray_start_params_labels(params: dict, group_labels: dict | None):
    if not group_labels:
        return
    params["labels"] = ",".join(f"{k}={group_labels[k]}" for k in sorted(group_labels, key=_enc))


def set_missing_ray_start_params(params: dict, node_type: str, head_port: str, fqdn_ray_ip: str):
    if node_type == "worker" and "address" not in params:
        params["address"] = f"{fqdn_ray_ip}:{head_port}"
    if node_type == "head" and "dashboard-host" not in params:
        params["dashboard-host"] = "0.0.0.0"
    params.setdefault("metrics-export-port", "8080")
    params["block"] = "true"
    params.setdefault("dashboard-agent-listen-port", "52365")


def convert_param_map(params: dict) -> str:
    out = ""
    for k in sorted(params, key=_enc):
        v = params[k]
        if v.lower() in ("true", "false") and k not in ("log-color", "include-dashboard"):
            if v.lower() == "true":
                out += f" --{k} "
        else:
            out += f" --{k}={v} "
    return out


def generate_ray_start_command(node_type: str, params: dict, limits: dict | None, requests: dict | None) -> str:
    limits, requests = limits or {}, requests or {}

    def nonzero(m, k):
        q = parse_quantity(m[k]) if k in m else None
        return q if q else None   # None for absent, unparsable or zero
    if "num-cpus" not in params:
        q = nonzero(limits, "cpu") or nonzero(requests, "cpu")
        if q:
            params["num-cpus"] = str(quantity_value(q))
    if "memory" not in params:
        q = nonzero(limits, "memory")
        if q:
            params["memory"] = str(quantity_value(q))
    if limits:
        res, ok = {}, True
        if "resources" in params:
            try:
                res = json.loads(params["resources"].strip("'\"` "))
                ok = res is None or (isinstance(res, dict) and all(v is None or (isinstance(v, (int, float)) and not isinstance(v, bool)) for v in res.values()))
                res = {k: float(v) for k, v in (res or {}).items() if v is not None} if ok else {}
            except ValueError:
                ok = False
        if ok:
            have_custom = any(n in res for n in CUSTOM_ACCELERATORS.values())
            for key in sorted(limits, key=_enc):
                q = nonzero(limits, key)
                if "num-gpus" not in params and is_gpu_resource_key(key) and q:
                    params["num-gpus"] = str(quantity_value(q))
                if not have_custom and key in CUSTOM_ACCELERATORS and q:
                    if CUSTOM_ACCELERATORS[key] not in res:
                        res[CUSTOM_ACCELERATORS[key]] = quantity_float(limits[key])
                        params["resources"] = "'" + marshal_float_map(res) + "'"
                    have_custom = True
    if node_type == "head":
        return "ray start --head " + convert_param_map(params)
    if node_type == "worker":
        return "ray start " + convert_param_map(params)
    return ""


def ray_start_command(node_type: str, ray_start_params: dict | None = None, *, group_labels=None, group_resources=None, limits=None, requests=None,
                      command=None, args=None, head_port=None, fqdn_ray_ip="", autoscaling=False, overwrite_cmd=False, login_shell=False) -> dict:
    """The composed step: DefaultHead/WorkerPodTemplate's parameter handling + BuildPod's command assembly."""
    p = dict(ray_start_params or {})
    update_ray_start_params_resources(p, group_resources)
    update_ray_start_params_labels(p, group_labels)
    set_missing_ray_start_params(p, node_type, head_port or "6379", fqdn_ray_ip)
    if node_type == "head" and autoscaling:
        p["no-monitor"] = "true"
    line = generate_ray_start_command(node_type, p, limits, requests)
    cmd = "".join(f" {v} " for v in (command or [])) + "".join(f" {v} " for v in (args or []))
    generated = not overwrite_cmd and "ray start" not in cmd
    out_cmd, out_args = list(command or []), list(args or [])
    if generated:
        out_cmd = ["/bin/bash", "-c" + ("l" if login_shell else ""), "--"]
        gen = "ulimit -n 65536; " + line
        out_args = [f"{cmd} && {gen}" if cmd else gen]
    return {"rayStartParams": p, "rayStartCommand": line, "generated": generated, "command": out_cmd, "args": out_args}


def ray_container_env(node_type: str, *, existing=None, default_envs=None, fqdn_ray_ip="", head_port="6379", ray_start_cmd="", crd_type="RayCluster",
                      kuberay_version="v1.5.0", init_container=False) -> list[dict]:
    """setContainerEnvVars (common/pod.go:815-933) / setInitContainerEnvVars (:801-813): the EnvVars appended, in order."""
    names = list(existing or [])
    out = []

    def value(n, v):
        out.append({"name": n, **({"value": v} if v != "" else {})})
        names.append(n)

    def field(n, path):
        out.append({"name": n, "valueFrom": {"fieldRef": {"fieldPath": path}}})
        names.append(n)
    short = fqdn_ray_ip.split(".")[0]
    if init_container:
        value("FQ_RAY_IP", fqdn_ray_ip); value("RAY_IP", short)
        return out
    for n, v in (default_envs or {}).items():
        if n not in names:
            value(n, v)
    ip = "127.0.0.1"
    if node_type == "worker":
        ip = fqdn_ray_ip
        value("FQ_RAY_IP", ip); value("RAY_IP", short)
    field("RAY_CLUSTER_NAME", "metadata.labels['ray.io/cluster']")
    field("RAY_CLUSTER_NAMESPACE", "metadata.namespace")
    field("RAY_CLOUD_INSTANCE_ID", "metadata.name")
    field("RAY_NODE_TYPE_NAME", "metadata.labels['ray.io/group']")
    value("KUBERAY_GEN_RAY_START_CMD", ray_start_cmd)
    if "RAY_PORT" not in names:
        value("RAY_PORT", head_port)
    if crd_type == "RayService":
        for n, v in (("RAY_timeout_ms_task_wait_for_death_info", "0"), ("RAY_gcs_server_request_timeout_seconds", "5"), ("RAY_SERVE_KV_TIMEOUT_S", "5")):
            if n not in names:
                value(n, v)
    if "RAY_ADDRESS" not in names:
        value("RAY_ADDRESS", f"{ip}:{head_port}")
    if "RAY_USAGE_STATS_KUBERAY_IN_USE" not in names:
        value("RAY_USAGE_STATS_KUBERAY_IN_USE", "1")
    if node_type == "head":
        value("RAY_USAGE_STATS_EXTRA_TAGS", f"kuberay_version={kuberay_version};kuberay_crd={crd_type if crd_type in ('RayJob', 'RayService') else 'RayCluster'}")
    if "RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE" not in names:
        value("RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE", "1")
    return out


def ray_version_at_least(text: str, minimum=(2, 53, 0)) -> bool:
    """supportsUnifiedHealthCheck (common/pod.go:466-475) through version.ParseGeneric / AtLeast."""
    m = re.match(r"^\s*v?([0-9]+(?:\.[0-9]+)*)", text)
    if not m:
        return False
    comp = m.group(1).split(".")
    if len(comp) < 2 or (comp[0].startswith("0") and comp[0] != "0"):
        return False
    nums = [int(c) for c in comp]
    n = max(len(nums), len(minimum))
    return tuple(nums + [0] * (n - len(nums))) >= tuple(list(minimum) + [0] * (n - len(minimum)))


def ray_probes(node_type: str, ray_start_params=None, *, crd_type="RayCluster", ray_version="", has_liveness=False, has_readiness=False, serving_port=0) -> dict:
    """initLivenessAndReadinessProbe (common/pod.go:477-573)."""
    p = ray_start_params or {}

    def port(key, dflt):
        s = p.get(key)
        if s is None or not re.fullmatch(r"[+-]?[0-9]+", s):
            return dflt
        v = int(s)
        return v if -(1 << 31) <= v < (1 << 31) else dflt
    head = node_type == "head"
    agent, dash = port("dashboard-agent-listen-port", 52365), port("dashboard-port", 8265)
    http = ray_version_at_least(ray_version)
    wget = lambda t, prt, path: f"wget --tries 1 -T {t} -q -O- http://localhost:{prt}/{path} | grep success"   # noqa: E731
    commands = [wget(2, agent, "api/local_raylet_healthz")] + ([wget(10, dash, "api/gcs_healthz")] if head else [])

    def probe(use_http, cmds, delay, timeout, period, success, failure):
        handler = {"httpGet": {"path": "/api/healthz", "port": agent}} if use_http else {"exec": {"command": ["bash", "-c", " && ".join(cmds)]}}
        return {**handler, "initialDelaySeconds": delay, "timeoutSeconds": timeout, "periodSeconds": period, "successThreshold": success, "failureThreshold": failure}
    out = {}
    if not has_liveness:
        out["livenessProbe"] = probe(http, commands, 30, 5 if head else 2, 5, 1, 120)
    if not has_readiness:
        use_http, failure = http, 10
        if crd_type == "RayService" and not head:
            failure, use_http = 1, False
            commands = commands + [wget(10, serving_port if serving_port > 0 else 8000, "-/healthz")]
        out["readinessProbe"] = probe(use_http, commands, 10, 5 if head else 2, 5, 1, failure)
    return out


# ---------------------------------------------------------------------------------------------------------------- template surgery
# Object-level restatements (they mutate plain-dict corev1 objects the way the Go code mutates its structs); the native builders
# (kuberay_b200/csrc/kr_raytemplate.cpp) return fragments to append, and tests/test_raytemplate.py checks that applying those
# fragments gives the same objects.
def _env_exists(name: str, env: list) -> bool:
    """utils.EnvVarExists (utils/util.go:691-698)."""
    return any(e.get("name") == name for e in env)


def _env_var(name: str, value: str = "", value_from=None) -> dict:
    e = {"name": name}
    if value != "":
        e["value"] = value
    if value_from is not None:
        e["valueFrom"] = value_from
    return e


def configure_gcs_fault_tolerance(pod_template: dict, instance: dict, node_type: str, ft_enabled: bool) -> None:
    """common/pod.go:77-163.  instance: {uid, annotations, spec{gcsFaultToleranceOptions, headGroupSpec{rayStartParams}}}."""
    ann = pod_template.setdefault("metadata", {}).setdefault("annotations", {})
    if node_type == "head":
        ann["ray.io/ft-enabled"] = "true" if ft_enabled else "false"
    if not ft_enabled:
        return
    options = instance["spec"].get("gcsFaultToleranceOptions")
    container = pod_template["spec"]["containers"][0]
    env = container.setdefault("env", [])
    if not _env_exists("RAY_gcs_rpc_server_reconnect_timeout_s", env) and node_type == "worker":
        env.append(_env_var("RAY_gcs_rpc_server_reconnect_timeout_s", "600"))
    if node_type != "head":
        return
    storage_ns = instance.get("uid", "")
    if "ray.io/external-storage-namespace" in (instance.get("annotations") or {}):
        storage_ns = instance["annotations"]["ray.io/external-storage-namespace"]
    if options is not None and options.get("externalStorageNamespace", "") != "":
        storage_ns = options["externalStorageNamespace"]
    ann["ray.io/external-storage-namespace"] = storage_ns
    if not _env_exists("RAY_external_storage_namespace", env):
        env.append(_env_var("RAY_external_storage_namespace", storage_ns))
    params = instance["spec"]["headGroupSpec"].setdefault("rayStartParams", {})
    if options is not None:
        env.append(_env_var("RAY_REDIS_ADDRESS", options.get("redisAddress", "")))
        if options.get("redisUsername") is not None:
            params["redis-username"] = "$REDIS_USERNAME"
            env.append(_env_var("REDIS_USERNAME", options["redisUsername"].get("value", ""), options["redisUsername"].get("valueFrom")))
        if options.get("redisPassword") is not None:
            params["redis-password"] = "$REDIS_PASSWORD"
            env.append(_env_var("REDIS_PASSWORD", options["redisPassword"].get("value", ""), options["redisPassword"].get("valueFrom")))
    elif not _env_exists("REDIS_PASSWORD", env) and "redis-password" in params:
        env.append(_env_var("REDIS_PASSWORD", params["redis-password"]))


def is_k8s_auth_enabled(auth_options: dict | None) -> bool:
    """utils/util.go:763-765."""
    return auth_options is not None and auth_options.get("enableK8sTokenAuth") is True


def set_container_token_auth_env_vars(cluster_name: str, container: dict, auth_options: dict | None) -> None:
    """common/pod.go:296-335."""
    env = container.setdefault("env", [])
    if not _env_exists("RAY_AUTH_MODE", env):
        env.append(_env_var("RAY_AUTH_MODE", "token"))
    if is_k8s_auth_enabled(auth_options):
        if not _env_exists("RAY_ENABLE_K8S_TOKEN_AUTH", env):
            env.append(_env_var("RAY_ENABLE_K8S_TOKEN_AUTH", "true"))
        mounts = container.setdefault("volumeMounts", [])
        if not any(m.get("name") == "ray-token" for m in mounts):
            mounts.append({"name": "ray-token", "readOnly": True, "mountPath": "/var/run/secrets/ray.io/serviceaccount"})
        return
    secret = _dec(check_name(_enc(cluster_name)))
    if auth_options is not None and auth_options.get("secretName"):
        secret = auth_options["secretName"]
    if not _env_exists("RAY_AUTH_TOKEN", env):
        env.append(_env_var("RAY_AUTH_TOKEN", "", {"secretKeyRef": {"name": secret, "key": "auth_token"}}))


def add_ray_token_volume(pod_spec: dict) -> None:
    """common/pod.go:274-293."""
    vols = pod_spec.setdefault("volumes", [])
    if any(v.get("name") == "ray-token" for v in vols):
        return
    vols.append({"name": "ray-token", "projected": {"sources": [{"serviceAccountToken": {"path": "token"}}]}})


def configure_token_auth(cluster_name: str, pod_template: dict, auth_options: dict | None) -> None:
    """common/pod.go:254-271: the Ray container, the pod's token volume, and the wait-gcs-ready init container when present."""
    set_container_token_auth_env_vars(cluster_name, pod_template["spec"]["containers"][0], auth_options)
    if is_k8s_auth_enabled(auth_options):
        add_ray_token_volume(pod_template["spec"])
    for c in pod_template["spec"].get("initContainers") or []:
        if c.get("name") == "wait-gcs-ready":
            set_container_token_auth_env_vars(cluster_name, c, auth_options)


def container_command(login_shell: bool = False) -> list:
    """utils.GetContainerCommand([]string{}) (utils/util.go:884-892)."""
    return ["/bin/bash", "-c" + ("l" if login_shell else ""), "--"]


def build_autoscaler_container(image: str, login_shell: bool = False) -> dict:
    """common/pod.go:673-724."""
    small = {"cpu": "500m", "memory": "512Mi"}
    c = {"name": "autoscaler"}
    if image != "":
        c["image"] = image
    c.update({
        "command": container_command(login_shell),
        "args": ["ray kuberay-autoscaler --cluster-name $(RAY_CLUSTER_NAME) --cluster-namespace $(RAY_CLUSTER_NAMESPACE)"],
        "env": [_env_var("RAY_CLUSTER_NAME", "", {"fieldRef": {"fieldPath": "metadata.labels['ray.io/cluster']"}}),
                _env_var("RAY_CLUSTER_NAMESPACE", "", {"fieldRef": {"fieldPath": "metadata.namespace"}}),
                _env_var("RAY_HEAD_POD_NAME", "", {"fieldRef": {"fieldPath": "metadata.name"}}),
                _env_var("KUBERAY_CRD_VER", "v1")],
        "resources": {"limits": dict(small), "requests": dict(small)},
        "imagePullPolicy": "IfNotPresent",
    })
    return c


def merge_autoscaler_overrides(container: dict, options: dict | None) -> None:
    """common/pod.go:727-751."""
    if options is None:
        return
    if options.get("resources") is not None:
        container["resources"] = options["resources"]
    for key in ("image", "imagePullPolicy"):  # a pointer that is set overrides, even with ""; the JSON form then drops the empty field
        if options.get(key) is not None:
            container[key] = options[key]
            if options[key] == "":
                del container[key]
    if options.get("env"):
        container["env"] = container.get("env", []) + list(options["env"])
    if options.get("envFrom"):
        container["envFrom"] = container.get("envFrom", []) + list(options["envFrom"])
    if options.get("volumeMounts"):
        container["volumeMounts"] = container.get("volumeMounts", []) + list(options["volumeMounts"])
    if options.get("securityContext") is not None:
        container["securityContext"] = options["securityContext"]


def head_autoscaler_sidecar(instance: dict, pod_template: dict, login_shell: bool = False) -> None:
    """The autoscaling block of DefaultHeadPodTemplate (common/pod.go:194-220) on a head template; instance: {name, spec{authOptions,
    autoscalerOptions}}; the caller has established that autoscaling is on."""
    spec = pod_template["spec"]
    sa = spec.get("serviceAccountName") or instance["name"]  # utils.GetHeadGroupServiceAccountName
    spec["serviceAccountName"] = _dec(check_name(_enc(sa)))
    c = build_autoscaler_container(spec["containers"][0].get("image", ""), login_shell)
    auth = instance["spec"].get("authOptions")
    if auth is not None and auth.get("mode") == "token":
        set_container_token_auth_env_vars(instance["name"], c, auth)
    options = instance["spec"].get("autoscalerOptions")
    merge_autoscaler_overrides(c, options)
    spec["containers"].append(c)
    if options is not None and options.get("version") == "v2":  # utils.IsAutoscalingV2Enabled as far as the spec goes
        spec["containers"][0].setdefault("env", []).append(_env_var("RAY_enable_autoscaler_v2", "true"))
        spec["restartPolicy"] = "Never"


def wait_gcs_ready_script(fqdn_ray_ip: str, head_port: str) -> str:
    """The polling loop of common/pod.go:372-390, with the indentation its Go raw string carries."""
    addr = f"{fqdn_ray_ip}:{head_port}"
    lines = [
        (5, "SECONDS=0"),
        (5, "while true; do"),
        (6, "if (( SECONDS <= 120 )); then"),
        (7, f"if ray health-check --address {addr} > /dev/null 2>&1; then"),
        (8, 'echo "GCS is ready."'),
        (8, "break"),
        (7, "fi"),
        (7, 'echo "$SECONDS seconds elapsed: Waiting for GCS to be ready."'),
        (6, "else"),
        (7, f"if ray health-check --address {addr}; then"),
        (8, 'echo "GCS is ready. Any error messages above can be safely ignored."'),
        (8, "break"),
        (7, "fi"),
        (7, 'echo "$SECONDS seconds elapsed: Still waiting for GCS to be ready. For troubleshooting, refer to the FAQ at '
            'https://docs.ray.io/en/master/cluster/kubernetes/troubleshooting.html."'),
        (6, "fi"),
        (6, "sleep 5"),
        (5, "done"),
    ]
    return "\n" + "".join("\t" * n + text + "\n" for n, text in lines) + "\t" * 4


def wait_gcs_ready_container(ray_container: dict, fqdn_ray_ip: str, head_port: str, login_shell: bool = False) -> dict:
    """The init container DefaultWorkerPodTemplate appends (common/pod.go:363-415)."""
    import copy
    small = {"cpu": "200m", "memory": "256Mi"}
    c = {"name": "wait-gcs-ready"}
    if ray_container.get("image"):
        c["image"] = ray_container["image"]
    c["command"] = container_command(login_shell)
    c["args"] = [wait_gcs_ready_script(fqdn_ray_ip, head_port)]
    if ray_container.get("env"):
        c["env"] = copy.deepcopy(ray_container["env"])
    c["resources"] = {"limits": dict(small), "requests": dict(small)}
    if ray_container.get("volumeMounts"):
        c["volumeMounts"] = copy.deepcopy(ray_container["volumeMounts"])
    if ray_container.get("imagePullPolicy"):
        c["imagePullPolicy"] = ray_container["imagePullPolicy"]
    if ray_container.get("securityContext") is not None:
        c["securityContext"] = copy.deepcopy(ray_container["securityContext"])
    return c


# ---------------------------------------------------------------------------------------------------------------- the whole Pod
def add_empty_dir(container: dict, pod_spec: dict, volume_name: str, mount_path: str, memory: bool, canonical=lambda q: q) -> None:
    """addEmptyDir / makeEmptyDirVolume / findMemoryReqOrLimit (common/pod.go:1137-1217).  `canonical`: resource.Quantity's String()."""
    mounts = container.get("volumeMounts") or []
    if any(m.get("mountPath") == mount_path for m in mounts):
        return
    vols = pod_spec.get("volumes") or []
    if not any(v.get("name") == volume_name for v in vols):
        empty = {}
        if memory:
            empty["medium"] = "Memory"
            res = container.get("resources") or {}
            q = (res.get("limits") or {}).get("memory", (res.get("requests") or {}).get("memory"))
            if q is not None:
                empty["sizeLimit"] = canonical(q)
        pod_spec["volumes"] = vols + [{"name": volume_name, "emptyDir": empty}]
    container["volumeMounts"] = mounts + [{"name": volume_name, "mountPath": mount_path}]


def build_pod(cluster: dict, create: tuple, *, kuberay_version="v1.5.0", deterministic_head_name=False, multihost_indexing_gate=True, login_shell=False,
              init_container_injection=True, probes_injection=True, cluster_domain="cluster.local", default_container_envs=None, cluster_hash=None,
              canonical=lambda q: q) -> dict:
    """buildHeadPod / buildWorkerPod (raycluster_controller.go:1387-1433) = DefaultHead/WorkerPodTemplate (common/pod.go:166-239,
    352-464) + BuildPod (:577-669), followed statement by statement on plain dicts (maps shared and mutated as the Go code does)."""
    import copy
    g = create[0]
    head = g < 0
    node = "head" if head else "worker"
    instance = copy.deepcopy(cluster)
    spec = instance.setdefault("spec", {})
    head_spec = spec.setdefault("headGroupSpec", {})
    head_spec.setdefault("rayStartParams", {})
    grp = head_spec if head else spec["workerGroupSpecs"][g]
    params = grp.setdefault("rayStartParams", {})          # the map BuildPod receives; the template functions mutate it in place
    annots = instance.get("annotations") or {}
    head_port = head_spec["rayStartParams"].get("port", "6379")
    svc = ((head_spec.get("headService") or {}).get("metadata") or {}).get("name") or (head_spec.get("headService") or {}).get("name") or f"{instance['name']}-head-svc"
    fqdn = f"{svc}.{instance.get('namespace', 'default')}.svc.{cluster_domain}"
    autoscaling = spec.get("enableInTreeAutoscaling") is True
    auto_opts = spec.get("autoscalerOptions")
    auto_v2 = auto_opts is not None and auto_opts.get("version") == "v2"
    auth = spec.get("authOptions")
    auth_on = auth is not None and auth.get("mode") == "token"
    ft_opts = spec.get("gcsFaultToleranceOptions")
    ft = ("ray.io/ft-enabled" in annots and annots["ray.io/ft-enabled"].lower() == "true") or ft_opts is not None
    crd = (instance.get("labels") or {}).get("ray.io/originated-from-crd")
    crd = crd if crd in ("RayJob", "RayService") else "RayCluster"

    template = {"metadata": {}, "spec": copy.deepcopy((grp.get("template") or {}).get("spec") or {})}
    pspec = template["spec"]
    ray = pspec["containers"][0]
    if not head and init_container_injection:                                                 # :359-415
        pspec["initContainers"] = (pspec.get("initContainers") or []) + [wait_gcs_ready_container(ray, fqdn, head_port, login_shell)]
    update_ray_start_params_resources(params, grp.get("resources"))                           # :181 / :421
    update_ray_start_params_labels(params, grp.get("labels"))                                 # :184 / :424
    set_missing_ray_start_params(params, node, head_port, "" if head else fqdn)               # :190 / :440
    if head and autoscaling:                                                                  # :194-220
        params["no-monitor"] = "true"
        head_autoscaler_sidecar(instance, template, login_shell)
    configure_gcs_fault_tolerance(template, instance, node, ft)                               # :222 / :443
    if not any(p.get("name") == "metrics" for p in ray.get("ports") or []):                   # :224-232 / :445-453
        ray["ports"] = (ray.get("ports") or []) + [{"name": "metrics", "containerPort": 8080}]
    if not head and autoscaling and auto_v2:                                                  # :455-457
        pspec["restartPolicy"] = "Never"
    if auth_on:                                                                               # :234-236 / :459-461
        configure_token_auth(instance["name"], template, auth)

    # BuildPod (:577-669)
    if "plasma-directory" not in params:
        add_empty_dir(ray, pspec, "shared-mem", "/dev/shm", True, canonical)
    if head and autoscaling:
        side = next(c for c in pspec["containers"] if c.get("name") == "autoscaler")          # getAutoscalerContainerIndex (it panics when absent)
        add_empty_dir(ray, pspec, "ray-logs", "/tmp/ray", False)
        add_empty_dir(side, pspec, "ray-logs", "/tmp/ray", False)
    meta = pod_meta(cluster, create, kuberay_version=kuberay_version, deterministic_head_name=deterministic_head_name,
                    multihost_indexing_gate=multihost_indexing_gate, cluster_hash=cluster_hash)
    cmd = "".join(f" {v} " for v in ray.get("command") or []) + "".join(f" {v} " for v in ray.get("args") or [])
    res = ray.get("resources") or {}
    line = generate_ray_start_command(node, params, res.get("limits"), res.get("requests"))
    overwrite = meta["annotations"].get("ray.io/overwrite-container-cmd", "").lower() == "true"
    if not overwrite and "ray start" not in cmd:
        generated = "ulimit -n 65536; " + line
        ray["command"] = container_command(login_shell)
        ray["args"] = [f"{cmd} && {generated}" if cmd else generated]
    for c in pspec.get("initContainers") or []:
        c["env"] = (c.get("env") or []) + ray_container_env(node, fqdn_ray_ip=fqdn, init_container=True)
    ray["env"] = (ray.get("env") or []) + ray_container_env(node, existing=[e.get("name") for e in ray.get("env") or []], default_envs=default_container_envs,
                                                             fqdn_ray_ip=fqdn, head_port=head_port, ray_start_cmd=line, crd_type=crd, kuberay_version=kuberay_version)
    if probes_injection:
        serve = next((p.get("containerPort", 8000) for p in ray.get("ports") or [] if p.get("name") == "serve"), 8000)
        ray.update(ray_probes(node, params, crd_type=crd, ray_version=spec.get("rayVersion", ""), has_liveness=ray.get("livenessProbe") is not None,
                              has_readiness=ray.get("readinessProbe") is not None, serving_port=serve))
    # ObjectMeta: podTemplateSpec.ObjectMeta (:598) — everything else the template's metadata carries rides along; DefaultWorkerPodTemplate
    # clears the worker's name (:418), DefaultHeadPodTemplate sets name OR generateName and leaves the other as the template had it (:171-175)
    tmeta = copy.deepcopy((grp.get("template") or {}).get("metadata") or {})
    if not head:
        tmeta.pop("name", None)
    tmeta = {k: v for k, v in tmeta.items() if v is not None}
    return {"apiVersion": "v1", "kind": "Pod", "metadata": {**tmeta, **meta}, "spec": pspec}
