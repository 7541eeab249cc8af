"""Pod metadata builder over the C ABI (kr_pod_* in include/kr_engine.h; kuberay_b200/csrc/kr_podmeta.cpp) — SURVEY §8 f3.

Host-side mirror of the reference's names for this step:
  pod_name / check_name / check_label      utils.PodName / CheckName / CheckLabel (controllers/ray/utils/util.go:198-265)
  build_pod_meta(cluster, creates, env)    the ObjectMeta of buildHeadPod / buildWorkerPod (raycluster_controller.go:1387-1433) for the
                                           engine's create tuples: one dict per Pod to create
  expand_creates(...)                      engine results -> (group, replicaIndex, hostIndex, replicaGrpName) tuples in Create order

Everything is computed by the native library; this file only marshals dicts into the C structs.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass

import numpy as np

from . import abi
from .engine import EngineError, lib

OVERWRITE_CMD_ANNOT = "ray.io/overwrite-container-cmd"
FT_ENABLED_ANNOT = "ray.io/ft-enabled"
STORAGE_NS_ANNOT = "ray.io/external-storage-namespace"
ORIGINATED_FROM_CRD_LABEL = "ray.io/originated-from-crd"
CRD_TYPES = {"RayJob": abi.CRD_RAYJOB, "RayService": abi.CRD_RAYSERVICE}  # utils.GetCRDType (util.go:59-64): anything else is RayCluster


@dataclass
class PodMetaEnv:
    """Process-level inputs of the builder (env vars / feature gates / build constants of the operator)."""
    kuberay_version: str = "v1.5.0"
    deterministic_head_name: bool = False    # ENABLE_DETERMINISTIC_HEAD_POD_NAME
    multihost_indexing_gate: bool = True     # features.RayMultiHostIndexing


_bound = False


def _lib():
    global _bound
    L = lib()
    if not _bound:
        P = C.POINTER
        for f in (L.kr_check_name, L.kr_check_label):
            f.argtypes = [abi.kr_str, C.c_char_p, C.c_uint64]
            f.restype = C.c_int64
        L.kr_pod_name.argtypes = [abi.kr_str, C.c_uint8, C.c_uint8, C.c_char_p, C.c_uint64]
        L.kr_pod_name.restype = C.c_int64
        L.kr_pod_meta_build.argtypes = [P(abi.kr_podmeta_cluster), P(abi.kr_podmeta_group), P(abi.kr_podmeta_group), C.c_uint32,
                                        P(abi.kr_podmeta_create), C.c_uint32, C.c_void_p, C.c_uint64, P(C.c_uint64), P(C.c_uint64)]
        L.kr_pod_creates_expand.argtypes = [C.c_void_p, P(abi.kr_podmeta_group), C.c_uint32, C.c_void_p, C.c_uint8, C.c_uint8, C.c_uint64,
                                            P(abi.kr_podmeta_create), C.c_uint32, C.c_char_p, C.c_uint64, P(C.c_uint32)]
        L.kr_pod_meta_last_error.restype = C.c_char_p
        _bound = True
    return L


class _Keep:
    """Owns the byte strings / arrays the C structs point into for the duration of one call."""

    def __init__(self):
        self.refs = []

    def s(self, v) -> abi.kr_str:
        if v is None:
            return abi.kr_str(None, 0)
        b = v if isinstance(v, bytes) else str(v).encode("utf-8", "surrogateescape")
        self.refs.append(b)
        return abi.kr_str(b, len(b))

    def kvs(self, m: dict | None):
        items = list((m or {}).items())
        arr = (abi.kr_kv * max(len(items), 1))()
        for i, (k, v) in enumerate(items):
            arr[i].key, arr[i].value = self.s(k), self.s(v)
        self.refs.append(arr)
        return arr, len(items)


def _err(L, rc):
    return EngineError(int(rc), (L.kr_pod_meta_last_error() or b"").decode())


def _text(fn, s: str, *extra) -> str:
    L = _lib()
    b = s.encode("utf-8", "surrogateescape")
    buf = C.create_string_buffer(len(b) + 16)
    n = fn(abi.kr_str(b, len(b)), *extra, buf, len(buf))
    if n < 0:
        raise _err(L, n)
    return buf.raw[:n].decode("utf-8", "surrogateescape")


def check_name(s: str) -> str:
    return _text(_lib().kr_check_name, s)


def check_label(s: str) -> str:
    return _text(_lib().kr_check_label, s)


def pod_name(prefix: str, node_type: str, is_generate_name: bool) -> str:
    nt = abi.NT_HEAD if node_type == "head" else abi.NT_WORKER
    return _text(_lib().kr_pod_name, prefix, nt, 1 if is_generate_name else 0)


def _group_struct(keep: _Keep, grp: dict, head: bool) -> abi.kr_podmeta_group:
    g = abi.kr_podmeta_group()
    tmeta = ((grp.get("template") or {}).get("metadata") or {})
    g.group_name = keep.s("headgroup" if head else grp.get("groupName", ""))
    g.num_of_hosts = 1 if head else int(grp.get("numOfHosts", 1))
    tl, g.n_template_labels = keep.kvs(tmeta.get("labels"))
    gl, g.n_group_labels = keep.kvs(grp.get("labels"))
    ta, g.n_template_annotations = keep.kvs(tmeta.get("annotations"))
    g.template_labels, g.group_labels, g.template_annotations = tl, gl, ta
    return g


def _cluster_struct(keep: _Keep, cluster: dict, env: PodMetaEnv, cluster_hash: str | None) -> abi.kr_podmeta_cluster:
    spec = cluster.get("spec") or {}
    annots = cluster.get("annotations") or {}
    c = abi.kr_podmeta_cluster()
    c.name, c.ns, c.uid = keep.s(cluster["name"]), keep.s(cluster.get("namespace", "default")), keep.s(cluster.get("uid", ""))
    c.cluster_hash = keep.s(cluster_hash or None)
    c.kuberay_version = keep.s(env.kuberay_version)
    c.storage_ns_annotation = keep.s(annots.get(STORAGE_NS_ANNOT))
    ft = spec.get("gcsFaultToleranceOptions")
    c.storage_ns_option = keep.s((ft or {}).get("externalStorageNamespace") or None)
    c.overwrite_container_cmd = 1 if str(annots.get(OVERWRITE_CMD_ANNOT, "")).lower() == "true" and OVERWRITE_CMD_ANNOT in annots else 0
    c.ft_enabled = 1 if (FT_ENABLED_ANNOT in annots and str(annots[FT_ENABLED_ANNOT]).lower() == "true") or ft is not None else 0
    c.crd_type = CRD_TYPES.get((cluster.get("labels") or {}).get(ORIGINATED_FROM_CRD_LABEL), abi.CRD_RAYCLUSTER)
    c.deterministic_head_name = 1 if env.deterministic_head_name else 0
    c.gate_multihost_indexing = 1 if env.multihost_indexing_gate else 0
    return c


def build_pod_meta(cluster: dict, creates: list[tuple[int, int, int, str]], env: PodMetaEnv | None = None, cluster_hash: str | None = None,
                   raw: bool = False) -> list:
    """ObjectMeta of every Pod to create.  creates: (group index or -1 for the head, replicaIndex, hostIndex, replicaGrpName).

    raw=True returns the JSON bytes per create exactly as the library wrote them (Go map/field order)."""
    L = _lib()
    env = env or PodMetaEnv()
    keep = _Keep()
    spec = cluster.get("spec") or {}
    c = _cluster_struct(keep, cluster, env, cluster_hash)
    head = _group_struct(keep, spec.get("headGroupSpec") or {}, True)
    wgs = spec.get("workerGroupSpecs") or []
    groups = (abi.kr_podmeta_group * max(len(wgs), 1))(*[_group_struct(keep, g, False) for g in wgs])
    tuples = (abi.kr_podmeta_create * max(len(creates), 1))()
    for i, (g, ri, hi, rn) in enumerate(creates):
        tuples[i].group, tuples[i].replica_index, tuples[i].host_index = int(g), int(ri), int(hi)
        tuples[i].replica_name = keep.s(rn or "")
    off = (C.c_uint64 * (len(creates) + 1))()
    need = C.c_uint64()
    rc = L.kr_pod_meta_build(C.byref(c), C.byref(head), groups, len(wgs), tuples, len(creates), None, 0, off, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise _err(L, rc)
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = L.kr_pod_meta_build(C.byref(c), C.byref(head), groups, len(wgs), tuples, len(creates), buf, need.value, off, C.byref(need))
    if rc:
        raise _err(L, rc)
    b = bytes(buf)
    parts = [b[off[i]:off[i + 1]] for i in range(len(creates))]
    return parts if raw else [json.loads(p) for p in parts]


def expand_creates(group_results: np.ndarray, create_idx: np.ndarray, groups: list[dict], head_create: bool, env: PodMetaEnv | None = None,
                   seed: int = 0) -> list[tuple[int, int, int, str]]:
    """The cluster's create tuples in the order the reference issues the Create calls (kr_pod_creates_expand).
    group_results: the cluster's rows of Results.groups; create_idx: Results.create_idx (the whole arena)."""
    L = _lib()
    env = env or PodMetaEnv()
    keep = _Keep()
    gr = np.ascontiguousarray(group_results)
    ci = np.ascontiguousarray(create_idx, dtype=np.int32)
    gs = (abi.kr_podmeta_group * max(len(groups), 1))(*[_group_struct(keep, g, False) for g in groups])
    n = C.c_uint32()
    args = (gr.ctypes.data if len(gr) else None, gs, len(groups), ci.ctypes.data if len(ci) else None, 1 if head_create else 0,
            1 if env.multihost_indexing_gate else 0, seed & (2 ** 64 - 1))
    rc = L.kr_pod_creates_expand(*args, None, 0, None, 0, C.byref(n))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise _err(L, rc)
    out = (abi.kr_podmeta_create * max(n.value, 1))()
    name_cap = sum((len(str(g.get("groupName", "")).encode()) + 6) * int(r["n_create"]) for g, r in zip(groups, gr)) + 8
    names = C.create_string_buffer(name_cap)
    rc = L.kr_pod_creates_expand(*args, out, n.value, names, name_cap, C.byref(n))
    if rc:
        raise _err(L, rc)
    res = []
    for i in range(n.value):
        t = out[i]
        rn = t.replica_name.p[:t.replica_name.n].decode() if t.replica_name.n else ""
        res.append((t.group, t.replica_index, t.host_index, rn))
    return res


# ------------------------------------------------------------------------------------------------ `ray start` command (kr_raystart.cpp)
def _bind_raystart():
    L = _lib()
    if not getattr(L, "_kr_rs_bound", False):
        P = C.POINTER
        L.kr_ray_start_command.argtypes = [P(abi.kr_raystart_in), C.c_void_p, C.c_uint64, P(C.c_uint64)]
        L.kr_quantity_value.argtypes = [abi.kr_str, P(C.c_int64), P(C.c_double), P(C.c_uint8)]
        L.kr_quantity_value.restype = C.c_int64
        L.kr_ray_start_last_error.restype = C.c_char_p
        L._kr_rs_bound = True
    return L


def quantity_value(text: str) -> tuple[int, float, bool]:
    """resource.Quantity as the builder reads it: (Value() rounded up, AsApproximateFloat64(), IsZero())."""
    L = _bind_raystart()
    b = text.encode()
    v, f, z = C.c_int64(), C.c_double(), C.c_uint8()
    rc = L.kr_quantity_value(abi.kr_str(b, len(b)), C.byref(v), C.byref(f), C.byref(z))
    if rc:
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    return v.value, f.value, bool(z.value)


def ray_start_command(node_type: str, ray_start_params: dict | None = None, *, group_labels: dict | None = None, group_resources: dict | None = None,
                      limits: dict | None = None, requests: dict | None = None, command: list[str] | None = None, args: list[str] | None = None,
                      head_port: str | None = None, fqdn_ray_ip: str = "", autoscaling: bool = False, overwrite_cmd: bool = False, login_shell: bool = False,
                      steps: int = 0) -> dict:
    """kr_ray_start_command: the group's final rayStartParams, the `ray start` line and the Ray container's command / args
    (DefaultHeadPodTemplate / DefaultWorkerPodTemplate + BuildPod, common/pod.go:179-200, 420-440, 617-650, 935-1135)."""
    L = _bind_raystart()
    keep = _Keep()
    a = abi.kr_raystart_in()
    a.node_type = abi.NT_HEAD if node_type == "head" else abi.NT_WORKER if node_type == "worker" else abi.NT_NONE
    a.autoscaling_enabled, a.overwrite_container_cmd, a.login_shell, a.steps = int(autoscaling), int(overwrite_cmd), int(login_shell), steps
    a.head_port, a.fqdn_ray_ip = keep.s(head_port), keep.s(fqdn_ray_ip)
    for field, m in (("ray_start_params", ray_start_params), ("group_labels", group_labels), ("group_resources", group_resources),
                     ("container_limits", limits), ("container_requests", requests)):
        arr, n = keep.kvs(m)
        setattr(a, field, arr)
        setattr(a, "n_" + field, n)
    for field, lst in (("command", command), ("args", args)):
        lst = lst or []
        arr = (abi.kr_str * max(len(lst), 1))(*[keep.s(x) for x in lst])
        keep.refs.append(arr)
        setattr(a, field, arr)
        setattr(a, "n_" + field, len(lst))
    need = C.c_uint64()
    rc = L.kr_ray_start_command(C.byref(a), None, 0, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = L.kr_ray_start_command(C.byref(a), buf, need.value, C.byref(need))
    if rc:
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    return json.loads(bytes(buf)[:need.value])


def ray_container_env(node_type: str, *, existing: list[str] | None = None, default_envs: dict | None = None, fqdn_ray_ip: str = "", head_port: str = "6379",
                      ray_start_cmd: str = "", crd_type: str = "RayCluster", kuberay_version: str = "v1.5.0", init_container: bool = False) -> list[dict]:
    """kr_ray_container_env: the EnvVars BuildPod appends (setContainerEnvVars / setInitContainerEnvVars, common/pod.go:801-933)."""
    L = _bind_raystart()
    if not getattr(L, "_kr_env_bound", False):
        L.kr_ray_container_env.argtypes = [C.POINTER(abi.kr_rayenv_in), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L._kr_env_bound = True
    keep = _Keep()
    a = abi.kr_rayenv_in()
    a.node_type = abi.NT_HEAD if node_type == "head" else abi.NT_WORKER if node_type == "worker" else abi.NT_NONE
    a.crd_type, a.init_container = CRD_TYPES.get(crd_type, abi.CRD_RAYCLUSTER), int(init_container)
    a.fqdn_ray_ip, a.head_port, a.ray_start_cmd, a.kuberay_version = keep.s(fqdn_ray_ip), keep.s(head_port), keep.s(ray_start_cmd), keep.s(kuberay_version)
    ex = existing or []
    arr = (abi.kr_str * max(len(ex), 1))(*[keep.s(x) for x in ex])
    keep.refs.append(arr)
    a.existing, a.n_existing = arr, len(ex)
    a.default_envs, a.n_default_envs = keep.kvs(default_envs)
    need = C.c_uint64()
    rc = L.kr_ray_container_env(C.byref(a), None, 0, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = L.kr_ray_container_env(C.byref(a), buf, need.value, C.byref(need))
    if rc:
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    return json.loads(bytes(buf)[:need.value])


def ray_probes(node_type: str, ray_start_params: dict | None = None, *, crd_type: str = "RayCluster", ray_version: str = "", has_liveness: bool = False,
               has_readiness: bool = False, serving_port: int = 0) -> dict:
    """kr_ray_probes: the probes BuildPod injects (initLivenessAndReadinessProbe, common/pod.go:477-573)."""
    L = _bind_raystart()
    if not getattr(L, "_kr_probe_bound", False):
        L.kr_ray_probes.argtypes = [C.POINTER(abi.kr_rayprobe_in), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L._kr_probe_bound = True
    keep = _Keep()
    a = abi.kr_rayprobe_in()
    a.node_type = abi.NT_HEAD if node_type == "head" else abi.NT_WORKER if node_type == "worker" else abi.NT_NONE
    a.crd_type, a.has_liveness_probe, a.has_readiness_probe, a.serving_port = CRD_TYPES.get(crd_type, abi.CRD_RAYCLUSTER), int(has_liveness), int(has_readiness), serving_port
    a.ray_version = keep.s(ray_version)
    a.ray_start_params, a.n_ray_start_params = keep.kvs(ray_start_params)
    need = C.c_uint64()
    rc = L.kr_ray_probes(C.byref(a), None, 0, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = L.kr_ray_probes(C.byref(a), buf, need.value, C.byref(need))
    if rc:
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    return json.loads(bytes(buf)[:need.value])


def ray_volumes(node_type: str, *, autoscaling: bool = False, plasma_directory_set: bool = False, memory_limit: str | None = None, memory_request: str | None = None,
                volume_names: list[str] | None = None, ray_mount_paths: list[str] | None = None, autoscaler_mount_paths: list[str] | None = None) -> dict:
    """kr_ray_volumes: the emptyDir volumes / mounts BuildPod adds (common/pod.go:600-615, 1137-1217)."""
    L = _bind_raystart()
    if not getattr(L, "_kr_vol_bound", False):
        L.kr_ray_volumes.argtypes = [C.POINTER(abi.kr_rayvol_in), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L._kr_vol_bound = True
    keep = _Keep()
    a = abi.kr_rayvol_in()
    a.node_type = abi.NT_HEAD if node_type == "head" else abi.NT_WORKER
    a.autoscaling_enabled, a.plasma_directory_set = int(autoscaling), int(plasma_directory_set)
    a.memory_limit, a.memory_request = keep.s(memory_limit), keep.s(memory_request)
    for field, lst in (("volume_names", volume_names), ("ray_mount_paths", ray_mount_paths), ("autoscaler_mount_paths", autoscaler_mount_paths)):
        lst = lst or []
        arr = (abi.kr_str * max(len(lst), 1))(*[keep.s(x) for x in lst])
        keep.refs.append(arr)
        setattr(a, field, arr)
        setattr(a, "n_" + field, len(lst))
    need = C.c_uint64()
    rc = L.kr_ray_volumes(C.byref(a), None, 0, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = L.kr_ray_volumes(C.byref(a), buf, need.value, C.byref(need))
    if rc:
        raise EngineError(int(rc), (L.kr_ray_start_last_error() or b"").decode())
    return json.loads(bytes(buf)[:need.value])


# ------------------------------------------------------------------------------------------------ template surgery (kr_raytemplate.cpp)
def _template_call(name: str, arg) -> dict:
    L = _lib()
    fn = getattr(L, name)
    if not getattr(L, "_kr_tpl_bound_" + name, False):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.kr_ray_template_last_error.restype = C.c_char_p
        setattr(L, "_kr_tpl_bound_" + name, True)
    need = C.c_uint64()
    rc = fn(C.byref(arg), None, 0, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise EngineError(int(rc), (L.kr_ray_template_last_error() or b"").decode())
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = fn(C.byref(arg), buf, need.value, C.byref(need))
    if rc:
        raise EngineError(int(rc), (L.kr_ray_template_last_error() or b"").decode())
    return json.loads(bytes(buf)[:need.value])


def _raw(keep: _Keep, obj) -> abi.kr_str:
    """A corev1 fragment as the Go side would hand it over: json.Marshal text (None: absent)."""
    if obj is None:
        return abi.kr_str(None, 0)
    return keep.s(obj if isinstance(obj, (str, bytes)) else json.dumps(obj, separators=(",", ":")))


def _names(keep: _Keep, a, field: str, lst):
    lst = lst or []
    arr = (abi.kr_str * max(len(lst), 1))(*[keep.s(x) for x in lst])
    keep.refs.append(arr)
    setattr(a, field, arr)
    setattr(a, "n_" + field, len(lst))


def ray_ft_env(node_type: str, *, ft_enabled: bool, cluster_uid: str = "", storage_ns_annotation: str | None = None, options: dict | None = None,
               head_redis_password_param: str | None = None, existing: list[str] | None = None) -> dict:
    """kr_ray_ft_env: configureGCSFaultTolerance's env / rayStartParams additions (common/pod.go:77-163).  `options` mirrors
    spec.gcsFaultToleranceOptions: redisAddress, externalStorageNamespace, redisUsername / redisPassword {value, valueFrom}."""
    keep = _Keep()
    a = abi.kr_rayft_in()
    a.node_type = abi.NT_HEAD if node_type == "head" else abi.NT_WORKER
    a.ft_enabled, a.has_options = int(ft_enabled), int(options is not None)
    a.cluster_uid, a.storage_ns_annotation = keep.s(cluster_uid), keep.s(storage_ns_annotation)
    o = options or {}
    a.storage_ns_option, a.redis_address = keep.s(o.get("externalStorageNamespace")), keep.s(o.get("redisAddress", ""))
    for key, flag, val, frm in (("redisUsername", "has_redis_username", "redis_username_value", "redis_username_value_from"),
                                ("redisPassword", "has_redis_password", "redis_password_value", "redis_password_value_from")):
        cred = o.get(key)
        setattr(a, flag, int(cred is not None))
        if cred is not None:
            setattr(a, val, keep.s(cred.get("value", "")))
            setattr(a, frm, _raw(keep, cred.get("valueFrom")))
    a.head_redis_password_param = keep.s(head_redis_password_param)
    _names(keep, a, "existing", existing)
    return _template_call("kr_ray_ft_env", a)


def ray_auth(cluster_name: str, *, k8s_token_auth: bool = False, secret_name: str | None = None, existing_env: list[str] | None = None,
             existing_mount_names: list[str] | None = None, existing_volume_names: list[str] | None = None) -> dict:
    """kr_ray_auth: the token-auth env / mount / volume for one container (common/pod.go:254-335)."""
    keep = _Keep()
    a = abi.kr_rayauth_in()
    a.k8s_token_auth = int(k8s_token_auth)
    a.cluster_name, a.secret_name = keep.s(cluster_name), keep.s(secret_name)
    _names(keep, a, "existing_env", existing_env)
    _names(keep, a, "existing_mount_names", existing_mount_names)
    _names(keep, a, "existing_volume_names", existing_volume_names)
    return _template_call("kr_ray_auth", a)


def ray_autoscaler_container(cluster_name: str, ray_image: str, *, options: dict | None = None, autoscaler_v2: bool = False, auth_enabled: bool = False,
                             k8s_token_auth: bool = False, secret_name: str | None = None, head_service_account: str | None = None, login_shell: bool = False) -> dict:
    """kr_ray_autoscaler_container: the head's autoscaler sidecar (common/pod.go:194-220, 673-751).  `options` mirrors spec.autoscalerOptions:
    image, imagePullPolicy, resources, env, envFrom, volumeMounts, securityContext."""
    keep = _Keep()
    a = abi.kr_rayautoscaler_in()
    a.login_shell, a.autoscaler_v2, a.auth_enabled, a.k8s_token_auth, a.has_options = int(login_shell), int(autoscaler_v2), int(auth_enabled), int(k8s_token_auth), int(options is not None)
    a.cluster_name, a.secret_name, a.head_service_account, a.ray_image = keep.s(cluster_name), keep.s(secret_name), keep.s(head_service_account), keep.s(ray_image)
    o = options or {}
    a.image, a.image_pull_policy = keep.s(o.get("image")), keep.s(o.get("imagePullPolicy"))
    a.resources_json, a.env_json, a.env_from_json = _raw(keep, o.get("resources")), _raw(keep, o.get("env")), _raw(keep, o.get("envFrom"))
    a.volume_mounts_json, a.security_context_json = _raw(keep, o.get("volumeMounts")), _raw(keep, o.get("securityContext"))
    return _template_call("kr_ray_autoscaler_container", a)


def ray_init_container(image: str, fqdn_ray_ip: str, head_port: str = "6379", *, image_pull_policy: str | None = None, env: list | None = None,
                       volume_mounts: list | None = None, security_context: dict | None = None, login_shell: bool = False) -> dict:
    """kr_ray_init_container: the worker's wait-gcs-ready init container (common/pod.go:359-415)."""
    keep = _Keep()
    a = abi.kr_rayinit_in()
    a.login_shell = int(login_shell)
    a.image, a.image_pull_policy, a.fqdn_ray_ip, a.head_port = keep.s(image), keep.s(image_pull_policy), keep.s(fqdn_ray_ip), keep.s(head_port)
    a.env_json, a.volume_mounts_json, a.security_context_json = _raw(keep, env), _raw(keep, volume_mounts), _raw(keep, security_context)
    return _template_call("kr_ray_init_container", a)
