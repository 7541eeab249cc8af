"""buildHeadPod / buildWorkerPod (raycluster_controller.go:1387-1433) assembled from the native builders — the host-side mirror of
the glue the Go shim keeps once DefaultHeadPodTemplate / DefaultWorkerPodTemplate / BuildPod (common/pod.go:166-239, 352-464,
577-669) are answered by libkrengine.so: deep-copy the group's template, append the fragments each C call returns, hand the Pod to
client.Create.  Everything that decides a byte of the manifest runs behind the C ABI (kr_podmeta.cpp, kr_raystart.cpp,
kr_raytemplate.cpp); this file only moves their answers into place, in the reference's order:

  worker: wait-gcs-ready init container (copies the Ray container BEFORE anything is added to it)      common/pod.go:359-415
  head:   autoscaler sidecar, service account, autoscaler-v2 env + restartPolicy                       :194-220
  GCS fault tolerance env (+ the head's redis-* rayStartParams)                                          :222 / :443
  the default metrics port                                                                               :224-232 / :445-453
  worker: restartPolicy Never under autoscaler v2                                                        :455-457
  token auth: Ray container, token volume, wait-gcs-ready                                                :234-236 / :459-461
  operator-configured sidecars                                                 raycluster_controller.go:1397-1399, 1424-1426
  BuildPod: emptyDir volumes, `ray start` command, init-container env, Ray container env, probes         :577-669
  ObjectMeta + the controller ownerReference (kr_pod_meta_build)                                         :1404, 1431

Object model: the plain-dict RayCluster of tests/golden (name, namespace, uid, labels, annotations, spec{...} with corev1 templates in
their JSON form)."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

from . import podmeta

FT_ENABLED_ANNOT = "ray.io/ft-enabled"
STORAGE_NS_ANNOT = "ray.io/external-storage-namespace"
OVERWRITE_CMD_ANNOT = "ray.io/overwrite-container-cmd"
ORIGINATED_FROM_CRD_LABEL = "ray.io/originated-from-crd"


@dataclass
class BuilderEnv:
    """What the operator process contributes: its configuration and the environment switches the builders read."""
    kuberay_version: str = "v1.5.0"
    deterministic_head_name: bool = False          # utils.IsDeterministicHeadPodNameEnabled
    multihost_indexing_gate: bool = True           # features.RayMultiHostIndexing
    login_shell: bool = False                      # ENABLE_LOGIN_SHELL
    init_container_injection: bool = True          # ENABLE_INIT_CONTAINER_INJECTION
    probes_injection: bool = True                  # ENABLE_PROBES_INJECTION
    cluster_domain: str = "cluster.local"          # CLUSTER_DOMAIN
    default_container_envs: dict = field(default_factory=dict)     # configuration DefaultContainerEnvs (name -> value)
    head_sidecar_containers: list = field(default_factory=list)    # configuration HeadSidecarContainers
    worker_sidecar_containers: list = field(default_factory=list)  # configuration WorkerSidecarContainers


def head_service_fqdn(cluster: dict, domain: str = "cluster.local") -> str:
    """utils.GenerateFQDNServiceName (utils/util.go:313-337): <head service>.<namespace>.svc.<cluster domain>."""
    head = (cluster.get("spec") or {}).get("headGroupSpec") or {}
    svc = ((head.get("headService") or {}).get("metadata") or {}).get("name") or (head.get("headService") or {}).get("name") or f"{cluster['name']}-head-svc"
    return f"{svc}.{cluster.get('namespace', 'default')}.svc.{domain}"


def _names(container: dict) -> list[str]:
    return [e.get("name", "") for e in container.get("env") or []]


def _extend(obj: dict, key: str, items: list) -> None:
    if items:
        obj.setdefault(key, [])
        obj[key] = (obj[key] or []) + items


def build_pod(cluster: dict, create: tuple[int, int, int, str], env: BuilderEnv | None = None, cluster_hash: str | None = None) -> dict:
    """The corev1.Pod (JSON form) the reference would Create for one engine create tuple: (group index or -1 for the head,
    replicaIndex, hostIndex, replicaGrpName)."""
    env = env or BuilderEnv()
    g = int(create[0])
    head = g < 0
    node = "head" if head else "worker"
    spec = cluster.get("spec") or {}
    head_spec = spec.get("headGroupSpec") or {}
    grp = head_spec if head else spec["workerGroupSpecs"][g]
    annots = cluster.get("annotations") or {}
    crd = (cluster.get("labels") or {}).get(ORIGINATED_FROM_CRD_LABEL)
    crd = crd if crd in ("RayJob", "RayService") else "RayCluster"
    head_params = head_spec.get("rayStartParams") or {}
    head_port = head_params.get("port", "6379")                                   # common.GetHeadPort
    fqdn = head_service_fqdn(cluster, env.cluster_domain)
    autoscaling = spec.get("enableInTreeAutoscaling") is True
    auto_opts = spec.get("autoscalerOptions")
    auto_v2 = bool(auto_opts is not None and auto_opts.get("version") == "v2")
    auth = spec.get("authOptions")
    auth_on = bool(auth is not None and auth.get("mode") == "token")
    k8s_auth = bool(auth is not None and auth.get("enableK8sTokenAuth") is True)
    ft_opts = spec.get("gcsFaultToleranceOptions")
    ft = (FT_ENABLED_ANNOT in annots and str(annots[FT_ENABLED_ANNOT]).lower() == "true") or ft_opts is not None

    meta = podmeta.build_pod_meta(cluster, [create], podmeta.PodMetaEnv(kuberay_version=env.kuberay_version, deterministic_head_name=env.deterministic_head_name,
                                                                        multihost_indexing_gate=env.multihost_indexing_gate), cluster_hash=cluster_hash)[0]
    pspec = copy.deepcopy((grp.get("template") or {}).get("spec") or {})
    ray = pspec["containers"][0]
    params = dict(grp.get("rayStartParams") or {})

    # ---- DefaultWorkerPodTemplate / DefaultHeadPodTemplate
    if not head and env.init_container_injection:
        pspec["initContainers"] = (pspec.get("initContainers") or []) + [podmeta.ray_init_container(
            ray.get("image", ""), fqdn, head_port, image_pull_policy=ray.get("imagePullPolicy"), env=ray.get("env"), volume_mounts=ray.get("volumeMounts"),
            security_context=ray.get("securityContext"), login_shell=env.login_shell)]
    if head and autoscaling:
        side = podmeta.ray_autoscaler_container(cluster["name"], ray.get("image", ""), options=auto_opts, autoscaler_v2=auto_v2, auth_enabled=auth_on, k8s_token_auth=k8s_auth,
                                                secret_name=(auth or {}).get("secretName"), head_service_account=pspec.get("serviceAccountName"), login_shell=env.login_shell)
        pspec["serviceAccountName"] = side["serviceAccountName"]
        pspec["containers"].append(side["container"])
        _extend(ray, "env", side["rayContainerEnv"])
        if side["restartPolicy"]:
            pspec["restartPolicy"] = side["restartPolicy"]
    got = podmeta.ray_ft_env(node, ft_enabled=ft, cluster_uid=cluster.get("uid", ""), storage_ns_annotation=annots.get(STORAGE_NS_ANNOT), options=ft_opts,
                             head_redis_password_param=head_params.get("redis-password"), existing=_names(ray))
    _extend(ray, "env", got["env"])
    if head:
        params.update(got["rayStartParams"])
    if not any(p.get("name") == "metrics" for p in ray.get("ports") or []):
        ray["ports"] = (ray.get("ports") or []) + [{"name": "metrics", "containerPort": 8080}]
    if not head and autoscaling and auto_v2:
        pspec["restartPolicy"] = "Never"
    if auth_on:
        targets = [ray] + [c for c in pspec.get("initContainers") or [] if c.get("name") == "wait-gcs-ready"]
        for c in targets:
            got = podmeta.ray_auth(cluster["name"], k8s_token_auth=k8s_auth, secret_name=(auth or {}).get("secretName"), existing_env=_names(c),
                                   existing_mount_names=[m.get("name", "") for m in c.get("volumeMounts") or []],
                                   existing_volume_names=[v.get("name", "") for v in pspec.get("volumes") or []])
            _extend(c, "env", got["env"])
            _extend(c, "volumeMounts", got["volumeMounts"])
            _extend(pspec, "volumes", got["volumes"])
    _extend(pspec, "containers", copy.deepcopy(env.head_sidecar_containers if head else env.worker_sidecar_containers))

    # ---- BuildPod
    res = ray.get("resources") or {}
    limits, requests = res.get("limits") or {}, res.get("requests") or {}
    side = next((c for c in pspec["containers"] if c.get("name") == "autoscaler"), None) if head and autoscaling else None
    vols = podmeta.ray_volumes(node, autoscaling=autoscaling, plasma_directory_set="plasma-directory" in params, memory_limit=limits.get("memory"),
                               memory_request=requests.get("memory"), volume_names=[v.get("name", "") for v in pspec.get("volumes") or []],
                               ray_mount_paths=[m.get("mountPath", "") for m in ray.get("volumeMounts") or []],
                               autoscaler_mount_paths=[m.get("mountPath", "") for m in (side or {}).get("volumeMounts") or []])
    _extend(pspec, "volumes", vols["volumes"])
    _extend(ray, "volumeMounts", vols["rayContainerVolumeMounts"])
    if side is not None:
        _extend(side, "volumeMounts", vols["autoscalerVolumeMounts"])
    overwrite = str((meta.get("annotations") or {}).get(OVERWRITE_CMD_ANNOT, "")).lower() == "true"
    rs = podmeta.ray_start_command(node, params, group_labels=grp.get("labels"), group_resources=grp.get("resources"), limits=limits, requests=requests,
                                   command=ray.get("command"), args=ray.get("args"), head_port=head_port, fqdn_ray_ip=fqdn, autoscaling=autoscaling,
                                   overwrite_cmd=overwrite, login_shell=env.login_shell)
    if rs["generated"]:
        ray["command"], ray["args"] = rs["command"], rs["args"]
    for c in pspec.get("initContainers") or []:
        _extend(c, "env", podmeta.ray_container_env(node, fqdn_ray_ip=fqdn, init_container=True))
    _extend(ray, "env", podmeta.ray_container_env(node, existing=_names(ray), default_envs=env.default_container_envs, fqdn_ray_ip=fqdn, head_port=head_port,
                                                  ray_start_cmd=rs["rayStartCommand"], crd_type=crd, kuberay_version=env.kuberay_version))
    if env.probes_injection:
        serve = next((p.get("containerPort", 8000) for p in ray.get("ports") or [] if p.get("name") == "serve"), 8000)
        ray.update(podmeta.ray_probes(node, rs["rayStartParams"], crd_type=crd, ray_version=spec.get("rayVersion", ""), has_liveness=ray.get("livenessProbe") is not None,
                                      has_readiness=ray.get("readinessProbe") is not None, serving_port=int(serve)))
    # ObjectMeta: the template's own metadata with the builders' decisions laid over it (common/pod.go:598); the worker's name is cleared (:418)
    tmeta = {k: v for k, v in ((grp.get("template") or {}).get("metadata") or {}).items() if v is not None and not (k == "name" and not head)}
    return {"apiVersion": "v1", "kind": "Pod", "metadata": {**copy.deepcopy(tmeta), **meta}, "spec": pspec}


def build_pods_native(cluster: dict, creates: list[tuple[int, int, int, str]], env: BuilderEnv | None = None, cluster_hash: str | None = None, raw: bool = False) -> list:
    """kr_pod_build: the same manifests from ONE native call for all of a RayCluster's create tuples (the container half assembled once per group
    inside the library).  raw=True returns each Pod's JSON bytes exactly as written (Go field order)."""
    import ctypes as C
    import json

    from . import abi
    from .engine import EngineError
    env = env or BuilderEnv()
    L = podmeta._lib()
    if not getattr(L, "_kr_pb_bound", False):
        L.kr_pod_build.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(abi.kr_podbuild_env), C.POINTER(abi.kr_podmeta_create), C.c_uint32, C.c_void_p, C.c_uint64,
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.kr_pod_build_last_error.restype = C.c_char_p
        L._kr_pb_bound = True
    keep = podmeta._Keep()
    # the RayCluster in its Kubernetes shape (what json.Marshal(instance) gives the Go shim)
    doc = json.dumps({"metadata": {k: cluster[k] for k in ("name", "namespace", "uid", "labels", "annotations") if k in cluster}, "spec": cluster.get("spec") or {}}).encode()
    e = abi.kr_podbuild_env()
    e.kuberay_version, e.cluster_domain, e.cluster_hash = keep.s(env.kuberay_version), keep.s(env.cluster_domain), keep.s(cluster_hash or None)
    e.deterministic_head_name, e.gate_multihost_indexing, e.login_shell = int(env.deterministic_head_name), int(env.multihost_indexing_gate), int(env.login_shell)
    e.no_init_container_injection, e.no_probes_injection = int(not env.init_container_injection), int(not env.probes_injection)
    e.default_envs, e.n_default_envs = keep.kvs(env.default_container_envs)
    e.head_sidecars_json = keep.s(json.dumps(env.head_sidecar_containers) if env.head_sidecar_containers else None)
    e.worker_sidecars_json = keep.s(json.dumps(env.worker_sidecar_containers) if env.worker_sidecar_containers else None)
    tuples = (abi.kr_podmeta_create * max(len(creates), 1))()
    for i, (g, ri, hi, rn) in enumerate(creates):
        tuples[i].group, tuples[i].replica_index, tuples[i].host_index = int(g), int(ri), int(hi)
        tuples[i].replica_name = keep.s(rn or "")
    off = (C.c_uint64 * (len(creates) + 1))()
    need = C.c_uint64()
    rc = L.kr_pod_build(doc, len(doc), C.byref(e), tuples, len(creates), None, 0, off, C.byref(need))
    if rc not in (0, abi.KR_E_CAPACITY):
        raise EngineError(int(rc), (L.kr_pod_build_last_error() or b"").decode())
    buf = (C.c_uint8 * max(need.value, 1))()
    rc = L.kr_pod_build(doc, len(doc), C.byref(e), tuples, len(creates), buf, need.value, off, C.byref(need))
    if rc:
        raise EngineError(int(rc), (L.kr_pod_build_last_error() or b"").decode())
    b = bytes(buf)
    parts = [b[off[i]:off[i + 1]] for i in range(len(creates))]
    return parts if raw else [json.loads(p) for p in parts]
