#!/usr/bin/env python
"""Throughput of the host-side builders behind the C ABI (no GPU): Pod ObjectMeta patches, `ray start` command lines, container env lists
and the muted-spec JSON emitter — timed through ctypes with the arguments marshalled once (development aid; prints one JSON line)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kuberay_b200 import abi, podmeta as pm  # noqa: E402
from kuberay_b200.engine import lib, spec_json_emit  # noqa: E402

L = pm._bind_raystart()
L.kr_ray_container_env.argtypes = [C.POINTER(abi.kr_rayenv_in), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
keep = pm._Keep()


def timed(fn, n, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return n / best


# ---- Pod metadata: one RayCluster with 4 groups, 1000 creates per call
cluster = {"name": "raycluster-sample", "namespace": "default", "uid": "0f5b5f0c-7c1e-4f0e-9a0a-3f1d8a4c2b11",
           "spec": {"headGroupSpec": {"template": {"metadata": {"labels": {"team": "ml", "zone": "a"}, "annotations": {"a": "b"}}}},
                    "workerGroupSpecs": [{"groupName": f"group-{i}", "numOfHosts": 4 if i == 3 else 1, "labels": {"pool": f"p{i}"},
                                          "template": {"metadata": {"labels": {"team": "ml", "tier": str(i)}}}} for i in range(4)]}}
c = pm._cluster_struct(keep, cluster, pm.PodMetaEnv(), "0123456789ABCDEFGHIJKLMNOPQRSTUV")
head = pm._group_struct(keep, cluster["spec"]["headGroupSpec"], True)
groups = (abi.kr_podmeta_group * 4)(*[pm._group_struct(keep, g, False) for g in cluster["spec"]["workerGroupSpecs"]])
N = 1000
tuples = (abi.kr_podmeta_create * N)()
for i in range(N):
    tuples[i].group, tuples[i].replica_index, tuples[i].host_index = i % 4, i // 4, i % 4
    tuples[i].replica_name = keep.s(f"group-3-abcde")
off = (C.c_uint64 * (N + 1))()
need = C.c_uint64()
buf = (C.c_uint8 * (1 << 21))()


def meta():
    for _ in range(50):
        assert L.kr_pod_meta_build(C.byref(c), C.byref(head), groups, 4, tuples, N, buf, len(buf), off, C.byref(need)) == 0


# ---- ray start command + env: one call per group
rs = abi.kr_raystart_in()
rs.node_type = abi.NT_WORKER
rs.fqdn_ray_ip, rs.head_port = keep.s("raycluster-sample-head-svc.default.svc.cluster.local"), keep.s("6379")
rs.ray_start_params, rs.n_ray_start_params = keep.kvs({"num-cpus": "4", "object-store-memory": "1000000000"})
rs.group_labels, rs.n_group_labels = keep.kvs({"zone": "us-central2", "spot": "true"})
rs.container_limits, rs.n_container_limits = keep.kvs({"cpu": "4", "memory": "16Gi", "nvidia.com/gpu": "1", "google.com/tpu": "4"})
ev = abi.kr_rayenv_in()
ev.node_type = abi.NT_WORKER
ev.fqdn_ray_ip, ev.head_port, ev.ray_start_cmd, ev.kuberay_version = rs.fqdn_ray_ip, rs.head_port, keep.s("ray start  --block "), keep.s("v1.5.0")
small = (C.c_uint8 * 8192)()


def start_cmd():
    for _ in range(20000):
        assert L.kr_ray_start_command(C.byref(rs), small, len(small), C.byref(need)) == 0


def env():
    for _ in range(20000):
        assert L.kr_ray_container_env(C.byref(ev), small, len(small), C.byref(need)) == 0


spec = json.dumps({"rayVersion": "2.46.0", "headGroupSpec": {"rayStartParams": {"dashboard-host": "0.0.0.0"}, "template": {"spec": {"containers": [
    {"name": "ray", "image": "rayproject/ray:2.46.0", "env": [{"name": f"E{i}", "value": str(i)} for i in range(40)],
     "resources": {"limits": {"cpu": "2", "memory": "4Gi"}, "requests": {"cpu": "2", "memory": "4Gi"}}}]}}},
    "workerGroupSpecs": [{"groupName": "g", "replicas": 3, "minReplicas": 1, "maxReplicas": 9, "rayStartParams": {}, "template": {"spec": {"containers": [
        {"name": "ray", "image": "rayproject/ray:2.46.0", "env": [{"name": f"E{i}", "value": str(i)} for i in range(40)]}]}}}]}).encode()


def emit():
    for _ in range(2000):
        spec_json_emit(spec)


# ---- template surgery + the whole manifest (the assembly glue is Python here, Go in the shim: its cost is reported, not hidden)
from kuberay_b200 import podbuilder  # noqa: E402

ini = abi.kr_rayinit_in()
ini.image, ini.fqdn_ray_ip, ini.head_port = keep.s("rayproject/ray:2.46.0"), rs.fqdn_ray_ip, rs.head_port
ini.env_json = keep.s(json.dumps([{"name": f"E{i}", "value": str(i)} for i in range(10)], separators=(",", ":")))
au = abi.kr_rayautoscaler_in()
au.cluster_name, au.ray_image, au.auth_enabled = keep.s("raycluster-sample"), keep.s("rayproject/ray:2.46.0"), 1
L.kr_ray_init_container.argtypes = L.kr_ray_autoscaler_container.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]


def init_container():
    for _ in range(20000):
        assert L.kr_ray_init_container(C.byref(ini), small, len(small), C.byref(need)) == 0


def autoscaler():
    for _ in range(20000):
        assert L.kr_ray_autoscaler_container(C.byref(au), small, len(small), C.byref(need)) == 0


full = {"name": "raycluster-sample", "namespace": "default", "uid": "u", "spec": {"enableInTreeAutoscaling": True, **json.loads(spec)}}


def whole_pods():
    for i in range(300):
        podbuilder.build_pod(full, (-1, 0, 0, "") if i % 4 == 0 else (0, i, 0, ""))


# kr_pod_build straight through ctypes (arguments marshalled once, like the other lines): every create tuple of one RayCluster per call
pb_doc = json.dumps({"metadata": {"name": "raycluster-sample", "namespace": "default", "uid": "u"}, "spec": full["spec"]}).encode()
pb_env = abi.kr_podbuild_env()
pb_env.kuberay_version, pb_env.gate_multihost_indexing = keep.s("v1.5.0"), 1
pb_tuples = (abi.kr_podmeta_create * 1000)()
for i in range(1000):
    pb_tuples[i].group, pb_tuples[i].replica_index, pb_tuples[i].replica_name = (-1 if i == 0 else 0), i, keep.s("")
pb_off = (C.c_uint64 * 1001)()
pb_buf = (C.c_uint8 * (8 << 20))()
L.kr_pod_build.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]


def pod_build(n, reps):
    def run():
        for _ in range(reps):
            assert L.kr_pod_build(pb_doc, len(pb_doc), C.byref(pb_env), pb_tuples, n, pb_buf, len(pb_buf), pb_off, C.byref(need)) == 0
    return run


if "--quick" in sys.argv:  # bench.py's extra: the one-call whole-Pod builder and the muted-spec emitter only
    print(json.dumps({"whole_pod_manifests_per_s_kr_pod_build_1000_per_call": round(timed(pod_build(1000, 5), 5 * 1000)),
                      "whole_pod_manifests_per_s_kr_pod_build_10_per_call": round(timed(pod_build(10, 100), 100 * 10)),
                      "whole_pod_manifest_bytes": len(podbuilder.build_pods_native(full, [(0, 0, 0, "")], raw=True)[0]),
                      "muted_spec_json_emits_per_s": round(timed(lambda: [spec_json_emit(spec) for _ in range(500)], 500)), "muted_spec_json_input_bytes": len(spec),
                      "threads_used": 1}))
    sys.exit(0)

out = {"host": os.uname().nodename, "cpus": os.cpu_count(), "threads_used": 1,
       "pod_meta_patches_per_s": round(timed(meta, 50 * N)), "pod_meta_bytes_per_patch": need.value and int(off[N] / N),
       "ray_start_commands_per_s": round(timed(start_cmd, 20000)), "container_env_lists_per_s": round(timed(env, 20000)),
       "muted_spec_json_emits_per_s": round(timed(emit, 2000)), "muted_spec_json_input_bytes": len(spec),
       "init_containers_per_s": round(timed(init_container, 20000)), "autoscaler_containers_per_s": round(timed(autoscaler, 20000)),
       "whole_pod_manifests_per_s_python_glue": round(timed(whole_pods, 300)),
       "whole_pod_manifests_per_s_kr_pod_build_1000_per_call": round(timed(pod_build(1000, 20), 20 * 1000)),
       "whole_pod_manifests_per_s_kr_pod_build_10_per_call": round(timed(pod_build(10, 500), 500 * 10)),
       "whole_pod_manifests_per_s_kr_pod_build_1_per_call": round(timed(pod_build(1, 2000), 2000)),
       "whole_pod_manifest_bytes": len(podbuilder.build_pods_native(full, [(0, 0, 0, "")], raw=True)[0]),
       "note": "one host thread through ctypes, arguments marshalled once; each call does all of its own work (no caching across calls)"}
print(json.dumps(out))
