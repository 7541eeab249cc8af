"""The whole Pod: buildHeadPod / buildWorkerPod (raycluster_controller.go:1387-1433) assembled from the native builders
(kuberay_b200/podbuilder.py over the C ABI) against
  * the reference's BuildPod / DefaultHead|WorkerPodTemplate tests, transcribed from ray-operator/controllers/ray/common/pod_test.go
    (fixture `instance` :29-105; TestBuildPod :641-728, _WithPlasmaDirectory :730-777, _WithNoCPULimits :882-924, _WithOverwriteCommand
    :926-955, _WithAutoscalerEnabled :957-986, _WithCreatedByRayService :988-1012, _WithLoginBash :1014-1077, TestHeadPodTemplate_* :1157-1311,
    TestDefault*PodTemplateWithConfigurablePorts :1365-1407, TestDefault*PodTemplate_Autoscaling :1176-1234, 1409-1447),
  * the statement-by-statement CPU restatement oracle/podmeta.py:build_pod on randomised RayClusters."""
import copy
import random

import pytest

from kuberay_b200 import engine, podbuilder as pb
from oracle import podmeta as ref

HEAD, WORKER = (-1, 0, 0, ""), (0, 0, 0, "")
FQDN = "raycluster-sample-head-svc.default.svc.cluster.local"


def instance():
    """pod_test.go:29-105."""
    gi = {"cpu": "1", "memory": "1Gi"}
    return {"name": "raycluster-sample", "namespace": "default", "uid": "uid-0",
            "spec": {"headGroupSpec": {"rayStartParams": {}, "template": {"metadata": {"namespace": "default"}, "spec": {"containers": [
                {"name": "ray-head", "image": "repo/image:custom", "env": [{"name": "TEST_ENV_NAME", "value": "TEST_ENV_VALUE"}],
                 "resources": {"requests": dict(gi), "limits": dict(gi)}}]}}},
                "workerGroupSpecs": [{"replicas": 3, "minReplicas": 0, "maxReplicas": 10000, "groupName": "small-group", "rayStartParams": {"port": "6379"},
                                      "template": {"metadata": {"namespace": "default"}, "spec": {"containers": [
                                          {"name": "ray-worker", "image": "repo/image:custom", "resources": {"limits": {"cpu": "1", "memory": "1Gi", "nvidia.com/gpu": "3"}},
                                           "env": [{"name": "TEST_ENV_NAME", "value": "TEST_ENV_VALUE"}]}]}}}]}}


def env_value(container, name):
    """checkContainerEnv's reading: the value, or the fieldRef path."""
    e = next(e for e in container["env"] if e["name"] == name)
    return e["value"] if e.get("value", "") != "" else e["valueFrom"]["fieldRef"]["fieldPath"]


def split_sorted(s):
    return sorted(x for x in s.split(" ") if x)


def build_both(cluster, create, **kw):
    """The native assembly and the restatement on the same input; returns the native Pod after checking they agree."""
    env = pb.BuilderEnv(**kw)
    got = pb.build_pod(cluster, create, env)
    want = ref.build_pod(cluster, create, kuberay_version=env.kuberay_version, deterministic_head_name=env.deterministic_head_name,
                         multihost_indexing_gate=env.multihost_indexing_gate, login_shell=env.login_shell, init_container_injection=env.init_container_injection,
                         probes_injection=env.probes_injection, cluster_domain=env.cluster_domain, default_container_envs=env.default_container_envs,
                         canonical=engine.quantity_canonical)
    assert got == want
    return got


def test_build_pod():
    """TestBuildPod (:641-728)."""
    cluster = instance()
    pod = build_both(cluster, HEAD, default_container_envs={"TEST_DEFAULT_ENV_NAME": "TEST_ENV_VALUE"})
    ray = pod["spec"]["containers"][0]
    assert env_value(ray, "RAY_ADDRESS") == "127.0.0.1:6379" and env_value(ray, "RAY_USAGE_STATS_KUBERAY_IN_USE") == "1"
    assert env_value(ray, "RAY_CLUSTER_NAME") == "metadata.labels['ray.io/cluster']" and env_value(ray, "RAY_CLUSTER_NAMESPACE") == "metadata.namespace"
    assert env_value(ray, "RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE") == "1" and env_value(ray, "RAY_NODE_TYPE_NAME") == "metadata.labels['ray.io/group']"
    assert env_value(ray, "RAY_USAGE_STATS_EXTRA_TAGS") == "kuberay_version=v1.5.0;kuberay_crd=RayCluster"
    assert "ray start" in env_value(ray, "KUBERAY_GEN_RAY_START_CMD")
    labels = pod["metadata"]["labels"]
    assert labels["ray.io/cluster"] == "raycluster-sample" and labels["ray.io/node-type"] == "head" and labels["ray.io/group"] == "headgroup"
    assert pod["spec"]["volumes"] == [{"name": "shared-mem", "emptyDir": {"medium": "Memory", "sizeLimit": "1Gi"}}]      # volumesNoAutoscaler
    assert ray["volumeMounts"] == [{"name": "shared-mem", "mountPath": "/dev/shm"}]                                   # volumeMountsNoAutoscaler
    assert "initContainers" not in pod["spec"] and pod["metadata"]["generateName"] == "raycluster-sample-head-"

    pod = build_both(cluster, WORKER, default_container_envs={"TEST_DEFAULT_ENV_NAME": "TEST_ENV_VALUE"})
    ray = pod["spec"]["containers"][0]
    assert ray["resources"]["limits"] == {"cpu": "1", "memory": "1Gi", "nvidia.com/gpu": "3"}
    assert env_value(ray, "RAY_ADDRESS") == FQDN + ":6379" and env_value(ray, "FQ_RAY_IP") == FQDN and env_value(ray, "RAY_IP") == "raycluster-sample-head-svc"
    assert env_value(ray, "RAY_CLUSTER_NAME") == "metadata.labels['ray.io/cluster']" and env_value(ray, "RAY_NODE_TYPE_NAME") == "metadata.labels['ray.io/group']"
    assert "ray start" in env_value(ray, "KUBERAY_GEN_RAY_START_CMD")
    assert split_sorted(ray["args"][0]) == split_sorted("ulimit -n 65536; ray start --block --dashboard-agent-listen-port=52365 --memory=1073741824 --num-cpus=1 --num-gpus=3 "
                                                         f"--address={FQDN}:6379 --port=6379 --metrics-export-port=8080")
    assert env_value(ray, "TEST_ENV_NAME") == "TEST_ENV_VALUE" and env_value(ray, "TEST_DEFAULT_ENV_NAME") == "TEST_ENV_VALUE"
    # the injected init container waits for THIS head service, got the Ray container's env as it was in the template, then FQ_RAY_IP / RAY_IP (:655-660)
    init = pod["spec"]["initContainers"][-1]
    assert init["name"] == "wait-gcs-ready" and [e["name"] for e in init["env"]] == ["TEST_ENV_NAME", "FQ_RAY_IP", "RAY_IP"] and env_value(init, "FQ_RAY_IP") == FQDN
    assert f"ray health-check --address {FQDN}:6379" in init["args"][0]


@pytest.mark.parametrize("params,expect", [({}, True), ({"plasma-directory": "/dev/shm"}, False), ({"plasma-directory": "/dev/shm/"}, False), ({"plasma-directory": "/tmp/ray/plasma"}, False)])
def test_plasma_directory_skips_the_shared_memory_mount(params, expect):
    """TestBuildPod_WithPlasmaDirectory (:730-777)."""
    cluster = instance()
    cluster["spec"]["headGroupSpec"]["rayStartParams"] = dict(params)
    pod = build_both(cluster, HEAD)
    assert any(m["mountPath"] == "/dev/shm" for m in pod["spec"]["containers"][0].get("volumeMounts", [])) == expect
    assert any(v["name"] == "shared-mem" for v in pod["spec"].get("volumes", [])) == expect


def test_no_cpu_limits():
    """TestBuildPod_WithNoCPULimits (:882-924): num-cpus falls back to the request."""
    cluster = instance()
    cluster["spec"]["headGroupSpec"]["template"]["spec"]["containers"][0]["resources"] = {"requests": {"cpu": "2", "memory": "1Gi"}, "limits": {"memory": "1Gi"}}
    cluster["spec"]["workerGroupSpecs"][0]["template"]["spec"]["containers"][0]["resources"] = {"requests": {"cpu": "2", "memory": "1Gi"}, "limits": {"memory": "1Gi", "nvidia.com/gpu": "3"}}
    head = build_both(cluster, HEAD)
    assert split_sorted(head["spec"]["containers"][0]["args"][0]) == split_sorted(
        "ulimit -n 65536; ray start --head --block --dashboard-agent-listen-port=52365 --memory=1073741824 --num-cpus=2 --metrics-export-port=8080 --dashboard-host=0.0.0.0")
    worker = build_both(cluster, WORKER)
    assert split_sorted(worker["spec"]["containers"][0]["args"][0]) == split_sorted(
        f"ulimit -n 65536; ray start --block --dashboard-agent-listen-port=52365 --memory=1073741824 --num-cpus=2 --num-gpus=3 --address={FQDN}:6379 --port=6379 --metrics-export-port=8080")


def test_overwrite_command():
    """TestBuildPod_WithOverwriteCommand (:926-955)."""
    cluster = instance()
    cluster["annotations"] = {"ray.io/overwrite-container-cmd": "true"}
    cluster["spec"]["headGroupSpec"]["template"]["spec"]["containers"][0].update(command=["I am head"], args=["I am head again"])
    cluster["spec"]["workerGroupSpecs"][0]["template"]["spec"]["containers"][0].update(command=["I am worker"], args=["I am worker again"])
    h = build_both(cluster, HEAD)["spec"]["containers"][0]
    w = build_both(cluster, WORKER)["spec"]["containers"][0]
    assert (h["command"], h["args"], w["command"], w["args"]) == (["I am head"], ["I am head again"], ["I am worker"], ["I am worker again"])
    # without the annotation the user's command runs first, then the generated one (:641-647)
    del cluster["annotations"]
    h = build_both(cluster, HEAD)["spec"]["containers"][0]
    assert h["command"] == ["/bin/bash", "-c", "--"] and h["args"][0].startswith(" I am head  I am head again  && ulimit -n 65536; ray start --head ")


def test_autoscaler_enabled():
    """TestBuildPod_WithAutoscalerEnabled (:957-986), TestHeadPodTemplate_WithAutoscalingEnabled (:1157-1174)."""
    cluster = instance()
    cluster["spec"]["enableInTreeAutoscaling"] = True
    pod = build_both(cluster, HEAD)
    assert "--no-monitor" in pod["spec"]["containers"][0]["args"][0]
    assert pod["spec"]["volumes"] == [{"name": "shared-mem", "emptyDir": {"medium": "Memory", "sizeLimit": "1Gi"}}, {"name": "ray-logs", "emptyDir": {}}]
    assert pod["spec"]["containers"][0]["volumeMounts"] == [{"name": "shared-mem", "mountPath": "/dev/shm"}, {"name": "ray-logs", "mountPath": "/tmp/ray"}]
    assert len(pod["spec"]["containers"]) == 2 and pod["spec"]["serviceAccountName"] == "raycluster-sample"
    small = {"cpu": "500m", "memory": "512Mi"}
    field = lambda p: {"fieldRef": {"fieldPath": p}}  # noqa: E731
    assert pod["spec"]["containers"][1] == {                                                                        # `autoscalerContainer` (:145-218)
        "name": "autoscaler", "image": "repo/image:custom", "imagePullPolicy": "IfNotPresent",
        "env": [{"name": "RAY_CLUSTER_NAME", "valueFrom": field("metadata.labels['ray.io/cluster']")}, {"name": "RAY_CLUSTER_NAMESPACE", "valueFrom": field("metadata.namespace")},
                {"name": "RAY_HEAD_POD_NAME", "valueFrom": field("metadata.name")}, {"name": "KUBERAY_CRD_VER", "value": "v1"}],
        "command": ["/bin/bash", "-c", "--"], "args": ["ray kuberay-autoscaler --cluster-name $(RAY_CLUSTER_NAME) --cluster-namespace $(RAY_CLUSTER_NAMESPACE)"],
        "resources": {"limits": small, "requests": small}, "volumeMounts": [{"name": "ray-logs", "mountPath": "/tmp/ray"}]}
    cluster["name"] = "x" * 200                                                                                     # longString / shortString
    assert build_both(cluster, HEAD)["spec"]["serviceAccountName"] == "x" * 50


@pytest.mark.parametrize("sa,autoscaling,want", [(None, False, None), ("head-service-account", False, "head-service-account"), ("head-service-account", True, "head-service-account")])
def test_head_service_account(sa, autoscaling, want):
    """TestHeadPodTemplate_WithNoServiceAccount / _WithServiceAccountNoAutoscaling / _WithServiceAccount (:1268-1311)."""
    cluster = instance()
    if sa:
        cluster["spec"]["headGroupSpec"]["template"]["spec"]["serviceAccountName"] = sa
    cluster["spec"]["enableInTreeAutoscaling"] = autoscaling
    assert build_both(cluster, HEAD)["spec"].get("serviceAccountName") == want


@pytest.mark.parametrize("mode,containers,v2env,restart", [("off", 1, False, None), ("v1", 2, False, None), ("v2", 2, True, "Never")])
def test_autoscaling_versions(mode, containers, v2env, restart):
    """TestDefaultHeadPodTemplate_Autoscaling (:1176-1234) and TestDefaultWorkerPodTemplate_Autoscaling (:1409-1447)."""
    cluster = instance()
    if mode != "off":
        cluster["spec"]["enableInTreeAutoscaling"] = True
    if mode == "v2":
        cluster["spec"]["autoscalerOptions"] = {"version": "v2"}
    head = build_both(cluster, HEAD)
    assert len(head["spec"]["containers"]) == containers and head["spec"].get("restartPolicy") == restart
    assert ({"name": "RAY_enable_autoscaler_v2", "value": "true"} in head["spec"]["containers"][0]["env"]) == v2env
    assert build_both(cluster, WORKER)["spec"].get("restartPolicy") == restart


def test_created_by_rayservice():
    """TestBuildPod_WithCreatedByRayService (:988-1012): the serve label, the three RayService timeouts, the worker's Serve proxy readiness check."""
    cluster = instance()
    cluster["labels"] = {"ray.io/originated-from-crd": "RayService"}
    cluster["spec"]["enableInTreeAutoscaling"] = True
    head, worker = build_both(cluster, HEAD), build_both(cluster, WORKER)
    assert head["metadata"]["labels"]["ray.io/serve"] == "false" and worker["metadata"]["labels"]["ray.io/serve"] == "true"
    for pod in (head, worker):
        assert env_value(pod["spec"]["containers"][0], "RAY_timeout_ms_task_wait_for_death_info") == "0"
    assert "localhost:8000/-/healthz" in worker["spec"]["containers"][0]["readinessProbe"]["exec"]["command"][2]
    assert "-/healthz" not in head["spec"]["containers"][0]["readinessProbe"]["exec"]["command"][2]
    assert env_value(head["spec"]["containers"][0], "RAY_USAGE_STATS_EXTRA_TAGS").endswith("kuberay_crd=RayService")


@pytest.mark.parametrize("login", [True, False])
def test_login_bash(login):
    """TestBuildPod_WithLoginBash (:1014-1077): Ray container, autoscaler sidecar, worker, init container."""
    want = ["/bin/bash", "-cl" if login else "-c", "--"]
    cluster = instance()
    cluster["labels"] = {"ray.io/originated-from-crd": "RayService"}
    cluster["spec"]["enableInTreeAutoscaling"] = True
    head, worker = build_both(cluster, HEAD, login_shell=login), build_both(cluster, WORKER, login_shell=login)
    assert head["spec"]["containers"][0]["command"] == want and head["spec"]["containers"][1]["command"] == want
    assert worker["spec"]["containers"][0]["command"] == want and worker["spec"]["initContainers"][0]["command"] == want


@pytest.mark.parametrize("create", [HEAD, WORKER])
def test_configurable_metrics_port(create):
    """TestDefaultHeadPodTemplateWithConfigurablePorts / ...Worker... (:1365-1407)."""
    cluster = instance()
    grp = cluster["spec"]["headGroupSpec"] if create[0] < 0 else cluster["spec"]["workerGroupSpecs"][0]
    grp["template"]["spec"]["containers"][0]["ports"] = []
    assert {"name": "metrics", "containerPort": 8080} in build_both(cluster, create)["spec"]["containers"][0]["ports"]
    grp["template"]["spec"]["containers"][0]["ports"] = [{"name": "metrics", "containerPort": 8081}]
    assert build_both(cluster, create)["spec"]["containers"][0]["ports"] == [{"name": "metrics", "containerPort": 8081}]


def test_switches_and_sidecars():
    """ENABLE_INIT_CONTAINER_INJECTION / ENABLE_PROBES_INJECTION off (:337-349), operator sidecars (raycluster_controller.go:1397-1399, 1424-1426), a
    template that brings its own init container (it gets FQ_RAY_IP too, :651-653), user probes left alone, a custom head service name and cluster domain."""
    cluster = instance()
    w = cluster["spec"]["workerGroupSpecs"][0]["template"]["spec"]
    w["initContainers"] = [{"name": "mine", "image": "busybox"}]
    w["containers"][0]["livenessProbe"] = {"exec": {"command": ["true"]}}
    cluster["spec"]["headGroupSpec"]["headService"] = {"metadata": {"name": "my-head"}}
    pod = build_both(cluster, WORKER, init_container_injection=False, cluster_domain="example.org")
    assert [c["name"] for c in pod["spec"]["initContainers"]] == ["mine"] and env_value(pod["spec"]["initContainers"][0], "FQ_RAY_IP") == "my-head.default.svc.example.org"
    ray = pod["spec"]["containers"][0]
    assert ray["livenessProbe"] == {"exec": {"command": ["true"]}} and "readinessProbe" in ray and env_value(ray, "RAY_IP") == "my-head"
    pod = build_both(cluster, WORKER, probes_injection=False)
    assert "readinessProbe" not in pod["spec"]["containers"][0] and [c["name"] for c in pod["spec"]["initContainers"]] == ["mine", "wait-gcs-ready"]
    env = pb.BuilderEnv(head_sidecar_containers=[{"name": "fluentbit", "image": "fb"}], worker_sidecar_containers=[{"name": "w-side", "image": "ws"}])
    assert [c["name"] for c in pb.build_pod(cluster, HEAD, env)["spec"]["containers"]] == ["ray-head", "fluentbit"]
    assert [c["name"] for c in pb.build_pod(cluster, WORKER, env)["spec"]["containers"]] == ["ray-worker", "w-side"]
    assert cluster == (lambda c: (c["spec"]["workerGroupSpecs"][0]["template"]["spec"].update(initContainers=[{"name": "mine", "image": "busybox"}]), c)[1])(copy.deepcopy(cluster))  # input untouched


# ---- randomised RayClusters: assembly == restatement ---------------------------------------------------------------------------------
def random_cluster(rng):
    pick = lambda *xs: rng.choice(xs)  # noqa: E731
    maybe = lambda x, p=0.5: x if rng.random() < p else None  # noqa: E731

    def container(name):
        c = {"name": name, "image": pick("rayproject/ray:2.9.0", "repo/image:custom")}
        res = {}
        if rng.random() < 0.8:
            res["limits"] = {k: v for k, v in {"cpu": pick("1", "500m", "2"), "memory": pick("1Gi", "1024Mi", "2G", "512Mi"), "nvidia.com/gpu": maybe("2", 0.3),
                                               "google.com/tpu": maybe("4", 0.15)}.items() if v is not None and rng.random() < 0.85}
        if rng.random() < 0.5:
            res["requests"] = {"cpu": pick("250m", "1"), "memory": pick("256Mi", "1Gi")}
        if res:
            c["resources"] = res
        env = [{"name": n, "value": "v"} for n in rng.sample(["RAY_PORT", "RAY_ADDRESS", "REDIS_PASSWORD", "RAY_AUTH_MODE", "FOO", "RAY_USAGE_STATS_KUBERAY_IN_USE",
                                                              "RAY_external_storage_namespace", "RAY_gcs_rpc_server_reconnect_timeout_s"], rng.randint(0, 3))]
        if env or rng.random() < 0.3:
            c["env"] = env
        if rng.random() < 0.3:
            c["ports"] = [p for p in ({"name": "metrics", "containerPort": 9090}, {"name": "serve", "containerPort": 8123}, {"name": "dash", "containerPort": 8265}) if rng.random() < 0.5]
        if rng.random() < 0.2:
            c["command"], c["args"] = pick(["python"], ["ray start --head"]), pick(["x.py"], [])
        if rng.random() < 0.25:
            c["volumeMounts"] = [pick({"name": "shm", "mountPath": "/dev/shm"}, {"name": "logs", "mountPath": "/tmp/ray"}, {"name": "ray-token", "mountPath": "/tok"})]
        if rng.random() < 0.2:
            c["readinessProbe"] = {"exec": {"command": ["true"]}}
        if rng.random() < 0.2:
            c["securityContext"] = {"runAsUser": 1000}
        if rng.random() < 0.2:
            c["imagePullPolicy"] = pick("Always", "Never")
        return c

    def template(name):
        t = {"spec": {"containers": [container(name)] + ([{"name": "side", "image": "s"}] if rng.random() < 0.2 else [])}}
        if rng.random() < 0.4:
            t["metadata"] = {"labels": {"team": "a", "ray.io/group": "spoof"}, "annotations": {k: v for k, v in {"a": "b", "ray.io/overwrite-container-cmd": maybe("true", 0.2)}.items() if v}}
        if rng.random() < 0.2:
            t["spec"]["volumes"] = [pick({"name": "shared-mem", "emptyDir": {}}, {"name": "ray-logs", "emptyDir": {}}, {"name": "other", "emptyDir": {}})]
        if rng.random() < 0.2:
            t["spec"]["serviceAccountName"] = pick("sa", "9-bad-name")
        if rng.random() < 0.15:
            t["spec"]["initContainers"] = [{"name": pick("wait-gcs-ready", "setup"), "image": "busybox"}]
        return t

    def params(head):
        p = {k: v for k, v in {"port": maybe("6380", 0.2) if head else None, "num-cpus": maybe("3", 0.2), "dashboard-port": maybe("8266", 0.2), "block": maybe("false", 0.1),
                               "plasma-directory": maybe("/tmp/p", 0.1), "redis-password": maybe("pw", 0.2) if head else None, "resources": maybe("'{\"a\": 1}'", 0.15),
                               "dashboard-agent-listen-port": maybe("5000", 0.1), "address": maybe("h:1", 0.1) if not head else None}.items() if v is not None}
        return p

    cred = lambda n: pick(None, {"value": n}, {"valueFrom": {"secretKeyRef": {"name": "s", "key": n}}})  # noqa: E731
    ft = maybe({k: v for k, v in {"redisAddress": "redis:6379", "externalStorageNamespace": maybe("ns-opt"), "redisUsername": cred("u"), "redisPassword": cred("p")}.items() if v is not None}, 0.3)
    auto = maybe({k: v for k, v in {"version": maybe(pick("v1", "v2")), "image": maybe("auto:1", 0.3), "imagePullPolicy": maybe("Always", 0.3), "env": maybe([{"name": "E", "value": "1"}], 0.3),
                                     "resources": maybe({"limits": {"cpu": "1"}}, 0.3), "volumeMounts": maybe([{"name": "x", "mountPath": "/x"}], 0.2),
                                     "securityContext": maybe({"privileged": False}, 0.2), "envFrom": maybe([{"prefix": "P"}], 0.2)}.items() if v is not None}, 0.5)
    spec = {"headGroupSpec": {"rayStartParams": params(True), "template": template("ray-head")},
            "workerGroupSpecs": [{"groupName": f"g{i}", "numOfHosts": pick(1, 1, 2), "rayStartParams": params(False), "template": template("ray-worker")} for i in range(rng.randint(1, 3))]}
    for grp in [spec["headGroupSpec"]] + spec["workerGroupSpecs"]:
        if rng.random() < 0.2:
            grp["labels"] = {"zone": "a", "ray.io/cluster": "spoof"}
        if rng.random() < 0.2:
            grp["resources"] = {k: v for k, v in {"CPU": "4", "memory": maybe("2Gi"), "GPU": maybe("1"), "custom": maybe("0.5")}.items() if v is not None}
    for k, v in (("enableInTreeAutoscaling", maybe(True)), ("autoscalerOptions", auto), ("gcsFaultToleranceOptions", ft), ("rayVersion", maybe(pick("2.9.0", "2.53.0", "nightly"))),
                 ("authOptions", maybe(pick({"mode": "token"}, {"mode": "token", "enableK8sTokenAuth": True}, {"mode": "token", "secretName": "sec"}, {"mode": "disabled"}), 0.4))):
        if v is not None:
            spec[k] = v
    if rng.random() < 0.15:
        spec["headGroupSpec"]["headService"] = {"metadata": {"name": "custom-svc"}}
    cluster = {"name": pick("rc", "raycluster-sample", "9starts-with-digit", "n" * 70), "namespace": pick("default", "ml"), "uid": "uid-" + str(rng.randint(0, 99)), "spec": spec}
    ann = {k: v for k, v in {"ray.io/ft-enabled": maybe(pick("true", "True", "false"), 0.3), "ray.io/external-storage-namespace": maybe("ns-ann", 0.2),
                             "ray.io/overwrite-container-cmd": maybe(pick("true", "TRUE", "no"), 0.15)}.items() if v is not None}
    if ann:
        cluster["annotations"] = ann
    if rng.random() < 0.3:
        cluster["labels"] = {"ray.io/originated-from-crd": pick("RayService", "RayJob", "RayCluster")}
    return cluster


@pytest.mark.parametrize("seed", range(150))
def test_random_clusters_assemble_like_the_restatement(seed):
    rng = random.Random(seed)
    cluster = random_cluster(rng)
    before = copy.deepcopy(cluster)
    kw = dict(login_shell=rng.random() < 0.2, init_container_injection=rng.random() < 0.85, probes_injection=rng.random() < 0.85,
              deterministic_head_name=rng.random() < 0.3, multihost_indexing_gate=rng.random() < 0.7,
              default_container_envs=rng.choice([{}, {"RAY_PORT": "1", "DEF": "d"}]))
    build_both(cluster, HEAD, **kw)
    for gi, grp in enumerate(cluster["spec"]["workerGroupSpecs"]):
        hosts = int(grp.get("numOfHosts", 1))
        build_both(cluster, (gi, rng.randint(0, 5), rng.randint(0, hosts - 1), f"{grp['groupName']}-abcde" if hosts > 1 else ""), **kw)
    assert cluster == before     # neither builder touches its input (the Go code does mutate the cached object's maps; the shim must not rely on that)


# ---- the one-call native builder (kr_pod_build) ---------------------------------------------------------------------------------------
def go_form(obj, key=None):
    """What a typed round trip through corev1 does to the Python assembly's dicts: empty slices vanish (omitempty), quantities print canonically."""
    if isinstance(obj, dict):
        out = {}
        for k, v in obj.items():
            if isinstance(v, list) and not v or v is None:
                continue
            if k in ("limits", "requests", "labels", "annotations", "nodeSelector") and v == {} and key != "metadata":
                continue                                   # an empty map under omitempty
            if k in ("limits", "requests") and isinstance(v, dict):
                out[k] = {rk: engine.quantity_canonical(str(rv)) for rk, rv in v.items()}
            elif k == "sizeLimit":
                out[k] = engine.quantity_canonical(str(v))
            else:
                out[k] = go_form(v, k)
        return out
    if isinstance(obj, list):
        return [go_form(v) for v in obj]
    return obj


def creates_of(cluster, rng):
    out = [HEAD]
    for gi, grp in enumerate(cluster["spec"]["workerGroupSpecs"]):
        hosts = int(grp.get("numOfHosts", 1))
        for _ in range(rng.randint(1, 3)):
            out.append((gi, rng.randint(0, 7), rng.randint(0, hosts - 1), f"{grp['groupName']}-abcde" if hosts > 1 else ""))
    rng.shuffle(out)
    return out


@pytest.mark.parametrize("seed", range(150))
def test_native_one_call_builder_agrees_with_the_assembly(seed):
    rng = random.Random(1000 + seed)
    cluster = random_cluster(rng)
    env = pb.BuilderEnv(login_shell=rng.random() < 0.2, init_container_injection=rng.random() < 0.85, probes_injection=rng.random() < 0.85,
                        deterministic_head_name=rng.random() < 0.3, multihost_indexing_gate=rng.random() < 0.7, cluster_domain=rng.choice(["cluster.local", "corp.example"]),
                        default_container_envs=rng.choice([{}, {"RAY_PORT": "1", "DEF": "d"}]),
                        head_sidecar_containers=rng.choice([[], [{"name": "fluentbit", "image": "fb:1"}]]), worker_sidecar_containers=rng.choice([[], [{"name": "w", "image": "w:1"}]]))
    creates = creates_of(cluster, rng)
    chash = rng.choice([None, "0123456789ABCDEFGHIJKLMNOPQRSTUV"])
    got = pb.build_pods_native(cluster, creates, env, cluster_hash=chash)
    assert len(got) == len(creates)
    for pod, create in zip(got, creates):
        assert pod.pop("status") == {}
        want = go_form(pb.build_pod(cluster, create, env, cluster_hash=chash))
        for c in want["spec"]["containers"] + want["spec"].get("initContainers", []):
            c.setdefault("resources", {})          # a struct value: encoding/json writes it even when empty
        assert pod == want, create


def test_native_builder_writes_go_field_order():
    """Bytes, not just the decoded value: TypeMeta first, corev1.PodSpec / Container fields in declaration order, maps sorted, alphabetical API-server input reordered."""
    import json
    cluster = instance()
    cluster["spec"]["enableInTreeAutoscaling"] = True
    # the API server serves keys alphabetically: feed the builder that order
    cluster = json.loads(json.dumps(cluster, sort_keys=True))
    raw_head, raw_worker = pb.build_pods_native(cluster, [HEAD, WORKER], raw=True)
    assert raw_head.startswith(b'{"kind":"Pod","apiVersion":"v1","metadata":{"generateName":"raycluster-sample-head-","namespace":"default","labels":{')
    assert raw_head.endswith(b',"status":{}}')
    head = json.loads(raw_head)
    assert list(head) == ["kind", "apiVersion", "metadata", "spec", "status"]
    assert list(head["spec"]) == ["volumes", "containers", "serviceAccountName"]
    assert list(head["spec"]["containers"][0]) == ["name", "image", "command", "args", "ports", "env", "resources", "volumeMounts", "livenessProbe", "readinessProbe"]
    assert list(head["spec"]["containers"][1]) == ["name", "image", "command", "args", "env", "resources", "volumeMounts", "imagePullPolicy"]
    assert list(head["spec"]["containers"][0]["livenessProbe"]) == ["exec", "initialDelaySeconds", "timeoutSeconds", "periodSeconds", "successThreshold", "failureThreshold"]
    worker = json.loads(raw_worker)
    assert list(worker["spec"]) == ["volumes", "initContainers", "containers"]
    assert list(worker["spec"]["initContainers"][0]) == ["name", "image", "command", "args", "env", "resources"]
    assert list(worker["spec"]["containers"][0]["resources"]["limits"]) == ["cpu", "memory", "nvidia.com/gpu"]
    assert b'"volumes":[{"name":"shared-mem","emptyDir":{"medium":"Memory","sizeLimit":"1Gi"}}]' in raw_worker
    # one call, many tuples of the same group: the container half is byte-identical, only the ObjectMeta differs
    many = pb.build_pods_native(cluster, [(0, i, 0, "") for i in range(50)], raw=True)
    tails = {m[m.index(b',"spec":'):] for m in many}
    assert len(tails) == 1 and len({m for m in many}) == 50


def test_native_builder_errors():
    from kuberay_b200.engine import EngineError
    cluster = instance()
    with pytest.raises(EngineError, match="worker group"):
        pb.build_pods_native(cluster, [(5, 0, 0, "")])
    broken = copy.deepcopy(cluster)
    broken["spec"]["headGroupSpec"]["template"]["spec"]["containers"] = []
    with pytest.raises(EngineError, match="Ray container"):
        pb.build_pods_native(broken, [HEAD])
    assert pb.build_pods_native(cluster, []) == []


@pytest.mark.parametrize("deterministic", [False, True])
def test_template_metadata_rides_along(deterministic):
    """ObjectMeta: podTemplateSpec.ObjectMeta (common/pod.go:598): whatever else the template's metadata carries stays; the worker's name is cleared
    (:418, TestDefaultWorkerPodTemplateWithName pod_test.go:1313-1327); the head gets name OR generateName and keeps the other as the template had it (:171-175)."""
    cluster = instance()
    for grp in (cluster["spec"]["headGroupSpec"], cluster["spec"]["workerGroupSpecs"][0]):
        grp["template"]["metadata"] = {"name": "from-template", "generateName": "tpl-", "namespace": "elsewhere", "finalizers": ["example.com/hold"],
                                       "labels": {"team": "a"}, "annotations": {"note": "x"}}
    head, worker = build_both(cluster, HEAD, deterministic_head_name=deterministic), build_both(cluster, WORKER, deterministic_head_name=deterministic)
    native = pb.build_pods_native(cluster, [HEAD, WORKER], pb.BuilderEnv(deterministic_head_name=deterministic))
    for got, want in zip(native, (head, worker)):
        order = ["name", "generateName", "namespace", "labels", "annotations", "ownerReferences", "finalizers"]      # metav1.ObjectMeta's declaration order
        assert got["metadata"] == want["metadata"] and list(got["metadata"]) == [k for k in order if k in got["metadata"]]
    assert worker["metadata"]["generateName"] == "raycluster-sample-small-group-worker-" and "name" not in worker["metadata"]
    assert worker["metadata"]["finalizers"] == ["example.com/hold"] and worker["metadata"]["namespace"] == "default" and worker["metadata"]["labels"]["team"] == "a"
    if deterministic:
        assert head["metadata"]["name"] == "raycluster-sample-head" and head["metadata"]["generateName"] == "tpl-"
    else:
        assert head["metadata"]["name"] == "from-template" and head["metadata"]["generateName"] == "raycluster-sample-head-"
