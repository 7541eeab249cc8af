"""The last Pod-builder pieces: GCS fault tolerance env, token auth, the autoscaler sidecar, the wait-gcs-ready init
container — native builders behind the C ABI (kuberay_b200/csrc/kr_raytemplate.cpp) against
  * the reference's own unit-test tables, transcribed (ray-operator/controllers/ray/common/pod_test.go: TestConfigureGCSFaultTolerance
    WithAnnotations :299-484, ...WithGcsFTOptions :486-640, TestBuildPod_WithEnableK8sTokenAuth :779-880, TestBuildPodWithAutoscalerOptions
    :1080-1155, TestHeadPodTemplate_AutoscalerImage :1236-1266, TestDefaultInitContainer[ImagePullPolicy] :1449-1525,
    TestSetAutoscalerV2EnvVars :2132-2186),
  * the object-level CPU restatement in oracle/podmeta.py (apply the native fragments to the template == what the restatement does),
  * the polling script as extracted byte for byte from the reference source (tests/golden/wait_gcs_ready_script.json)."""
import copy
import hashlib
import itertools
import json
import os
import random

import pytest

from kuberay_b200 import podmeta as pm
from kuberay_b200.engine import EngineError
from oracle import podmeta as ref

HERE = os.path.dirname(os.path.abspath(__file__))


def env_of(container, name):
    return next((e for e in container.get("env", []) if e["name"] == name), None)


def names(container):
    return [e["name"] for e in container.get("env", [])]


# ---- GCS fault tolerance ----------------------------------------------------------------------------------------------------------
def apply_ft(template, instance, node_type, ft_enabled):
    """configureGCSFaultTolerance through the native builder: its fragments appended to the template the way the shim would."""
    container = template["spec"]["containers"][0]
    got = pm.ray_ft_env(node_type, ft_enabled=ft_enabled, cluster_uid=instance.get("uid", ""),
                        storage_ns_annotation=(instance.get("annotations") or {}).get("ray.io/external-storage-namespace"),
                        options=instance["spec"].get("gcsFaultToleranceOptions"),
                        head_redis_password_param=instance["spec"]["headGroupSpec"].get("rayStartParams", {}).get("redis-password"),
                        existing=names(container))
    container.setdefault("env", []).extend(got["env"])
    instance["spec"]["headGroupSpec"].setdefault("rayStartParams", {}).update(got["rayStartParams"])
    return got


ANNOTATION_CASES = [  # pod_test.go:310-374: (name, storageNS, userEnv, passEnv, userParam, passParam, head, enabled)
    ("GCS FT enabled", "", "", "", "", "", True, True),
    ("GCS FT enabled with external storage", "test-ns", "", "", "", "", True, True),
    ("GCS FT enabled with redis password env", "", "", "test-password", "", "", True, True),
    ("GCS FT enabled with redis username and password env", "", "test-username", "test-password", "", "", True, True),
    ("GCS FT enabled with redis password ray start params", "", "", "", "", "test-password", True, True),
    ("GCS FT enabled with redis username and password ray start params", "", "", "", "test-username", "test-password", True, True),
    ("password env and params referring to env", "", "", "test-password", "", "$REDIS_PASSWORD", False, True),
    ("username and password env and params referring to env", "", "test-username", "test-password", "$REDIS_USERNAME", "$REDIS_PASSWORD", False, True),
    ("GCS FT enabled / worker Pod", "", "", "", "", "", False, True),
    ("GCS FT disabled", "", "", "", "", "", True, False),
]


@pytest.mark.parametrize("case", ANNOTATION_CASES, ids=[c[0] for c in ANNOTATION_CASES])
def test_gcs_ft_configured_by_annotation(case):
    _name, storage_ns, user_env, pass_env, user_param, pass_param, is_head, enabled = case
    head_env = [{"name": "RAY_REDIS_ADDRESS", "value": "redis:6379"}]
    if user_env:
        head_env.append({"name": "REDIS_USERNAME", "value": user_env})
    if pass_env:
        head_env.append({"name": "REDIS_PASSWORD", "value": pass_env})
    params = {}
    if user_param:
        params["redis-username"] = user_param
    if pass_param:
        params["redis-password"] = pass_param
    instance = {"uid": "uid-1234", "annotations": {"ray.io/ft-enabled": "true" if enabled else "false"},
                "spec": {"headGroupSpec": {"rayStartParams": params}}}
    if storage_ns:
        instance["annotations"]["ray.io/external-storage-namespace"] = storage_ns
    template = {"spec": {"containers": [{"env": head_env if is_head else []}]}}
    want_t, want_i = copy.deepcopy(template), copy.deepcopy(instance)
    ref.configure_gcs_fault_tolerance(want_t, want_i, "head" if is_head else "worker", enabled)
    apply_ft(template, instance, "head" if is_head else "worker", enabled)
    container = template["spec"]["containers"][0]
    assert container == want_t["spec"]["containers"][0] and instance["spec"] == want_i["spec"]
    if is_head:  # the reference test's assertions
        assert env_of(container, "RAY_gcs_rpc_server_reconnect_timeout_s") is None
        if storage_ns:
            assert env_of(container, "RAY_external_storage_namespace")["value"] == storage_ns
        if enabled and not storage_ns:
            assert env_of(container, "RAY_external_storage_namespace")["value"] == "uid-1234"
        if user_env:
            assert env_of(container, "REDIS_USERNAME")["value"] == user_env
        if pass_env:
            assert env_of(container, "REDIS_PASSWORD")["value"] == pass_env
        elif pass_param:
            assert env_of(container, "REDIS_PASSWORD")["value"] == pass_param
        if not enabled:
            assert names(container) == ["RAY_REDIS_ADDRESS"]
    else:  # assertWorkerGCSFaultToleranceConfig
        assert env_of(container, "RAY_gcs_rpc_server_reconnect_timeout_s")["value"] == "600"
        for n in ("RAY_external_storage_namespace", "RAY_REDIS_ADDRESS", "REDIS_PASSWORD", "REDIS_USERNAME"):
            assert env_of(container, n) is None


FIELD = lambda path: {"fieldRef": {"fieldPath": path}}  # noqa: E731
OPTION_CASES = [  # pod_test.go:492-573
    ("GCS FT enabled", {"redisAddress": "redis:6379"}, True),
    ("redis password", {"redisAddress": "redis:6379", "redisPassword": {"value": "test-password"}}, True),
    ("redis username and password", {"redisAddress": "redis:6379", "redisUsername": {"value": "test-username"}, "redisPassword": {"value": "test-password"}}, True),
    ("redis password in secret", {"redisAddress": "redis:6379", "redisPassword": {"valueFrom": FIELD("spec.redisPassword")}}, True),
    ("redis username and password in secret", {"redisAddress": "redis:6379", "redisUsername": {"valueFrom": FIELD("spec.redisUsername")},
                                               "redisPassword": {"valueFrom": FIELD("spec.redisPassword")}}, True),
    ("external storage namespace", {"redisAddress": "redis:6379", "externalStorageNamespace": "test-ns"}, True),
    ("worker Pod", {"redisAddress": "redis:6379"}, False),
]


@pytest.mark.parametrize("case", OPTION_CASES, ids=[c[0] for c in OPTION_CASES])
def test_gcs_ft_configured_by_options(case):
    _name, options, is_head = case
    instance = {"uid": "", "spec": {"gcsFaultToleranceOptions": options, "headGroupSpec": {"rayStartParams": {}}}}
    template = {"spec": {"containers": [{"env": []}]}}
    want_t, want_i = copy.deepcopy(template), copy.deepcopy(instance)
    ref.configure_gcs_fault_tolerance(want_t, want_i, "head" if is_head else "worker", True)
    apply_ft(template, instance, "head" if is_head else "worker", True)
    container = template["spec"]["containers"][0]
    assert container == want_t["spec"]["containers"][0] and instance["spec"] == want_i["spec"]
    if not is_head:
        assert names(container) == ["RAY_gcs_rpc_server_reconnect_timeout_s"]
        return
    assert env_of(container, "RAY_gcs_rpc_server_reconnect_timeout_s") is None
    assert env_of(container, "RAY_REDIS_ADDRESS")["value"] == "redis:6379"
    for key, env_name, param in (("redisUsername", "REDIS_USERNAME", "redis-username"), ("redisPassword", "REDIS_PASSWORD", "redis-password")):
        if key in options:
            e = env_of(container, env_name)
            assert e.get("value", "") == options[key].get("value", "") and e.get("valueFrom") == options[key].get("valueFrom")
            assert instance["spec"]["headGroupSpec"]["rayStartParams"][param] == "$" + env_name
    if options.get("externalStorageNamespace"):
        assert env_of(container, "RAY_external_storage_namespace")["value"] == "test-ns"


def test_gcs_ft_namespace_precedence_and_existing_env():
    """UID < annotation < option (:107-113); a namespace env the user already set is kept (:115-118); an EMPTY annotation still wins over the UID."""
    base = dict(ft_enabled=True, cluster_uid="the-uid")
    val = lambda r: next(e.get("value", "") for e in r["env"] if e["name"] == "RAY_external_storage_namespace")  # noqa: E731
    assert val(pm.ray_ft_env("head", **base)) == "the-uid"
    assert val(pm.ray_ft_env("head", storage_ns_annotation="ann", **base)) == "ann"
    assert val(pm.ray_ft_env("head", storage_ns_annotation="ann", options={"redisAddress": "r", "externalStorageNamespace": "opt"}, **base)) == "opt"
    assert val(pm.ray_ft_env("head", storage_ns_annotation="ann", options={"redisAddress": "r", "externalStorageNamespace": ""}, **base)) == "ann"
    assert val(pm.ray_ft_env("head", storage_ns_annotation="", **base)) == ""
    assert pm.ray_ft_env("head", existing=["RAY_external_storage_namespace"], **base)["env"] == []
    assert pm.ray_ft_env("worker", existing=["RAY_gcs_rpc_server_reconnect_timeout_s"], **base)["env"] == []
    assert pm.ray_ft_env("head", ft_enabled=False, options=None) == {"env": [], "rayStartParams": {}}
    with pytest.raises(EngineError):
        pm.ray_ft_env("head", ft_enabled=True, options={"redisAddress": "r", "redisPassword": {"valueFrom": "[1]"}})


# ---- token auth -------------------------------------------------------------------------------------------------------------------
def apply_auth(cluster_name, template, auth_options):
    """configureTokenAuth through the native builder."""
    spec = template["spec"]
    k8s = ref.is_k8s_auth_enabled(auth_options)
    targets = [spec["containers"][0]] + [c for c in spec.get("initContainers") or [] if c.get("name") == "wait-gcs-ready"]
    for c in targets:
        got = pm.ray_auth(cluster_name, k8s_token_auth=k8s, secret_name=(auth_options or {}).get("secretName"), existing_env=names(c),
                          existing_mount_names=[m["name"] for m in c.get("volumeMounts", [])], existing_volume_names=[v["name"] for v in spec.get("volumes", [])])
        c.setdefault("env", []).extend(got["env"])
        if got["volumeMounts"]:
            c.setdefault("volumeMounts", []).extend(got["volumeMounts"])
        if got["volumes"]:
            spec.setdefault("volumes", []).extend(got["volumes"])


@pytest.mark.parametrize("k8s", [True, False, None])
def test_token_auth_on_the_head_template(k8s):
    """TestBuildPod_WithEnableK8sTokenAuth (:779-844)."""
    auth = {"mode": "token", "enableK8sTokenAuth": k8s}
    template = {"spec": {"containers": [{"name": "ray-head", "env": [{"name": "TEST_ENV_NAME", "value": "TEST_ENV_VALUE"}]}]}}
    want = copy.deepcopy(template)
    ref.configure_token_auth("raycluster-sample", want, auth)
    apply_auth("raycluster-sample", template, auth)
    assert template == want
    ray = template["spec"]["containers"][0]
    assert env_of(ray, "RAY_AUTH_MODE")["value"] == "token"
    if k8s:
        assert env_of(ray, "RAY_ENABLE_K8S_TOKEN_AUTH")["value"] == "true" and env_of(ray, "RAY_AUTH_TOKEN") is None
        assert {"name": "ray-token", "readOnly": True, "mountPath": "/var/run/secrets/ray.io/serviceaccount"} in ray["volumeMounts"]
        assert any(v["name"] == "ray-token" and "projected" in v for v in template["spec"]["volumes"])
    else:
        assert env_of(ray, "RAY_ENABLE_K8S_TOKEN_AUTH") is None and "volumes" not in template["spec"] and "volumeMounts" not in ray
        assert env_of(ray, "RAY_AUTH_TOKEN")["valueFrom"] == {"secretKeyRef": {"name": "raycluster-sample", "key": "auth_token"}}


def test_token_auth_reaches_the_init_container_and_adds_one_volume():
    """TestBuildPod_WithEnableK8sTokenAuth_InitContainer (:846-880): the mount lands on wait-gcs-ready too; the pod gets ONE token volume."""
    auth = {"mode": "token", "enableK8sTokenAuth": True}
    template = {"spec": {"containers": [{"name": "ray-worker"}], "initContainers": [{"name": "other"}, {"name": "wait-gcs-ready", "env": [{"name": "RAY_AUTH_MODE", "value": "token"}]}]}}
    want = copy.deepcopy(template)
    ref.configure_token_auth("c", want, auth)
    apply_auth("c", template, auth)
    assert template == want
    init = template["spec"]["initContainers"][1]
    assert any(m["name"] == "ray-token" and m["readOnly"] for m in init["volumeMounts"]) and names(init).count("RAY_AUTH_MODE") == 1
    assert len(template["spec"]["volumes"]) == 1 and "env" not in template["spec"]["initContainers"][0]


def test_token_auth_secret_name_rules():
    tok = lambda **kw: next(e for e in pm.ray_auth(**kw)["env"] if e["name"] == "RAY_AUTH_TOKEN")["valueFrom"]["secretKeyRef"]["name"]  # noqa: E731
    assert tok(cluster_name="my-cluster", secret_name="custom") == "custom"
    assert tok(cluster_name="my-cluster", secret_name="") == "my-cluster"
    long = "9" + "x" * 70
    assert tok(cluster_name=long) == ref.check_name(long.encode()).decode() and len(tok(cluster_name=long)) == 50
    assert pm.ray_auth("c", existing_env=["RAY_AUTH_MODE", "RAY_AUTH_TOKEN"]) == {"env": [], "volumeMounts": [], "volumes": []}
    assert pm.ray_auth("c", k8s_token_auth=True, existing_env=["RAY_ENABLE_K8S_TOKEN_AUTH"], existing_mount_names=["ray-token"], existing_volume_names=["ray-token"]) == \
        {"env": [{"name": "RAY_AUTH_MODE", "value": "token"}], "volumeMounts": [], "volumes": []}


# ---- autoscaler sidecar -----------------------------------------------------------------------------------------------------------
def apply_autoscaler(instance, template, login_shell=False):
    spec = template["spec"]
    auth = instance["spec"].get("authOptions")
    options = instance["spec"].get("autoscalerOptions")
    got = pm.ray_autoscaler_container(instance["name"], spec["containers"][0].get("image", ""), options=options,
                                      autoscaler_v2=bool(options and options.get("version") == "v2"),
                                      auth_enabled=bool(auth and auth.get("mode") == "token"), k8s_token_auth=ref.is_k8s_auth_enabled(auth),
                                      secret_name=(auth or {}).get("secretName"), head_service_account=spec.get("serviceAccountName"), login_shell=login_shell)
    spec["serviceAccountName"] = got["serviceAccountName"]
    spec["containers"].append(got["container"])
    if got["rayContainerEnv"]:
        spec["containers"][0].setdefault("env", []).extend(got["rayContainerEnv"])
    if got["restartPolicy"]:
        spec["restartPolicy"] = got["restartPolicy"]
    return got


CUSTOM_OPTIONS = {  # TestBuildPodWithAutoscalerOptions (:1086-1141)
    "upscalingMode": "Aggressive", "idleTimeoutSeconds": 100, "image": "custom-autoscaler-xxx", "imagePullPolicy": "IfNotPresent",
    "resources": {"limits": {"cpu": "1", "memory": "1Gi"}, "requests": {"cpu": "1", "memory": "1Gi"}},
    "env": [{"name": "fooEnv", "value": "fooValue"}], "envFrom": [{"prefix": "Pre"}],
    "volumeMounts": [{"name": "ca-tls", "readOnly": True, "mountPath": "/etc/ca/tls"}, {"name": "ray-tls", "mountPath": "/etc/ray/tls"}],
    "securityContext": {"capabilities": {"drop": ["ALL"]}, "runAsNonRoot": True, "allowPrivilegeEscalation": False, "seccompProfile": {"type": "RuntimeDefault"}},
}


def test_autoscaler_container_defaults():
    """BuildAutoscalerContainer (:673-724) as TestBuildPod_WithAutoscalerEnabled / TestHeadPodTemplate_AutoscalerImage see it."""
    instance = {"name": "raycluster-sample", "spec": {}}
    template = {"spec": {"containers": [{"name": "ray-head", "image": "repo/image:custom"}]}}
    want = copy.deepcopy(template)
    ref.head_autoscaler_sidecar(instance, want)
    got = apply_autoscaler(instance, template)
    assert template == want
    c = template["spec"]["containers"][1]
    assert c["name"] == "autoscaler" and c["image"] == "repo/image:custom" and c["imagePullPolicy"] == "IfNotPresent"
    assert c["command"] == ["/bin/bash", "-c", "--"]
    assert c["args"] == ["ray kuberay-autoscaler --cluster-name $(RAY_CLUSTER_NAME) --cluster-namespace $(RAY_CLUSTER_NAMESPACE)"]
    assert names(c) == ["RAY_CLUSTER_NAME", "RAY_CLUSTER_NAMESPACE", "RAY_HEAD_POD_NAME", "KUBERAY_CRD_VER"]
    assert c["resources"] == {"limits": {"cpu": "500m", "memory": "512Mi"}, "requests": {"cpu": "500m", "memory": "512Mi"}}
    assert list(c) == ["name", "image", "command", "args", "env", "resources", "imagePullPolicy"]          # corev1.Container's field order
    assert template["spec"]["serviceAccountName"] == "raycluster-sample" and got["restartPolicy"] == "" and got["rayContainerEnv"] == []


def test_autoscaler_container_overrides():
    """TestBuildPodWithAutoscalerOptions (:1080-1155): every override lands; env / mounts are appended after the built-in ones."""
    instance = {"name": "raycluster-sample", "spec": {"autoscalerOptions": CUSTOM_OPTIONS}}
    template = {"spec": {"containers": [{"name": "ray-head", "image": "repo/image:custom"}], "serviceAccountName": "head-service-account"}}
    want = copy.deepcopy(template)
    ref.head_autoscaler_sidecar(instance, want)
    apply_autoscaler(instance, template)
    assert template == want
    c = template["spec"]["containers"][1]
    assert c["image"] == "custom-autoscaler-xxx" and c["resources"] == CUSTOM_OPTIONS["resources"] and c["envFrom"] == [{"prefix": "Pre"}]
    assert names(c)[-1] == "fooEnv" and len(c["env"]) == 5 and c["volumeMounts"] == CUSTOM_OPTIONS["volumeMounts"] and c["securityContext"] == CUSTOM_OPTIONS["securityContext"]
    assert list(c) == ["name", "image", "command", "args", "envFrom", "env", "resources", "volumeMounts", "imagePullPolicy", "securityContext"]
    assert template["spec"]["serviceAccountName"] == "head-service-account"      # TestHeadPodTemplate_WithServiceAccount (:1290-1311)


def test_autoscaler_v2_auth_and_login_shell():
    """setAutoscalerV2EnvVars (:242-251, TestSetAutoscalerV2EnvVars :2132-2186), restartPolicy Never (:216-219), token auth on the sidecar
    BEFORE the overrides (:207-213), ENABLE_LOGIN_SHELL."""
    instance = {"name": "c", "spec": {"autoscalerOptions": {"version": "v2", "env": [{"name": "RAY_AUTH_MODE", "value": "user"}], "volumeMounts": [{"name": "m", "mountPath": "/m"}]},
                                      "authOptions": {"mode": "token", "enableK8sTokenAuth": True}}}
    template = {"spec": {"containers": [{"name": "ray-head", "image": "img", "env": [{"name": "A", "value": "b"}]}]}}
    want = copy.deepcopy(template)
    ref.head_autoscaler_sidecar(instance, want, login_shell=True)
    got = apply_autoscaler(instance, template, login_shell=True)
    assert template == want
    c = template["spec"]["containers"][1]
    assert c["command"] == ["/bin/bash", "-cl", "--"]
    assert names(c) == ["RAY_CLUSTER_NAME", "RAY_CLUSTER_NAMESPACE", "RAY_HEAD_POD_NAME", "KUBERAY_CRD_VER", "RAY_AUTH_MODE", "RAY_ENABLE_K8S_TOKEN_AUTH", "RAY_AUTH_MODE"]
    assert [m["name"] for m in c["volumeMounts"]] == ["ray-token", "m"]
    assert template["spec"]["containers"][0]["env"][-1] == {"name": "RAY_enable_autoscaler_v2", "value": "true"} and template["spec"]["restartPolicy"] == "Never"
    assert got["serviceAccountName"] == "c"
    # secret-based auth on the sidecar; an image override that is SET to "" empties the field (a nil pointer would not)
    g = pm.ray_autoscaler_container("clu", "img", auth_enabled=True, options={"image": ""})
    assert env_of(g["container"], "RAY_AUTH_TOKEN")["valueFrom"]["secretKeyRef"]["name"] == "clu" and "image" not in g["container"]
    with pytest.raises(EngineError):
        pm.ray_autoscaler_container("clu", "img", options={"env": "{}"})


# ---- the wait-gcs-ready init container --------------------------------------------------------------------------------------------
def test_init_script_is_the_reference_literal():
    """The script text against the bytes extracted from the reference source (tests/golden/gen_init_script.py)."""
    doc = json.load(open(os.path.join(HERE, "golden", "wait_gcs_ready_script.json")))
    assert hashlib.sha256(doc["format"].encode()).hexdigest() == doc["sha256"] and doc["format"].count("%s") == 4
    fqdn, port = "raycluster-sample-head-svc.default.svc.cluster.local", "6379"
    want = doc["format"].replace("%s:%s", f"{fqdn}:{port}")
    assert ref.wait_gcs_ready_script(fqdn, port) == want
    assert pm.ray_init_container("img", fqdn, port)["args"] == [want]


@pytest.mark.parametrize("pull", ["Always", "IfNotPresent", "Never", None])
def test_init_container_copies_the_ray_container(pull):
    """TestDefaultInitContainer / ...ImagePullPolicy (:1449-1525): env (values and fieldRefs), mounts, security context and the pull policy
    of the Ray container; fixed small resources; nothing else."""
    ray = {"name": "ray-worker", "image": "repo/image:custom", "command": ["echo"], "args": ["hi"],
           "env": [{"name": "TEST_ENV_NAME", "value": "TEST_ENV_VALUE"}, {"name": "MY_POD_IP", "valueFrom": {"fieldRef": {"fieldPath": "status.podIP"}}}],
           "resources": {"limits": {"cpu": "1", "memory": "1Gi", "nvidia.com/gpu": "3"}}, "volumeMounts": [{"name": "tls", "readOnly": True, "mountPath": "/etc/tls"}],
           "securityContext": {"runAsUser": 1000, "allowPrivilegeEscalation": False}}
    if pull:
        ray["imagePullPolicy"] = pull
    want = ref.wait_gcs_ready_container(ray, "svc.ns.svc.cluster.local", "6380")
    got = pm.ray_init_container(ray["image"], "svc.ns.svc.cluster.local", "6380", image_pull_policy=pull, env=ray["env"], volume_mounts=ray["volumeMounts"],
                                security_context=ray["securityContext"])
    assert got == want and list(got) == list(want)
    assert got["env"] == ray["env"] and got.get("imagePullPolicy") == pull
    assert got["resources"] == {"limits": {"cpu": "200m", "memory": "256Mi"}, "requests": {"cpu": "200m", "memory": "256Mi"}}
    assert "svc.ns.svc.cluster.local:6380 > /dev/null" in got["args"][0] and got["command"] == ["/bin/bash", "-c", "--"]
    bare = pm.ray_init_container("i", "f", "1", env=[], volume_mounts=None, security_context=None, login_shell=True)
    assert list(bare) == ["name", "image", "command", "args", "resources"] and bare["command"][1] == "-cl"


# ---- randomised agreement with the restatement -----------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(60))
def test_random_templates_agree_with_the_restatement(seed):
    rng = random.Random(seed)
    pick = lambda *xs: rng.choice(xs)  # noqa: E731
    cred = lambda n: pick(None, {"value": n}, {"valueFrom": FIELD("spec." + n)}, {"value": n, "valueFrom": {"secretKeyRef": {"name": "s\"q", "key": "k"}}})  # noqa: E731
    options = pick(None, {"redisAddress": pick("", "redis:6379"), "externalStorageNamespace": pick("", "opt-ns"), "redisUsername": cred("u"), "redisPassword": cred("p")})
    if options:
        options = {k: v for k, v in options.items() if v is not None}
    pool = ["RAY_external_storage_namespace", "RAY_gcs_rpc_server_reconnect_timeout_s", "REDIS_PASSWORD", "RAY_AUTH_MODE", "RAY_AUTH_TOKEN", "RAY_ENABLE_K8S_TOKEN_AUTH", "X"]
    env = [{"name": n, "value": "v"} for n in rng.sample(pool, rng.randint(0, 4))]
    params = pick({}, {"redis-password": pick("", "pw")})
    ann = pick({}, {"ray.io/external-storage-namespace": pick("", "ann-ns")})
    auth = pick(None, {"mode": "token"}, {"mode": "token", "enableK8sTokenAuth": pick(True, False)}, {"mode": "token", "secretName": pick("", "sec")})
    auto = pick(None, {}, {"version": pick("v1", "v2"), "image": pick("", "auto:1"), "env": pick([], [{"name": "E", "value": "é <"}]),
                           "envFrom": pick([], [{"prefix": "P"}]), "volumeMounts": pick([], [{"name": "ray-token", "mountPath": "/x"}]), "securityContext": pick(None, {"privileged": True})})
    if auto:
        auto = {k: v for k, v in auto.items() if v is not None}
    instance = {"name": pick("c", "9cluster", "a" * 60), "uid": "u-1", "annotations": ann,
                "spec": {"gcsFaultToleranceOptions": options, "authOptions": auth, "autoscalerOptions": auto, "headGroupSpec": {"rayStartParams": params}}}
    node = pick("head", "worker")
    ft = options is not None or rng.random() < 0.5
    template = {"spec": {"containers": [{"name": "ray", "image": pick("", "ray:2.9"), "env": env}], "volumes": pick([], [{"name": "ray-token", "emptyDir": {}}])}}
    if rng.random() < 0.5:
        template["spec"]["serviceAccountName"] = pick("", "sa-1")
    if node == "worker":
        template["spec"]["initContainers"] = [ref.wait_gcs_ready_container(template["spec"]["containers"][0], "f.q.d.n", "6379")]
        got_init = pm.ray_init_container(template["spec"]["containers"][0]["image"], "f.q.d.n", "6379", env=env)
        assert got_init == template["spec"]["initContainers"][0]
    want_t, want_i = copy.deepcopy(template), copy.deepcopy(instance)
    # the reference's order: autoscaler sidecar (head, :194-220), GCS FT (:222 / :443), token auth (:234-236 / :459-461)
    if node == "head" and auto is not None:
        ref.head_autoscaler_sidecar(want_i, want_t)
        apply_autoscaler(instance, template)
    ref.configure_gcs_fault_tolerance(want_t, want_i, node, ft)
    apply_ft(template, instance, node, ft)
    if auth is not None:
        ref.configure_token_auth(want_i["name"], want_t, auth)
        apply_auth(instance["name"], template, auth)
    want_t.pop("metadata", None)   # the two annotations belong to kr_pod_meta_build (tests/test_podmeta.py)
    assert template == want_t and instance["spec"] == want_i["spec"]
