"""`ray start` command builder (kr_ray_start_command, kuberay_b200/csrc/kr_raystart.cpp; SURVEY §8 f3, second part) — host-side, no GPU.

1. the reference's own tables, transcribed, through the NATIVE path one step at a time (kr_raystart_in.steps) and through the CPU
   restatement (oracle/podmeta.py): common/pod_test.go:1986-2131 (TestGenerateRayStartCommand), :2229-2283
   (TestUpdateRayStartParamsLabels), :2285-2361 (TestUpdateRayStartParamsResources), :1526-1771 (TestSetMissingRayStartParams*),
   :258-297 (TestGetHeadPort), :926-955 (TestBuildPod_WithOverwriteCommand), :1014-1078 (TestBuildPod_WithLoginBash),
   :882-924 (TestBuildPod_WithNoCPULimits)
2. resource.Quantity and encoding/json float formatting cases
3. fuzz: native composed step == restatement, byte for byte on the JSON it writes"""
import json
import random

import pytest

from kuberay_b200 import abi
from kuberay_b200 import podmeta as pm
from kuberay_b200.engine import EngineError
from oracle import podmeta as ref

GEN = abi.RS_GENERATE

GENERATE_CASES = [  # pod_test.go:1994-2122: (node type, rayStartParams, limits, expected)
    ("worker", {}, {"nvidia.com/gpu": "1"}, "ray start  --num-gpus=1 "),
    ("worker", {}, {"nvidia.com/mig-2g.32gb": "1"}, "ray start  --num-gpus=1 "),
    ("worker", {}, {"google.com/tpu": "4"}, """ray start  --resources='{"TPU":4}' """),
    ("head", {}, {"aws.amazon.com/neuroncore": "4"}, """ray start --head  --resources='{"neuron_cores":4}' """),
    ("head", {}, {"aws.amazon.com/neuroncore": "4", "nvidia.com/gpu": "1"}, """ray start --head  --num-gpus=1  --resources='{"neuron_cores":4}' """),
    ("head", {}, {"google.com/tpu": "8", "aws.amazon.com/neuroncore": "4", "nvidia.com/gpu": "1"}, """ray start --head  --num-gpus=1  --resources='{"neuron_cores":4}' """),
    ("head", {"resources": '"{"custom_resource":2}"'}, {"aws.amazon.com/neuroncore": "4"}, """ray start --head  --resources='{"custom_resource":2,"neuron_cores":4}' """),
    ("head", {"resources": """'{"custom_resource":2,"neuron_cores":3}'"""}, {"aws.amazon.com/neuroncore": "4"}, """ray start --head  --resources='{"custom_resource":2,"neuron_cores":3}' """),
    ("head", {"resources": """'{"custom_resource":2,"TPU":4}'"""}, {"google.com/tpu": "8"}, """ray start --head  --resources='{"custom_resource":2,"TPU":4}' """),
    ("head", {"resources": "{"}, {"aws.amazon.com/neuroncore": "4"}, "ray start --head  --resources={ "),
]


@pytest.mark.parametrize("node_type,params,limits,want", GENERATE_CASES)
def test_generate_ray_start_command_vectors(node_type, params, limits, want):
    got = pm.ray_start_command(node_type, params, limits=limits, steps=GEN)
    assert got["rayStartCommand"] == want
    assert ref.generate_ray_start_command(node_type, dict(params), limits, None) == want


def test_invalid_node_type_is_refused():
    assert ref.generate_ray_start_command("InvalidType", {}, None, None) == ""   # pod_test.go:2116-2121: the reference returns ""
    with pytest.raises(EngineError):                                            # the C ABI refuses the call instead
        pm.ray_start_command("InvalidType", {})


LABEL_CASES = [  # pod_test.go:2235-2268
    ({}, {"topology.kubernetes.io/zone": "us-central2", "ray.io/node-group": "worker-group-1", "cloud.google.com/gke-spot": "true"},
     {"labels": "cloud.google.com/gke-spot=true,ray.io/node-group=worker-group-1,topology.kubernetes.io/zone=us-central2"}),
    ({"labels": "old=label,to-be=replaced", "resources": "some-resources"}, {"new": "label"}, {"labels": "new=label", "resources": "some-resources"}),
    ({"labels": "some=labels"}, None, {"labels": "some=labels"}),
    ({"labels": "some=labels"}, {}, {"labels": "some=labels"}),
]


@pytest.mark.parametrize("initial,labels,want", LABEL_CASES)
def test_update_ray_start_params_labels_vectors(initial, labels, want):
    assert pm.ray_start_command("worker", initial, group_labels=labels, steps=abi.RS_UPDATE_LABELS)["rayStartParams"] == want
    p = dict(initial)
    ref.update_ray_start_params_labels(p, labels)
    assert p == want


RESOURCE_CASES = [  # pod_test.go:2294-2345
    ({"existing": "true"}, None, {"existing": "true"}),
    ({}, {"cpu": "2", "memory": "4Gi"}, {"num-cpus": "2", "memory": "4294967296"}),
    ({}, {"CPU": "2", "GPU": "4"}, {"num-cpus": "2", "num-gpus": "4"}),
    ({}, {"nvidia.com/gpu": "1", "TPU": "4"}, {"num-gpus": "1", "resources": "'{\"TPU\":4}'"}),
    ({"num-cpus": "1", "memory": "1000", "resources": "'{\"Custom-Resource\": 10}'"}, {"cpu": "4", "Custom-Resource": "5"},
     {"num-cpus": "4", "memory": "1000", "resources": "'{\"Custom-Resource\":5}'"}),
]


@pytest.mark.parametrize("initial,resources,want", RESOURCE_CASES)
def test_update_ray_start_params_resources_vectors(initial, resources, want):
    assert pm.ray_start_command("worker", initial, group_resources=resources, steps=abi.RS_UPDATE_RESOURCES)["rayStartParams"] == want
    p = dict(initial)
    ref.update_ray_start_params_resources(p, resources)
    assert p == want


def test_set_missing_ray_start_params_vectors():
    """pod_test.go:1526-1771: address only for workers (<fqdn>:<head port>) and never overwritten; metrics-export-port default 8080 kept
    when given; block always "true" (even when the user said false); dashboard-host 0.0.0.0 for the head only, kept when given."""
    fqdn, custom = "raycluster-kuberay-head-svc.default.svc.cluster.local", "custom-address:1234"
    sm = abi.RS_SET_MISSING
    head = pm.ray_start_command("head", {}, steps=sm)["rayStartParams"]
    assert "address" not in head and head["dashboard-host"] == "0.0.0.0" and head["metrics-export-port"] == "8080" and head["block"] == "true"
    assert head["dashboard-agent-listen-port"] == "52365"
    assert pm.ray_start_command("head", {"address": custom}, steps=sm)["rayStartParams"]["address"] == custom
    w = pm.ray_start_command("worker", {}, fqdn_ray_ip=fqdn, head_port="6379", steps=sm)["rayStartParams"]
    assert w["address"] == f"{fqdn}:6379" and "dashboard-host" not in w
    assert pm.ray_start_command("worker", {"address": custom}, fqdn_ray_ip=fqdn, head_port="6379", steps=sm)["rayStartParams"]["address"] == custom
    for nt in ("head", "worker"):
        assert pm.ray_start_command(nt, {"metrics-export-port": "9999"}, steps=sm)["rayStartParams"]["metrics-export-port"] == "9999"
        assert pm.ray_start_command(nt, {"block": "false"}, steps=sm)["rayStartParams"]["block"] == "true"
    assert pm.ray_start_command("head", {"dashboard-host": "localhost"}, steps=sm)["rayStartParams"]["dashboard-host"] == "localhost"
    assert "dashboard-host" not in pm.ray_start_command("worker", {}, steps=sm)["rayStartParams"]
    for nt, kw in (("head", {}), ("worker", {"fqdn_ray_ip": fqdn, "head_port": "6379"})):
        p = {}
        ref.set_missing_ray_start_params(p, nt, kw.get("head_port", "6379"), kw.get("fqdn_ray_ip", ""))
        assert p == pm.ray_start_command(nt, {}, steps=sm, **kw)["rayStartParams"]
    # GetHeadPort (pod_test.go:258-297): the head's own "port" parameter, else 6379 — the worker dials what it is given
    assert pm.ray_start_command("worker", {}, fqdn_ray_ip="svc", head_port="9999", steps=sm)["rayStartParams"]["address"] == "svc:9999"
    assert pm.ray_start_command("worker", {}, fqdn_ray_ip="svc", steps=sm)["rayStartParams"]["address"] == "svc:6379"


def test_container_command_assembly():
    """BuildPod (common/pod.go:617-650)."""
    # generated: /bin/bash -c -- "ulimit -n 65536; ray start ..."
    r = pm.ray_start_command("head", {}, limits={"cpu": "1", "memory": "1Gi"})
    assert r["generated"] and r["command"] == ["/bin/bash", "-c", "--"]
    assert r["args"] == ["ulimit -n 65536; " + r["rayStartCommand"]] and r["rayStartCommand"].startswith("ray start --head ")
    assert " --num-cpus=1 " in r["rayStartCommand"] and " --memory=1073741824 " in r["rayStartCommand"] and " --block " in r["rayStartCommand"]
    # a user command goes in front (pod_test.go:1014-1078 also turns the login shell on: -cl)
    r = pm.ray_start_command("worker", {}, command=["echo", "hi"], args=["there"], fqdn_ray_ip="svc", login_shell=True)
    assert r["command"] == ["/bin/bash", "-cl", "--"] and r["args"] == [" echo  hi  there  && ulimit -n 65536; " + r["rayStartCommand"]]
    # overwrite annotation: the template's command / args stay (pod_test.go:926-955)
    r = pm.ray_start_command("head", {}, command=["I am head"], args=["I am head again"], overwrite_cmd=True)
    assert not r["generated"] and r["command"] == ["I am head"] and r["args"] == ["I am head again"]
    # a command that already runs "ray start" is left alone
    r = pm.ray_start_command("worker", {}, command=["/bin/sh", "-c"], args=["ray start --address=x --block"])
    assert not r["generated"] and r["args"] == ["ray start --address=x --block"]
    # no CPU limit: the request is used (pod_test.go:882-924); neither: no --num-cpus
    assert " --num-cpus=2 " in pm.ray_start_command("worker", {}, requests={"cpu": "1500m"})["rayStartCommand"]
    assert "--num-cpus" not in pm.ray_start_command("worker", {}, limits={"memory": "1Gi"})["rayStartCommand"]
    assert " --num-cpus=7 " in pm.ray_start_command("worker", {"num-cpus": "7"}, limits={"cpu": "1"})["rayStartCommand"]
    # the head under the autoscaler does not run the monitor (common/pod.go:196-200)
    assert " --no-monitor " in pm.ray_start_command("head", {}, autoscaling=True)["rayStartCommand"]
    assert "no-monitor" not in pm.ray_start_command("worker", {}, autoscaling=True)["rayStartParams"]
    # booleans: true -> bare flag, false -> dropped, except the two options whose argument may be true / false
    r = pm.ray_start_command("head", {"include-dashboard": "false", "log-color": "True", "verbose": "TRUE", "quiet": "False"}, steps=GEN)
    assert r["rayStartCommand"] == "ray start --head  --include-dashboard=false  --log-color=True  --verbose "


@pytest.mark.parametrize("text,value,approx,zero", [
    ("1", 1, 1.0, False), ("0", 0, 0.0, True), ("500m", 1, 0.5, False), ("1500m", 2, 1.5, False), ("4Gi", 4294967296, 4294967296.0, False),
    ("1.5Gi", 1610612736, 1610612736.0, False), ("2k", 2000, 2000.0, False), ("1e3", 1000, 1000.0, False), ("1E", 10 ** 18, 1e18, False),
    ("100n", 1, 1e-7, False), ("0.0", 0, 0.0, True), (".5", 1, 0.5, False), ("5.", 5, 5.0, False), ("12E-1", 2, 1.2, False), ("+3", 3, 3.0, False),
])
def test_quantity_value(text, value, approx, zero):
    v, f, z = pm.quantity_value(text)
    assert (v, z) == (value, zero) and f == pytest.approx(approx, rel=1e-12)
    q = ref.parse_quantity(text)
    assert ref.quantity_value(q) == value and ref.quantity_float(text) == f


@pytest.mark.parametrize("text", ["", "abc", "1K", "1Kii", "1 Gi", "--1", "1e", "1.2.3", "Gi", "0x10"])
def test_not_a_quantity(text):
    assert ref.parse_quantity(text) is None
    with pytest.raises(EngineError):
        pm.quantity_value(text)


def test_float_formatting_like_encoding_json():
    for x, want in ((4.0, "4"), (0.5, "0.5"), (1e21, "1e+21"), (1e20, "100000000000000000000"), (1e-7, "1e-7"), (1.5e-6, "0.0000015"), (123456789.0, "123456789"),
                    (0.1, "0.1"), (2.0 ** 70, "1.1805916207174113e+21"), (1e-6, "0.000001"), (9.999e-7, "9.999e-7")):
        assert ref.go_float(x) == want
    # through the native path: quantity text -> float64(unscaled) * math.Pow10(-scale) -> json.Marshal.  (Go multiplies by a rounded power
    # of ten, so "100n" is 100 * 1e-9 = 1.0000000000000001e-7, not 1e-7: both sides follow that formula.)
    for text, want in (("4", "4"), ("500m", "0.5"), ("0.000001", "0.000001"), ("100n", "1.0000000000000001e-7"), ("1E3", "1000"), ("1e21", "1e+21"),
                       ("100E", "100000000000000000000"), ("0.1", "0.1"), ("3", "3"), ("2.5Gi", "2684354560"), ("123456789", "123456789")):
        got = pm.ray_start_command("worker", {}, group_resources={"r": text}, steps=abi.RS_UPDATE_RESOURCES)["rayStartParams"]["resources"]
        assert got == "'{\"r\":" + want + "}'", (text, got)
        assert ref.go_float(ref.quantity_float(text)) == want, text
    rng = random.Random(3)
    for _ in range(400):
        text = rng.choice(["", "0."]) + str(rng.randint(0, 10 ** rng.randint(1, 12))) + rng.choice(["", "", "m", "u", "n", "k", "M", "Ki", "Mi", "Gi", "e-3", "e5", "E-12"])
        if ref.parse_quantity(text) is None:
            continue
        got = pm.ray_start_command("worker", {}, group_resources={"r": text}, steps=abi.RS_UPDATE_RESOURCES)["rayStartParams"]["resources"]
        assert got == "'{\"r\":" + ref.go_float(ref.quantity_float(text)) + "}'", (text, got)
        assert pm.quantity_value(text)[0] == ref.quantity_value(ref.parse_quantity(text)), text


NAMES = ["cpu", "CPU", "memory", "nvidia.com/gpu", "GPU", "nvidia.com/mig-1g.5gb", "google.com/tpu", "aws.amazon.com/neuroncore", "TPU", "neuron_cores", "custom/a<b", "x"]
QTYS = ["0", "1", "2", "500m", "1500m", "4Gi", "1.5Gi", "8", "100n", "abc", "", "1e3", "250M"]


@pytest.mark.parametrize("seed", range(6))
def test_native_builder_matches_restatement(seed):
    rng = random.Random(seed)
    for _ in range(150):
        def kv(keys, vals, n):
            return {rng.choice(keys): rng.choice(vals) for _ in range(rng.randint(0, n))}
        params = kv(["num-cpus", "num-gpus", "memory", "block", "address", "dashboard-host", "metrics-export-port", "port", "log-color", "verbose", "labels"],
                    ["1", "true", "False", "x=y", "8080", "0.0.0.0"], 4)
        if rng.random() < 0.4:
            params["resources"] = rng.choice(["'{\"TPU\":4}'", "\"{\"a\":1.5}\"", "{", "'{\"neuron_cores\":2,\"b\":0.25}'", "`{}`", "null", "'{\"a\":\"x\"}'", "'{\"a\":null}'"])
        node = rng.choice(["head", "worker"])
        args = dict(group_labels=kv(["zone", "ray.io/x", "a", "Z"], ["v1", "spot", ""], 3) or None, group_resources=kv(NAMES, QTYS, 4) or None,
                    limits=kv(NAMES, QTYS, 4) or None, requests=kv(["cpu", "memory"], QTYS, 2) or None,
                    command=rng.choice([None, [], ["echo", "x"], ["ray start --foo"]]), args=rng.choice([None, ["a b"], ["sleep", "1"]]),
                    head_port=rng.choice([None, "6379", "1234"]), fqdn_ray_ip=rng.choice(["", "svc.ns.svc.cluster.local"]),
                    autoscaling=rng.random() < 0.3, overwrite_cmd=rng.random() < 0.2, login_shell=rng.random() < 0.2)
        got = pm.ray_start_command(node, dict(params), **args)
        want = ref.ray_start_command(node, dict(params), **args)
        assert got == want, (node, params, args)


# ------------------------------------------------------------------------------------------------------------ container env (BuildPod)
def _env(lst):
    return {e["name"]: (e.get("value", "") if "valueFrom" not in e else e["valueFrom"]["fieldRef"]["fieldPath"]) for e in lst}


def test_container_env_vectors():
    """pod_test.go:641-728 (TestBuildPod), :988-1012 (created by RayService)."""
    fqdn = "raycluster-sample-head-svc.default.svc.cluster.local"
    head = pm.ray_container_env("head", existing=["TEST_ENV_NAME"], default_envs={"TEST_DEFAULT_ENV_NAME": "TEST_ENV_VALUE"}, head_port="6379",
                                ray_start_cmd="ray start --head  --block ", kuberay_version="v9")
    e = _env(head)
    assert e["RAY_ADDRESS"] == "127.0.0.1:6379" and e["RAY_USAGE_STATS_KUBERAY_IN_USE"] == "1" and e["RAY_CLUSTER_NAME"] == "metadata.labels['ray.io/cluster']"
    assert e["RAY_CLUSTER_NAMESPACE"] == "metadata.namespace" and e["RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE"] == "1" and e["RAY_NODE_TYPE_NAME"] == "metadata.labels['ray.io/group']"
    assert e["RAY_USAGE_STATS_EXTRA_TAGS"] == "kuberay_version=v9;kuberay_crd=RayCluster" and "ray start" in e["KUBERAY_GEN_RAY_START_CMD"]
    assert e["TEST_DEFAULT_ENV_NAME"] == "TEST_ENV_VALUE" and e["RAY_CLOUD_INSTANCE_ID"] == "metadata.name" and e["RAY_PORT"] == "6379"
    assert "FQ_RAY_IP" not in e and "RAY_IP" not in e
    worker = _env(pm.ray_container_env("worker", fqdn_ray_ip=fqdn, head_port="6379", ray_start_cmd="ray start  --block "))
    assert worker["RAY_ADDRESS"] == fqdn + ":6379" and worker["FQ_RAY_IP"] == fqdn and worker["RAY_IP"] == "raycluster-sample-head-svc"
    assert "RAY_USAGE_STATS_EXTRA_TAGS" not in worker
    init = pm.ray_container_env("worker", fqdn_ray_ip=fqdn, init_container=True)
    assert init == [{"name": "FQ_RAY_IP", "value": fqdn}, {"name": "RAY_IP", "value": "raycluster-sample-head-svc"}]
    svc = _env(pm.ray_container_env("head", crd_type="RayService", existing=["RAY_SERVE_KV_TIMEOUT_S"]))
    assert svc["RAY_timeout_ms_task_wait_for_death_info"] == "0" and svc["RAY_gcs_server_request_timeout_seconds"] == "5" and "RAY_SERVE_KV_TIMEOUT_S" not in svc
    assert svc["RAY_USAGE_STATS_EXTRA_TAGS"].endswith("kuberay_crd=RayService")
    # names the template already sets are left alone; a default env with such a name is skipped; a default named like a managed one counts as set
    e = _env(pm.ray_container_env("head", existing=["RAY_ADDRESS", "RAY_PORT", "X"], default_envs={"X": "1", "RAY_USAGE_STATS_KUBERAY_IN_USE": "0"}))
    assert "RAY_ADDRESS" not in e and "RAY_PORT" not in e and "X" not in e and e["RAY_USAGE_STATS_KUBERAY_IN_USE"] == "0"
    # value omitted when empty (omitempty), as Go marshals it
    raw = pm.ray_container_env("head", head_port="", ray_start_cmd="")
    assert {"name": "RAY_PORT"} in raw and {"name": "KUBERAY_GEN_RAY_START_CMD"} in raw


@pytest.mark.parametrize("seed", range(4))
def test_container_env_matches_restatement(seed):
    rng = random.Random(seed)
    pool = ["RAY_ADDRESS", "RAY_PORT", "RAY_USAGE_STATS_KUBERAY_IN_USE", "RAY_DASHBOARD_ENABLE_K8S_DISK_USAGE", "RAY_SERVE_KV_TIMEOUT_S", "RAY_timeout_ms_task_wait_for_death_info",
            "FOO", "BAR", "RAY_CLUSTER_NAME", "a<b"]
    for _ in range(200):
        kw = dict(existing=rng.sample(pool, rng.randint(0, 4)), default_envs={rng.choice(pool): rng.choice(["", "1", "x&y"]) for _ in range(rng.randint(0, 3))} or None,
                  fqdn_ray_ip=rng.choice(["", "svc", "svc.ns.svc.cluster.local"]), head_port=rng.choice(["", "6379"]), ray_start_cmd=rng.choice(["", "ray start --head  --block "]),
                  crd_type=rng.choice(["RayCluster", "RayJob", "RayService"]), kuberay_version=rng.choice(["v1.5.0", "nightly"]), init_container=rng.random() < 0.15)
        node = rng.choice(["head", "worker"])
        assert pm.ray_container_env(node, **kw) == ref.ray_container_env(node, **kw), (node, kw)


# ------------------------------------------------------------------------------------------------------------ probes (BuildPod)
def _cmd(pr):
    return " ".join(pr["exec"]["command"]) if "exec" in pr else ""


def test_probe_vectors():
    """pod_test.go:1824-1985 (TestInitLivenessAndReadinessProbe), the numbered cases."""
    # 1: the user's own probes are left alone
    assert pm.ray_probes("head", {}, has_liveness=True, has_readiness=True) == {}
    # 2: RayService worker — exec probes, the Serve proxy check only in the readiness probe, 2 s timeouts
    r = pm.ray_probes("worker", {}, crd_type="RayService")
    assert "exec" in r["livenessProbe"] and "exec" in r["readinessProbe"] and "-/healthz" not in _cmd(r["livenessProbe"]) and "-/healthz" in _cmd(r["readinessProbe"])
    assert r["livenessProbe"]["timeoutSeconds"] == 2 and r["readinessProbe"]["timeoutSeconds"] == 2 and r["readinessProbe"]["failureThreshold"] == 1
    # 3: RayService head — no proxy check, 5 s timeouts
    r = pm.ray_probes("head", {}, crd_type="RayService")
    assert "-/healthz" not in _cmd(r["livenessProbe"]) + _cmd(r["readinessProbe"]) and r["livenessProbe"]["timeoutSeconds"] == 5 and r["readinessProbe"]["timeoutSeconds"] == 5
    # 4: custom ports, head
    r = pm.ray_probes("head", {"dashboard-agent-listen-port": "8266", "dashboard-port": "8365"})
    for pr in r.values():
        assert ":8266" in _cmd(pr) and ":8365" in _cmd(pr)
    # 5: custom ports, worker: no dashboard-port check
    r = pm.ray_probes("worker", {"dashboard-agent-listen-port": "9000"})
    for pr in r.values():
        assert ":9000" in _cmd(pr) and ":8265" not in _cmd(pr)
    # 6: RayService worker with a custom agent port and the serve port
    r = pm.ray_probes("worker", {"dashboard-agent-listen-port": "8500"}, crd_type="RayService", serving_port=8000)
    assert ":8500" in _cmd(r["readinessProbe"]) and "-/healthz" in _cmd(r["readinessProbe"]) and r["readinessProbe"]["failureThreshold"] == 1
    # 8: invalid ports fall back to the defaults
    r = pm.ray_probes("head", {"dashboard-agent-listen-port": "invalid-port", "dashboard-port": "not-a-number"})
    assert ":52365" in _cmd(r["livenessProbe"]) and ":8265" in _cmd(r["livenessProbe"])
    # 9: Ray >= 2.53.0: one HTTP check; a RayService worker's readiness stays exec; 2.52.0: exec
    r = pm.ray_probes("head", {}, ray_version="2.53.0")
    assert r["livenessProbe"]["httpGet"] == {"path": "/api/healthz", "port": 52365} and "exec" not in r["livenessProbe"] and "httpGet" in r["readinessProbe"]
    r = pm.ray_probes("worker", {}, ray_version="2.53.0")
    assert "httpGet" in r["livenessProbe"] and "httpGet" in r["readinessProbe"]
    r = pm.ray_probes("worker", {}, crd_type="RayService", ray_version="2.53.0")
    assert "httpGet" in r["livenessProbe"] and "exec" in r["readinessProbe"] and "httpGet" not in r["readinessProbe"] and "-/healthz" in _cmd(r["readinessProbe"])
    r = pm.ray_probes("head", {}, ray_version="2.52.0")
    assert "exec" in r["livenessProbe"] and "exec" in r["readinessProbe"]
    # the full head probe, literally
    assert pm.ray_probes("head", {}, has_readiness=True) == {"livenessProbe": {"exec": {"command": ["bash", "-c",
        "wget --tries 1 -T 2 -q -O- http://localhost:52365/api/local_raylet_healthz | grep success && wget --tries 1 -T 10 -q -O- http://localhost:8265/api/gcs_healthz | grep success"]},
        "initialDelaySeconds": 30, "timeoutSeconds": 5, "periodSeconds": 5, "successThreshold": 1, "failureThreshold": 120}}


@pytest.mark.parametrize("text,want", [("2.53.0", True), ("2.53", True), ("2.52.9", False), ("v2.53.1", True), ("3.0.0", True), ("2.53.0rc1", True), (" 2.100.0 ", True),
                                       ("", False), ("nightly", False), ("2", False), ("02.53.0", False), ("2.053.0", True), ("2.53.0.1", True), ("1.99.99", False), ("2.9.0", False)])
def test_ray_version_gate(text, want):
    assert ref.ray_version_at_least(text) == want
    assert ("httpGet" in pm.ray_probes("head", {}, ray_version=text)["livenessProbe"]) == want


@pytest.mark.parametrize("seed", range(3))
def test_probes_match_restatement(seed):
    rng = random.Random(seed)
    for _ in range(300):
        params = {k: rng.choice(["8266", "abc", "", "-5", "99999999999", "+80"]) for k in rng.sample(["dashboard-agent-listen-port", "dashboard-port", "x"], rng.randint(0, 3))}
        kw = dict(crd_type=rng.choice(["RayCluster", "RayJob", "RayService"]), ray_version=rng.choice(["", "2.52.0", "2.53.0", "v2.60", "x"]), has_liveness=rng.random() < 0.3,
                  has_readiness=rng.random() < 0.3, serving_port=rng.choice([0, 8000, 9001]))
        node = rng.choice(["head", "worker"])
        assert pm.ray_probes(node, params, **kw) == ref.ray_probes(node, params, **kw), (node, params, kw)


# ------------------------------------------------------------------------------------------------------------ volumes (BuildPod)
def test_volume_vectors():
    """pod_test.go:222-256 (TestAddEmptyDirVolumes), :730-777 (TestBuildPod_WithPlasmaDirectory), :957-986 (autoscaler log volume)."""
    shm = {"name": "shared-mem", "mountPath": "/dev/shm"}
    r = pm.ray_volumes("head", memory_limit="1Gi")
    assert r == {"volumes": [{"name": "shared-mem", "emptyDir": {"medium": "Memory", "sizeLimit": "1Gi"}}], "rayContainerVolumeMounts": [shm], "autoscalerVolumeMounts": []}
    assert pm.ray_volumes("worker", memory_request="1500M")["volumes"][0]["emptyDir"] == {"medium": "Memory", "sizeLimit": "1500M"}   # limit absent: the request
    assert pm.ray_volumes("worker", memory_limit="2048Mi", memory_request="1Gi")["volumes"][0]["emptyDir"]["sizeLimit"] == "2Gi"    # canonical form, limit wins
    assert pm.ray_volumes("worker")["volumes"][0]["emptyDir"] == {"medium": "Memory"}                                               # no memory resource: no limit
    # any plasma-directory (even /dev/shm itself) skips the shared-memory mount
    assert pm.ray_volumes("head", plasma_directory_set=True, memory_limit="1Gi") == {"volumes": [], "rayContainerVolumeMounts": [], "autoscalerVolumeMounts": []}
    # the path already mounted: nothing; the volume name already there: only the mount
    assert pm.ray_volumes("worker", ray_mount_paths=["/dev/shm"], volume_names=["shared-mem"]) == {"volumes": [], "rayContainerVolumeMounts": [], "autoscalerVolumeMounts": []}
    assert pm.ray_volumes("worker", volume_names=["shared-mem"]) == {"volumes": [], "rayContainerVolumeMounts": [shm], "autoscalerVolumeMounts": []}
    # head with the autoscaler sidecar: one ray-logs volume, mounted in both containers
    r = pm.ray_volumes("head", autoscaling=True, memory_limit="1Gi")
    logs = {"name": "ray-logs", "mountPath": "/tmp/ray"}
    assert r["volumes"] == [{"name": "shared-mem", "emptyDir": {"medium": "Memory", "sizeLimit": "1Gi"}}, {"name": "ray-logs", "emptyDir": {}}]
    assert r["rayContainerVolumeMounts"] == [shm, logs] and r["autoscalerVolumeMounts"] == [logs]
    assert pm.ray_volumes("worker", autoscaling=True)["autoscalerVolumeMounts"] == []
    with pytest.raises(EngineError):
        pm.ray_volumes("head", memory_limit="lots")
