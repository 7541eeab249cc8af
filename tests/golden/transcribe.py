#!/usr/bin/env python
"""Transcription of the reference's own known-answer tests for the hot path into language-neutral JSON fixtures.

The Go reference (ray-project/kuberay) cannot be built or run in this image (no Go toolchain, un-vendored k8s modules), so
these fixtures are NOT outputs of the reference: they are hand transcriptions of the inputs and the asserted outcomes of
its Go tests, each entry citing file:line under ray-operator/controllers/ray/.  Running this script rewrites the JSON files
next to it (they are committed); tests/test_golden_*.py replay them against the CPU oracle (always) and the CUDA engine (-m gpu).
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
I32MAX = 2147483647

# ---------------------------------------------------------------------------------------------- replica arithmetic
replica_arithmetic = {
    "_source": "utils/util_test.go",
    "desired_replicas": {  # TestGetWorkerGroupDesiredReplicas utils/util_test.go:555-601
        "_cite": "utils/util_test.go:555-601",
        "cases": [
            {"replicas": None, "min": 1, "max": 5, "hosts": 1, "suspend": False, "want": 1},
            {"replicas": 3, "min": 1, "max": 5, "hosts": 1, "suspend": False, "want": 3},
            {"replicas": 6, "min": 1, "max": 5, "hosts": 1, "suspend": False, "want": 5},
            {"replicas": 0, "min": 1, "max": 5, "hosts": 1, "suspend": False, "want": 1},
            {"replicas": 0, "min": 5, "max": 1, "hosts": 1, "suspend": True, "want": 0},
            {"replicas": 5, "min": 1, "max": 5, "hosts": 4, "suspend": False, "want": 20},
        ],
    },
    "min_max": {  # TestCalculateMinAndMaxReplicas utils/util_test.go:603-710
        "_cite": "utils/util_test.go:603-710",
        "cases": [
            {"name": "Single group with one host", "groups": [{"hosts": 1, "min": 2, "max": 3}], "want_min": 2, "want_max": 3},
            {"name": "Single group with four hosts", "groups": [{"hosts": 4, "min": 2, "max": 3}], "want_min": 8, "want_max": 12},
            {"name": "Two worker groups: one with a single host, one with two hosts",
             "groups": [{"hosts": 1, "min": 4, "max": 4}, {"hosts": 2, "min": 3, "max": 3}], "want_min": 10, "want_max": 10},
            {"name": "Two groups with suspended",
             "groups": [{"hosts": 1, "min": 3, "max": 3, "suspend": True}, {"hosts": 1, "min": 1, "max": 1, "suspend": True}], "want_min": 0, "want_max": 0},
        ],
    },
    "desired_cluster": {  # TestCalculateDesiredReplicas utils/util_test.go:712-800
        "_cite": "utils/util_test.go:712-800",
        "cases": [
            {"name": "Both groups' Replicas are nil", "groups": [{"replicas": None, "hosts": 1, "min": 1, "max": 5}, {"replicas": None, "hosts": 1, "min": 2, "max": 5}], "want": 3},
            {"name": "Group1 < Min, Group2 > Max", "groups": [{"replicas": 0, "hosts": 1, "min": 2, "max": 5}, {"replicas": 6, "hosts": 1, "min": 2, "max": 5}], "want": 7},
            {"name": "Group1 > Max", "groups": [{"replicas": 6, "hosts": 1, "min": 2, "max": 5}, {"replicas": 3, "hosts": 1, "min": 2, "max": 5}], "want": 8},
            {"name": "Group1 NumOfHosts 4", "groups": [{"replicas": 3, "hosts": 4, "min": 1, "max": 6}, {"replicas": 3, "hosts": 1, "min": 2, "max": 5}], "want": 15},
        ],
    },
    "max_overflow": {  # TestCalculateMaxReplicasOverflow utils/util_test.go:802-894
        "_cite": "utils/util_test.go:802-894",
        "cases": [
            {"name": "issue report replicas=1 min=3 hosts=4", "groups": [{"replicas": 1, "min": 3, "max": I32MAX, "hosts": 4}], "want_max": I32MAX},
            {"name": "default maxReplicas hosts=4", "groups": [{"hosts": 4, "min": 3, "max": I32MAX}], "want_max": I32MAX},
            {"name": "large values", "groups": [{"hosts": 1000, "min": 1, "max": I32MAX}], "want_max": I32MAX},
            {"name": "multiple groups overflow when summed", "groups": [{"hosts": 2, "min": 1, "max": 1500000000}, {"hosts": 1, "min": 1, "max": 1000000000}], "want_max": I32MAX},
            {"name": "no overflow", "groups": [{"hosts": 4, "min": 2, "max": 100}], "want_max": 400},
            {"name": "exactly at max int32", "groups": [{"hosts": 1, "min": 1, "max": I32MAX}], "want_max": I32MAX},
        ],
    },
    "available_ready": {  # TestCalculateAvailableReplicas utils/util_test.go:408-475
        "_cite": "utils/util_test.go:408-475",
        "pods": [
            {"name": "pod1", "nodeType": "head", "phase": "Running"},
            {"name": "pod2", "nodeType": "worker", "phase": "Running", "ready": "True"},
            {"name": "pod3", "nodeType": "worker", "phase": "Pending", "ready": "False"},
            {"name": "pod4", "nodeType": "worker", "phase": "Failed"},
        ],
        "want_available": 1, "want_ready": 1,
    },
    "check_all_pods_running": {  # TestCheckAllPodsRunning utils/util_test.go:58-130
        "_cite": "utils/util_test.go:58-130",
        "cases": [
            {"name": "all running", "pods": [{"phase": "Running"}, {"phase": "Running"}], "want": True},
            {"name": "no pods", "pods": [], "want": False},
            {"name": "one pending", "pods": [{"phase": "Pending"}, {"phase": "Running"}], "want": False},
            {"name": "Ready condition not True", "pods": [{"phase": "Running", "ready": "False"}], "want": False},
        ],
    },
}

# ---------------------------------------------------------------------------------------------- shouldDeletePod
should_delete_pod = {  # Test_ShouldDeletePod raycluster_controller_unit_test.go:2380-2503
    "_cite": "raycluster_controller_unit_test.go:2380-2503",
    "cases": [
        {"restartPolicy": "Always", "phase": "Failed", "terminated": False, "want": True},
        {"restartPolicy": "Always", "phase": "Running", "terminated": False, "want": False},
        {"restartPolicy": "Always", "phase": "Running", "terminated": True, "want": False},
        {"restartPolicy": "Never", "phase": "Failed", "terminated": False, "want": True},
        {"restartPolicy": "Never", "phase": "Succeeded", "terminated": False, "want": True},
        {"restartPolicy": "Never", "phase": "Running", "terminated": False, "want": False},
        {"restartPolicy": "Never", "phase": "Running", "terminated": True, "want": True},
    ],
}

# ---------------------------------------------------------------------------------------------- head-ready condition
head_pod_ready = {  # TestFindHeadPodReadyCondition / TestFindHeadPodReadyMessage utils/util_test.go:934-1038
    "_cite": "utils/util_test.go:331-355,934-1038",
    "status_cases": [
        {"phase": "Running", "ready": "True", "want_status": "True"},
        {"phase": "Pending", "ready": "False", "want_status": "False"},
        {"phase": "Running", "ready": "False", "want_status": "False"},
    ],
    "message_cases": [
        {"name": "no message no status", "message": "", "containerStatuses": [], "want_reason": "ContainersNotReady", "want_message": ""},
        {"name": "no container status", "message": "TooEarlyInTheMorning", "containerStatuses": [], "want_reason": "ContainersNotReady", "want_message": "TooEarlyInTheMorning"},
        {"name": "one reason one status", "message": "containers not ready",
         "containerStatuses": [{"name": "ray", "state": {"waiting": {"reason": "ImagePullBackOff", "message": "Back-off pulling image royproject/roy:latest: ErrImagePull: rpc error: code = NotFound"}}}],
         "want_reason": "ImagePullBackOff", "want_message": "containers not ready; ray: Back-off pulling image royproject/roy:latest: ErrImagePull: rpc error: code = NotFound"},
        {"name": "two statuses only copy first", "message": "aesthetic problems",
         "containerStatuses": [{"name": "indigo", "state": {"waiting": {"reason": "BadColor", "message": "too blue"}}},
                               {"name": "circle", "state": {"terminated": {"reason": "BadGeometry", "message": "too round"}}}],
         "want_reason": "BadColor", "want_message": "aesthetic problems; indigo: too blue"},
        {"name": "no reason one status", "message": "",
         "containerStatuses": [{"name": "my-image", "state": {"terminated": {"reason": "Crashed", "message": "bash not found"}}}],
         "want_reason": "Crashed", "want_message": "my-image: bash not found"},
    ],
}

# ---------------------------------------------------------------------------------------------- base fixture F0
# setupTest raycluster_controller_unit_test.go:90-415: namespace default; RayCluster raycluster-sample, autoscaling on,
# one worker group small-group {replicas 3, min 0, max 10000, numOfHosts 1, workersToDelete [pod1,pod2]}; head `headNode`
# (node-type head, ray.io/group=head-group, :95,107-112) and workers pod1..pod5 carrying NO
# ray.io/node-type label; everything Running with a live ray container.
def f0_cluster():
    return {"namespace": "default", "name": "raycluster-sample",
            "spec": {"enableInTreeAutoscaling": True, "rayVersion": "2.46.0",
                     "headGroupSpec": {"rayStartParams": {}, "template": {"spec": {"containers": [{"name": "ray-head", "image": "rayproject/ray:2.46.0"}]}}},
                     "workerGroupSpecs": [{"groupName": "small-group", "replicas": 3, "minReplicas": 0, "maxReplicas": 10000, "numOfHosts": 1,
                                           "rayStartParams": {}, "workersToDelete": ["pod1", "pod2"],
                                           "template": {"spec": {"containers": [{"name": "ray-worker", "image": "rayproject/ray:2.46.0"}]}}}]},
            "status": {}}


def f0_pod(name, group, node_type=None, **kw):
    labels = {"ray.io/is-ray-node": "yes", "ray.io/cluster": "raycluster-sample", "ray.io/group": group}
    if node_type:
        labels["ray.io/node-type"] = node_type
    p = {"namespace": "default", "name": name, "labels": labels, "phase": "Running", "restartPolicy": "Always",
         "containers": [{"name": "ray-head" if node_type == "head" else "ray-worker"}],
         "containerStatuses": [{"name": "ray-head" if node_type == "head" else "ray-worker", "state": {}}]}
    p.update(kw)
    return p


def f0_pods():
    return [f0_pod("headNode", "head-group", "head", podIP="1.2.3.4")] + [f0_pod(f"pod{i}", "small-group") for i in range(1, 6)]


reconcile_scenarios = {
    "_source": "raycluster_controller_unit_test.go",
    "base": {"cluster": f0_cluster(), "pods": f0_pods()},
    "scenarios": [
        # TestReconcile_RemoveWorkersToDelete_RandomDelete :417-538
        *[{"name": f"RemoveWorkersToDelete_RandomDelete[{','.join(w)}]", "cite": ":417-538", "env": {"enable_random_pod_delete": True},
           "patch_group": {"workersToDelete": w}, "want_err": False, "want_workers": 3, "want_gone": [x for x in w if x.startswith("pod")],
           "want_random_deletes": n}
          for w, n in ((["pod1", "pod2"], 0), (["pod3", "pod4"], 0), (["pod1", "pod5"], 0), (["pod2", "NonExistentPod"], 1), (["NonExistentPod1", "NonExistentPod2"], 2))],
        # TestReconcile_RemoveWorkersToDelete_NoRandomDelete :540-631
        *[{"name": f"RemoveWorkersToDelete_NoRandomDelete[{','.join(w)}]", "cite": ":540-631", "env": {},
           "patch_group": {"workersToDelete": w}, "want_err": False, "want_workers": n, "want_gone": [x for x in w if x.startswith("pod")], "want_random_deletes": 0}
          for w, n in ((["pod2", "pod3"], 3), (["pod2", "NonExistentPod"], 4), (["NonExistentPod1", "NonExistentPod2"], 5))],
        # TestReconcile_RandomDelete_OK :633-678 — autoscaling nil, replicas 2, WTD [pod1,pod2] => pod3 goes too
        {"name": "RandomDelete_OK", "cite": ":633-678", "env": {}, "patch_spec": {"enableInTreeAutoscaling": None},
         "patch_group": {"replicas": 2, "workersToDelete": ["pod1", "pod2"]}, "want_err": False, "want_workers": 2, "want_gone": ["pod1", "pod2", "pod3"]},
        # TestReconcile_PodDeleted_Diff0_OK :680-736 — two workers deleted externally beforehand
        {"name": "PodDeleted_Diff0_OK", "cite": ":680-736", "env": {}, "patch_group": {"workersToDelete": []}, "pre_delete": ["pod3", "pod4"],
         "want_err": False, "want_workers": 3, "want_creates": 0, "want_deletes": 0},
        # TestReconcile_PodDeleted_DiffLess0_OK :738-796 — autoscaling nil, one worker deleted externally => one prefix delete
        {"name": "PodDeleted_DiffLess0_OK", "cite": ":738-796", "env": {}, "patch_spec": {"enableInTreeAutoscaling": None}, "patch_group": {"workersToDelete": []},
         "pre_delete": ["pod3"], "want_err": False, "want_workers": 3, "want_gone": ["pod1"]},
        # TestReconcile_Diff0_WorkersToDelete_OK :798-852
        {"name": "Diff0_WorkersToDelete_OK", "cite": ":798-852", "env": {}, "patch_group": {"workersToDelete": ["pod3", "pod4"]},
         "want_err": False, "want_workers": 3, "want_gone": ["pod3", "pod4"]},
        # TestReconcile_PodCrash_DiffLess0_OK :854-947 — WTD [pod3]; env true => 3 remain, env false => 4 remain
        {"name": "PodCrash_DiffLess0_OK[env=true]", "cite": ":854-947", "env": {"enable_random_pod_delete": True}, "patch_group": {"workersToDelete": ["pod3"]},
         "want_err": False, "want_workers": 3, "want_gone": ["pod3", "pod1"]},
        {"name": "PodCrash_DiffLess0_OK[env=false]", "cite": ":854-947", "env": {}, "patch_group": {"workersToDelete": ["pod3"]},
         "want_err": False, "want_workers": 4, "want_gone": ["pod3"]},
        # TestReconcile_PodEvicted_DiffLess0_OK :949-1010 — head Failed, restartPolicy Always / OnFailure => error, head gone, workers untouched
        *[{"name": f"PodEvicted_DiffLess0_OK[{rp}]", "cite": ":949-1010", "env": {}, "patch_pods": {"headNode": {"phase": "Failed", "restartPolicy": rp}},
           "want_err": True, "want_heads": 0, "want_workers": 5} for rp in ("Always", "OnFailure")],
        # TestReconcile_Replicas_Optional :2873-2963 — autoscaling false, WTD []
        *[{"name": f"Replicas_Optional[{r},{mn},{mx}]", "cite": ":2873-2963", "env": {}, "patch_spec": {"enableInTreeAutoscaling": False},
           "patch_group": {"replicas": r, "minReplicas": mn, "maxReplicas": mx, "workersToDelete": []}, "want_err": False, "want_workers": n}
          for r, mn, mx, n in ((None, 1, 10000, 1), (0, 1, 10000, 1), (4, 1, 3, 3))],
        # TestReconcile_Multihost_Replicas :2965-3062 — RayMultiHostIndexing gate OFF, numOfHosts 4
        *[{"name": f"Multihost_Replicas[{r},{mn},{mx}]", "cite": ":2965-3062", "env": {"multihost_indexing_gate": False}, "patch_spec": {"enableInTreeAutoscaling": False},
           "patch_group": {"replicas": r, "minReplicas": mn, "maxReplicas": mx, "numOfHosts": 4, "workersToDelete": []}, "want_err": False, "want_workers": n}
          for r, mn, mx, n in ((None, 1, 10000, 4), (0, 1, 10000, 4), (4, 1, 3, 12))],
        # TestReconcile_NumOfHosts :3064-3137 — only the head exists; replicas 1; numOfHosts 1 / 4
        *[{"name": f"NumOfHosts[{h}]", "cite": ":3064-3137", "env": {}, "patch_spec": {"enableInTreeAutoscaling": False},
           "patch_group": {"replicas": 1, "minReplicas": 1, "maxReplicas": 10000, "numOfHosts": h, "workersToDelete": []},
           "pre_delete": ["pod1", "pod2", "pod3", "pod4", "pod5"], "want_err": False, "want_workers": h} for h in (1, 4)],
    ],
    # Test_TerminatedWorkers_NoAutoscaler :2093-2221 (multi-pass)
    "terminated_workers": {
        "cite": ":2093-2221", "patch_spec": {"enableInTreeAutoscaling": None}, "patch_group": {"workersToDelete": []},
        "passes": [
            {"want_err": False, "want_workers": 3},
            {"set_phase": {"first_worker": "Failed"}, "want_err": True, "want_workers": 2},
            {"want_err": False, "want_workers": 3},
            {"set_phase": {"first_worker": "Succeeded"}, "want_err": True, "want_workers": 2},
            {"want_err": False, "want_workers": 3},
        ],
    },
    # Test_TerminatedHead_RestartPolicy :2223-2309, Test_RunningPods_RayContainerTerminated :2311-2378 (no worker groups)
    "terminated_head": {
        "cite": ":2223-2378",
        "steps": [
            {"head": {"phase": "Failed", "restartPolicy": "Always"}, "want_err": True, "want_pods": 0},
            {"want_err": False, "want_pods": 1},
            {"head": {"phase": "Running", "restartPolicy": "Never", "rayContainerTerminated": True}, "want_err": True, "want_pods": 0},
            {"want_err": False, "want_pods": 1},
        ],
    },
}

# ---------------------------------------------------------------------------------------------- shouldRecreatePodsForUpgrade
recreate_upgrade = {  # TestShouldRecreatePodsForUpgrade raycluster_controller_unit_test.go:3680-3814
    "_cite": "raycluster_controller_unit_test.go:3680-3814",
    "cases": [
        {"name": "strategy nil", "upgradeStrategy": None, "head": None, "want": False},
        {"name": "type nil", "upgradeStrategy": {"type": None}, "head": None, "want": False},
        {"name": "type None", "upgradeStrategy": {"type": "None"}, "head": None, "want": False},
        {"name": "Recreate, no pods", "upgradeStrategy": {"type": "Recreate"}, "head": None, "want": False},
        {"name": "Recreate, hash equal", "upgradeStrategy": {"type": "Recreate"}, "head": {"hash": "<current>", "version": "<current>"}, "want": False},
        {"name": "Recreate, hash differs (same version)", "upgradeStrategy": {"type": "Recreate"}, "head": {"hash": "0123456789ABCDEFGHIJKLMNOPQRSTUV", "version": "<current>"}, "want": True},
        {"name": "Recreate, version v1.0.0, hash differs", "upgradeStrategy": {"type": "Recreate"}, "head": {"hash": "0123456789ABCDEFGHIJKLMNOPQRSTUV", "version": "v1.0.0"}, "want": False},
        {"name": "same version, hash differs", "upgradeStrategy": {"type": "Recreate"}, "head": {"hash": "VUTSRQPONMLKJIHGFEDCBA9876543210", "version": "<current>"}, "want": True},
        {"name": "same version, hash equal", "upgradeStrategy": {"type": "Recreate"}, "head": {"hash": "<current>", "version": "<current>"}, "want": False},
    ],
}

# ---------------------------------------------------------------------------------------------- calculateStatus
status_scenarios = {
    "_source": "raycluster_controller_unit_test.go",
    # TestCalculateStatus :1611-1723 — head + 3 workers with node-type labels, all Running+Ready; head Service ClusterIP set
    "calculate_status": {"cite": ":1611-1723", "headServiceIP": "aaa.bbb.ccc.ddd", "headNodeIP": "1.2.3.4", "workers": 3},
    # TestCalculateStatusWithoutDesiredReplicas :1727-1780 — only the head pod exists, desired 3
    "without_desired": {"cite": ":1727-1780", "want_state": "", "want_reason": ""},
    # TestCalculateStatusWithSuspendedWorkerGroups :1784-1847 — group suspend, min=max=100; only head pod
    "suspended_groups": {"cite": ":1784-1847", "want": {"desired": 0, "min": 0, "max": 0, "state": "ready"}},
    # TestCalculateStatusWithReconcileErrorBackAndForth :1851-1944 — err -> nil -> err
    "error_back_and_forth": {"cite": ":1851-1944", "want_states": ["", "ready", "ready"]},
    # TestRayClusterProvisionedCondition :1946-2042
    "provisioned": {"cite": ":1946-2042",
                    "steps": [
                        {"head_ready": "False", "worker_ready": "False", "want": ["False", "RayClusterPodsProvisioning"]},
                        {"head_ready": "True", "worker_ready": "True", "want": ["True", "AllPodRunningAndReadyFirstTime"]},
                        {"head_ready": "True", "worker_ready": "False", "want": ["True", "AllPodRunningAndReadyFirstTime"]},
                        {"head_ready": "False", "worker_ready": "False", "want": ["True", "AllPodRunningAndReadyFirstTime"]},
                    ]},
    # TestStateTransitionTimes_NoStateChange :2044-2091
    "no_state_change": {"cite": ":2044-2091"},
}

# ---------------------------------------------------------------------------------------------- InconsistentRayClusterStatus
inconsistent_status = {  # TestInconsistentRayClusterStatus utils/consistency_test.go:16-146
    "_cite": "utils/consistency_test.go:16-146",
    "old": {"state": "ready", "readyWorkerReplicas": 1, "availableWorkerReplicas": 1, "desiredWorkerReplicas": 1, "minWorkerReplicas": 1, "maxWorkerReplicas": 10,
            "lastUpdateTime": "t0", "endpoints": {"client": "10001", "dashboard": "8265", "gcs-server": "6379", "metrics": "8080"},
            "head": {"podIP": "10.244.0.6", "serviceIP": "10.96.140.249"}, "observedGeneration": 1, "reason": "test reason"},
    "cases": [
        {"name": "State", "set": {"state": "suspended"}, "want": True},
        {"name": "Reason", "set": {"reason": "new reason"}, "want": True},
        {"name": "ReadyWorkerReplicas", "set": {"readyWorkerReplicas": 2}, "want": True},
        {"name": "AvailableWorkerReplicas", "set": {"availableWorkerReplicas": 2}, "want": True},
        {"name": "DesiredWorkerReplicas", "set": {"desiredWorkerReplicas": 2}, "want": True},
        {"name": "MinWorkerReplicas", "set": {"minWorkerReplicas": 2}, "want": True},
        {"name": "MaxWorkerReplicas", "set": {"maxWorkerReplicas": 11}, "want": True},
        {"name": "Endpoints", "set_endpoint": {"fakeEndpoint": "10009"}, "want": True},
        {"name": "Head.PodIP", "set_head": {"podIP": "test head pod ip"}, "want": True},
        {"name": "RayClusterReplicaFailure condition", "set": {"conditions": [{"type": "ReplicaFailure", "status": "True"}]}, "want": True},
        {"name": "LastUpdateTime", "set": {"lastUpdateTime": "t0+1h"}, "want": False},
        {"name": "ObservedGeneration", "set": {"observedGeneration": 2}, "want": False},
    ],
}

# ---------------------------------------------------------------------------------------------- hash relations
hash_relations = {  # TestGenerateHashWithoutReplicasAndWorkersToDelete rayservice_controller_unit_test.go:39-97
    "_cite": "rayservice_controller_unit_test.go:39-97",
    "note": "relational only: the reference holds no literal digest anywhere (SURVEY.md §4)",
    "relations": ["h(replicas+1) == h", "h(rayVersion changed) != h", "h(+tolerations) == h", "h(+schedulingGates) == h"],
}

# ---------------------------------------------------------------------------------------------- SHA-1 / base32hex vectors
sha1_vectors = {
    "_cite": "FIPS 180-4 examples (SHA-1 of 'abc', of the 448-bit and 896-bit messages) and RFC 4648 §10 base32hex test vectors",
    "sha1": [
        {"msg": "abc", "hex": "a9993e364706816aba3e25717850c26c9cd0d89d"},
        {"msg": "", "hex": "da39a3ee5e6b4b0d3255bfef95601890afd80709"},
        {"msg": "abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq", "hex": "84983e441c3bd26ebaae4aa1f95129e5e54670f1"},
        {"msg": "abcdefghbcdefghicdefghijdefghijkefghijklfghijklmghijklmnhijklmnoijklmnopjklmnopqklmnopqrlmnopqrsmnopqrstnopqrstu", "hex": "a49b2446a02c645bf419f995b67091253a04a259"},
    ],
    "base32hex": [["", ""], ["f", "CO======"], ["fo", "CPNG===="], ["foo", "CPNMU==="], ["foob", "CPNMUOG="], ["fooba", "CPNMUOJ1"], ["foobar", "CPNMUOJ1E8======"]],
}

if __name__ == "__main__":
    for name, obj in (("replica_arithmetic", replica_arithmetic), ("should_delete_pod", should_delete_pod), ("head_pod_ready", head_pod_ready),
                      ("reconcile_scenarios", reconcile_scenarios), ("recreate_upgrade", recreate_upgrade), ("status_scenarios", status_scenarios),
                      ("inconsistent_status", inconsistent_status), ("hash_relations", hash_relations), ("sha1_vectors", sha1_vectors)):
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(obj, f, indent=1, sort_keys=False)
            f.write("\n")
    print("wrote fixtures to", HERE)
