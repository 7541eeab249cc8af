package krengine

/*
#include "kr_engine.h"
*/
import "C"

import (
	"encoding/json"
	"errors"
	"unsafe"

	corev1 "k8s.io/api/core/v1"
)

// CreateTuple is one Pod the engine decided to create: Group -1 is the head, else the index of the worker group in spec order.
type CreateTuple struct {
	Group        int32
	ReplicaIndex int32
	HostIndex    int32
	ReplicaName  string // replicaGrpName of a multi-host group; "" otherwise
}

// BuilderEnv is the operator process's contribution to a Pod manifest (kr_podbuild_env): configuration and environment switches.
type BuilderEnv struct {
	KubeRayVersion          string // utils.KUBERAY_VERSION
	ClusterDomain           string // CLUSTER_DOMAIN ("" = cluster.local)
	DeterministicHeadName   bool   // utils.IsDeterministicHeadPodNameEnabled()
	MultiHostIndexing       bool   // features.RayMultiHostIndexing
	LoginShell              bool   // ENABLE_LOGIN_SHELL
	NoInitContainer         bool   // ENABLE_INIT_CONTAINER_INJECTION == "false"
	NoProbes                bool   // ENABLE_PROBES_INJECTION == "false"
	DefaultContainerEnvs    []corev1.EnvVar
	HeadSidecarContainers   []corev1.Container
	WorkerSidecarContainers []corev1.Container
}

// BuildPods replaces buildHeadPod / buildWorkerPod (raycluster_controller.go:1387-1433) for every create tuple of one RayCluster:
// clusterJSON is json.Marshal(instance) (metadata + spec are read), clusterHash is createHeadPod's hash ("" for none).  One native
// call; the container half of the manifests is assembled once per group inside the library.
func BuildPods(clusterJSON []byte, env *BuilderEnv, clusterHash string, creates []CreateTuple) ([]corev1.Pod, error) {
	if len(creates) == 0 {
		return nil, nil
	}
	var s strs
	defer s.release()
	ce := C.kr_podbuild_env{kuberay_version: s.str(env.KubeRayVersion), cluster_domain: s.str(env.ClusterDomain), cluster_hash: s.str(clusterHash),
		deterministic_head_name: b2u(env.DeterministicHeadName), gate_multihost_indexing: b2u(env.MultiHostIndexing), login_shell: b2u(env.LoginShell),
		no_init_container_injection: b2u(env.NoInitContainer), no_probes_injection: b2u(env.NoProbes)}
	if n := len(env.DefaultContainerEnvs); n > 0 {
		kv := make([]C.kr_kv, n)
		for i, e := range env.DefaultContainerEnvs {
			kv[i] = C.kr_kv{key: s.str(e.Name), value: s.present(e.Value)}
		}
		s.pin.Pin(&kv[0])
		ce.default_envs, ce.n_default_envs = &kv[0], C.uint32_t(n)
	}
	if len(env.HeadSidecarContainers) > 0 {
		b, _ := json.Marshal(env.HeadSidecarContainers)
		ce.head_sidecars_json = s.bytes(b)
	}
	if len(env.WorkerSidecarContainers) > 0 {
		b, _ := json.Marshal(env.WorkerSidecarContainers)
		ce.worker_sidecars_json = s.bytes(b)
	}
	tuples := make([]C.kr_podmeta_create, len(creates))
	for i, t := range creates {
		tuples[i] = C.kr_podmeta_create{group: C.int32_t(t.Group), replica_index: C.int32_t(t.ReplicaIndex), host_index: C.int32_t(t.HostIndex), replica_name: s.present(t.ReplicaName)}
	}
	s.pin.Pin(&tuples[0])
	s.pin.Pin(&clusterJSON[0])
	off := make([]C.uint64_t, len(creates)+1)
	var need C.uint64_t
	cj, n := (*C.uint8_t)(unsafe.Pointer(&clusterJSON[0])), C.uint64_t(len(clusterJSON))
	rc := C.kr_pod_build(cj, n, &ce, &tuples[0], C.uint32_t(len(tuples)), nil, 0, &off[0], &need) // size
	if rc != C.KR_OK && rc != C.KR_E_CAPACITY {
		return nil, errors.New(C.GoString(C.kr_pod_build_last_error()))
	}
	buf := make([]byte, need)
	if rc = C.kr_pod_build(cj, n, &ce, &tuples[0], C.uint32_t(len(tuples)), (*C.uint8_t)(unsafe.Pointer(&buf[0])), need, &off[0], &need); rc != C.KR_OK {
		return nil, errors.New(C.GoString(C.kr_pod_build_last_error()))
	}
	pods := make([]corev1.Pod, len(creates))
	for i := range pods {
		if err := json.Unmarshal(buf[off[i]:off[i+1]], &pods[i]); err != nil {
			return nil, err
		}
	}
	return pods, nil
}

// SpecJSON writes json.Marshal(mute(spec)) — the bytes utils.GenerateHashWithoutReplicasAndWorkersToDelete hashes (utils/util.go:642-665) —
// for a RayClusterSpec given as JSON text in any key order.
func SpecJSON(spec []byte) ([]byte, error) {
	if len(spec) == 0 {
		return nil, errors.New("krengine: empty spec")
	}
	out := make([]byte, 2*len(spec)+256)
	for {
		var n C.uint64_t
		rc := C.kr_spec_json_emit((*C.uint8_t)(unsafe.Pointer(&spec[0])), C.uint64_t(len(spec)), 0, (*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(len(out)), &n)
		if rc == C.KR_E_CAPACITY {
			out = make([]byte, n)
			continue
		}
		if rc != C.KR_OK {
			return nil, errors.New(C.GoString(C.kr_spec_json_last_error()))
		}
		return out[:n], nil
	}
}
