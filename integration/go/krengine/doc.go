// Package krengine is the cgo shim between the KubeRay operator and libkrengine.so (include/kr_engine.h): the batched B200 engine for
// RayClusterReconciler.reconcilePods / calculateStatus, the event-driven packer that feeds it, and the host-side Pod builder.
//
// Where it goes: ray-operator/controllers/ray/krengine/ in the KubeRay tree, built with
//
//	CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/kuberay_b200 -lkrengine -Wl,-rpath,<repo>/kuberay_b200" go build ./...
//
// It needs Go >= 1.21 (runtime.Pinner).  This image has no Go toolchain, so the package is source that has not been compiled here; the
// executable statement of the same protocol — the same C calls in the same order, the same handling of every record — is the Python mirror
// kuberay_b200/{engine,packer,reconciler,podbuilder}.py, which the test suite drives against the CPU oracle and on the GPU.
//
// cgo pointer rules: every struct handed to C that carries string pointers (kr_str) is filled through a scope (see strs), which pins the
// strings' backing arrays for the duration of the call; C never retains a Go pointer — the packer interns what it needs before returning.
// The large arenas go the other way: they are allocated by C (cudaHostAlloc) and Go only holds slices over them.
//
// Threading: one goroutine (Batcher.run) owns the packer and the engine.  Informer event handlers and Reconcile(req) calls talk to it
// through channels; nothing else calls into C.
package krengine
