// kr_kernels.cuh — sm_100a kernels of the batched reconcile engine.
//
// Integer / hash work: no tensor cores.  What matters here: coalesced SoA column streaming,
// shared-memory staging (hash chunks, per-tile digit counters, per-warp group accumulators), warp-ballot /
// match_any group-by, grids sized in multiples of the SM count.
//
// Pipeline of one pass (engine stream M unless noted); one CUDA graph, programmatic dependent launch along the chain:
//   k_clear          per-pass clears (hash tables, workersToDelete resolutions, totals, bucket counters) in one launch
//   k_build_tables   cluster table (ns,name)->idx (+ per-cluster group record), workersToDelete-name table, head-aux table
//   k_match          per pod: label/selector match -> cluster idx + group slot, 16-byte pod row, bucket rank
//   k_place_fused    bucket starts (scan in shared memory) + pod -> slot of its cluster's bucket   [large: k_scan_counts + k_place]
//   k_decide_small   one warp per RayCluster (<= 256 pods): in-register bitonic sort (List order), warp-ballot group-by,
//                    head / group decisions, ordered deletes, status roll-up        | k_decide: general path, side stream
//   k_hash2          (stream H, concurrent) SHA-1 + base32hex of every muted-spec JSON
//   k_decide phase 1 clusters whose Recreate gate needs the hash
//   k_creates_fused  create offsets + lowest free replica indices + compact action list   [large: k_scan_* + k_create_fill ...]
//   k_jobs           RayJob -> RayCluster status roll-up join
//   k_patch_pods     (copy stream, incremental epochs) rewritten pod rows pulled from the mapped pinned arena
//   radix pipeline   (k_match<radix>, k_hist, k_scan_rows, k_scatter): stable LSD sort, taken when a RayCluster has > 1024 pods
//
// Reference semantics restated here are cited per function (paths relative to
// ray-operator/controllers/ray/ in ray-project/kuberay).
#pragma once

#include "kr_common.cuh"
#include "kr_match.cuh"
#include "kr_bucket.cuh"
#include "kr_decide.cuh"
#include "kr_emit.cuh"
#include "kr_bucket2.cuh"
#include "kr_hash.cuh"
