"""The multi-GPU coordinator of the C ABI (kr_group_*, kuberay_b200/csrc/kr_group.cpp).

kr_group_route is checked against the Python statement of the same routing (synthetic.shard_by_uid): every shard's pinned arena
must hold exactly the columns of that shard, byte for byte; every shard's pass must match the oracle on the shard's snapshot; the
all-gather of the per-group delta records must deliver every shard's records to the coordinator.  On a one-GPU box the shards
share device 0 (peer-copy exchange); with as many GPUs as shards each gets its own device and the exchange is an ncclAllGather
(tools/group_check.py runs that under `gpurun --gpus N`)."""
import numpy as np
import pytest

from kuberay_b200 import abi, synthetic
from kuberay_b200.engine import Group, lib

pytestmark = pytest.mark.gpu


def run_group(snap, flags, world, devices, oracle_mod):
    d = snap.dims
    cap = abi.kr_config(0, d["clusters"] + 1, d["groups"] + 1, d["wtd"] + 1, d["pods"] + 1, d["heads"] + 1, d["jobs"] + 1, max(1024, d["pods"]), d["json"] + 64)
    grp = Group(cap, devices)
    try:
        sizes, cs, cr, ps, pr = grp.route(snap)
        shards = [synthetic.shard_by_uid(snap, r, world) for r in range(world)]
        for r, sh in enumerate(shards):
            assert [getattr(sizes[r], f) for f, _ in abi.kr_sizes._fields_] == [getattr(sh.sizes(), f) for f, _ in abi.kr_sizes._fields_], r
        # where every global row went
        keep = [(snap.c_uid_hash % np.uint64(world)) == np.uint64(r) for r in range(world)]
        for r in range(world):
            assert np.array_equal(np.flatnonzero(cs == r), np.flatnonzero(keep[r])) and np.array_equal(cr[cs == r], np.arange(int(keep[r].sum())))
            assert np.array_equal(pr[ps == r], np.arange(int((ps == r).sum())))
        grp.commit()
        f0 = abi.kr_flags.from_buffer_copy(flags); f0.fetch_pod_lists = 0
        for f in (flags, f0):
            res = grp.reconcile(f)
            for r, sh in enumerate(shards):
                want = oracle_mod.run(sh, f, threads=4)
                dd = want.diff(res[r])
                assert not dd, (r, dd[:6])
        gathered, slot, used_nccl = grp.allgather_group_results()
        assert used_nccl == (len(set(devices)) == len(devices) and world > 1) or not used_nccl
        for r, sh in enumerate(shards):
            ng = sh.dims["groups"]
            got = gathered[r][:ng]
            for fld in ("expected", "n_list", "n_unhealthy", "n_running", "diff", "n_create", "flags"):
                assert np.array_equal(got[fld], res[r].groups[fld]), (r, fld)
            assert not gathered[r][ng:].view(np.uint8).any()
        return used_nccl
    finally:
        grp.close()


@pytest.mark.parametrize("world", [1, 3])
def test_group_routes_commits_reconciles_and_gathers_on_one_device(world, oracle_mod):
    snap, flags = synthetic.generate(synthetic.config("C5", wtd_group_frac=0.3, jobs=True, n_clusters=400, pods_per_cluster=40))
    run_group(snap, flags, world, [0] * world, oracle_mod)


def test_group_one_shard_per_device_when_the_box_has_several(oracle_mod):
    n = lib().kr_device_count()
    if n < 2:
        pytest.skip("one GPU: covered by the shared-device test; tools/group_check.py runs this under gpurun --gpus N")
    snap, flags = synthetic.generate(synthetic.config("C5", wtd_group_frac=0.3))
    used = run_group(snap, flags, n, list(range(n)), oracle_mod)
    assert used, "every shard on its own device: the exchange should be an ncclAllGather"
