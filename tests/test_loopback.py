"""All ranks of a world inside ONE process (simian_b200.distributed.
drive_windows_loopback): the multi-rank paths of the engine -- outboxes, ingest
of peer records, LBTS candidate, advance -- on a single GPU, where the NCCL /
NVLink tests (test_distributed.py) need two.  Every rank's entities are
compared bit-for-bit with a single-rank oracle run of the same model (results
are placement invariant)."""
import numpy as np
import pytest

from dist_worker import CASES, build, engine_options, read_all, shape
from simian_b200.distributed import drive_windows_loopback


def _reference(om, c, world):
    end, min_delay, classes, tie_mode = shape(c)
    ref = om.Oracle("ref", 0.0, end, min_delay, size=1, seed=c["seed"], tie_mode=tie_mode)
    build(ref, c, world)
    st = ref.run()
    cnt, h = read_all(ref, classes)
    ref.close()
    if tie_mode == 0:
        assert st["distinguishable_ties"] == 0
    return st, cnt, h


def _check(world, stats, results, owners, ref):
    st_ref, cnt_ref, h_ref = ref
    seen = np.zeros(len(cnt_ref), dtype=bool)
    for rank in range(world):
        cnt, h = results[rank]
        mine = owners[rank]
        assert mine.sum() > 0 and not (seen & mine).any()
        seen |= mine
        np.testing.assert_array_equal(cnt[mine], cnt_ref[mine])
        np.testing.assert_array_equal(h[mine], h_ref[mine])
        assert stats[rank]["windows"] == st_ref["windows"]
    assert seen.all()
    assert sum(s["total_events"] for s in stats) == st_ref["total_events"]
    assert sum(s["digest"] for s in stats) & 0xFFFFFFFFFFFFFFFF == st_ref["digest"]
    assert sum(s["emitted"] for s in stats) == st_ref["emitted"]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["ross", "hello"])
def test_loopback_driver_with_oracle_ranks(name, world, oracle_mod):
    """the loopback driver itself, on the CPU: one oracle world per rank, as the
    gloo processes have"""
    om = oracle_mod
    c = CASES[name]
    end, min_delay, classes, tie_mode = shape(c)
    ref = _reference(om, c, world)
    worlds, ranks = [], []
    for rank in range(world):
        o = om.Oracle("dist", 0.0, end, min_delay, size=world, seed=c["seed"], tie_mode=tie_mode)
        build(o, c, world)
        worlds.append(o)
        ranks.append(om.OracleRankEngine(o, rank, world))
    stats = drive_windows_loopback(ranks)
    results, owners = [], []
    for rank, o in enumerate(worlds):
        results.append(read_all(o, classes))
        gid0, mine = 0, []
        for k, n in classes:
            mine += [o.L.oracle_owner_rank(o.h, gid0 + g) == rank for g in range(n)]
            gid0 += n
        owners.append(np.array(mine))
        o.close()
    _check(world, stats, results, owners, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["ross", "other_group", "uniform_sparse", "lanl", "hello", "blobs"])
def test_cuda_ranks_share_one_gpu(name, world, oracle_mod):
    from simian_b200 import capi
    from simian_b200.distributed import GpuRankEngine
    c = CASES[name]
    end, min_delay, classes, tie_mode = shape(c)
    ref = _reference(oracle_mod, c, world)
    opts = engine_options(c)
    engines = []
    for rank in range(world):
        e = capi.Engine("dist", 0.0, end, min_delay, rank=rank, size=world, device=0,
                        seed=c["seed"], **opts)
        build(e, c, world)
        engines.append(e)
    stats = drive_windows_loopback([GpuRankEngine(e, "cuda:0") for e in engines])
    results, owners = [], []
    for rank, e in enumerate(engines):
        assert stats[rank]["error"] == 0
        results.append(read_all(e, classes))
        owners.append(np.concatenate([(e.base_rank(k) + np.arange(n)) % world == rank
                                      for k, n in classes]))
        e.close()
    _check(world, stats, results, owners, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("deferred", [0, 1], ids=["pull_all", "deferred_pull"])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("name", ["ross", "other_group", "uniform_sparse", "lanl", "hello", "blobs"])
def test_device_driven_exchange_all_ranks_on_one_gpu(name, world, deferred, oracle_mod):
    """The data plane the multi-GPU numbers come from -- k_window(fuse = 3: LBTS
    candidate published into the peers' exchange slots) -> k_pull (flags, then
    the peers' outboxes) -> k_advance(minimum over the slots) -- with every rank
    of a 2/3/4/8-rank world on ONE device (drive_windows_p2p_loopback: peers wired
    by device pointers, one stream).  Bit-exact per-entity counts and trace hashes
    against a single-rank oracle run.  deferred_pull: two-ended outboxes -- the
    near end pulled by k_pull before the advance step, the far end by the first
    CTAs of the next k_window launch (uniform_sparse and hello jump over idle
    time, so the near pull also takes far records below the next bound)."""
    import torch
    from simian_b200 import capi
    from simian_b200.distributed import drive_windows_p2p_loopback
    c = CASES[name]
    end, min_delay, classes, tie_mode = shape(c)
    ref = _reference(oracle_mod, c, world)
    opts = engine_options(c)
    engines = []
    for rank in range(world):
        # (worlds of 3 and 8: the LBTS candidate by a launch of its own, as the engine
        # does by default from 592 entity blocks on)
        e = capi.Engine("p2p", 0.0, end, min_delay, rank=rank, size=world, device=0,
                        seed=c["seed"], deferred_pull=deferred,
                        fused_advance=0 if world in (3, 8) else 1, **opts)
        build(e, c, world)
        engines.append(e)
    stats = drive_windows_p2p_loopback(engines, torch.device("cuda", 0))
    results, owners = [], []
    for rank, e in enumerate(engines):
        assert stats[rank]["error"] == 0
        results.append(read_all(e, classes))
        owners.append(np.concatenate([(e.base_rank(k) + np.arange(n)) % world == rank
                                      for k, n in classes]))
    for e in engines:
        e.close()
    _check(world, stats, results, owners, ref)


# ---- a model script's own getBaseRank (simian.js:192-194) ------------------------

def _owners_with_base(classes, bases, world):
    return [np.concatenate([(bases[k] + np.arange(n)) % world == rank for k, n in classes])
            for rank in range(world)]


@pytest.mark.parametrize("world", [2, 3])
def test_user_base_rank_with_oracle_ranks(world, oracle_mod):
    """results are placement invariant: any base rank gives the single-rank answer"""
    om = oracle_mod
    c = CASES["hello"]
    end, min_delay, classes, tie_mode = shape(c)
    ref = _reference(om, c, world)
    bases = {"Alice": 1, "Bob": world - 1}
    worlds, ranks = [], []
    for rank in range(world):
        o = om.Oracle("dist", 0.0, end, min_delay, size=world, seed=c["seed"], tie_mode=tie_mode)
        for k in bases:
            o.define_class(k, 0)
            o.set_base_rank(k, bases[k])
        build(o, c, world)
        assert [o.base_rank(k) for k in bases] == [bases[k] % world for k in bases]
        worlds.append(o)
        ranks.append(om.OracleRankEngine(o, rank, world))
    stats = drive_windows_loopback(ranks)
    results = [read_all(o, classes) for o in worlds]
    for o in worlds:
        o.close()
    _check(world, stats, results, _owners_with_base(classes, bases, world), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_user_base_rank_cuda(world, oracle_mod):
    from simian_b200 import capi
    from simian_b200.distributed import GpuRankEngine
    c = CASES["hello"]
    end, min_delay, classes, tie_mode = shape(c)
    ref = _reference(oracle_mod, c, world)
    bases = {"Alice": 1, "Bob": world - 1}
    engines = []
    for rank in range(world):
        e = capi.Engine("dist", 0.0, end, min_delay, rank=rank, size=world, device=0,
                        seed=c["seed"], bucket_width=1.0)
        for k in bases:
            e.define_class(k, 0)
            e.set_base_rank(k, bases[k])
        build(e, c, world)
        assert [e.base_rank(k) for k in bases] == [bases[k] % world for k in bases]
        engines.append(e)
    stats = drive_windows_loopback([GpuRankEngine(e, "cuda:0") for e in engines])
    results = [read_all(e, classes) for e in engines]
    with pytest.raises(capi.SimianError):   # only before the class's entities are added
        engines[0].set_base_rank("Alice", 0)
    for e in engines:
        e.close()
    _check(world, stats, results, _owners_with_base(classes, bases, world), ref)
