"""-m gpu: BASELINE config 2 at its FULL size (2^20 entities x 16 events, 1.7e8
committed events): the full run against the oracle (its ranks on the host's
threads: seconds), and size-independent properties.

  * the oracle itself: per-entity committed counts and order-sensitive trace
    hashes of all 2^20 entities, digest, window count, emitted / dropped / tie
    counts, f64 final time -- bit for bit;
  * a checksum of checksums: the digest the device reduces equals the sum of
    simian_digest_term over the per-entity counts and trace hashes read back,
    recomputed here (include/simian_spec.h, transcribed to numpy);
  * conservation: every committed PHOLD event sends exactly one event, which is
    committed later, still pending, or dropped past endTime;
  * determinism and schedule independence: a second run, and a run with another
    entity-block size (other CTAs, other append order into the calendar pages),
    give the same per-entity counts and order-sensitive trace hashes;
  * placement invariance: the same model on a world of two ranks (both engines on
    this GPU, all ranks stepped in one process) gives the same entities.
The small-size bit-exact comparisons against the oracle are in test_parity_gpu.py."""
import numpy as np
import pytest

from simian_b200.workloads import CONFIGS, build_phold

pytestmark = pytest.mark.gpu

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
C1 = np.uint64(0xD1342543DE82EF95)


def mix64(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def digest_of(counts, hashes, first_id=0):
    """sum over entities of simian_digest_term (wrapping)"""
    with np.errstate(over="ignore"):
        e = np.arange(first_id, first_id + len(counts), dtype=np.uint64)
        term = (mix64(hashes ^ mix64((e << np.uint64(1)) | np.uint64(1))) +
                mix64(counts + C1 * (e + np.uint64(1))))
        return int(term.sum(dtype=np.uint64))


def run(cfg, **options):
    from simian_b200 import capi
    e = capi.Engine("full", 0.0, cfg["end"], cfg["min_delay"], seed=1, **options)
    build_phold(e, cfg["n"], cfg["per_entity"], cfg["lookahead"], cfg["p_remote"], 1.0,
                cfg["mode"], cfg["groups"])
    st = e.run()
    counts = e.read_counts("Node", cfg["n"])
    hashes = e.read_trace_hashes("Node", cfg["n"])
    e.close()
    return st, counts, hashes


@pytest.fixture(scope="module")
def full():
    cfg = dict(CONFIGS["config2_phold_1m"])
    return cfg, run(cfg)


def test_digest_is_the_checksum_of_the_per_entity_checksums(full):
    cfg, (st, counts, hashes) = full
    assert st["error"] == 0 and st["distinguishable_ties"] == 0
    assert st["total_events"] > 1.6e8
    assert int(counts.sum()) == st["total_events"]
    assert digest_of(counts, hashes) == st["digest"]


def test_full_size_against_the_oracle(full, oracle_mod):
    """config 2 at BASELINE's own size against the CPU restatement of the
    reference algorithm (one rank per host thread, proven equal to the
    sequential oracle in test_oracle.py): every entity, bit for bit"""
    import os
    cfg, (st, counts, hashes) = full
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 32))
    o = oracle_mod.Oracle("full", 0.0, cfg["end"], cfg["min_delay"], size=cores, seed=1,
                          threads=cores)
    build_phold(o, cfg["n"], cfg["per_entity"], cfg["lookahead"], cfg["p_remote"], 1.0,
                cfg["mode"], cfg["groups"])
    so = o.run()
    co, ho = o.read_counts("Node", cfg["n"]), o.read_trace_hashes("Node", cfg["n"])
    o.close()
    assert so["distinguishable_ties"] == 0 and st["distinguishable_ties"] == 0
    for k in ("total_events", "windows", "digest", "emitted", "dropped_past_end", "ties"):
        assert st[k] == so[k], k
    assert float(st["last_time"]).hex() == float(so["last_time"]).hex()
    np.testing.assert_array_equal(counts, co)
    np.testing.assert_array_equal(hashes, ho)


def test_event_conservation(full):
    cfg, (st, counts, hashes) = full
    initial = cfg["n"] * cfg["per_entity"]
    # one send per committed event (phold-noop-noMPI.js:36-42): sent = kept + dropped
    assert st["emitted"] + st["dropped_past_end"] == st["total_events"]
    # committed events are initial or emitted ones; what is left is still pending
    pending = initial + st["emitted"] - st["total_events"]
    assert 0 <= pending <= initial + st["emitted"]
    # nothing at or before endTime is left behind: the run ends with gmin > endTime
    assert st["last_time"] <= cfg["end"]
    assert st["windows"] == 100  # endTime / minDelay windows of a dense model


def test_same_results_on_another_schedule(full):
    cfg, (st, counts, hashes) = full
    for options in ({}, {"block_entities": 480}, {"pdl": 0, "launch_batch": 1}):
        st2, c2, h2 = run(cfg, **options)
        assert st2["total_events"] == st["total_events"], options
        assert st2["digest"] == st["digest"], options
        assert float(st2["last_time"]).hex() == float(st["last_time"]).hex()
        np.testing.assert_array_equal(c2, counts)
        np.testing.assert_array_equal(h2, hashes)


def test_placement_invariance_two_ranks_on_one_gpu(full):
    from simian_b200 import capi
    from simian_b200.distributed import GpuRankEngine, drive_windows_loopback
    cfg, (st, counts, hashes) = full
    world, engines = 2, []
    for rank in range(world):
        e = capi.Engine("full2", 0.0, cfg["end"], cfg["min_delay"], rank=rank, size=world,
                        device=0, seed=1)
        build_phold(e, cfg["n"], cfg["per_entity"], cfg["lookahead"], cfg["p_remote"], 1.0,
                    cfg["mode"], cfg["groups"])
        engines.append(e)
    stats = drive_windows_loopback([GpuRankEngine(e, "cuda:0") for e in engines])
    assert sum(s["total_events"] for s in stats) == st["total_events"]
    assert sum(s["digest"] for s in stats) & 0xFFFFFFFFFFFFFFFF == st["digest"]
    for rank, e in enumerate(engines):
        assert stats[rank]["error"] == 0 and stats[rank]["windows"] == st["windows"]
        loc_c, first, stride = e.read_local("Node", 0)
        loc_h, _, _ = e.read_local("Node", 1)
        np.testing.assert_array_equal(loc_c, counts[first::stride])
        np.testing.assert_array_equal(loc_h, hashes[first::stride])
        e.close()
