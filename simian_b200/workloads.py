"""Workload (model-script) builders for the shapes BASELINE.json names.

Each builder plays the role of a reference model script
(SimianJS/Examples/phold-noop-noMPI.js:44-49: addEntity loop + schedService
loop) against any low-level engine handle exposing
define_class / set_class_params / add_entities / attach_service /
sched_service_bulk -- i.e. simian_b200.capi.Engine.
"""
import struct

PHOLD_UNIFORM, PHOLD_ROSS_REMOTE, PHOLD_OTHER_GROUP = 0, 1, 2


def phold_params_blob(first_id, n_entities, p_remote, lam, lookahead, mode,
                      groups):
    """Pack simian_phold_params_t (include/simian_models.h), 48 bytes."""
    blob = struct.pack("<IIdddIIII", first_id, n_entities, p_remote, lam,
                       lookahead, mode, groups, 0, 0)
    assert len(blob) == 48
    return blob


def build_phold(eng, n, per_entity, lookahead, p_remote=0.0, lam=1.0,
                mode=PHOLD_UNIFORM, groups=1, class_name="Node",
                service="generate", t0=0.0):
    """`count` Node entities, `per_entity` initial generate events each at t0."""
    eng.define_class(class_name, 0)
    eng.add_entities(class_name, 0, n)
    eng.attach_service(class_name, service, "phold_generate")
    eng.set_class_params(
        class_name,
        phold_params_blob(eng.first_id(class_name), n, p_remote, lam,
                          lookahead, mode, groups))
    eng.sched_service_bulk(t0, service, 0, class_name, 0, n, per_entity)
    return eng


# the five BASELINE.json configs
CONFIGS = {
    "config1_phold_1k": dict(n=1000, per_entity=1, min_delay=1e-4,
                             lookahead=1e-4, p_remote=0.0, mode=PHOLD_UNIFORM,
                             groups=1, end=1000.0),
    "config2_phold_1m": dict(n=1 << 20, per_entity=16, min_delay=0.1,
                             lookahead=0.1, p_remote=0.1,
                             mode=PHOLD_ROSS_REMOTE, groups=1, end=10.0),
    "config3_phold_16m": dict(n=1 << 24, per_entity=16, min_delay=0.1,
                              lookahead=0.1, p_remote=0.5,
                              mode=PHOLD_OTHER_GROUP, groups=8, end=5.0),
    # LANL PDES benchmark V8-style: per-event compute over
    # an 8-KB list per entity, geometric hot-spot destinations
    "config4_lanl": dict(kind="lanl", n_ent=1 << 20, s_ent=100.0, q_avg=1.0,
                         p_receive=1e-5, m_ent=1000.0, ops_ent=1000.0,
                         ops_sigma=0.1, cache_friendliness=0.5, end_time=100.0,
                         min_delay=1.0, time_bins=10,
                         # the geometric destinations make the first entity blocks hot
                         # spots (block 0 of 512: ten times the average events per
                         # window, several passes of one CTA): blocks of 256 halve the
                         # critical CTA -- measured 566 -> 421 ms per run (128: 474 ms)
                         options=dict(block_entities=256)),
    # tiny windows: 0.16 due events per entity per window.  A CTA's time is a fixed
    # chain of memory round trips, so blocks of 1024 entities (~160 due events each)
    # and two calendar rows per window instead of eight: measured 3.75 -> 5.1 G
    # events/s on one B200
    "config5_phold_4m_small_lookahead": dict(
        n=1 << 22, per_entity=16, min_delay=0.01, lookahead=0.01, p_remote=0.1,
        mode=PHOLD_ROSS_REMOTE, groups=1, end=2.0,
        options=dict(bucket_width=0.005, block_entities=1024)),
}


# ---- LANL PDES benchmark (SimianJS/Examples/pdes_lanl_benchmarkV8.js) --------

import math  # noqa: E402

import numpy as np  # noqa: E402

LANL_CLASS = "PDES_LANL_Node"
LANL_TIME_BINS = 16  # SIMIAN_LANL_TIME_BINS
# simian_lanl_state_t (include/simian_models.h), 288 bytes; the list follows it
LANL_STATE_DTYPE = np.dtype([
    ("local_intersend_delay", "<f8"), ("q_target", "<f8"),
    ("last_scheduled", "<f8"), ("ops", "<f8"), ("ops_max", "<f8"),
    ("ops_min", "<f8"), ("ops_mean", "<f8"), ("value_sink", "<f8"),
    ("q_size", "<i8"), ("send_count", "<u8"), ("receive_count", "<u8"),
    ("ops_count", "<u8"), ("list_size", "<u4"), ("active_elements", "<u4"),
    ("time_sends", "<u4", (LANL_TIME_BINS,)),
    # the global statistics entity 0 collects when the fan-in runs on the event path
    ("stats_received", "<u8"), ("gsend_count", "<u8"), ("greceive_count", "<u8"),
    ("gops_count", "<u8"), ("gops_max", "<f8"), ("gops_min", "<f8"), ("gops_mean", "<f8"),
    ("gtime_sends", "<u4", (LANL_TIME_BINS,))])
LANL_STATE_BYTES = LANL_STATE_DTYPE.itemsize
assert LANL_STATE_BYTES == 288


def lanl_list_size(n_ent, m_ent, p_list=0.0):
    """The LARGEST this.list_size (pdes_lanl_benchmarkV8.js:219-221): every
    entity's state has room for it.  p_list == 0: the uniform size; otherwise
    entity 0's (prob = p_list * (1-p_list)**0).  Evaluated in the same order as
    the handler: (prob * n_ent) * m_ent in f64."""
    prob = (1.0 / n_ent) if p_list == 0 else p_list
    return int(math.floor(prob * n_ent * m_ent))


def lanl_state_bytes(n_ent, m_ent, p_list=0.0):
    return LANL_STATE_DTYPE.itemsize + 8 * lanl_list_size(n_ent, m_ent, p_list)


def lanl_engine_end(end_time, min_delay):
    """max(endTime, 2) + 3*minDelay  (pdes_lanl_benchmarkV8.js:349)"""
    return max(end_time, 2) + 3 * min_delay


def lanl_params_blob(first_id, n_ent, s_ent, q_avg, p_receive, m_ent, ops_ent,
                     ops_sigma, cache_friendliness, end_time, min_delay,
                     time_bins, svc_send, svc_receive, svc_finish,
                     p_send=0.0, p_list=0.0, invert=False, svc_output=0,
                     stats_on_path=False):
    """Pack simian_lanl_params_t (include/simian_models.h), 120 bytes."""
    r_min = math.pow(1.0 - p_receive, n_ent)  # :196
    blob = struct.pack("<II12dIHHHHHH", first_id, n_ent, s_ent, q_avg, p_receive,
                       m_ent, ops_ent, ops_sigma, cache_friendliness, end_time,
                       min_delay, r_min, p_send, p_list, time_bins, svc_send,
                       svc_receive, svc_finish, svc_output, 1 if invert else 0,
                       1 if stats_on_path else 0)
    assert len(blob) == 120
    return blob


def build_lanl(eng, n_ent, s_ent=100.0, q_avg=1.0, p_receive=0.0, m_ent=1000.0,
               ops_ent=1000.0, ops_sigma=0.1, cache_friendliness=0.5,
               end_time=100.0, min_delay=1.0, time_bins=10, p_send=0.0,
               p_list=0.0, invert=False, stats_on_path=False):
    """The "MAIN" of pdes_lanl_benchmarkV8.js:346-356: one addEntity loop whose
    constructors schedule FinishHandler and run the first SendHandler.  The
    engine must have been created with end = lanl_engine_end(end_time,
    min_delay).  p_send / invert / p_list: the geometric source and list-size
    distributions (:209-221); every entity's state has room for the largest list.
    stats_on_path: FinishHandler sends the statistics message to OutputHandler of
    entity 0 as an event with a payload (:311-342; the engine needs option
    payloads = 1; models of a few dozen entities: all messages are due at entity 0
    in one window)."""
    assert time_bins <= LANL_TIME_BINS
    eng.define_class(LANL_CLASS, lanl_state_bytes(n_ent, m_ent, p_list))
    eng.add_entities(LANL_CLASS, 0, n_ent)
    eng.attach_service(LANL_CLASS, "SendHandler", "lanl_send")
    eng.attach_service(LANL_CLASS, "ReceiveHandler", "lanl_receive")
    eng.attach_service(LANL_CLASS, "FinishHandler", "lanl_finish")
    svc_output = 0
    if stats_on_path:
        eng.attach_service(LANL_CLASS, "OutputHandler", "lanl_output")
        svc_output = eng.intern_service("OutputHandler")
    ids = [eng.intern_service(s) for s in
           ("SendHandler", "ReceiveHandler", "FinishHandler")]
    eng.set_class_params(LANL_CLASS, lanl_params_blob(
        eng.first_id(LANL_CLASS), n_ent, s_ent, q_avg, p_receive, m_ent,
        ops_ent, ops_sigma, cache_friendliness, end_time, min_delay, time_bins,
        *ids, p_send=p_send, p_list=p_list, invert=invert, svc_output=svc_output,
        stats_on_path=stats_on_path))
    # each constructor sends FinishHandler + up to ceil(q_target) timers + 1 request;
    # with a geometric source distribution the busiest entity's q_target is
    # q_avg * n_ent * p_send (the sum over the entities stays about q_avg * n_ent)
    per = 3 + int(math.ceil(q_avg))
    extra = int(math.ceil(q_avg * n_ent * p_send)) if p_send else 0
    if stats_on_path:  # the messages of the fan-in: header + 7 + time_bins payload words each
        extra += n_ent * (8 + time_bins)
    eng.construct_entities(LANL_CLASS, "lanl_construct", n_ent * per + extra)
    return eng


def lanl_states(raw, n_ent, m_ent, p_list=0.0):
    """split simian_read_state bytes into (struct array, lists[n, max list_size])"""
    sb = lanl_state_bytes(n_ent, m_ent, p_list)
    a = np.frombuffer(bytes(raw), dtype=np.uint8).reshape(n_ent, sb)
    st = a[:, :LANL_STATE_BYTES].copy().view(LANL_STATE_DTYPE).reshape(n_ent)
    lists = a[:, LANL_STATE_BYTES:].copy().view("<f8").reshape(n_ent, -1)
    return st, lists


def lanl_report(state, time_bins, out):
    """The statistics table of the LANL PDES benchmark, in the reference's
    on-disk format (OutputHandler, pdes_lanl_benchmarkV8.js:317-341).  The
    reference fans the per-entity statistics in to entity 0 with n_ent events
    at one timestamp (an N-way distinguishable tie, order = heap shape); here
    the fan-in is done off the event path from the entity state read back from
    HBM (simian_read_state), in entity-id order.  `state`: the struct array of
    lanl_states(); `out`: a file-like object."""
    n = len(state)
    out.write("%10s %10s %10s %10s %10s %10s    '%s'\n" % (
        "#EntityID", "#Sends", "#Receives", "Ops(Min", "Avg", "Max)", "Time Bin Sends"))
    g_send = g_recv = g_ops = 0
    g_mean, g_min, g_max = 0.0, float("inf"), 0.0
    g_bins = np.zeros(time_bins, dtype=np.int64)
    for i in range(n):
        r = state[i]
        sends, recvs = int(r["send_count"]), int(r["receive_count"])
        bins = [int(x) for x in r["time_sends"][:time_bins]]
        g_mean = (recvs * float(r["ops_mean"]) + g_mean * g_recv) / max(1, g_recv + recvs)
        g_send += sends
        g_recv += recvs
        g_ops += int(r["ops_count"])
        g_min = min(g_min, float(r["ops_min"]))
        g_max = max(g_max, float(r["ops_max"]))
        g_bins += np.array(bins, dtype=np.int64)
        out.write("%10d %10d %10d %10d %10.5g %10d    %s\n" % (
            i, sends, recvs, _js_int(r["ops_min"]), float(r["ops_mean"]), _js_int(r["ops_max"]),
            "[" + ",".join(map(str, bins)) + "]"))
    if n and int(state[0]["stats_received"]) == n:
        # the fan-in ran ON the event path (stats_on_path): the totals are the ones
        # OutputHandler of entity 0 accumulated (pdes_lanl_benchmarkV8.js:321-330)
        e0 = state[0]
        g_send, g_recv, g_ops = (int(e0["gsend_count"]), int(e0["greceive_count"]),
                                 int(e0["gops_count"]))
        g_mean, g_min, g_max = float(e0["gops_mean"]), float(e0["gops_min"]), float(e0["gops_max"])
        g_bins = np.array([int(x) for x in e0["gtime_sends"][:time_bins]], dtype=np.int64)
    out.write("===================== LANL PDES BENCHMARK  Collected Stats from All Ranks "
              "=======================\n")
    out.write("%10s %10s %10s %10s %10s %10s %10s    '%s'\n" % (
        "#Entities", "#Sends", "#Receives", "OpsCount", "Ops(Min", "Avg", "Max)", "Time Bin Sends"))
    out.write("%10d %10d %10d %10d %10d %10.5g %10d    %s\n" % (
        n, g_send, g_recv, g_ops, _js_int(g_min), g_mean, _js_int(g_max),
        "[" + ",".join(str(int(x)) for x in g_bins) + "]"))
    out.write("=" * 97 + "\n")
    return dict(sends=g_send, receives=g_recv, ops_count=g_ops, ops_min=g_min,
                ops_mean=g_mean, ops_max=g_max, time_sends=g_bins)


def _js_int(x):
    """sprintf.js %d of a double: truncation (parseInt); +/-inf and NaN print 0"""
    x = float(x)
    return int(x) if math.isfinite(x) else 0


# ---- hello.js: two entity classes, data-carrying events, replies to tx/txId ----

HELLO_SERVICES = ("generate", "square", "sqrt", "result")


def hello_params_blob(alice_first, bob_first, count, ids):
    """Pack simian_hello_params_t (include/simian_models.h), 24 bytes."""
    blob = struct.pack("<IIIIHHHH", alice_first, bob_first, count, 0, *ids)
    assert len(blob) == 24
    return blob


def build_hello(eng, count, t_alice=0.0, t_bob=50.0):
    """SimianJS/Examples/hello.js:96-109: `count` Alice and `count` Bob
    entities, one initial `generate` each (Alice at 0, Bob at 50)."""
    ids = [eng.intern_service(s) for s in HELLO_SERVICES]
    for cls in ("Alice", "Bob"):
        eng.define_class(cls, 0)
        eng.add_entities(cls, 0, count)
    eng.attach_service("Alice", "generate", "hello_generate")
    eng.attach_service("Alice", "square", "hello_square")
    eng.attach_service("Alice", "result", "hello_result")
    eng.attach_service("Bob", "generate", "hello_generate")
    eng.attach_service("Bob", "sqrt", "hello_sqrt")
    eng.attach_service("Bob", "result", "hello_result")
    blob = hello_params_blob(eng.first_id("Alice"), eng.first_id("Bob"), count, ids)
    eng.set_class_params("Alice", blob)
    eng.set_class_params("Bob", blob)
    eng.sched_service_bulk(t_alice, "generate", 0, "Alice", 0, count, 1)
    eng.sched_service_bulk(t_bob, "generate", 0, "Bob", 0, count, 1)
    return eng


# ---- payloads beyond 8 bytes (entity.js:52-60: `data` is any object) ------------

BLOB_SERVICES = ("send", "recv", "ack")


def blob_params_blob(first_id, n, lookahead, max_words, ids):
    """Pack simian_blob_params_t (include/simian_models.h), 32 bytes."""
    blob = struct.pack("<IIdIIHHHH", first_id, n, lookahead, max_words, 0, *ids, 0)
    assert len(blob) == 32
    return blob


def build_blobs(eng, n, lookahead, max_words):
    """The payload test model: every entity keeps sending a random-length vector of
    64-bit words to a random entity, which folds it and acknowledges the fold.
    CUDA engines need option payloads=1."""
    ids = [eng.intern_service(s) for s in BLOB_SERVICES]
    eng.define_class("Node", 0)
    eng.add_entities("Node", 0, n)
    eng.attach_service("Node", "send", "blob_send")
    eng.attach_service("Node", "recv", "blob_recv")
    eng.attach_service("Node", "ack", "blob_ack")
    eng.set_class_params("Node", blob_params_blob(eng.first_id("Node"), n, lookahead,
                                                   max_words, ids))
    eng.sched_service_bulk(0.0, "send", 0, "Node", 0, n, 1)
    return eng
