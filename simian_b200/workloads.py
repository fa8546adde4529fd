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
    "config5_phold_4m_small_lookahead": dict(
        n=1 << 22, per_entity=16, min_delay=0.01, lookahead=0.01, p_remote=0.1,
        mode=PHOLD_ROSS_REMOTE, groups=1, end=2.0),
}
