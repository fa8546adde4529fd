// napi_shim.cc -- N-API addon `simian_b200_napi`: every exported function is a
// thin 1:1 wrapper over one C symbol of include/simian_b200.h.  It plays the
// role the `masala` object plays for SimianJS (MasalaChai/masala.cpp:39-146):
// natives there return false + a pending exception; here a non-zero status
// becomes a thrown Error carrying simian_last_error().
//
// Source only: node_api.h / Node are absent from the build image (SURVEY.md
// appendix B), so this file is compiled where Node exists:
//   node-gyp (binding.gyp: sources [napi_shim.cc], libraries [-lsimian_b200])
#include <node_api.h>

#include <string>
#include <vector>

#include "../../include/simian_b200.h"

namespace {

#define NAPI_OK(call) \
  if ((call) != napi_ok) { napi_throw_error(env, nullptr, #call " failed"); return nullptr; }

std::string str(napi_env env, napi_value v) {
  size_t n = 0;
  napi_get_value_string_utf8(env, v, nullptr, 0, &n);
  std::string s(n, '\0');
  napi_get_value_string_utf8(env, v, &s[0], n + 1, &n);
  return s;
}
double num(napi_env env, napi_value v) { double d = 0; napi_get_value_double(env, v, &d); return d; }
simian_engine *eng(napi_env env, napi_value v) { void *p = nullptr; napi_get_value_external(env, v, &p); return (simian_engine *)p; }
napi_value undef(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }
napi_value check(napi_env env, simian_engine *e, int rc) {
  if (rc) { napi_throw_error(env, nullptr, simian_last_error(e)); return nullptr; }
  return undef(env);
}
#define ARGS(n) size_t argc = n; napi_value a[n]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr))

napi_value Create(napi_env env, napi_callback_info info) {  // simian_create
  ARGS(7);
  simian_engine *e = simian_create(str(env, a[0]).c_str(), num(env, a[1]), num(env, a[2]), num(env, a[3]),
                                   (int)num(env, a[4]), (int)num(env, a[5]), (int)num(env, a[6]));
  if (!e) { napi_throw_error(env, nullptr, simian_last_error(nullptr)); return nullptr; }
  napi_value ext; NAPI_OK(napi_create_external(env, e, nullptr, nullptr, &ext)); return ext;
}
napi_value Destroy(napi_env env, napi_callback_info info) { ARGS(1); simian_destroy(eng(env, a[0])); return undef(env); }
napi_value DefineClass(napi_env env, napi_callback_info info) {
  ARGS(3); simian_engine *e = eng(env, a[0]);
  int rc = simian_define_class(e, str(env, a[1]).c_str(), (uint32_t)num(env, a[2]));
  return check(env, e, rc < 0 ? -rc : 0);
}
napi_value SetBaseRank(napi_env env, napi_callback_info info) {  // a script's own getBaseRank
  ARGS(3); simian_engine *e = eng(env, a[0]);
  return check(env, e, simian_set_base_rank(e, str(env, a[1]).c_str(), (int)num(env, a[2])));
}
napi_value SetClassParams(napi_env env, napi_callback_info info) {
  ARGS(3); simian_engine *e = eng(env, a[0]); void *p; size_t n;
  NAPI_OK(napi_get_buffer_info(env, a[2], &p, &n));
  return check(env, e, simian_set_class_params(e, str(env, a[1]).c_str(), p, (uint32_t)n));
}
napi_value AddEntities(napi_env env, napi_callback_info info) {
  ARGS(4); simian_engine *e = eng(env, a[0]);
  return check(env, e, simian_add_entities(e, str(env, a[1]).c_str(), (uint32_t)num(env, a[2]), (uint32_t)num(env, a[3]), nullptr));
}
napi_value AttachService(napi_env env, napi_callback_info info) {
  ARGS(4); simian_engine *e = eng(env, a[0]);
  return check(env, e, simian_attach_service(e, str(env, a[1]).c_str(), str(env, a[2]).c_str(), str(env, a[3]).c_str()));
}
napi_value InternService(napi_env env, napi_callback_info info) {
  ARGS(2); napi_value r; napi_create_int32(env, simian_intern_service(eng(env, a[0]), str(env, a[1]).c_str()), &r); return r;
}
napi_value FirstId(napi_env env, napi_callback_info info) {
  ARGS(2); napi_value r; napi_create_int64(env, simian_first_id(eng(env, a[0]), str(env, a[1]).c_str()), &r); return r;
}
napi_value SchedService(napi_env env, napi_callback_info info) {
  ARGS(6); simian_engine *e = eng(env, a[0]);
  return check(env, e, simian_sched_service(e, num(env, a[1]), str(env, a[2]).c_str(), (uint64_t)num(env, a[3]),
                                            str(env, a[4]).c_str(), (uint32_t)num(env, a[5])));
}
napi_value SchedServiceBulk(napi_env env, napi_callback_info info) {
  ARGS(8); simian_engine *e = eng(env, a[0]);
  return check(env, e, simian_sched_service_bulk(e, num(env, a[1]), str(env, a[2]).c_str(), (uint64_t)num(env, a[3]),
                                                 str(env, a[4]).c_str(), (uint32_t)num(env, a[5]),
                                                 (uint32_t)num(env, a[6]), (uint32_t)num(env, a[7])));
}
napi_value SchedEvent(napi_env env, napi_callback_info info) {  // one POD record -> simian_sched_events
  ARGS(7); simian_engine *e = eng(env, a[0]); simian_event_t ev;
  ev.time = num(env, a[1]); ev.dst = (uint32_t)num(env, a[2]); ev.src = (uint32_t)num(env, a[3]);
  ev.handler = (uint16_t)num(env, a[4]); ev.flags = 0; ev.seq = (uint32_t)num(env, a[5]); ev.data = (uint64_t)num(env, a[6]);
  return check(env, e, simian_sched_events(e, &ev, 1));
}
napi_value SchedEvents(napi_env env, napi_callback_info info) {  // ArrayBuffer of 32-byte records (borrowed)
  ARGS(2); simian_engine *e = eng(env, a[0]); void *p; size_t n;
  NAPI_OK(napi_get_arraybuffer_info(env, a[1], &p, &n));
  return check(env, e, simian_sched_events(e, (const simian_event_t *)p, n / sizeof(simian_event_t)));
}
napi_value Run(napi_env env, napi_callback_info info) {
  ARGS(1); simian_engine *e = eng(env, a[0]); simian_stats st;
  if (simian_run(e, &st)) { napi_throw_error(env, nullptr, simian_last_error(e)); return nullptr; }
  napi_value o, v; napi_create_object(env, &o);
  napi_create_double(env, (double)st.total_events, &v); napi_set_named_property(env, o, "totalEvents", v);
  napi_create_double(env, (double)st.windows, &v); napi_set_named_property(env, o, "windows", v);
  napi_create_double(env, st.last_time, &v); napi_set_named_property(env, o, "lastTime", v);
  napi_create_double(env, st.run_seconds, &v); napi_set_named_property(env, o, "runSeconds", v);
  napi_create_double(env, st.device_seconds, &v); napi_set_named_property(env, o, "deviceSeconds", v);
  napi_create_bigint_uint64(env, st.digest, &v); napi_set_named_property(env, o, "digest", v);
  return o;
}
napi_value ReadCounts(napi_env env, napi_callback_info info) {  // -> BigUint64Array contents in an ArrayBuffer
  ARGS(4); simian_engine *e = eng(env, a[0]); uint32_t first = (uint32_t)num(env, a[2]), n = (uint32_t)num(env, a[3]);
  void *p; napi_value ab; NAPI_OK(napi_create_arraybuffer(env, (size_t)n * 8, &p, &ab));
  if (simian_read_counts(e, str(env, a[1]).c_str(), first, n, (uint64_t *)p)) { napi_throw_error(env, nullptr, simian_last_error(e)); return nullptr; }
  return ab;
}
napi_value ReadTraceHashes(napi_env env, napi_callback_info info) {
  ARGS(4); simian_engine *e = eng(env, a[0]); uint32_t first = (uint32_t)num(env, a[2]), n = (uint32_t)num(env, a[3]);
  void *p; napi_value ab; NAPI_OK(napi_create_arraybuffer(env, (size_t)n * 8, &p, &ab));
  if (simian_read_trace_hashes(e, str(env, a[1]).c_str(), first, n, (uint64_t *)p)) { napi_throw_error(env, nullptr, simian_last_error(e)); return nullptr; }
  return ab;
}

napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
      {"create", 0, Create, 0, 0, 0, napi_default, 0}, {"destroy", 0, Destroy, 0, 0, 0, napi_default, 0},
      {"defineClass", 0, DefineClass, 0, 0, 0, napi_default, 0}, {"setClassParams", 0, SetClassParams, 0, 0, 0, napi_default, 0},
      {"setBaseRank", 0, SetBaseRank, 0, 0, 0, napi_default, 0},
      {"addEntities", 0, AddEntities, 0, 0, 0, napi_default, 0}, {"attachService", 0, AttachService, 0, 0, 0, napi_default, 0},
      {"internService", 0, InternService, 0, 0, 0, napi_default, 0}, {"firstId", 0, FirstId, 0, 0, 0, napi_default, 0},
      {"schedService", 0, SchedService, 0, 0, 0, napi_default, 0}, {"schedServiceBulk", 0, SchedServiceBulk, 0, 0, 0, napi_default, 0},
      {"schedEvent", 0, SchedEvent, 0, 0, 0, napi_default, 0}, {"schedEvents", 0, SchedEvents, 0, 0, 0, napi_default, 0},
      {"run", 0, Run, 0, 0, 0, napi_default, 0}, {"readCounts", 0, ReadCounts, 0, 0, 0, napi_default, 0},
      {"readTraceHashes", 0, ReadTraceHashes, 0, 0, 0, napi_default, 0}};
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
