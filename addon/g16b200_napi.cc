// g16b200_napi.cc -- Node N-API addon over the C ABI of libg16b200.so (include/g16b200.h).
//
// The host driver of the reference is TypeScript: client/proof.helper.ts:28-72 writes Prover.toml, runs
// `nargo execute`, then shells out to `sunspot prove` (line 64) and reads <name>.proof / <name>.pw back.
// This addon replaces that child process by in-process calls; addon/client/proof.helper.ts is the drop-in
// module with the reference's `generateProof(config, inputs)` signature.
//
// STATUS: source only.  The build image has neither `node` nor `node_api.h` (probed), so this file has not
// been compiled; it uses only the stable C N-API (v8) and the functions declared in g16b200.h.
//
// JS surface:
//   loadCircuit(ccs: Buffer, pk: Buffer, device = 0)                -> Circuit handle (key resident on the GPU)
//   proveSync(circuit, witnessGz: Buffer)                           -> { proof: Buffer, publicWitness: Buffer }
//   prove(circuit, witnessGz: Buffer)                               -> Promise of the same (libuv worker thread)
//   proveBatch(circuit, assignments: Buffer, n: number)             -> Promise<{ proofs: Buffer, publicWitnesses: Buffer, pwStride }>
//   verify(vk: Buffer, proof: Buffer, publicWitness: Buffer)        -> boolean   (host only, `sunspot verify`)
//   setup(ccs: Buffer, seed: Buffer, device = 0)                    -> { pk: Buffer, vk: Buffer }   (`sunspot setup`)
#include <node_api.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "g16b200.h"

namespace {

std::mutex g_ctx_mutex;
std::map<int, g16_ctx*> g_ctx;   // one context per device, created on first use

g16_ctx* context_for(int device, std::string* err) {
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) return it->second;
  g16_ctx* ctx = nullptr;
  if (g16_init(&device, 1, &ctx) != G16_OK) {
    *err = g16_last_error();
    return nullptr;
  }
  g_ctx[device] = ctx;
  return ctx;
}

napi_value throw_error(napi_env env, const std::string& msg) {
  napi_throw_error(env, "G16B200", msg.c_str());
  return nullptr;
}

bool buffer_arg(napi_env env, napi_value v, const uint8_t** data, size_t* len) {
  bool is_buf = false;
  if (napi_is_buffer(env, v, &is_buf) != napi_ok || !is_buf) return false;
  void* p = nullptr;
  if (napi_get_buffer_info(env, v, &p, len) != napi_ok) return false;
  *data = static_cast<const uint8_t*>(p);
  return true;
}

napi_value make_buffer(napi_env env, const uint8_t* data, size_t len) {
  napi_value out;
  void* copy = nullptr;
  napi_create_buffer_copy(env, len, data, &copy, &out);
  return out;
}

struct Circuit {
  g16_circuit* handle = nullptr;
  size_t n_values = 0;   // public (without ONE) + secret values of one assignment
  size_t pw_len = 0;
  size_t proof_len = G16_PROOF_LEN;
  std::mutex busy;       // a circuit handle runs one proving call at a time
};

void circuit_finalize(napi_env, void* data, void*) {
  Circuit* c = static_cast<Circuit*>(data);
  if (c->handle) g16_circuit_free(c->handle);
  delete c;
}

napi_value LoadCircuit(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *ccs, *pk;
  size_t nccs, npk;
  if (argc < 2 || !buffer_arg(env, argv[0], &ccs, &nccs) || !buffer_arg(env, argv[1], &pk, &npk))
    return throw_error(env, "loadCircuit(ccs: Buffer, pk: Buffer, device?: number)");
  int32_t device = 0;
  if (argc > 2) napi_get_value_int32(env, argv[2], &device);
  std::string err;
  g16_ctx* ctx = context_for(device, &err);
  if (!ctx) return throw_error(env, err);
  Circuit* c = new Circuit();
  if (g16_circuit_load(ctx, ccs, nccs, pk, npk, nullptr, &c->handle) != G16_OK) {
    delete c;
    return throw_error(env, g16_last_error());
  }
  uint64_t what[16];
  g16_circuit_info(c->handle, what);
  c->n_values = (size_t)(what[2] - 1 + what[3]);
  c->pw_len = 12 + 32 * (size_t)(what[2] - 1);
  c->proof_len = what[5] ? 388 : 324;
  napi_value out;
  napi_create_external(env, c, circuit_finalize, nullptr, &out);
  return out;
}

struct ProveResult {
  int rc = G16_OK;
  std::string err;
  std::vector<uint8_t> proof, pw;
};

void run_prove(Circuit* c, const std::vector<uint8_t>& gz, ProveResult* r) {
  r->proof.resize(c->proof_len);
  r->pw.resize(c->pw_len);
  size_t pl = r->proof.size(), wl = r->pw.size();
  std::lock_guard<std::mutex> lock(c->busy);
  r->rc = g16_prove(c->handle, gz.data(), gz.size(), nullptr, r->proof.data(), &pl, r->pw.data(), &wl);
  if (r->rc != G16_OK) r->err = g16_last_error();
  r->proof.resize(pl);
  r->pw.resize(wl);
}

napi_value result_object(napi_env env, const ProveResult& r) {
  napi_value obj;
  napi_create_object(env, &obj);
  napi_set_named_property(env, obj, "proof", make_buffer(env, r.proof.data(), r.proof.size()));
  napi_set_named_property(env, obj, "publicWitness", make_buffer(env, r.pw.data(), r.pw.size()));
  return obj;
}

bool circuit_and_buffer(napi_env env, napi_callback_info info, Circuit** c, std::vector<uint8_t>* buf, double* extra) {
  size_t argc = 3;
  napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* p;
  size_t n;
  void* ext = nullptr;
  if (argc < 2 || napi_get_value_external(env, argv[0], &ext) != napi_ok || !ext || !buffer_arg(env, argv[1], &p, &n)) return false;
  *c = static_cast<Circuit*>(ext);
  buf->assign(p, p + n);
  if (extra && argc > 2) napi_get_value_double(env, argv[2], extra);
  return true;
}

napi_value ProveSync(napi_env env, napi_callback_info info) {
  Circuit* c;
  std::vector<uint8_t> gz;
  if (!circuit_and_buffer(env, info, &c, &gz, nullptr)) return throw_error(env, "proveSync(circuit, witnessGz: Buffer)");
  ProveResult r;
  run_prove(c, gz, &r);
  if (r.rc != G16_OK) return throw_error(env, r.err);
  return result_object(env, r);
}

// ---- async: napi_async_work keeps the event loop free (execSync blocks it today, payroll-demo.ts:326-350)
struct Job {
  napi_async_work work = nullptr;
  napi_deferred deferred = nullptr;
  Circuit* circuit = nullptr;
  std::vector<uint8_t> input;
  size_t n = 0;          // 0: single proof from a witness; > 0: batch of n assignments
  ProveResult single;
  std::vector<uint8_t> proofs, pws;
  int rc = G16_OK;
  std::string err;
};

void job_execute(napi_env, void* data) {
  Job* j = static_cast<Job*>(data);
  if (j->n == 0) {
    run_prove(j->circuit, j->input, &j->single);
    j->rc = j->single.rc;
    j->err = j->single.err;
    return;
  }
  Circuit* c = j->circuit;
  j->proofs.resize(j->n * c->proof_len);
  j->pws.resize(j->n * c->pw_len);
  std::lock_guard<std::mutex> lock(c->busy);
  j->rc = g16_prove_batch(c->handle, j->n, j->input.data(), c->n_values, nullptr, j->proofs.data(), j->pws.data(), c->pw_len);
  if (j->rc != G16_OK) j->err = g16_last_error();
}

void job_complete(napi_env env, napi_status, void* data) {
  Job* j = static_cast<Job*>(data);
  if (j->rc != G16_OK) {
    napi_value msg, error;
    napi_create_string_utf8(env, j->err.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &error);
    napi_reject_deferred(env, j->deferred, error);
  } else if (j->n == 0) {
    napi_resolve_deferred(env, j->deferred, result_object(env, j->single));
  } else {
    napi_value obj, stride;
    napi_create_object(env, &obj);
    napi_set_named_property(env, obj, "proofs", make_buffer(env, j->proofs.data(), j->proofs.size()));
    napi_set_named_property(env, obj, "publicWitnesses", make_buffer(env, j->pws.data(), j->pws.size()));
    napi_create_uint32(env, (uint32_t)j->circuit->pw_len, &stride);
    napi_set_named_property(env, obj, "pwStride", stride);
    napi_resolve_deferred(env, j->deferred, obj);
  }
  napi_delete_async_work(env, j->work);
  delete j;
}

napi_value queue_job(napi_env env, Job* j) {
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  napi_create_string_utf8(env, "g16b200.prove", NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, job_execute, job_complete, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}

napi_value Prove(napi_env env, napi_callback_info info) {
  Job* j = new Job();
  if (!circuit_and_buffer(env, info, &j->circuit, &j->input, nullptr)) {
    delete j;
    return throw_error(env, "prove(circuit, witnessGz: Buffer)");
  }
  return queue_job(env, j);
}

napi_value ProveBatch(napi_env env, napi_callback_info info) {
  Job* j = new Job();
  double n = 0;
  if (!circuit_and_buffer(env, info, &j->circuit, &j->input, &n) || n < 1 ||
      j->input.size() != (size_t)n * j->circuit->n_values * 32) {
    delete j;
    return throw_error(env, "proveBatch(circuit, assignments: Buffer /* n * nValues * 32 B big-endian */, n: number)");
  }
  j->n = (size_t)n;
  return queue_job(env, j);
}

napi_value Verify(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *vk, *proof, *pw;
  size_t nvk, nproof, npw;
  if (argc < 3 || !buffer_arg(env, argv[0], &vk, &nvk) || !buffer_arg(env, argv[1], &proof, &nproof) ||
      !buffer_arg(env, argv[2], &pw, &npw))
    return throw_error(env, "verify(vk: Buffer, proof: Buffer, publicWitness: Buffer)");
  int ok = 0;
  if (g16_verify(vk, nvk, proof, nproof, pw, npw, &ok) != G16_OK) return throw_error(env, g16_last_error());
  napi_value out;
  napi_get_boolean(env, ok != 0, &out);
  return out;
}

napi_value Setup(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *ccs, *seed;
  size_t nccs, nseed;
  if (argc < 2 || !buffer_arg(env, argv[0], &ccs, &nccs) || !buffer_arg(env, argv[1], &seed, &nseed) || nseed < 16)
    return throw_error(env, "setup(ccs: Buffer, seed: Buffer /* >= 16 bytes of fresh entropy */, device?: number)");
  int32_t device = 0;
  if (argc > 2) napi_get_value_int32(env, argv[2], &device);
  std::string err;
  g16_ctx* ctx = context_for(device, &err);
  if (!ctx) return throw_error(env, err);
  size_t pl = 0, vl = 0;
  if (g16_setup(ctx, ccs, nccs, seed, nseed, nullptr, &pl, nullptr, &vl) != G16_OK) return throw_error(env, g16_last_error());
  std::vector<uint8_t> pk(pl), vk(vl);
  if (g16_setup(ctx, ccs, nccs, seed, nseed, pk.data(), &pl, vk.data(), &vl) != G16_OK) return throw_error(env, g16_last_error());
  napi_value obj;
  napi_create_object(env, &obj);
  napi_set_named_property(env, obj, "pk", make_buffer(env, pk.data(), pl));
  napi_set_named_property(env, obj, "vk", make_buffer(env, vk.data(), vl));
  return obj;
}

// execute(ccs: Buffer, programJson: Buffer, proverToml: Buffer) -> Buffer (the witness file `nargo execute` writes);
// host only.  For circuits whose constraints determine their witnesses (the withdraw circuit): include/g16b200.h.
napi_value Execute(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *ccs, *acir, *toml;
  size_t nccs, nacir, ntoml;
  if (argc < 3 || !buffer_arg(env, argv[0], &ccs, &nccs) || !buffer_arg(env, argv[1], &acir, &nacir) ||
      !buffer_arg(env, argv[2], &toml, &ntoml))
    return throw_error(env, "execute(ccs: Buffer, programJson: Buffer, proverToml: Buffer)");
  size_t n = 0;
  if (g16_execute(ccs, nccs, (const char*)acir, nacir, (const char*)toml, ntoml, nullptr, &n) != G16_OK)
    return throw_error(env, g16_last_error());
  std::vector<uint8_t> gz(n);
  if (g16_execute(ccs, nccs, (const char*)acir, nacir, (const char*)toml, ntoml, gz.data(), &n) != G16_OK)
    return throw_error(env, g16_last_error());
  return make_buffer(env, gz.data(), n);
}

}  // namespace

NAPI_MODULE_INIT() {
  struct { const char* name; napi_callback fn; } fns[] = {
      {"loadCircuit", LoadCircuit}, {"proveSync", ProveSync}, {"prove", Prove},
      {"proveBatch", ProveBatch},   {"verify", Verify},       {"setup", Setup},
      {"execute", Execute},
  };
  for (auto& f : fns) {
    napi_value v;
    napi_create_function(env, f.name, NAPI_AUTO_LENGTH, f.fn, nullptr, &v);
    napi_set_named_property(env, exports, f.name, v);
  }
  return exports;
}
