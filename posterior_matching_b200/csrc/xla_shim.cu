// XLA custom-call targets over the C ABI (include/pmvae.h): what jax.ffi / xla_client registers so that the
// reference's jitted loss_fn / eval_fn (train_pm_vae.py:58-72, eval_pm_vae_uci.py:82-96) reach the CUDA path.
// Status-returning legacy signature (XLA CustomCallApiVersion API_VERSION_STATUS_RETURNING = 2: the trailing
// XlaCustomCallStatus* argument; 1 = API_VERSION_ORIGINAL has none); failures go through XlaCustomCallStatusSetFailure,
// which is resolved from the hosting process (jaxlib) at run time so that this library links without XLA.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef void (*SetFailureFn)(void* status, const char* msg, size_t len);

void report(void* status, const char* what) {
  static SetFailureFn fn = reinterpret_cast<SetFailureFn>(dlsym(RTLD_DEFAULT, "XlaCustomCallStatusSetFailure"));
  const char* detail = pmvae_last_error();
  std::string msg = std::string(what) + ": " + (detail ? detail : "?");
  if (fn && status) fn(status, msg.c_str(), msg.size());
  else fprintf(stderr, "pmvae xla target failed: %s\n", msg.c_str());
}

const pmvae_xla_opaque* decode(const char* opaque, size_t len, void* status) {
  if (opaque == nullptr || len != sizeof(pmvae_xla_opaque)) {
    pmvae::set_error("opaque descriptor has the wrong size (expected one pmvae_xla_opaque)");
    report(status, "pmvae_xla");
    return nullptr;
  }
  return reinterpret_cast<const pmvae_xla_opaque*>(opaque);
}

// XLA aligns its buffers to 256 bytes at most; the tensor path wants 1024 (TMA / swizzle atoms).  The workspace
// buffer XLA allocates is therefore ws_bytes + PMVAE_XLA_WS_SLACK bytes and every target rounds the pointer up
// (the same offset in the forward and the backward, which receive the same buffer).
void* align_ws(void* p) { return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(p) + 1023u) & ~uintptr_t(1023)); }

int maybe_prepare(const pmvae_xla_opaque* o, const float* params, void* ws, pmvae_stream_t s) {
  return o->prepare ? pmvae_prepare_params(&o->cfg, params, ws, o->ws_bytes, s) : 0;
}

}  // namespace

extern "C" {

uint64_t pmvae_xla_opaque_size(void) { return sizeof(pmvae_xla_opaque); }

void pmvae_xla_forward(pmvae_stream_t s, void** buf, const char* opaque, size_t len, void* status) {
  const pmvae_xla_opaque* o = decode(opaque, len, status);
  if (!o) return;
  const float* params = static_cast<const float*>(buf[0]);
  void* ws = align_ws(buf[7]);
  if (maybe_prepare(o, params, ws, s) != 0 ||
      pmvae_forward(&o->cfg, params, static_cast<const float*>(buf[1]), static_cast<const float*>(buf[2]),
                    static_cast<const float*>(buf[3]), o->B, static_cast<float*>(buf[4]), static_cast<float*>(buf[5]),
                    static_cast<float*>(buf[6]), ws, o->ws_bytes, s) != 0)
    report(status, "pmvae_xla_forward");
}

void pmvae_xla_backward(pmvae_stream_t s, void** buf, const char* opaque, size_t len, void* status) {
  const pmvae_xla_opaque* o = decode(opaque, len, status);
  if (!o) return;
  // operands 0..7 = params, x, b, eps, g_rec, g_kl, g_match, ws; results 8..9 = grads, ws (same buffer as 7)
  if (buf[9] != buf[7]) {
    pmvae::set_error("the workspace result must alias the workspace operand (operand_output_aliases={7: 1})");
    report(status, "pmvae_xla_backward");
    return;
  }
  if (pmvae_backward(&o->cfg, static_cast<const float*>(buf[0]), static_cast<const float*>(buf[1]),
                     static_cast<const float*>(buf[2]), static_cast<const float*>(buf[3]), o->B,
                     static_cast<const float*>(buf[4]), static_cast<const float*>(buf[5]),
                     static_cast<const float*>(buf[6]), static_cast<float*>(buf[8]), align_ws(buf[7]), o->ws_bytes, s) != 0)
    report(status, "pmvae_xla_backward");
}

void pmvae_xla_is_log_prob(pmvae_stream_t s, void** buf, const char* opaque, size_t len, void* status) {
  const pmvae_xla_opaque* o = decode(opaque, len, status);
  if (!o) return;
  const float* params = static_cast<const float*>(buf[0]);
  void* ws = align_ws(buf[5]);
  if (maybe_prepare(o, params, ws, s) != 0 ||
      pmvae_is_log_prob(&o->cfg, params, static_cast<const float*>(buf[1]), static_cast<const float*>(buf[2]), o->B,
                        o->K, o->key0, o->key1, o->B_total, o->row_start, static_cast<float*>(buf[3]),
                        static_cast<float*>(buf[4]), ws, o->ws_bytes, s) != 0)
    report(status, "pmvae_xla_is_log_prob");
}

void pmvae_xla_impute_mean(pmvae_stream_t s, void** buf, const char* opaque, size_t len, void* status) {
  const pmvae_xla_opaque* o = decode(opaque, len, status);
  if (!o) return;
  const float* params = static_cast<const float*>(buf[0]);
  void* ws = align_ws(buf[4]);
  if (maybe_prepare(o, params, ws, s) != 0 ||
      pmvae_impute_mean(&o->cfg, params, static_cast<const float*>(buf[1]), static_cast<const float*>(buf[2]), o->B,
                        o->K, o->key0, o->B_total, o->row_start, static_cast<float*>(buf[3]), ws, o->ws_bytes, s) != 0)
    report(status, "pmvae_xla_impute_mean");
}

void pmvae_xla_mask_bernoulli(pmvae_stream_t s, void** buf, const char* opaque, size_t len, void* status) {
  const pmvae_xla_opaque* o = decode(opaque, len, status);
  if (!o) return;
  if (pmvae_mask_bernoulli(o->key0, o->p, (uint64_t)o->B_total, (uint64_t)o->row_start, (uint64_t)o->B, o->D,
                           static_cast<float*>(buf[0]), s) != 0)
    report(status, "pmvae_xla_mask_bernoulli");
}

}  // extern "C"
