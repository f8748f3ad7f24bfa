// Host-side orchestration of the PM-VAE hot path: parameter arena layout, workspace
// plan, and the launch sequences behind the C ABI (include/pmvae.h).
//
// Reference: posterior_matching/models/vae.py:120-144 (__call__), :146-169 (impute),
// :171-226 (is_log_prob); networks.py:111-135 (ResidualMLP); train_pm_vae.py:58-83.
#include <math.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "kernels.h"
#include "model.h"
#include "tc_gemm.h"

namespace pmvae {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error() { return g_last_error.c_str(); }
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---------------------------------------------------------------- layout
static int check_cfg(const pmvae_config* c) {
  PMVAE_CHECK(c != nullptr, "null config");
  PMVAE_CHECK(c->D >= 1 && c->D <= 4096, "D out of range");
  PMVAE_CHECK(c->d >= 1 && c->d <= 64, "latent_dim must be in [1, 64]");
  PMVAE_CHECK(c->H >= 4 && c->H % 4 == 0 && c->H <= 1024, "hidden_units must be a multiple of 4 in [4, 1024]");
  PMVAE_CHECK(c->R_enc >= 0 && c->R_enc <= kMaxBlocks && c->R_dec >= 0 && c->R_dec <= kMaxBlocks &&
                  c->R_part >= 0 && c->R_part <= kMaxBlocks,
              "residual_blocks out of range");
  PMVAE_CHECK(c->precision == PMVAE_PREC_F32 || c->precision == PMVAE_PREC_BF16, "unknown precision");
  return 0;
}

static uint64_t pad64(uint64_t n) { return align_up(n, 64); }  // 256-byte leaf alignment

int build_layout(const pmvae_config* c, Layout* L) {
  PMVAE_TRY(check_cfg(c));
  uint64_t off = 0;
  auto leaf = [&](Leaf& lf, int rows, int cols) {
    lf.rows = rows; lf.cols = cols;
    lf.w = off; off += pad64((uint64_t)rows * cols);
    lf.b = off; off += pad64((uint64_t)cols);
  };
  auto net = [&](Net& n, int in_dim, int R, int ln) {
    n.in_dim = in_dim; n.R = R; n.ln = ln;
    leaf(n.lin[0], in_dim, c->H);
    for (int i = 1; i <= 2 * R; ++i) leaf(n.lin[i], c->H, c->H);
  };
  const int P = c->d + c->d * (c->d + 1) / 2;
  L->P = P;
  net(L->enc, c->D, c->R_enc, c->ln_enc);
  leaf(L->post, c->H, P);
  net(L->dec, c->d, c->R_dec, c->ln_dec);
  leaf(L->ddist, c->H, c->D);
  L->log_scale = off; off += 64;
  net(L->part, 2 * c->D, c->R_part, c->ln_part);
  leaf(L->ppost, c->H, P);
  L->total = off;
  return 0;
}

static void name_leaf(pmvae_leaf* o, const char* prefix, int i, const Leaf& lf) {
  if (i == 0) snprintf(o->name, sizeof(o->name), "%s/linear", prefix);
  else snprintf(o->name, sizeof(o->name), "%s/linear_%d", prefix, i);
  o->rows = lf.rows; o->cols = lf.cols; o->w_off = lf.w; o->b_off = lf.b;
}

int export_layout(const pmvae_config* c, pmvae_leaf* out, int cap) {
  Layout L;
  if (build_layout(c, &L) != 0) return -1;
  std::vector<pmvae_leaf> v;
  auto push_net = [&](const Net& n, const char* prefix) {
    for (int i = 0; i <= 2 * n.R; ++i) { pmvae_leaf l{}; name_leaf(&l, prefix, i, n.lin[i]); v.push_back(l); }
  };
  auto push_leaf = [&](const Leaf& lf, const char* prefix) { pmvae_leaf l{}; name_leaf(&l, prefix, 0, lf); v.push_back(l); };
  push_net(L.enc, "encoder_net");
  push_leaf(L.post, "posterior_dist");
  push_net(L.dec, "decoder_net");
  push_leaf(L.ddist, "decoder_dist");
  { pmvae_leaf l{}; snprintf(l.name, sizeof(l.name), "decoder_dist"); l.rows = 0; l.cols = 0; l.w_off = L.log_scale; l.b_off = ~0ull; v.push_back(l); }
  push_net(L.part, "partial_encoder_net");
  push_leaf(L.ppost, "partial_posterior_dist");
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

// ---------------------------------------------------------------- workspace plan
struct Bump {
  char* base; uint64_t off = 0;
  explicit Bump(void* b) : base(reinterpret_cast<char*>(b)) {}
  template <typename T> T* take(uint64_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(count * sizeof(T), 256);
    return p;
  }
};

struct NetSaved {            // fp32 path: pre-activation (post-LN) tensors, relu applied by the consumer
  float* H[kMaxBlocks + 1];
  float* U[kMaxBlocks];
  float* V[kMaxBlocks];      // LN nets only (xhat of the second linear)
  float* rstd0; float* rstdU[kMaxBlocks]; float* rstdV[kMaxBlocks];
};

struct TrainPlan {
  NetSaved enc, dec, part;
  float *par_e, *par_p, *z, *loc, *xob;
  float *dH, *tmp1, *tmp2, *dpar_e, *dpar_p, *dloc, *dz;
  uint64_t bytes;
};

static void plan_net(Bump& bp, const Net& n, int64_t B, int H, NetSaved& s, bool save_all) {
  const uint64_t e = (uint64_t)B * H;
  for (int r = 0; r <= n.R; ++r) s.H[r] = bp.take<float>(e);
  for (int r = 0; r < n.R; ++r) s.U[r] = bp.take<float>(e);
  (void)save_all;
  if (n.ln) {
    for (int r = 0; r < n.R; ++r) s.V[r] = bp.take<float>(e);
    s.rstd0 = bp.take<float>(B);
    for (int r = 0; r < n.R; ++r) { s.rstdU[r] = bp.take<float>(B); s.rstdV[r] = bp.take<float>(B); }
  }
}

static TrainPlan plan_train(const pmvae_config* c, const Layout& L, int64_t B, void* ws) {
  TrainPlan p{};
  Bump bp(ws);
  plan_net(bp, L.enc, B, c->H, p.enc, true);
  plan_net(bp, L.dec, B, c->H, p.dec, true);
  plan_net(bp, L.part, B, c->H, p.part, true);
  p.par_e = bp.take<float>((uint64_t)B * L.P);
  p.par_p = bp.take<float>((uint64_t)B * L.P);
  p.z = bp.take<float>((uint64_t)B * c->d);
  p.loc = bp.take<float>((uint64_t)B * c->D);
  p.xob = bp.take<float>((uint64_t)B * 2 * c->D);
  p.dH = bp.take<float>((uint64_t)B * c->H);
  p.tmp1 = bp.take<float>((uint64_t)B * c->H);
  p.tmp2 = bp.take<float>((uint64_t)B * c->H);
  p.dpar_e = bp.take<float>((uint64_t)B * L.P);
  p.dpar_p = bp.take<float>((uint64_t)B * L.P);
  p.dloc = bp.take<float>((uint64_t)B * c->D);
  p.dz = bp.take<float>((uint64_t)B * c->d);
  p.bytes = bp.off;
  return p;
}

constexpr int64_t kEvalChunkRows = 1 << 16;  // decoder rows (K * rows) per evaluator chunk

struct EvalPlan {
  NetSaved enc, part;           // B rows
  float *par_e, *par_p, *xob;
  NetSaved dec;                 // chunk rows
  float *z, *base, *loc, *llA, *llC;
  int64_t rows_per_chunk;       // data rows per chunk
  uint64_t bytes;
};

static EvalPlan plan_eval(const pmvae_config* c, const Layout& L, int64_t B, int64_t K, void* ws) {
  EvalPlan p{};
  Bump bp(ws);
  plan_net(bp, L.enc, B, c->H, p.enc, false);
  plan_net(bp, L.part, B, c->H, p.part, false);
  p.par_e = bp.take<float>((uint64_t)B * L.P);
  p.par_p = bp.take<float>((uint64_t)B * L.P);
  p.xob = bp.take<float>((uint64_t)B * 2 * c->D);
  int64_t rpc = kEvalChunkRows / (K > 0 ? K : 1);
  if (rpc < 1) rpc = 1;
  if (rpc > B) rpc = B > 0 ? B : 1;
  p.rows_per_chunk = rpc;
  const int64_t M = rpc * K;
  plan_net(bp, L.dec, M, c->H, p.dec, false);
  p.z = bp.take<float>((uint64_t)M * c->d);
  p.base = bp.take<float>((uint64_t)M);
  p.loc = bp.take<float>((uint64_t)M * c->D);
  p.llA = bp.take<float>((uint64_t)M);
  p.llC = bp.take<float>((uint64_t)M);
  p.bytes = bp.off;
  return p;
}

uint64_t workspace_bytes(const pmvae_config* c, int64_t B, int64_t K) {
  Layout L;
  if (build_layout(c, &L) != 0) return 0;
  if (B < 1) B = 1;
  uint64_t t = plan_train(c, L, B, nullptr).bytes;
  uint64_t e = K > 0 ? plan_eval(c, L, B, K, nullptr).bytes : 0;
  return (t > e ? t : e) + 256;
}

// ---------------------------------------------------------------- fp32 network passes
static int lin_fwd(const float* params, const Leaf& lf, const float* in, int64_t ld_in, int relu_in, int64_t B,
                   float* out, const float* resid, cudaStream_t s) {
  GemmF32Args a{};
  a.M = B; a.N = lf.cols; a.K = lf.rows;
  a.A = in; a.lda = ld_in;
  a.B = params + lf.w; a.ldb = lf.cols;
  a.C = out; a.ldc = lf.cols;
  a.bias = params + lf.b;
  a.relu_a = relu_in;
  a.resid = resid; a.ldresid = lf.cols;
  return gemm_f32(a, false, false, s);
}

// Head Linear restricted to the diagonal blocks of a (row block, column block) partition: rows [i rb, (i+1) rb) only
// need columns [i cb, (i+1) cb).  The AutoregressiveGMM batches its d steps as d row blocks and reads, for step i, only
// the 3 K parameters of dimension i out of the 3 K d the Linear produces (distributions.py:153-166), so 1/d of the head
// GEMM is ever used; `only` >= 0 computes that one block (the sampler, one step at a time).  Entries outside the
// blocks are not written.
static int head_fwd_blocks(const float* params, const Leaf& head, const float* act, int H, int64_t B, int blocks, int only,
                           float* head_out, cudaStream_t s) {
  PMVAE_CHECK(blocks >= 1 && head.cols % blocks == 0 && only < blocks && (only >= 0 || B % blocks == 0),
              "block-diagonal head: sizes do not divide");
  const int64_t rb = only >= 0 ? B : B / blocks;
  const int cb = head.cols / blocks;
  const int i0 = only >= 0 ? only : 0;
  GemmF32Args a{};                         // one launch: batch member i = block i (or the single block `only`)
  a.M = rb; a.N = cb; a.K = head.rows;
  a.A = act; a.lda = H;
  a.B = params + head.w + (int64_t)i0 * cb; a.ldb = head.cols;
  a.C = head_out + (int64_t)i0 * cb; a.ldc = head.cols;
  a.bias = params + head.b + (int64_t)i0 * cb;
  a.relu_a = 1;
  if (only < 0) {
    a.batch = blocks;
    a.sA = rb * H; a.sB = cb; a.sC = rb * head.cols + cb; a.sBias = cb;
  }
  return gemm_f32(a, false, false, s);
}

// Optional tensor-core route for the hidden H x H Linears (H = 256, no LayerNorm) of a float32-composed net: bf16 images
// of lin[1 .. 2R] (wn[l - 1], wt[l - 1]) and two bf16 scratch tensors of [rows, 256] (tensor.cu::hidden_*_tc).
struct HiddenTc {
  const __nv_bfloat16* wn[2 * kMaxBlocks];
  const __nv_bfloat16* wt[2 * kMaxBlocks];
  __nv_bfloat16 *xb, *dyb;
};

static int net_fwd_f32(const float* params, const Net& n, const Leaf& head, int H, const float* in, int64_t B,
                       const NetSaved& sv, float* head_out, cudaStream_t s, int head_blocks = 1, int head_only = -1,
                       const HiddenTc* tcp = nullptr) {
  PMVAE_TRY(lin_fwd(params, n.lin[0], in, n.in_dim, 0, B, sv.H[0], nullptr, s));
  if (n.ln) PMVAE_TRY(ln_fwd(sv.H[0], sv.rstd0, nullptr, nullptr, B, H, s));
  for (int r = 0; r < n.R; ++r) {
    if (tcp) {
      PMVAE_TRY(hidden_fwd_tc(sv.H[r], tcp->wt[2 * r], params + n.lin[2 * r + 1].b, nullptr, B, sv.U[r], tcp->xb, s));
      PMVAE_TRY(hidden_fwd_tc(sv.U[r], tcp->wt[2 * r + 1], params + n.lin[2 * r + 2].b, sv.H[r], B, sv.H[r + 1], tcp->xb, s));
      continue;
    }
    PMVAE_TRY(lin_fwd(params, n.lin[2 * r + 1], sv.H[r], H, 1, B, sv.U[r], nullptr, s));
    if (n.ln) {
      PMVAE_TRY(ln_fwd(sv.U[r], sv.rstdU[r], nullptr, nullptr, B, H, s));
      PMVAE_TRY(lin_fwd(params, n.lin[2 * r + 2], sv.U[r], H, 1, B, sv.V[r], nullptr, s));
      PMVAE_TRY(ln_fwd(sv.V[r], sv.rstdV[r], sv.H[r], sv.H[r + 1], B, H, s));
    } else {
      PMVAE_TRY(lin_fwd(params, n.lin[2 * r + 2], sv.U[r], H, 1, B, sv.H[r + 1], sv.H[r], s));
    }
  }
  if (head_blocks > 1) return head_fwd_blocks(params, head, sv.H[n.R], H, B, head_blocks, head_only, head_out, s);
  return lin_fwd(params, head, sv.H[n.R], H, 1, B, head_out, nullptr, s);
}

static int split_for(int64_t out_rows, int64_t out_cols, int64_t B) {
  const int64_t tiles = ceil_div(out_rows, 64) * ceil_div(out_cols, 64);
  int64_t sp = ceil_div(148 * 4, tiles);
  const int64_t maxsp = ceil_div(B, 256);
  if (sp > maxsp) sp = maxsp;
  if (sp < 1) sp = 1;
  if (sp > 65535) sp = 65535;
  return (int)sp;
}

// grads for one Linear: gW += relu?(in)^T dY ; gb += colsum(dY)
static int lin_bwd_params(float* grads, const Leaf& lf, const float* in, int64_t ld_in, int relu_in, const float* dY,
                          int64_t B, cudaStream_t s) {
  GemmF32Args a{};
  a.M = lf.rows; a.N = lf.cols; a.K = B;
  a.A = in; a.lda = ld_in;     // A'(m,k) = in[k*ld + m]  (TA)
  a.B = dY; a.ldb = lf.cols;   // B'(k,n) = dY[k*ld + n]
  a.C = grads + lf.w; a.ldc = lf.cols;
  a.relu_a = relu_in; a.atomic = 1; a.split_k = split_for(lf.rows, lf.cols, B);
  PMVAE_TRY(gemm_f32(a, true, false, s));
  return colsum_add(dY, lf.cols, grads + lf.b, B, lf.cols, s);
}

// dIn = (dY @ W^T) [* (mask > 0)] [+ resid]
static int lin_bwd_input(const float* params, const Leaf& lf, const float* dY, int64_t B, float* dIn,
                         const float* mask, const float* resid, cudaStream_t s) {
  GemmF32Args a{};
  a.M = B; a.N = lf.rows; a.K = lf.cols;
  a.A = dY; a.lda = lf.cols;
  a.B = params + lf.w; a.ldb = lf.cols;  // B'(k,n) = W[n*cols + k]  (TB)
  a.C = dIn; a.ldc = lf.rows;
  a.mask = mask; a.ldmask = lf.rows;
  a.resid = resid; a.ldresid = lf.rows;
  return gemm_f32(a, false, true, s);
}

// VJP of head_fwd_blocks (all blocks): dHead is read on the diagonal blocks only.
static int head_bwd_blocks(const float* params, float* grads, const Leaf& head, const float* act, int H, int64_t B,
                           int blocks, const float* dHead, float* dH, cudaStream_t s) {
  PMVAE_CHECK(blocks >= 1 && B % blocks == 0 && head.cols % blocks == 0, "block-diagonal head: sizes do not divide");
  const int64_t rb = B / blocks;
  const int cb = head.cols / blocks;
  GemmF32Args w{};                         // gW[:, block i] += relu(act_i)^T dy_i, all blocks in one launch
  w.M = head.rows; w.N = cb; w.K = rb;
  w.A = act; w.lda = H;
  w.B = dHead; w.ldb = head.cols;
  w.C = grads + head.w; w.ldc = head.cols;
  w.relu_a = 1; w.atomic = 1; w.split_k = split_for(head.rows, (int64_t)cb * blocks, rb);
  w.batch = blocks; w.sA = rb * H; w.sB = rb * head.cols + cb; w.sC = cb;
  PMVAE_TRY(gemm_f32(w, true, false, s));
  for (int i = 0; i < blocks; ++i)
    PMVAE_TRY(colsum_add(dHead + i * rb * head.cols + (int64_t)i * cb, head.cols, grads + head.b + (int64_t)i * cb, rb, cb, s));
  GemmF32Args x{};                         // dH_i = (dy_i W[:, block i]^T) * (act_i > 0)
  x.M = rb; x.N = head.rows; x.K = cb;
  x.A = dHead; x.lda = head.cols;
  x.B = params + head.w; x.ldb = head.cols;
  x.C = dH; x.ldc = head.rows;
  x.mask = act; x.ldmask = H;
  x.batch = blocks; x.sA = rb * head.cols + cb; x.sB = cb; x.sC = rb * head.rows; x.sMask = rb * H;
  return gemm_f32(x, false, true, s);
}

static int net_bwd_f32(const float* params, float* grads, const Net& n, const Leaf& head, int H, const float* in,
                       int64_t B, const NetSaved& sv, const float* dHead, float* dH, float* tmp1, float* tmp2,
                       float* dIn, cudaStream_t s, int head_blocks = 1, const HiddenTc* tcp = nullptr) {
  if (head_blocks > 1) {
    PMVAE_TRY(head_bwd_blocks(params, grads, head, sv.H[n.R], H, B, head_blocks, dHead, dH, s));
  } else {
    PMVAE_TRY(lin_bwd_params(grads, head, sv.H[n.R], H, 1, dHead, B, s));
    PMVAE_TRY(lin_bwd_input(params, head, dHead, B, dH, sv.H[n.R], nullptr, s));
  }
  for (int r = n.R - 1; r >= 0; --r) {
    if (tcp) {
      const Leaf& l2 = n.lin[2 * r + 2];
      const Leaf& l1 = n.lin[2 * r + 1];
      PMVAE_TRY(colsum_add(dH, l2.cols, grads + l2.b, B, l2.cols, s));
      PMVAE_TRY(hidden_bwd_tc(sv.U[r], dH, tcp->wn[2 * r + 1], grads + l2.w, tmp2, nullptr, B, tcp->xb, tcp->dyb, s));
      PMVAE_TRY(colsum_add(tmp2, l1.cols, grads + l1.b, B, l1.cols, s));
      PMVAE_TRY(hidden_bwd_tc(sv.H[r], tmp2, tcp->wn[2 * r], grads + l1.w, dH, dH, B, tcp->xb, tcp->dyb, s));
      continue;
    }
    const float* dV = dH;
    if (n.ln) { PMVAE_TRY(ln_bwd(dH, sv.V[r], sv.rstdV[r], tmp1, B, H, s)); dV = tmp1; }
    PMVAE_TRY(lin_bwd_params(grads, n.lin[2 * r + 2], sv.U[r], H, 1, dV, B, s));
    PMVAE_TRY(lin_bwd_input(params, n.lin[2 * r + 2], dV, B, tmp2, sv.U[r], nullptr, s));
    if (n.ln) PMVAE_TRY(ln_bwd(tmp2, sv.U[r], sv.rstdU[r], tmp2, B, H, s));
    PMVAE_TRY(lin_bwd_params(grads, n.lin[2 * r + 1], sv.H[r], H, 1, tmp2, B, s));
    PMVAE_TRY(lin_bwd_input(params, n.lin[2 * r + 1], tmp2, B, dH, sv.H[r], dH, s));
  }
  if (n.ln) PMVAE_TRY(ln_bwd(dH, sv.H[0], sv.rstd0, dH, B, H, s));
  PMVAE_TRY(lin_bwd_params(grads, n.lin[0], in, n.in_dim, 0, dH, B, s));
  if (dIn) PMVAE_TRY(lin_bwd_input(params, n.lin[0], dH, B, dIn, nullptr, nullptr, s));
  return 0;
}

// ---------------------------------------------------------------- public sequences
int forward(const pmvae_config* c, const float* params, const float* x, const float* b, const float* eps, int64_t B,
            float* out_rec, float* out_kl, float* out_match, void* ws, uint64_t ws_bytes, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  PMVAE_CHECK(B >= 0, "negative batch");
  if (B == 0) return 0;            // an empty batch has no rows to point at
  PMVAE_CHECK(params && x && b && eps && out_rec && out_kl && out_match && ws, "null pointer");
  if (c->precision == PMVAE_PREC_BF16) return forward_bf16(c, L, params, x, b, eps, B, out_rec, out_kl, out_match, ws, ws_bytes, s);
  TrainPlan p = plan_train(c, L, B, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_workspace_bytes)");
  PMVAE_TRY(net_fwd_f32(params, L.enc, L.post, c->H, x, B, p.enc, p.par_e, s));
  PMVAE_TRY(latent_fwd(p.par_e, eps, p.z, out_kl, B, c->d, s));
  PMVAE_TRY(net_fwd_f32(params, L.dec, L.ddist, c->H, p.z, B, p.dec, p.loc, s));
  PMVAE_TRY(rec_ll(x, p.loc, c->D, params + L.log_scale, nullptr, out_rec, B, c->D, s));
  PMVAE_TRY(concat_masked(x, b, p.xob, B, c->D, s));
  PMVAE_TRY(net_fwd_f32(params, L.part, L.ppost, c->H, p.xob, B, p.part, p.par_p, s));
  PMVAE_TRY(match_fwd(p.par_p, p.z, out_match, B, c->d, s));
  return 0;
}

int backward_staged(const pmvae_config* c, const float* params, const float* x, const float* b, const float* eps, int64_t B,
                    const float* g_rec, const float* g_kl, const float* g_match, float* grads, int stages, void* ws,
                    uint64_t ws_bytes, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  PMVAE_CHECK(grads != nullptr && B >= 0, "null gradient arena / negative batch");
  if (stages & 1) PMVAE_CUDA(cudaMemsetAsync(grads, 0, L.total * sizeof(float), s));
  if (B == 0) return 0;            // an empty batch contributes zero gradients
  PMVAE_CHECK(params && x && b && eps && g_rec && g_kl && g_match && ws, "null pointer");
  if (c->precision == PMVAE_PREC_BF16)
    return backward_bf16(c, L, params, x, b, eps, B, g_rec, g_kl, g_match, grads, stages, ws, ws_bytes, s);
  TrainPlan p = plan_train(c, L, B, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_workspace_bytes)");
  if (stages & 1) {
    PMVAE_TRY(rec_ll_bwd(x, p.loc, c->D, params + L.log_scale, g_rec, p.dloc, nullptr, c->D, grads + L.log_scale, B, c->D, s));
    PMVAE_TRY(net_bwd_f32(params, grads, L.dec, L.ddist, c->H, p.z, B, p.dec, p.dloc, p.dH, p.tmp1, p.tmp2, p.dz, s));
    PMVAE_TRY(latent_bwd(p.par_e, p.par_p, eps, p.z, p.dz, g_kl, g_match, c->stop_grad, p.dpar_e, p.dpar_p, nullptr, nullptr, B, c->d, s));
  }
  if (stages & 2) PMVAE_TRY(net_bwd_f32(params, grads, L.enc, L.post, c->H, x, B, p.enc, p.dpar_e, p.dH, p.tmp1, p.tmp2, nullptr, s));
  if (stages & 4) PMVAE_TRY(net_bwd_f32(params, grads, L.part, L.ppost, c->H, p.xob, B, p.part, p.dpar_p, p.dH, p.tmp1, p.tmp2, nullptr, s));
  return 0;
}

int backward(const pmvae_config* c, const float* params, const float* x, const float* b, const float* eps, int64_t B,
             const float* g_rec, const float* g_kl, const float* g_match, float* grads, void* ws, uint64_t ws_bytes,
             cudaStream_t s) {
  return backward_staged(c, params, x, b, eps, B, g_rec, g_kl, g_match, grads, 7, ws, ws_bytes, s);
}

// Bias leaves of the arena (ndim == 1: no weight decay, train_pm_vae.py:77-79), bound-checked against AdamSegs.
static_assert(kAdamSegCap >= 3 * (2 * kMaxBlocks + 1) + 3, "AdamSegs must hold every bias leaf kMaxBlocks allows");
static int bias_ranges(const Layout& L, AdamSegs* seg) {
  seg->n = 0;
  bool ok = true;
  auto add = [&](const Leaf& lf) {
    if (seg->n >= kAdamSegCap) { ok = false; return; }
    seg->beg[seg->n] = (uint32_t)lf.b; seg->end[seg->n] = (uint32_t)(lf.b + pad64(lf.cols)); ++seg->n;
  };
  auto addnet = [&](const Net& n) { for (int i = 0; i <= 2 * n.R; ++i) add(n.lin[i]); };
  addnet(L.enc); add(L.post); addnet(L.dec); add(L.ddist); addnet(L.part); add(L.ppost);
  PMVAE_CHECK(ok, "too many bias leaves for AdamSegs");
  return 0;
}

int adamw_step(const pmvae_config* c, float* params, const float* grads, float* m, float* v, int64_t count, float lr,
               float wd, float b1, float b2, float eps, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  PMVAE_CHECK(params && grads && m && v && count >= 0, "bad arguments");
  AdamSegs seg{};
  PMVAE_TRY(bias_ranges(L, &seg));
  const double t = (double)count + 1.0;
  const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2 = (float)(1.0 - pow((double)b2, t));
  return adamw(params, grads, m, v, L.total, seg, lr, wd, b1, b2, eps, bc1, bc2, s);
}

// evaluators (fp32 path): decoder over K*rows latent samples in chunks
static int eval_common(const pmvae_config* c, const Layout& L, const float* params, const float* x, const float* b,
                       int64_t B, const EvalPlan& p, bool need_enc, cudaStream_t s) {
  if (need_enc) PMVAE_TRY(net_fwd_f32(params, L.enc, L.post, c->H, x, B, p.enc, p.par_e, s));
  PMVAE_TRY(concat_masked(x, b, p.xob, B, c->D, s));
  PMVAE_TRY(net_fwd_f32(params, L.part, L.ppost, c->H, p.xob, B, p.part, p.par_p, s));
  return 0;
}

int is_log_prob(const pmvae_config* c, const float* params, const float* x, const float* b, int64_t B, int64_t K,
                const uint32_t key_z[2], const uint32_t key_zxo[2], int64_t B_total, int64_t row_start,
                float* out_log_p_x, float* out_cond, void* ws, uint64_t ws_bytes, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  if (B == 0) return 0;
  PMVAE_CHECK(params && x && b && key_z && key_zxo && ws, "null pointer");
  PMVAE_CHECK(K >= 1 && B >= 0 && row_start >= 0 && row_start + B <= B_total, "bad K / row range");
  if (B == 0) return 0;
  if (c->precision == PMVAE_PREC_BF16)
    return is_log_prob_bf16(c, L, params, x, b, B, K, key_z, key_zxo, B_total, row_start, out_log_p_x, out_cond, ws, ws_bytes, s);
  EvalPlan p = plan_eval(c, L, B, K, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_workspace_bytes)");
  PMVAE_TRY(eval_common(c, L, params, x, b, B, p, true, s));
  const float* ls = params + L.log_scale;
  for (int64_t r0 = 0; r0 < B; r0 += p.rows_per_chunk) {
    const int64_t nb = (B - r0 < p.rows_per_chunk) ? (B - r0) : p.rows_per_chunk;
    const int64_t M = nb * K;
    // z ~ q(z|x)
    PMVAE_TRY(sample_latents(p.par_e + r0 * L.P, Key2{key_z[0], key_z[1]}, nb, K, B_total, row_start + r0, c->d, p.z, p.base, s));
    PMVAE_TRY(net_fwd_f32(params, L.dec, L.ddist, c->H, p.z, M, p.dec, p.loc, s));
    PMVAE_TRY(eval_rows_ll(x + r0 * c->D, nullptr, p.loc, c->D, ls, p.base, p.llA, nb, K, c->D, s));
    if (out_log_p_x) PMVAE_TRY(logmeanexp_rows(p.llA, nullptr, out_log_p_x + r0, nb, K, s));
    if (out_cond) {
      // z' ~ q(z|x_o), observed dims only
      PMVAE_TRY(sample_latents(p.par_p + r0 * L.P, Key2{key_zxo[0], key_zxo[1]}, nb, K, B_total, row_start + r0, c->d, p.z, p.base, s));
      PMVAE_TRY(net_fwd_f32(params, L.dec, L.ddist, c->H, p.z, M, p.dec, p.loc, s));
      PMVAE_TRY(eval_rows_ll(x + r0 * c->D, b + r0 * c->D, p.loc, c->D, ls, p.base, p.llC, nb, K, c->D, s));
      PMVAE_TRY(logmeanexp_rows(p.llA, p.llC, out_cond + r0, nb, K, s));
    }
  }
  return 0;
}

int impute_mean_seq(const pmvae_config* c, const float* params, const float* x, const float* b, int64_t B, int64_t K,
                    const uint32_t key[2], int64_t B_total, int64_t row_start, float* out, float* out_samples, void* ws,
                    uint64_t ws_bytes, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  if (B == 0) return 0;
  PMVAE_CHECK(params && x && b && key && (out || out_samples) && ws, "null pointer");
  PMVAE_CHECK(K >= 1 && B >= 0 && row_start >= 0 && row_start + B <= B_total, "bad K / row range");
  if (B == 0) return 0;
  if (c->precision == PMVAE_PREC_BF16)
    return impute_mean_bf16(c, L, params, x, b, B, K, key, B_total, row_start, out, out_samples, ws, ws_bytes, s);
  EvalPlan p = plan_eval(c, L, B, K, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_workspace_bytes)");
  PMVAE_TRY(eval_common(c, L, params, x, b, B, p, false, s));
  for (int64_t r0 = 0; r0 < B; r0 += p.rows_per_chunk) {
    const int64_t nb = (B - r0 < p.rows_per_chunk) ? (B - r0) : p.rows_per_chunk;
    PMVAE_TRY(sample_latents(p.par_p + r0 * L.P, Key2{key[0], key[1]}, nb, K, B_total, row_start + r0, c->d, p.z, p.base, s));
    PMVAE_TRY(net_fwd_f32(params, L.dec, L.ddist, c->H, p.z, nb * K, p.dec, p.loc, s));
    if (out) PMVAE_TRY(impute_mean(x + r0 * c->D, b + r0 * c->D, p.loc, c->D, out + r0 * c->D, nb, K, c->D, s));
    if (out_samples)
      PMVAE_TRY(impute_samples(x + r0 * c->D, b + r0 * c->D, p.loc, c->D, out_samples + r0 * c->D, nb, B, K, c->D, s));
  }
  return 0;
}

int adamw_step_dev(const pmvae_config* c, float* params, const float* grads, float* m, float* v, float wd, float b1,
                   float b2, float eps, const StepState* st, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  AdamSegs seg{};
  PMVAE_TRY(bias_ranges(L, &seg));
  return adamw(params, grads, m, v, L.total, seg, 0.f, wd, b1, b2, eps, 1.f, 1.f, s, st);
}

// ---------------------------------------------------------------- AutoregressiveGMM (MNIST partial posterior)
struct ArgmmLayout { Net net; Leaf head; uint64_t total; int F, cols; };

static int build_argmm_layout(const pmvae_argmm_config* c, ArgmmLayout* L) {
  PMVAE_CHECK(c != nullptr, "null config");
  PMVAE_CHECK(c->d >= 1 && c->d <= 64 && c->n_comp >= 1 && c->n_comp <= 32 && c->C >= 0 && c->C <= 4096,
              "AutoregressiveGMM: event_size in [1, 64], num_components in [1, 32], context in [0, 4096]");
  PMVAE_CHECK(c->H >= 4 && c->H % 4 == 0 && c->H <= 1024 && c->R >= 0 && c->R <= kMaxBlocks, "bad hidden_units / residual_blocks");
  uint64_t off = 0;
  auto leaf = [&](Leaf& lf, int rows, int cols) {
    lf.rows = rows; lf.cols = cols;
    lf.w = off; off += pad64((uint64_t)rows * cols);
    lf.b = off; off += pad64((uint64_t)cols);
  };
  L->F = 2 * c->d + c->C;
  L->cols = 3 * c->n_comp * c->d;
  L->net.in_dim = L->F; L->net.R = c->R; L->net.ln = 0;
  leaf(L->net.lin[0], L->F, c->H);
  for (int i = 1; i <= 2 * c->R; ++i) leaf(L->net.lin[i], c->H, c->H);
  leaf(L->head, c->H, L->cols);
  L->total = off;
  return 0;
}

struct ArgmmPlan {
  NetSaved sv;
  float *X, *out, *dout, *dH, *tmp1, *tmp2, *dX, *dzd;
  void* img;                 // bf16 images of the hidden Linears (tensor-core route only)
  HiddenTc tc;
  bool use_tc;
  uint64_t bytes;
};

// pmvae_argmm_config.reserved[0] = 1: hidden Linears on the tcgen05 GEMMs with bf16 operands (the first Linear, the
// block-diagonal head and all mixture algebra stay float32)
static bool argmm_tc(const pmvae_argmm_config* c) { return c->reserved[0] == 1 && c->H == 256 && c->R >= 1; }

static ArgmmPlan plan_argmm(const pmvae_argmm_config* c, const ArgmmLayout& L, int64_t B, void* ws) {
  ArgmmPlan p{};
  Bump bp(ws);
  const int64_t M = (int64_t)c->d * B;
  plan_net(bp, L.net, M, c->H, p.sv, true);
  p.X = bp.take<float>((uint64_t)M * L.F);
  p.out = bp.take<float>((uint64_t)M * L.cols);
  p.dout = bp.take<float>((uint64_t)M * L.cols);
  p.dH = bp.take<float>((uint64_t)M * c->H);
  p.tmp1 = bp.take<float>((uint64_t)M * c->H);
  p.tmp2 = bp.take<float>((uint64_t)M * c->H);
  p.dX = bp.take<float>((uint64_t)M * L.F);
  p.dzd = bp.take<float>((uint64_t)B * c->d);
  p.use_tc = argmm_tc(c);
  if (p.use_tc) {
    p.img = bp.take<char>(hidden_images_bytes(2 * c->R));
    p.tc.xb = bp.take<__nv_bfloat16>((uint64_t)M * c->H);
    p.tc.dyb = bp.take<__nv_bfloat16>((uint64_t)M * c->H);
  }
  p.bytes = bp.off;
  return p;
}

int argmm_log_prob(const pmvae_argmm_config* c, const float* params, const float* z, const float* ctx, int64_t B,
                   float* out, void* ws, uint64_t ws_bytes, cudaStream_t s) {
  ArgmmLayout L;
  PMVAE_TRY(build_argmm_layout(c, &L));
  if (B == 0) return 0;
  PMVAE_CHECK(params && z && (ctx || c->C == 0) && out && ws && B >= 0, "null pointer");
  ArgmmPlan p = plan_argmm(c, L, B, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_argmm_workspace_bytes)");
  PMVAE_TRY(argmm_input(z, ctx, B, c->d, c->C, p.X, s));
  if (p.use_tc) PMVAE_TRY(hidden_images_pack(params, &L.net.lin[1], 2 * c->R, p.img, p.tc.wn, p.tc.wt, s));
  PMVAE_TRY(net_fwd_f32(params, L.net, L.head, c->H, p.X, (int64_t)c->d * B, p.sv, p.out, s, c->d, -1,
                        p.use_tc ? &p.tc : nullptr));
  return argmm_lp(p.out, z, B, c->d, c->n_comp, out, s);
}

// VJP of argmm_log_prob (whose intermediates are still in `ws`): cotangent g[B] -> parameter gradients
// (overwritten) and, optionally, d/dz [B, d] and d/dcontext [B, C].
int argmm_backward(const pmvae_argmm_config* c, const float* params, const float* z, const float* ctx, int64_t B,
                   const float* g, float* grads, float* dz, float* dctx, void* ws, uint64_t ws_bytes, cudaStream_t s) {
  ArgmmLayout L;
  PMVAE_TRY(build_argmm_layout(c, &L));
  PMVAE_CHECK(grads != nullptr && B >= 0, "null gradient arena / negative batch");
  PMVAE_CUDA(cudaMemsetAsync(grads, 0, L.total * sizeof(float), s));
  if (B == 0) return 0;
  PMVAE_CHECK(params && z && g && ws, "null pointer");
  ArgmmPlan p = plan_argmm(c, L, B, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_argmm_workspace_bytes)");
  const int64_t M = (int64_t)c->d * B;
  // (argmm_lp_bwd writes, and the block-diagonal head VJP reads, only the 3 K columns of step i in row block i)
  PMVAE_TRY(argmm_lp_bwd(p.out, z, g, B, c->d, c->n_comp, p.dout, p.dzd, s));
  // (tensor-core route: the bf16 weight images packed by argmm_log_prob are still in `ws`, like the activations)
  if (p.use_tc)
    for (int l = 0; l < 2 * c->R; ++l) {
      p.tc.wn[l] = reinterpret_cast<const __nv_bfloat16*>(p.img) + (uint64_t)(2 * l) * 256 * 256;
      p.tc.wt[l] = reinterpret_cast<const __nv_bfloat16*>(p.img) + (uint64_t)(2 * l + 1) * 256 * 256;
    }
  PMVAE_TRY(net_bwd_f32(params, grads, L.net, L.head, c->H, p.X, M, p.sv, p.dout, p.dH, p.tmp1, p.tmp2, p.dX, s, c->d,
                        p.use_tc ? &p.tc : nullptr));
  if (dz || dctx) PMVAE_TRY(argmm_reduce_dx(p.dX, p.dzd, B, c->d, c->C, dz, dctx, s));
  return 0;
}

// _AutoregressiveDistribution._sample_n (distributions.py:168-189): n samples per context row, one ResidualMLP pass over
// all n * B rows per latent dimension.  Workspace: the net's activations for n * B rows, the step input, the head output
// and the two [n, d, K] noise draws.
struct ArgmmSamplePlan { NetSaved sv; float *X, *out, *eps, *u; uint64_t bytes; };
static ArgmmSamplePlan plan_argmm_sample(const pmvae_argmm_config* c, const ArgmmLayout& L, int64_t M, int64_t n, void* ws) {
  ArgmmSamplePlan p{};
  Bump bp(ws);
  plan_net(bp, L.net, M, c->H, p.sv, false);
  p.X = bp.take<float>((uint64_t)M * L.F);
  p.out = bp.take<float>((uint64_t)M * L.cols);
  p.eps = bp.take<float>((uint64_t)n * c->d * c->n_comp);
  p.u = bp.take<float>((uint64_t)n * c->d * c->n_comp);
  p.bytes = bp.off;
  return p;
}

int argmm_sample(const pmvae_argmm_config* c, const float* params, const float* ctx, int64_t B, int64_t n,
                 const uint32_t key[2], float* out, void* ws, uint64_t ws_bytes, cudaStream_t s) {
  ArgmmLayout L;
  PMVAE_TRY(build_argmm_layout(c, &L));
  PMVAE_CHECK(B >= 0 && n >= 1, "bad sample count / batch");
  if (B == 0) return 0;
  PMVAE_CHECK(params && (ctx || c->C == 0) && key && out && ws, "null pointer");
  const int64_t M = n * B;
  ArgmmSamplePlan p = plan_argmm_sample(c, L, M, n, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_argmm_sample_workspace_bytes)");
  // TFP's MixtureSameFamily.sample splits its seed in two [R]: components first, then the categorical
  uint32_t ks[4];
  PMVAE_TRY(pmvae_key_split_host(key, 2, ks));
  const uint64_t nn = (uint64_t)n * c->d * c->n_comp;
  PMVAE_TRY(pmvae_normal(ks, nn, 0, nn, p.eps, s));
  PMVAE_TRY(pmvae_uniform(ks + 2, nn, 0, nn, p.u, s));
  PMVAE_CUDA(cudaMemsetAsync(out, 0, (size_t)M * c->d * sizeof(float), s));
  for (int i = 0; i < c->d; ++i) {
    PMVAE_TRY(argmm_sample_input(out, ctx, M, B, c->d, c->C, i, p.X, s));
    PMVAE_TRY(net_fwd_f32(params, L.net, L.head, c->H, p.X, M, p.sv, p.out, s, c->d, i));
    PMVAE_TRY(argmm_sample_step(p.out, p.eps, p.u, M, B, c->d, c->n_comp, i, out, s));
  }
  return 0;
}

int net_apply(const pmvae_config* c, const float* params, int which, const float* in, const float* msk, int64_t B,
              float* out, void* ws, uint64_t ws_bytes, cudaStream_t s) {
  Layout L;
  PMVAE_TRY(build_layout(c, &L));
  const bool save = (which & PMVAE_NET_SAVE) != 0;     // training-mode forward (fp32 path: it always saves)
  which &= ~PMVAE_NET_SAVE;
  PMVAE_CHECK(which >= 0 && which <= 2, "net id must be 0 (encoder), 1 (decoder) or 2 (partial encoder)");
  if (B == 0) return 0;
  PMVAE_CHECK(params && in && out && ws && B >= 0 && (which != 2 || msk), "null pointer");
  if (c->precision == PMVAE_PREC_BF16) return net_apply_bf16(c, L, params, which, in, msk, B, out, save, ws, ws_bytes, s);
  TrainPlan p = plan_train(c, L, B, ws);
  PMVAE_CHECK(p.bytes <= ws_bytes, "workspace too small (see pmvae_workspace_bytes)");
  if (which == 0) return net_fwd_f32(params, L.enc, L.post, c->H, in, B, p.enc, out, s);
  if (which == 1) return net_fwd_f32(params, L.dec, L.ddist, c->H, in, B, p.dec, out, s);
  PMVAE_TRY(concat_masked(in, msk, p.xob, B, c->D, s));
  return net_fwd_f32(params, L.part, L.ppost, c->H, p.xob, B, p.part, out, s);
}

}  // namespace pmvae

using namespace pmvae;

extern "C" {

const char* pmvae_last_error(void) { return pmvae::last_error(); }
int pmvae_version(void) { return 1; }
uint64_t pmvae_launch_count(void) { return g_launches.load(); }

int pmvae_tc_gemm_nt(const void* A, int64_t lda, const void* Bt, int64_t ldb, const float* bias, int64_t M, int32_t N,
                     int32_t K, float* y, pmvae_stream_t stream) {
  PMVAE_CHECK(A && Bt && y, "null pointer");
  tc::TcGemmArgs e{};
  e.bias = bias; e.out_f32 = y; e.ld_out_f32 = N;
  return tc::gemm_nt(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(Bt), ldb, M,
                     N, K, e, as_stream(stream));
}
int pmvae_tc_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N, int64_t rows,
                     float* y, pmvae_stream_t stream) {
  PMVAE_CHECK(A && B && y, "null pointer");
  return tc::gemm_tn(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb, M,
                     N, rows, y, N, 1, 0, nullptr, as_stream(stream));
}

int pmvae_linear(int32_t precision, const float* x, const float* w, const float* bias, int64_t B, int32_t K, int32_t N,
                 int32_t relu_in, float* y, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  PMVAE_CHECK(x && w && y && B >= 0 && K > 0 && N > 0, "bad arguments");
  if (precision == PMVAE_PREC_BF16) return linear_bf16(x, w, bias, B, K, N, relu_in, y, ws, ws_bytes, as_stream(stream));
  PMVAE_CHECK(precision == PMVAE_PREC_F32, "unknown precision");
  GemmF32Args a{};
  a.M = B; a.N = N; a.K = K;
  a.A = x; a.lda = K; a.B = w; a.ldb = N; a.C = y; a.ldc = N;
  a.bias = bias; a.relu_a = relu_in;
  return gemm_f32(a, false, false, as_stream(stream));
}

uint64_t pmvae_param_count(const pmvae_config* cfg) {
  Layout L;
  return build_layout(cfg, &L) == 0 ? L.total : 0;
}
int pmvae_layout(const pmvae_config* cfg, pmvae_leaf* out, int cap) { return export_layout(cfg, out, cap); }

uint64_t pmvae_workspace_bytes(const pmvae_config* cfg, int64_t B, int64_t K) {
  if (cfg && cfg->precision == PMVAE_PREC_BF16) return workspace_bytes_bf16(cfg, B, K);
  return workspace_bytes(cfg, B, K);
}

int pmvae_prepare_params(const pmvae_config* cfg, const float* params, void* ws, uint64_t ws_bytes,
                         pmvae_stream_t stream) {
  PMVAE_CHECK(cfg && params && ws, "null pointer");
  if (cfg->precision == PMVAE_PREC_BF16) return prepare_params_bf16(cfg, params, ws, ws_bytes, as_stream(stream));
  return 0;
}

int pmvae_forward(const pmvae_config* cfg, const float* params, const float* x, const float* b, const float* eps,
                  int64_t B, float* out_rec, float* out_kl, float* out_match, void* ws, uint64_t ws_bytes,
                  pmvae_stream_t stream) {
  return forward(cfg, params, x, b, eps, B, out_rec, out_kl, out_match, ws, ws_bytes, as_stream(stream));
}

int pmvae_backward(const pmvae_config* cfg, const float* params, const float* x, const float* b, const float* eps,
                   int64_t B, const float* g_rec, const float* g_kl, const float* g_match, float* grads, void* ws,
                   uint64_t ws_bytes, pmvae_stream_t stream) {
  return backward(cfg, params, x, b, eps, B, g_rec, g_kl, g_match, grads, ws, ws_bytes, as_stream(stream));
}

int pmvae_loss_cotangents(int64_t B, int64_t B_global, float beta, float coef, const float* rec, const float* kl,
                          const float* match, float* g_rec, float* g_kl, float* g_match, float* out_sums,
                          pmvae_stream_t stream) {
  PMVAE_CHECK(out_sums && (B == 0 || (rec && kl && match && g_rec && g_kl && g_match)), "null pointer");
  return loss_cotangents(B, B_global, beta, coef, rec, kl, match, g_rec, g_kl, g_match, out_sums, as_stream(stream));
}

int pmvae_adamw(const pmvae_config* cfg, float* params, const float* grads, float* m, float* v, int64_t count, float lr,
                float wd, float b1, float b2, float eps, pmvae_stream_t stream) {
  return adamw_step(cfg, params, grads, m, v, count, lr, wd, b1, b2, eps, as_stream(stream));
}

int pmvae_is_log_prob(const pmvae_config* cfg, const float* params, const float* x, const float* b, int64_t B,
                      int64_t K, const uint32_t key_z[2], const uint32_t key_zxo[2], int64_t B_total,
                      int64_t row_start, float* out_log_p_x, float* out_log_p_xu_given_xo, void* ws, uint64_t ws_bytes,
                      pmvae_stream_t stream) {
  return is_log_prob(cfg, params, x, b, B, K, key_z, key_zxo, B_total, row_start, out_log_p_x, out_log_p_xu_given_xo,
                     ws, ws_bytes, as_stream(stream));
}

uint64_t pmvae_argmm_param_count(const pmvae_argmm_config* cfg) {
  ArgmmLayout L;
  return build_argmm_layout(cfg, &L) == 0 ? L.total : 0;
}
int pmvae_argmm_layout(const pmvae_argmm_config* cfg, pmvae_leaf* out, int cap) {
  ArgmmLayout L;
  if (build_argmm_layout(cfg, &L) != 0) return -1;
  const int n = 2 * L.net.R + 2;
  for (int i = 0; i < n && i < cap; ++i) {
    pmvae_leaf l{};
    if (i <= 2 * L.net.R) name_leaf(&l, "partial_posterior_dist/residual_mlp", i, L.net.lin[i]);
    else name_leaf(&l, "partial_posterior_dist/one_dimensional_gmm", 0, L.head);
    out[i] = l;
  }
  return n;
}
uint64_t pmvae_argmm_workspace_bytes(const pmvae_argmm_config* cfg, int64_t B) {
  ArgmmLayout L;
  if (build_argmm_layout(cfg, &L) != 0) return 0;
  return plan_argmm(cfg, L, B < 1 ? 1 : B, nullptr).bytes + 256;
}
int pmvae_argmm_log_prob(const pmvae_argmm_config* cfg, const float* params, const float* z, const float* context,
                         int64_t B, float* out, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  return argmm_log_prob(cfg, params, z, context, B, out, ws, ws_bytes, as_stream(stream));
}
uint64_t pmvae_argmm_sample_workspace_bytes(const pmvae_argmm_config* cfg, int64_t B, int64_t n) {
  ArgmmLayout L;
  if (build_argmm_layout(cfg, &L) != 0 || n < 1) return 0;
  return plan_argmm_sample(cfg, L, n * (B < 1 ? 1 : B), n, nullptr).bytes + 256;
}
int pmvae_argmm_sample(const pmvae_argmm_config* cfg, const float* params, const float* context, int64_t B, int64_t n,
                       const uint32_t key[2], float* out, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  return argmm_sample(cfg, params, context, B, n, key, out, ws, ws_bytes, as_stream(stream));
}
int pmvae_logmeanexp_rows(const float* a, const float* c, float* out, int64_t B, int64_t K, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && K >= 1 && (B == 0 || (a && out)), "bad arguments");
  if (B == 0) return 0;
  return logmeanexp_rows(a, c, out, B, K, as_stream(stream));
}
int pmvae_argmm_backward(const pmvae_argmm_config* cfg, const float* params, const float* z, const float* context,
                         int64_t B, const float* g, float* grads, float* dz, float* dcontext, void* ws,
                         uint64_t ws_bytes, pmvae_stream_t stream) {
  return argmm_backward(cfg, params, z, context, B, g, grads, dz, dcontext, ws, ws_bytes, as_stream(stream));
}
int pmvae_bernoulli_ll(const float* logits, const float* x, const float* w, int64_t B, int32_t D, float* out,
                       pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && D >= 1 && (B == 0 || (logits && x && out)), "bad arguments");
  return bernoulli_ll(logits, x, w, B, D, out, as_stream(stream));
}
int pmvae_bernoulli_ll_backward(const float* logits, const float* x, const float* w, const float* g, int64_t B,
                                int32_t D, float* dlogits, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && D >= 1 && (B == 0 || (logits && x && g && dlogits)), "bad arguments");
  return bernoulli_ll_bwd(logits, x, w, g, B, D, dlogits, as_stream(stream));
}

// ---- building blocks exposed for models composed on the host (the MNIST config) ------------------
int pmvae_linear_backward(const float* x, const float* w, const float* dy, int64_t B, int32_t K, int32_t N,
                          int32_t relu_in, float* dx, float* dw, float* db, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && K > 0 && N > 0, "bad shape");
  if (B == 0) return 0;
  PMVAE_CHECK(x && w && dy, "null pointer");
  cudaStream_t s = as_stream(stream);
  Leaf lf{};
  lf.rows = K; lf.cols = N; lf.w = 0; lf.b = 0;
  if (dw) {
    GemmF32Args a{};
    a.M = K; a.N = N; a.K = B;
    a.A = x; a.lda = K; a.B = dy; a.ldb = N; a.C = dw; a.ldc = N;
    a.relu_a = relu_in; a.atomic = 1; a.split_k = split_for(K, N, B);
    PMVAE_TRY(gemm_f32(a, true, false, s));
  }
  if (db) PMVAE_TRY(colsum_add(dy, N, db, B, N, s));
  if (dx) {
    PMVAE_CHECK(!relu_in, "dx through a fused input relu is not provided: apply the mask on the caller's side");
    GemmF32Args a{};
    a.M = B; a.N = K; a.K = N;
    a.A = dy; a.lda = N; a.B = w; a.ldb = N; a.C = dx; a.ldc = K;
    PMVAE_TRY(gemm_f32(a, false, true, s));
  }
  return 0;
}

int pmvae_tril_sample_kl(const float* par, const float* eps, int64_t B, int32_t d, float* z, float* kl,
                         pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && (B == 0 || (par && eps && z && kl)), "bad arguments");
  return latent_fwd(par, eps, z, kl, B, d, as_stream(stream));
}
int pmvae_tril_sample_kl_backward(const float* par, const float* eps, const float* dz, const float* g_kl, int64_t B,
                                  int32_t d, float* dpar, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && (B == 0 || (par && eps && dz && g_kl && dpar)), "bad arguments");
  return tril_sample_bwd(par, eps, dz, g_kl, dpar, B, d, as_stream(stream));
}
int pmvae_adamw_flat(float* params, const float* grads, float* m, float* v, uint64_t n, int64_t count, float lr,
                     float wd, float b1, float b2, float eps, pmvae_stream_t stream) {
  PMVAE_CHECK(params && grads && m && v && count >= 0, "bad arguments");
  AdamSegs seg{};
  const double t = (double)count + 1.0;
  const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2 = (float)(1.0 - pow((double)b2, t));
  return adamw(params, grads, m, v, n, seg, lr, wd, b1, b2, eps, bc1, bc2, as_stream(stream));
}

int pmvae_net_apply(const pmvae_config* cfg, const float* params, int32_t which, const float* in, const float* msk,
                    int64_t B, float* out, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  PMVAE_CHECK(cfg != nullptr, "null config");
  return net_apply(cfg, params, which, in, msk, B, out, ws, ws_bytes, as_stream(stream));
}

int pmvae_impute_mean(const pmvae_config* cfg, const float* params, const float* x, const float* b, int64_t B,
                      int64_t K, const uint32_t key[2], int64_t B_total, int64_t row_start, float* out, void* ws,
                      uint64_t ws_bytes, pmvae_stream_t stream) {
  return impute_mean_seq(cfg, params, x, b, B, K, key, B_total, row_start, out, nullptr, ws, ws_bytes, as_stream(stream));
}

int pmvae_impute(const pmvae_config* cfg, const float* params, const float* x, const float* b, int64_t B, int64_t K,
                 const uint32_t key[2], int64_t B_total, int64_t row_start, float* out_samples, float* out_mean, void* ws,
                 uint64_t ws_bytes, pmvae_stream_t stream) {
  return impute_mean_seq(cfg, params, x, b, B, K, key, B_total, row_start, out_mean, out_samples, ws, ws_bytes,
                         as_stream(stream));
}

}  // extern "C"
