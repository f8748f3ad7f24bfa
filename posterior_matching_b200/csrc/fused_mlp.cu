// Fused ResidualMLP kernels for sm_100a: one persistent CTA per SM walks 128-row tiles through every
// Linear of a net (networks.py:111-135) and its distribution-head Linear (distributions.py:44,104)
// without the activations leaving the SM.
//
//   warp 0      : TMA producer, streams the bf16 weight K-blocks ([256 n] x [64 k], SWIZZLE_128B) of the
//                 current Linear through a 3-stage mbarrier ring (the weights come from L2)
//   warp 1      : MMA issuer: tcgen05.mma kind::f16, M = 128, N = 256 (head: N <= 256), K = 16; the whole warp runs
//                 the loop converged and one elect.sync lane issues (a lane-0 branch costs an R2UR waterfall per MMA)
//   warps 2..9  : epilogue: tcgen05.ld (thread = row) -> + bias -> relu -> bf16 -> st.shared straight into
//                 the swizzled K-major A-operand buffer of the NEXT Linear, 64 columns (= one K-block) at a
//                 time, so the next Linear's MMAs start while the rest of the tile is still being drained
//
// TMEM: columns [256, 512) hold the residual stream h (fp32): the first Linear writes it and the second
// Linear of every block ACCUMULATES into it (h += relu(...) @ W2), so the residual add costs nothing and
// h never leaves TMEM; columns [0, 256) take the block-internal Linear.  Biases are added when the
// accumulator is read (for h: the running sum b0 + sum_r b2_r).  Head tiles alternate between the two
// regions.  In training mode every bf16 operand tile is also streamed to HBM by TMA (the backward's
// weight-gradient operands) together with one relu bit per element.
// By default two CTAs of a cluster work as one tcgen05 pair (cta_group::2, M = 256): each holds half of every weight
// K-block, which halves the weight bytes through each SM's shared-memory port (PMVAE_FUSED_CTA2=0: single CTAs).
#include "fused_mlp.cuh"

#include <stdlib.h>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace pmvae {
namespace fused {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kThreads = 320;
constexpr int kFwdThreads = kThreads + 32;      // + 1 helper warp (training mode: TMA stores of the saved activations)
constexpr int kFwdThreadsLN = kFwdThreads + 32; // LayerNorm training mode: + 1 helper warp for the normalised pre-activations
constexpr int kEpiWarps = 8;
constexpr int kChunkBytes = 16384;             // [128 rows] x [64 k] bf16
constexpr int kOpndBytes = 4 * kChunkBytes;
constexpr int kWStageBytes = 32768;            // [256 n] x [64 k] bf16
constexpr int kWStages = 3;
constexpr int kInPitch = 65;                   // floats per staged input row (up to 64 values: x | mask)
constexpr int kInBytes = 128 * kInPitch * 4;   // fp32 input rows of the NEXT tile (x | mask), filled by cp.async
constexpr int kMaxLayers = 2 * kMaxBlocks + 1;
constexpr int kOffW = kOpndBytes;
constexpr int kOffL0 = kOffW + kWStages * kWStageBytes;   // A operand of the first Linear (one 64-wide K-block)
constexpr int kOffIn = kOffL0 + kChunkBytes;
constexpr int kOffBias = kOffIn + kInBytes;
constexpr int kOffBar = kOffBias + kMaxLayers * 1024;
constexpr int kSmemBytes = kOffBar + 512 + 1024;   // 227 KB: the whole opt-in shared memory of an SM
static_assert(kSmemBytes <= 232448, "shared-memory plan exceeds 227 KB");

struct FwdArgs {
  const float* in; const float* msk;
  int D_in, in_kind, in_lo, k16_0, R;   // in_lo: the first-Linear operand carries the lo(v) columns too
  int64_t B; int num_tiles;
  const float* bias[kMaxLayers];
  const float* head_bias; int head_N, head_NT, head_tiles;
  float* out; int64_t ld_out;     // head output [B, ld_out] fp32
  int head_tma;                   // 1: head rows staged in the (idle) operand buffer and stored by TMA
  int64_t Bpad; uint32_t* masks;
  float* rstd;                    // LayerNorm training mode: 1 / sigma of every LayerNorm, [(2R+1), Bpad]
  int in_wide;                    // masked input too wide for one K-block: first Linear = (x*b) @ W_x then += b @ W_b
  long long* trace;   // PMVAE_FUSED_TRACE: per-phase clock64 stamps of block 0 (profiling only)
  int debug;   // PMVAE_FUSED_DEBUG bits (profiling only): 2 no weight loads, 4 no MMAs, 8 coarse trace stamps
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }
// relu on a packed bf16 pair (rounding is monotonic, so relu(bf16(v)) == bf16(relu(v)))
__device__ __forceinline__ uint32_t relu2(uint32_t v) {
  uint32_t o;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(o) : "r"(v), "r"(0u));
  return o;
}

// bf16 pair of relu(a), relu(b) in one conversion (a in the low half)
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
  uint32_t o;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(b), "f"(a));
  return o;
}
// (x0, x1) += (b0, b1) as one packed fp32x2 add
__device__ __forceinline__ void add2(float& x0, float& x1, float b0, float b1) {
  asm("{\n\t.reg .b64 ta, tb;\n\t"
      "mov.b64 ta, {%0, %1};\n\t"
      "mov.b64 tb, {%2, %3};\n\t"
      "add.rn.f32x2 ta, ta, tb;\n\t"
      "mov.b64 {%0, %1}, ta;\n\t}"
      : "+f"(x0), "+f"(x1)
      : "f"(b0), "f"(b1));
}

// CTA2: two CTAs of a cluster run as one tcgen05 pair (cta_group::2, M = 256): each owns a 128-row tile and
// half of every weight K-block, so the weight traffic from L2 is halved and the ring covers twice the latency.
// LN: hk.LayerNorm(-1, False, False) after every Linear (networks.py:117-118,123-124,128-129; the bsds config).  The
// residual stream then cannot be accumulated by the MMA, so the epilogue keeps it in TMEM columns [256, 512) with
// tcgen05.ld / tcgen05.st and every hidden Linear writes columns [0, 256); forward only (evaluators).
// LN && SAVE (bsds training): besides the operand tiles and relu bits, every LayerNorm's normalised pre-activation
// xhat (signed, bf16: map_x, same slab layout as the operand stack) and 1 / sigma (p.rstd) are kept for the backward.
// xhat chunks are staged in the first-Linear operand buffer, which is idle between a tile's first Linear and the next
// tile's prologue, and stored by a second helper warp.
template <bool SAVE, bool CTA2, bool LN>
__global__ void __launch_bounds__(kFwdThreadsLN, 1)
net_fwd_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h,
               const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_o,
               const __grid_constant__ CUtensorMap map_x, FwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t opnd = sbase, wring = sbase + kOffW, l0buf = sbase + kOffL0, inbuf = sbase + kOffIn, bar = sbase + kOffBar;
  const float* inbuf_gen = reinterpret_cast<const float*>(sgen + kOffIn);
  float* bias_tbl = reinterpret_cast<float*>(sgen + kOffBias);
  constexpr int kStg = CTA2 ? 2 * kWStages : kWStages;             // ring stages
  constexpr int kStgBytes = CTA2 ? kWStageBytes / 2 : kWStageBytes;  // [128 or 256 n] x [64 k]
  constexpr int kArrive = CTA2 ? 2 * kEpiWarps : kEpiWarps;         // epilogue warps feeding one MMA issuer
  auto w_full = [&](int s) { return bar + 8u * s; };
  auto w_empty = [&](int s) { return bar + 8u * (kStg + s); };
  auto opnd_ready = [&](int c) { return bar + 8u * (2 * kStg + c); };
  auto acc_full = [&](int r) { return bar + 8u * (2 * kStg + 4 + r); };
  auto acc_empty = [&](int r) { return bar + 8u * (2 * kStg + 6 + r); };
  const uint32_t tmem_slot = bar + 8u * (2 * kStg + 8);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + kOffBar + 8 * (2 * kStg + 8));
  const uint32_t in_ready = bar + 8u * (2 * kStg + 9);      // first-Linear operand of the next tile is in l0buf
  // training mode: written(j) = operand chunk j is complete (this CTA's 8 epilogue warps), chunk_free(j) = the
  // helper warp's TMA store of it (the saved activation tile) has been read out
  auto written = [&](int c) { return bar + 8u * (2 * kStg + 10 + c); };
  auto chunk_free = [&](int c) { return bar + 8u * (2 * kStg + 14 + c); };
  // LN training mode: xs_written = an xhat chunk is complete in the staging buffer (8 epilogue warps), xs_free = its
  // TMA store has been read out;  l0_free = the MMAs of the first half of a wide first Linear have read l0buf
  const uint32_t xs_written = bar + 8u * (2 * kStg + 18), xs_free = bar + 8u * (2 * kStg + 19);
  const uint32_t l0_free = bar + 8u * (2 * kStg + 20);
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;               // 0 = the CTA that issues the pair's MMAs
  // barriers the MMA issuer waits on live in the leader CTA; the peer's warps arrive there remotely
  auto arrive_leader = [&](uint32_t b) {
    if (CTA2) mbar_arrive_cluster(mapa_shared(b, 0));
    else mbar_arrive(b);
  };
  // persistent schedule: a CTA (or CTA pair) takes every stride-th tile (pair of tiles)
  const int it_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int it_count = CTA2 ? (p.num_tiles + 1) / 2 : p.num_tiles;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
  const int R = p.R;
  const int n_hidden = 2 * R + 1;                 // Linears that feed the operand buffer
  const int nkb0 = (p.k16_0 + 3) >> 2;
  // accumulator region of step s: hidden Linears alternate (h lives in region 1 and is accumulated in place), except
  // under LayerNorm where they all land in region 0; head tiles alternate
  auto region_of = [&](int s_idx) { return (LN && s_idx < n_hidden) ? 0 : ((s_idx & 1) ? 0 : 1); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_h); tma_prefetch_desc(&map_o);
    if (SAVE) tma_prefetch_desc(&map_s);
    if (SAVE && LN) tma_prefetch_desc(&map_x);
    for (int s = 0; s < kStg; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int c = 0; c < 4; ++c) mbar_init(opnd_ready(c), kArrive);
    for (int r = 0; r < 2; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), kArrive); }
    mbar_init(in_ready, kArrive);
    for (int c = 0; c < 4; ++c) { mbar_init(written(c), kEpiWarps); mbar_init(chunk_free(c), 1); }
    mbar_init(xs_written, kEpiWarps); mbar_init(xs_free, 1); mbar_init(l0_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) { tmem_alloc_cta2(tmem_slot, 512); tmem_relinquish_cta2(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  // bias table: row l = what the epilogue of Linear l adds (running sum for the residual stream)
  for (int c = threadIdx.x; c < 256; c += kThreads) {
    float run = p.bias[0][c];
    bias_tbl[c] = run;
    for (int r = 0; r < R; ++r) {
      bias_tbl[(2 * r + 1) * 256 + c] = p.bias[2 * r + 1][c];
      run = LN ? p.bias[2 * r + 2][c] : run + p.bias[2 * r + 2][c];
      bias_tbl[(2 * r + 2) * 256 + c] = run;
    }
    // head bias (single head tile): row 2R+1, zero beyond the valid columns
    if (p.head_tiles == 1) bias_tbl[n_hidden * 256 + c] = c < p.head_N ? p.head_bias[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();          // both CTAs' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // trace: role 0 = MMA thread, role 1 = epilogue warp 0 lane 0; entries (tag, clock) from slot 1, count in slot 0
  int tr_n = 0;
  auto stamp = [&](int role, int tag) {
    if (p.trace && blockIdx.x == 0 && tr_n < 1000) {
      long long* t = p.trace + role * 2048;
      t[1 + 2 * tr_n] = tag; t[2 + 2 * tr_n] = clock64(); ++tr_n; t[0] = tr_n;
    }
  };

  if (warp == 0) {
    // ===================== weight producer (converged warp, one elected lane issues the TMA) =====================
    {
      int stage = 0; uint32_t ph = 0;
      // one K-block of `rows` weight rows starting at row c1 (CTA2: this CTA's half of them)
      auto load = [&](const CUtensorMap* m, int c0, int c1, int rows) {
        mbar_wait_x<CTA2>(w_empty(stage), ph ^ 1u, 1);
        const uint32_t dst = wring + stage * kStgBytes;
        __syncwarp();
        if (elect_one()) {
          if (CTA2) {
            const int half_rows = rows >> 1;
            if (rank == 0) mbar_arrive_expect_tx(w_full(stage), (uint32_t)rows * 128u);
            tma_load_2d_cta2(dst, m, mapa_shared(w_full(stage), 0), c0, c1 + (int)rank * half_rows);
          } else if (p.debug & 2) {
            mbar_arrive(w_full(stage));
          } else {
            mbar_arrive_expect_tx(w_full(stage), (uint32_t)rows * 128u);
            tma_load_2d(dst, m, w_full(stage), c0, c1);
          }
        }
        __syncwarp();
        if (++stage == kStg) { stage = 0; ph ^= 1u; }
      };
      for (int it = it_first; it < it_count; it += it_stride) {
        for (int kb = 0; kb < (p.in_wide ? 2 : nkb0); ++kb) load(&map_w, kb * 64, 0, 256);
        for (int l = 1; l < n_hidden; ++l)
          for (int kb = 0; kb < 4; ++kb) load(&map_w, kb * 64, l * 256, 256);
        for (int t = 0; t < p.head_tiles; ++t)
          for (int kb = 0; kb < 4; ++kb) load(&map_h, kb * 64, t * p.head_NT, p.head_NT);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA2: the leader CTA issues for the pair) =====================
    // The whole warp runs the loop converged; one elected lane issues (see elect_one in tc_ptx.cuh).
    if (rank == 0) {
      int stage = 0; uint32_t ph = 0;
      uint32_t ready_par = 0, use_cnt0 = 0, use_cnt1 = 0, in_par = 0;
      // wait_mode: 0 = operand already announced, 1 = chunk by chunk (opnd_ready), 2 = first-Linear buffer (in_ready)
      // Everything the issuing thread does per K-block is exposed once it exceeds the tensor pipe's small run-ahead
      // (scripts/probes/mma_probe.cu), so the loop is kept lean: descriptor low words are precomputed (a K=16 slice is
      // +2 in the address field), both barrier polls are in flight together, debug / trace flags are read once.
      const bool no_mma = (p.debug & 4) != 0;
      const bool tracing = p.trace != nullptr;
      const bool fine = tracing && !(p.debug & 8);       // debug bit 8: per-Linear stamps only
      constexpr uint32_t kDescHi = 0x40004040u;      // SBO 1024 | version 1 | SWIZZLE_128B (see smem_desc)
      auto desc_lo = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
      auto mk = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
      // part: 0 = a whole contraction; 1 / 2 = first / second half of a wide first Linear (the operand buffer is
      // rebuilt between them: the first half hands l0buf back through l0_free instead of announcing the accumulator)
      auto step = [&](int s_idx, uint32_t abuf, int nk16, int N, bool accum, int wait_mode, int part) {
        const int region = region_of(s_idx);
        uint32_t& uc = region ? use_cnt1 : use_cnt0;
        if (tracing && lane == 0) stamp(0, 100 + s_idx);
        if (part != 2) {
          mbar_wait_x<CTA2>(acc_empty(region), (uc & 1u) ^ 1u, 2);
          ++uc;
        }
        tc_fence_after();
        if (tracing && lane == 0) stamp(0, 200 + s_idx);
        const uint32_t d_tmem = tmem_base + (uint32_t)(region * 256);
        const uint32_t idesc = instr_desc(CTA2 ? 256 : 128, N, 0, 0);
        const int nkb = (nk16 + 3) >> 2;
        uint32_t a_lo = desc_lo(abuf);
        for (int kb = 0; kb < nkb; ++kb, a_lo += kChunkBytes >> 4) {
          // poll both barriers of this K-block before looking at either answer
          const uint32_t b_w = w_full(stage);
          bool ok_w = CTA2 ? mbar_try_wait_cluster(b_w, ph) : mbar_try_wait(b_w, ph);
          if (wait_mode == 1) {
            mbar_wait_x<CTA2>(opnd_ready(kb), (ready_par >> kb) & 1u, 3);
            ready_par ^= 1u << kb;
          } else if (wait_mode == 2) {
            mbar_wait_x<CTA2>(in_ready, in_par, 9);
            in_par ^= 1u;
          }
          if (fine && lane == 0) stamp(0, 300 + kb);
          if (!ok_w) mbar_wait_x<CTA2>(b_w, ph, 4);
          tc_fence_after();
          if (fine && lane == 0) stamp(0, 400 + kb);
          const uint32_t b_lo = desc_lo(wring + stage * kStgBytes);
          const int ks = no_mma ? 0 : min(4, nk16 - 4 * kb);
          __syncwarp();
          if (elect_one()) {
            if (ks == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (accum || kb > 0 || k > 0) ? 1u : 0u;
                if (CTA2) umma_f16_cta2(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, acc);
                else umma_f16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, acc);
              }
            } else {
              for (int k = 0; k < ks; ++k) {
                const uint32_t acc = (accum || kb > 0 || k > 0) ? 1u : 0u;
                if (CTA2) umma_f16_cta2(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, acc);
                else umma_f16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, acc);
              }
            }
            if (CTA2) umma_commit_cta2(w_empty(stage), 3); else umma_commit(w_empty(stage));
            if (kb == nkb - 1) {
              const uint32_t done = part == 1 ? l0_free : acc_full(region);
              if (CTA2) umma_commit_cta2(done, 3); else umma_commit(done);
            }
          }
          __syncwarp();
          if (++stage == kStg) { stage = 0; ph ^= 1u; }
        }
        if (tracing && lane == 0) stamp(0, 500 + s_idx);
      };
      for (int it = it_first; it < it_count; it += it_stride) {
        if (p.in_wide) {
          step(0, l0buf, p.k16_0, 256, false, 2, 1);
          step(0, l0buf, p.k16_0, 256, true, 2, 2);
        } else {
          step(0, l0buf, p.k16_0, 256, false, 2, 0);
        }
        for (int l = 1; l < n_hidden; ++l) step(l, opnd, 16, 256, !LN && (l & 1) == 0, 1, 0);
        for (int t = 0; t < p.head_tiles; ++t) step(n_hidden + t, opnd, 16, p.head_NT, false, t == 0 ? 1 : 0, 0);
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = ew >> 2;          // which 32 of the 64 columns of a chunk
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t full_par = 0;
    uint32_t n_writes = 0;             // operand tiles written so far (training mode: chunk_free phase tracking)
    uint32_t n_xs = 0;                 // xhat chunks staged so far (LN training mode: xs_free phase tracking)
    uint32_t l0_par = 0;               // wide first Linear: parity of l0_free
    const bool skip_bits = (p.debug & 32) != 0;      // profiling only: no relu-bit extraction
    const int D = p.D_in;

    // cp.async of a tile's fp32 input rows (x | mask) into the staging slots: thread (q, half 0, lane) copies row 32q+lane
    auto prefetch_input = [&](int tile_n) {
      if (half != 0) return;
      const int64_t gn = (int64_t)tile_n * 128 + row;
      const uint32_t dst = inbuf + (uint32_t)(row * kInPitch) * 4u;
      if (gn < p.B) {
        for (int k = 0; k < D; ++k) cp_async4(dst + 4u * k, p.in + gn * D + k);
        if (p.msk) for (int k = 0; k < D; ++k) cp_async4(dst + 4u * (D + k), p.msk + gn * D + k);
      }
      cp_async_commit();
    };
    // First-Linear operand of tile_n from the staged rows: hi/lo bf16 split of the fp32 input, [x*b, b] built
    // here (vae.py:132-133).  Lane = operand column (its source element and kind are fixed per lane), loop over the
    // warp's 32 rows: conflict-free LDS / STS and no per-element branching.
    auto prologue = [&](int tile_n) {
      if (half == 0) cp_async_wait_all();
      named_bar_sync(5 + q, 64);                     // the half-1 warp of this quadrant reads the same rows
      const int kk = 32 * half + lane;               // operand column 0..63
      int src = 0, kind = 3;                         // kind 0: hi(v), 1: lo(v), 2: mask, 3: zero
      if (kk < D) { src = kk; kind = 0; }
      else if (p.in_lo && kk < 2 * D) { src = kk - D; kind = 1; }
      else if (p.msk && kk < 3 * D) { src = kk - 2 * D; kind = 2; }
      const bool has_m = p.msk != nullptr;
      // branch-free per row (eight independent rows in flight): value = a * b with a = x or the mask itself,
      // b = mask or 1; the lo columns subtract the bf16 rounding; rows past B and unused columns give 0
      const int ia = (kind == 2) ? D + src : src;
      const int ib = (has_m && kind < 2) ? D + src : -1;
      const bool live = kind != 3;
      const bool is_lo = kind == 1;
      const int64_t left = p.B - ((int64_t)tile_n * 128 + q * 32);
      const int nvalid = left < 0 ? 0 : (left > 32 ? 32 : (int)left);
      const float* sbase_row = inbuf_gen + (q * 32) * kInPitch;
      const uint32_t col_off = ((kk & 7) << 1);
      const int kslot = kk >> 3;
      // eight rows per batch: all loads first (the volatile stores below would otherwise pin each row's loads behind
      // the previous row's store and expose the shared-memory latency 32 times)
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 8) {
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float* srow = sbase_row + (r0 + u) * kInPitch;
          a[u] = srow[ia];
          b[u] = ib >= 0 ? srow[ib] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = r0 + u;
          float v = (live && rr < nvalid) ? a[u] * b[u] : 0.f;
          const float lo = v - bf16_round(v);
          v = is_lo ? lo : v;
          const int r = q * 32 + rr;
          const uint32_t addr = l0buf + r * 128 + ((kslot ^ (r & 7)) << 4) + col_off;
          st_shared_u16(addr, __bfloat16_as_ushort(__float2bfloat16(v)));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_leader(in_ready);
    };

    // Wide masked input (2 D > 64, e.g. bsds: D = 63): the first Linear runs as two K-blocks through the one operand
    // buffer, [x*b] @ W[:D] then += b @ W[D:] (part 0 / 1), built straight from global memory (lane = column, the rows
    // are contiguous, so a warp reads whole rows; the latency is exposed once per part and tile).
    auto prologue_wide = [&](int tile_n, int part) {
      const int kk = 32 * half + lane;
      const int64_t g0 = (int64_t)tile_n * 128 + q * 32;
      const bool live = kk < D;
      const uint32_t col_off = ((kk & 7) << 1);
      const int kslot = kk >> 3;
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 8) {
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t gr = g0 + r0 + u;
          const bool ok = live && gr < p.B;
          b[u] = ok ? __ldg(p.msk + gr * D + kk) : 0.f;
          a[u] = (ok && part == 0) ? __ldg(p.in + gr * D + kk) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = q * 32 + r0 + u;
          const uint32_t addr = l0buf + r * 128 + ((kslot ^ (r & 7)) << 4) + col_off;
          st_shared_u16(addr, __bfloat16_as_ushort(__float2bfloat16(a[u] * b[u])));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_leader(in_ready);
    };
    auto first_operand = [&](int tile_n) {
      if (SAVE && LN && !(p.debug & 64)) mbar_wait(xs_free, (n_xs & 1u) ^ 1u, 16);     // the last staged xhat chunk has been stored
      if (p.in_wide) prologue_wide(tile_n, 0);
      else prologue(tile_n);
    };

    if (it_first < it_count && !p.in_wide) prefetch_input(CTA2 ? 2 * it_first + (int)rank : it_first);
    if (it_first < it_count) first_operand(CTA2 ? 2 * it_first + (int)rank : it_first);
    for (int it = it_first; it < it_count; it += it_stride) {
      const int tile = CTA2 ? 2 * it + (int)rank : it;
      const bool has_next = it + it_stride < it_count;
      const int tile_next = CTA2 ? 2 * (it + it_stride) + (int)rank : it + it_stride;
      const int64_t g = (int64_t)tile * 128 + row;
      const bool row_ok = g < p.B;
      if (ew == 0 && lane == 0) stamp(1, 1000);
      // ---- hidden Linears: accumulator -> bf16 operand of the next Linear
      // LayerNorm epilogue of hidden Linear l: y = acc + b_l; xhat = (y - mean) / sqrt(var + 1e-5) over the 256
      // columns of the row (each of the two warps of a quadrant holds 128 of them; partial sums meet in shared memory);
      // l even: h (+)= xhat kept in TMEM, operand = relu(h);  l odd: operand = relu(xhat).
      auto epi_ln = [&](int l, int64_t g_row) {
        mbar_wait_x<CTA2>(acc_full(0), full_par & 1u, 5);
        full_par ^= 1u;
        tc_fence_after();
        if (l == 0 && p.head_tma) {
          if (lane == 0) tma_store_wait_read<0>();
          named_bar_sync(1 + q, 64);
        }
        // Row statistics in ONE pass over the accumulator: each of the quadrant's two warps sums its 128 columns shifted
        // by its own first value c (S = sum(v - c), Q = sum((v - c)^2): no cancellation), the halves are combined
        // exactly: mean = (S0 + S1 + 128 (c0 + c1)) / 256, sum (v - mean)^2 = sum_h Q_h - 2 (mean - c_h) S_h + 128 (mean - c_h)^2.
        float* xch = bias_tbl + 14 * 256;                       // [128 rows][2 halves][S, Q, c]
        const uint32_t t_u = t_lane + (uint32_t)(32 * half);
        const uint32_t t_h = t_lane + (uint32_t)(256 + 32 * half);
        const float4* bq = reinterpret_cast<const float4*>(bias_tbl + l * 256 + 32 * half);   // chunk j: bq[16 j + i]
        uint32_t ra[32], rb[32];
        tmem_ld32(t_u, ra);
        float S = 0.f, Q = 0.f, c = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t (&r)[32] = (j & 1) ? rb : ra;
          tmem_ld_wait();
          if (j < 3) tmem_ld32(t_u + 64 * (j + 1), (j & 1) ? ra : rb);
          if (j == 0) c = __uint_as_float(r[0]) + bq[0].x;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = bq[16 * j + i];
            const float d0 = __uint_as_float(r[4 * i]) + bv.x - c, d1 = __uint_as_float(r[4 * i + 1]) + bv.y - c;
            const float d2 = __uint_as_float(r[4 * i + 2]) + bv.z - c, d3 = __uint_as_float(r[4 * i + 3]) + bv.w - c;
            S += (d0 + d1) + (d2 + d3);
            Q = fmaf(d0, d0, Q); Q = fmaf(d1, d1, Q); Q = fmaf(d2, d2, Q); Q = fmaf(d3, d3, Q);
          }
        }
        float* mine = xch + (row * 2 + half) * 3;
        mine[0] = S; mine[1] = Q; mine[2] = c;
        tmem_ld32(t_u, ra);                                      // second pass: prefetch its first chunk across the exchange
        named_bar_sync(1 + q, 64);
        const float* other = xch + (row * 2 + (half ^ 1)) * 3;
        const float So = other[0], Qo = other[1], co = other[2];
        const float mean = (S + So + 128.f * (c + co)) * (1.0f / 256.0f);
        const float dm = mean - c, dmo = mean - co;
        float ss = (Q - 2.f * dm * S + 128.f * dm * dm) + (Qo - 2.f * dmo * So + 128.f * dmo * dmo);
        ss = fmaxf(ss, 0.f);
        const float rstd = rsqrtf(ss * (1.0f / 256.0f) + 1e-5f);
        uint32_t mw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t (&r)[32] = (j & 1) ? rb : ra;
          uint32_t hreg[32];
          if ((l & 1) == 0 && l > 0) tmem_ld32(t_h + 64 * j, hreg);
          tmem_ld_wait();
          if (j < 3) tmem_ld32(t_u + 64 * (j + 1), (j & 1) ? ra : rb);
          if (j == 3) {
            // the accumulator has been read for the last time: hand it back now, so that the next Linear's first
            // K-blocks (their operand chunks are already announced) run under this chunk's math and stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(acc_empty(0));
          }
          uint32_t pk[16];
          uint32_t xk[16];           // training mode: xhat itself (signed), kept for the LayerNorm backward
          uint32_t neg = 0;          // sign bits of what the relu sees, element 0 ends up in bit 31
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = bq[16 * j + i];
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
            float xh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = (__uint_as_float(r[4 * i + e]) + bb[e] - mean) * rstd;
              xh[e] = v;
              if ((l & 1) == 0) {
                if (l > 0) v += __uint_as_float(hreg[4 * i + e]);
                hreg[4 * i + e] = __float_as_uint(v);
              }
              r[4 * i + e] = __float_as_uint(v);
              if (SAVE) neg = __funnelshift_l(__float_as_uint(v), neg, 1);
            }
            if (SAVE) { xk[2 * i] = pack2(xh[0], xh[1]); xk[2 * i + 1] = pack2(xh[2], xh[3]); }
          }
          mw[j] = ~neg;
          if ((l & 1) == 0) tmem_st32(t_h + 64 * j, hreg);
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack2_relu(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          if (SAVE) mbar_wait(chunk_free(j), (n_writes & 1u) ^ 1u, 13);   // the previous tile in this chunk has been stored
          const uint32_t rowaddr = opnd + j * kChunkBytes + row * 128;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int slot = (half * 4 + i4) ^ (row & 7);
            st_shared_v4(rowaddr + slot * 16, pk[4 * i4], pk[4 * i4 + 1], pk[4 * i4 + 2], pk[4 * i4 + 3]);
          }
          if (SAVE && !(p.debug & 64)) {
            // xhat chunk -> staging (the idle first-Linear operand buffer), same swizzled row layout
            mbar_wait(xs_free, (n_xs & 1u) ^ 1u, 18);
            ++n_xs;
            const uint32_t xaddr = l0buf + row * 128;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int slot = (half * 4 + i4) ^ (row & 7);
              st_shared_v4(xaddr + slot * 16, xk[4 * i4], xk[4 * i4 + 1], xk[4 * i4 + 2], xk[4 * i4 + 3]);
            }
          }
          if (p.debug & 256) tmem_st_wait();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) arrive_leader(opnd_ready(j));
          if (SAVE && lane == 0) { mbar_arrive(written(j)); if (!(p.debug & 64)) mbar_arrive(xs_written); }
        }
        if (SAVE) {
          ++n_writes;
          if (p.masks && !(p.debug & 128))
            *reinterpret_cast<uint4*>(p.masks + (((int64_t)l * p.Bpad + g_row) * 8 + half * 4)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
          if (half == 0 && p.rstd && !(p.debug & 128)) p.rstd[(int64_t)l * p.Bpad + g_row] = rstd;
        }
        tmem_st_wait();
      };
      for (int l = 0; l < n_hidden; ++l) {
        const int region = region_of(l);
        // the next tile's input rows: requested a whole tile ahead when the staging area is free (one head tile),
        // else while the last hidden Linear is drained (several head tiles reuse the area as head staging)
        if (l == (p.head_tiles == 1 ? 0 : n_hidden - 1) && has_next && !p.in_wide) prefetch_input(tile_next);
        if (l == 0 && p.in_wide) {
          // second half of the wide first Linear: the first half's MMAs have read the operand buffer
          mbar_wait(l0_free, l0_par, 17);
          l0_par ^= 1u;
          prologue_wide(tile, 1);
        }
        if (LN) { epi_ln(l, g); continue; }
        if (l == 0 && p.head_tma) {
          // the previous tile's head rows were staged in the operand buffer: their TMA stores must have been read out
          if (lane == 0) tma_store_wait_read<0>();
          named_bar_sync(1 + q, 64);
        }
        if (ew == 0 && lane == 0) stamp(1, 1100 + l);
        mbar_wait_x<CTA2>(acc_full(region), (full_par >> region) & 1u, 5);
        if (ew == 0 && lane == 0) stamp(1, 1200 + l);
        full_par ^= 1u << region;
        tc_fence_after();
        const uint32_t t_acc = t_lane + (uint32_t)(region * 256 + 32 * half);
        uint32_t mw[4];
        uint32_t ra[32], rb[32];
        tmem_ld32(t_acc, ra);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // the accumulator columns of chunk j+1 are fetched while chunk j is converted
          uint32_t (&r)[32] = (j & 1) ? rb : ra;
          tmem_ld_wait();
          if (j < 3) tmem_ld32(t_acc + 64 * (j + 1), (j & 1) ? ra : rb);
          const float4* bp = reinterpret_cast<const float4*>(bias_tbl + l * 256 + 64 * j + 32 * half);
          uint32_t pk[16];
          uint32_t neg = 0;          // sign bits of the pre-activations, element 0 ends up in bit 31
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = bp[i];
            float v0 = __uint_as_float(r[4 * i]), v1 = __uint_as_float(r[4 * i + 1]);
            float v2 = __uint_as_float(r[4 * i + 2]), v3 = __uint_as_float(r[4 * i + 3]);
            add2(v0, v1, bv.x, bv.y);
            add2(v2, v3, bv.z, bv.w);
            if (SAVE && !skip_bits) {
              neg = __funnelshift_l(__float_as_uint(v0), neg, 1);
              neg = __funnelshift_l(__float_as_uint(v1), neg, 1);
              neg = __funnelshift_l(__float_as_uint(v2), neg, 1);
              neg = __funnelshift_l(__float_as_uint(v3), neg, 1);
            }
            pk[2 * i] = pack2_relu(v0, v1);
            pk[2 * i + 1] = pack2_relu(v2, v3);
          }
          mw[j] = ~neg;
          if (SAVE) mbar_wait(chunk_free(j), (n_writes & 1u) ^ 1u, 13);   // the previous tile in this chunk has been stored
          const uint32_t rowaddr = opnd + j * kChunkBytes + row * 128;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int slot = (half * 4 + i4) ^ (row & 7);
            st_shared_v4(rowaddr + slot * 16, pk[4 * i4], pk[4 * i4 + 1], pk[4 * i4 + 2], pk[4 * i4 + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) arrive_leader(opnd_ready(j));
          if (ew == 0 && lane == 0) stamp(1, 1300 + j);
          if (SAVE && lane == 0) mbar_arrive(written(j));
        }
        ++n_writes;
        if (SAVE && p.masks)
          *reinterpret_cast<uint4*>(p.masks + (((int64_t)l * p.Bpad + g) * 8 + half * 4)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(acc_empty(region));
      }
      // ---- the next tile's first Linear can run while this tile's head is stored
      if (has_next) first_operand(tile_next);
      // several head tiles are staged in the input area: every warp must be done reading the next tile's input rows
      // (prologue) before any warp overwrites them
      if (has_next && p.head_tiles > 1 && p.head_tma && !p.in_wide) named_bar_sync(9, kEpiWarps * 32);
      // ---- head Linear: accumulator + bias -> fp32 rows (thread = row, 128 contiguous bytes per 32 columns)
      for (int t = 0; t < p.head_tiles; ++t) {
        const int region = region_of(n_hidden + t);
        if (ew == 0 && lane == 0) stamp(1, 1400 + t);
        mbar_wait_x<CTA2>(acc_full(region), (full_par >> region) & 1u, 6);
        if (ew == 0 && lane == 0) stamp(1, 1500 + t);
        full_par ^= 1u << region;
        tc_fence_after();
        if (p.head_tma && t == 0) {
          // the operand buffer is idle now (every MMA that read it has retired); in training mode its own TMA
          // stores (the last activation tile) must have been read out before it is reused as staging
          if (SAVE) {
            mbar_wait(chunk_free(2 * half), (n_writes & 1u) ^ 1u, 14);
            mbar_wait(chunk_free(2 * half + 1), (n_writes & 1u) ^ 1u, 14);
          }
        }
        int kpiece = 0;
        for (int pc = half; pc * 32 < p.head_NT; pc += 2, ++kpiece) {
          const int nb = t * p.head_NT + pc * 32;
          if (nb >= p.head_N) break;
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(region * 256 + pc * 32), r);
          tmem_ld_wait();
          if (p.head_tma) {
            // One head tile: the operand buffer is idle, two 4 KB staging tiles per warp inside its quadrant's rows.
            // Several head tiles: later tiles still multiply the operand, so the (by now consumed) input staging
            // area is used instead, one tile per warp.
            uint32_t st;
            if (p.head_tiles == 1) {
              st = opnd + (uint32_t)((2 * half + (kpiece & 1)) * kChunkBytes + q * 4096);
              if (kpiece >= 2) {
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
              }
            } else {
              st = inbuf + (uint32_t)(ew * 4096);
              if (t > 0 || kpiece > 0) {
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
              }
            }
            if (p.head_tiles == 1) {
              // all eight bias vectors first (independent broadcast loads), then packed adds and the staging stores
              const float4* bp = reinterpret_cast<const float4*>(bias_tbl + n_hidden * 256 + nb);
              float4 bv[8];
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) bv[i4] = bp[i4];
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                float v0 = __uint_as_float(r[4 * i4]), v1 = __uint_as_float(r[4 * i4 + 1]);
                float v2 = __uint_as_float(r[4 * i4 + 2]), v3 = __uint_as_float(r[4 * i4 + 3]);
                add2(v0, v1, bv[i4].x, bv[i4].y);
                add2(v2, v3, bv[i4].z, bv[i4].w);
                const int slot = i4 ^ (lane & 7);
                st_shared_v4(st + lane * 128 + slot * 16, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2),
                             __float_as_uint(v3));
              }
            } else {
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int n = nb + 4 * i4 + i;
                  v[i] = __uint_as_float(r[4 * i4 + i]) + (n < p.head_N ? __ldg(p.head_bias + n) : 0.f);
                }
                const int slot = i4 ^ (lane & 7);
                st_shared_v4(st + lane * 128 + slot * 16, __float_as_uint(v[0]), __float_as_uint(v[1]),
                             __float_as_uint(v[2]), __float_as_uint(v[3]));
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_o, st, nb, (int)((int64_t)tile * 128 + q * 32));
              tma_store_commit();
            }
          } else if (row_ok) {
            float* orow = p.out + g * p.ld_out;
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const int n0 = nb + 4 * i4;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (n0 + i < p.head_N) orow[n0 + i] = __uint_as_float(r[4 * i4 + i]) + __ldg(p.head_bias + n0 + i);
            }
          }
          if (ew == 0 && lane == 0) stamp(1, 1600 + pc);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(acc_empty(region));
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (SAVE && warp == 2 + kEpiWarps) {
    // ===================== helper warp (training mode): streams every finished operand chunk to HBM =====================
    // Two stores are kept in flight: chunk j is handed back once the store of chunk j+1 has been issued and the one
    // before it has been read out (waiting for each store's read before issuing the next made this warp the
    // bottleneck of the training-mode forward); everything is flushed at the end of each Linear, long before the
    // next Linear's epilogue (or the head staging) needs the chunk.
    uint32_t cnt = 0;
    for (int it = it_first; it < it_count; it += it_stride) {
      const int tile = CTA2 ? 2 * it + (int)rank : it;
      for (int l = 0; l < n_hidden; ++l, ++cnt) {
        for (int j = 0; j < 4; ++j) {
          mbar_wait(written(j), cnt & 1u, 15);
          if (lane == 0) {
            if (!(p.debug & 16)) tma_store_2d(&map_s, opnd + j * kChunkBytes, 64 * j, (int)((int64_t)l * p.Bpad + (int64_t)tile * 128));
            tma_store_commit();
            if (j > 0) {
              tma_store_wait_read<1>();
              mbar_arrive(chunk_free(j - 1));
            }
          }
          __syncwarp();
        }
        if (lane == 0) {
          tma_store_wait_read<0>();
          mbar_arrive(chunk_free(3));
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (SAVE && LN && warp == 3 + kEpiWarps) {
    // ===================== second helper warp (LayerNorm training mode): xhat chunks, one staging buffer =====================
    uint32_t cnt = 0;
    for (int it = it_first; it < it_count && !(p.debug & 64); it += it_stride) {
      const int tile = CTA2 ? 2 * it + (int)rank : it;
      for (int l = 0; l < n_hidden; ++l) {
        for (int j = 0; j < 4; ++j, ++cnt) {
          mbar_wait(xs_written, cnt & 1u, 19);
          if (lane == 0) {
            tma_store_2d(&map_x, l0buf, 64 * j, (int)((int64_t)l * p.Bpad + (int64_t)tile * 128));
            tma_store_commit();
            tma_store_wait_read<0>();
            mbar_arrive(xs_free);
          }
          __syncwarp();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();          // the peer may still be signalling this CTA's barriers / reading its TMEM
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_cta2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- backward chain
// Input-gradient chain of one ResidualMLP + head over 128-row tiles (the VJP of net_fwd_kernel with respect
// to the activations).  Per tile, with s = dL/dh (the gradient of the residual stream) and M_l the relu bits
// the forward saved for Linear l:
//   s    = M_2R . (dHead @ W_head^T)
//   for r = R-1 .. 0:   dU = M_{2r+1} . (s @ W_{2r+2}^T);   s += M_{2r} . (dU @ W_{2r+1}^T)
//   dIn  = s @ W_0^T                                         (decoder only: dL/dz)
// s and dU live in two bf16 operand buffers in shared memory (each is the A operand of the next
// contraction, handed over 64 columns at a time exactly like the forward); every version of them is also
// streamed to HBM by TMA as dY_l, the operand of the weight-gradient GEMM of Linear l, and their column sums
// (the bias gradients) are accumulated in registers and flushed once per CTA.
constexpr int kBwdOffU = kOpndBytes;
constexpr int kBwdOffW = 2 * kOpndBytes;
constexpr int kBwdOffBar = kBwdOffW + kWStages * kWStageBytes;
constexpr int kBwdSmemBytes = kBwdOffBar + 256;
constexpr int kBwdThreads = kThreads + 4 * 32;   // + 4 helper warps (TMA stores of dY, bias-gradient sums)

struct BwdArgs {
  int64_t B; int num_tiles;
  int k16_h;                  // K = 16 steps of the head contraction (ceil(head_N / 16))
  int din_N, din_cols;        // dIn: MMA N (multiple of 16, 0 = none) and valid columns
  float* dIn;                 // [B, din_cols] fp32
  const uint32_t* masks; int64_t Bpad;
  float* db[kMaxLayers];      // bias-gradient destinations (atomicAdd), Linear 0..2R
  int debug;                  // PMVAE_FUSED_DEBUG bits (profiling only): 16 no column sums, 32 no dY stores
};

// (A version with R as a run-time argument and rolled per-Linear loops -- 27 KB of SASS instead of 82 KB -- measured 2 %
// slower on the power step, 1.281 vs 1.258 ms: the instruction caches are not what paces this kernel, unlike
// net_bwd_ln_kernel below.)
template <int R, bool DIN>
__global__ void __launch_bounds__(kBwdThreads, 1)
net_bwd_kernel(const __grid_constant__ CUtensorMap map_dh, const __grid_constant__ CUtensorMap map_wh,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w0,
               const __grid_constant__ CUtensorMap map_dy, BwdArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("pmvae net_bwd_kernel: shared memory base is not 1024-byte aligned\n"); __trap(); }
  const uint32_t sbuf = sbase, ubuf = sbase + kBwdOffU, wring = sbase + kBwdOffW, bar = sbase + kBwdOffBar;
  auto w_full = [&](int s) { return bar + 8u * s; };
  auto w_empty = [&](int s) { return bar + 8u * (kWStages + s); };
  auto opnd_ready = [&](int c) { return bar + 8u * (2 * kWStages + c); };
  auto acc_full = [&](int r) { return bar + 8u * (2 * kWStages + 4 + r); };
  auto acc_empty = [&](int r) { return bar + 8u * (2 * kWStages + 6 + r); };
  const uint32_t dh_full = bar + 8u * (2 * kWStages + 8);
  const uint32_t ubuf_free = bar + 8u * (2 * kWStages + 9);
  // chunk hand-offs to the helper warps: written[j] = chunk j of the buffer being filled is complete (8 epilogue
  // warps); chunk_free(buf, j) = its TMA store has been read out and its column sums taken (helper j)
  auto written = [&](int c) { return bar + 8u * (2 * kWStages + 10 + c); };
  auto chunk_free = [&](int buf, int c) { return bar + 8u * (2 * kWStages + 14 + 4 * buf + c); };
  const uint32_t tmem_slot = bar + 8u * (2 * kWStages + 22);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + kBwdOffBar + 8 * (2 * kWStages + 22));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
  constexpr int n_hidden = 2 * R + 1;
  const int nkb_h = (p.k16_h + 3) >> 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dh); tma_prefetch_desc(&map_wh); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_dy);
    if (DIN) tma_prefetch_desc(&map_w0);
    for (int s = 0; s < kWStages; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int c = 0; c < 4; ++c) mbar_init(opnd_ready(c), kEpiWarps);
    for (int r = 0; r < 2; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), kEpiWarps); }
    mbar_init(dh_full, 1);
    mbar_init(ubuf_free, 4);
    for (int c = 0; c < 4; ++c) { mbar_init(written(c), kEpiWarps); mbar_init(chunk_free(0, c), 1); mbar_init(chunk_free(1, c), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== producer: dHead tiles and weight K-blocks (converged warp, elected lane issues) =====================
    {
      int stage = 0; uint32_t ph = 0; uint32_t uf_cnt = 0;
      auto load = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes) {
        mbar_wait(w_empty(stage), ph ^ 1u, 1);
        __syncwarp();
        if (elect_one()) {
          mbar_arrive_expect_tx(w_full(stage), bytes);
          tma_load_2d(wring + stage * kWStageBytes, m, w_full(stage), c0, c1);
        }
        __syncwarp();
        if (++stage == kWStages) { stage = 0; ph ^= 1u; }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        // dHead rows of this tile go into the dU buffer once the previous tile no longer needs it
        mbar_wait(ubuf_free, (uf_cnt & 1u) ^ 1u, 7);
        ++uf_cnt;
        __syncwarp();
        if (elect_one()) {
          mbar_arrive_expect_tx(dh_full, (uint32_t)nkb_h * kChunkBytes);
          for (int kb = 0; kb < nkb_h; ++kb) tma_load_2d(ubuf + kb * kChunkBytes, &map_dh, dh_full, kb * 64, tile * 128);
        }
        __syncwarp();
        for (int kb = 0; kb < nkb_h; ++kb) load(&map_wh, kb * 64, 0, kWStageBytes);
        for (int l = 2 * R; l >= 1; --l)
          for (int kb = 0; kb < 4; ++kb) load(&map_w, kb * 64, (l - 1) * 256, kWStageBytes);
        if (DIN)
          for (int kb = 0; kb < 4; ++kb) load(&map_w0, kb * 64, 0, (uint32_t)p.din_N * 128u);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues; see elect_one) =====================
    {
      int stage = 0; uint32_t ph = 0;
      uint32_t ready_par = 0, use_cnt0 = 0, use_cnt1 = 0, dh_par = 0;
      constexpr uint32_t kDescHi = 0x40004040u;      // SBO 1024 | version 1 | SWIZZLE_128B (see smem_desc)
      auto desc_lo = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
      auto mk = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
      // lean K-loop (see net_fwd_kernel): precomputed descriptor words, both barrier polls in flight together
      auto step = [&](int s_idx, uint32_t abuf, int nk16, int N, bool wait_opnd) {
        const int region = s_idx & 1;
        uint32_t& uc = region ? use_cnt1 : use_cnt0;
        mbar_wait(acc_empty(region), (uc & 1u) ^ 1u, 2);
        ++uc;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(region * 256);
        const uint32_t idesc = instr_desc(128, N, 0, 0);
        const int nkb = (nk16 + 3) >> 2;
        uint32_t a_lo = desc_lo(abuf);
        for (int kb = 0; kb < nkb; ++kb, a_lo += kChunkBytes >> 4) {
          const uint32_t b_w = w_full(stage);
          const bool ok_w = mbar_try_wait(b_w, ph);
          if (wait_opnd) {
            mbar_wait(opnd_ready(kb), (ready_par >> kb) & 1u, 3);
            ready_par ^= 1u << kb;
          }
          if (!ok_w) mbar_wait(b_w, ph, 4);
          tc_fence_after();
          const uint32_t b_lo = desc_lo(wring + stage * kWStageBytes);
          const int ks = min(4, nk16 - 4 * kb);
          __syncwarp();
          if (elect_one()) {
            if (ks == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            } else {
              for (int k = 0; k < ks; ++k) umma_f16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(w_empty(stage));
            if (kb == nkb - 1) umma_commit(acc_full(region));
          }
          __syncwarp();
          if (++stage == kWStages) { stage = 0; ph ^= 1u; }
        }
      };
      int t = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(dh_full, dh_par, 8);
        dh_par ^= 1u;
        tc_fence_after();
        // t runs on across tiles: consecutive steps alternate TMEM regions, so the first contraction of a tile
        // overlaps the last epilogue of the previous one
        step(t++, ubuf, p.k16_h, 256, false);
        for (int r = R - 1; r >= 0; --r) {
          step(t++, sbuf, 16, 256, true);
          step(t++, ubuf, 16, 256, true);
        }
        if (DIN) step(t++, sbuf, 16, p.din_N, true);
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t full_par = 0;
    uint32_t nwr[2] = {0, 0};          // writes so far into each operand buffer (all four chunks move together)

    // One epilogue step: accumulator (region) . mask_l [+ old s] -> bf16 chunk of buffer `buf` (0 = s, 1 = dU).
    // The helper warps store the chunk as dY_l and take its column sums.
    auto epi_step = [&](int s_idx, int l, int buf, bool add, int64_t g, bool has_consumer) {
      const int region = s_idx & 1;
      const uint32_t dst = buf ? ubuf : sbuf;
      const uint4 mq = *reinterpret_cast<const uint4*>(p.masks + (((int64_t)l * p.Bpad + g) * 8 + half * 4));
      const uint32_t mw[4] = {mq.x, mq.y, mq.z, mq.w};
      mbar_wait(acc_full(region), (full_par >> region) & 1u, 5);
      full_par ^= 1u << region;
      tc_fence_after();
      const uint32_t t_acc = t_lane + (uint32_t)(region * 256 + 32 * half);
      const uint32_t fpar = (nwr[buf] & 1u) ^ 1u;      // previous contents of this buffer have been stored and summed
      ++nwr[buf];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t r[32];
        tmem_ld32(t_acc + 64 * j, r);
        const uint32_t rowaddr = dst + j * kChunkBytes + row * 128;
        mbar_wait(chunk_free(buf, j), fpar, 10);
        uint32_t old[16];
        if (add) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int slot = (half * 4 + i4) ^ (row & 7);
            ld_shared_v4(rowaddr + slot * 16, old[4 * i4], old[4 * i4 + 1], old[4 * i4 + 2], old[4 * i4 + 3]);
          }
        }
        tmem_ld_wait();
        const uint32_t m = mw[j];
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = ((int32_t)(m << (2 * i)) < 0) ? __uint_as_float(r[2 * i]) : 0.f;
          float a1 = ((int32_t)(m << (2 * i + 1)) < 0) ? __uint_as_float(r[2 * i + 1]) : 0.f;
          if (add) { a0 += bf16_lo(old[i]); a1 += bf16_hi(old[i]); }
          pk[i] = pack2(a0, a1);
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int slot = (half * 4 + i4) ^ (row & 7);
          st_shared_v4(rowaddr + slot * 16, pk[4 * i4], pk[4 * i4 + 1], pk[4 * i4 + 2], pk[4 * i4 + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (has_consumer) mbar_arrive(opnd_ready(j));       // a chunk nobody multiplies is not announced to the MMA warp
          mbar_arrive(written(j));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(region));
    };

    int t = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t g = (int64_t)tile * 128 + row;
      epi_step(t++, 2 * R, 0, false, g, true);
#pragma unroll
      for (int r = R - 1; r >= 0; --r) {
        epi_step(t++, 2 * r + 1, 1, false, g, true);
        epi_step(t++, 2 * r, 0, true, g, DIN || r > 0);
      }
      if (DIN) {
        const int region = t & 1;
        ++t;
        mbar_wait(acc_full(region), (full_par >> region) & 1u, 6);
        full_par ^= 1u << region;
        tc_fence_after();
        for (int pc = half; pc * 32 < p.din_N; pc += 2) {
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(region * 256 + pc * 32), r);
          tmem_ld_wait();
          if (g < p.B) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (pc * 32 + i < p.din_cols) p.dIn[g * p.din_cols + pc * 32 + i] = __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(region));
      }
    }
  } else {
    // ===================== helpers (4 warps): warp j owns chunk j of every operand tile =====================
    // As soon as the epilogue has written chunk j: one TMA store of the [128 rows x 64 cols] chunk as dY_l, and the
    // bias gradient (column sums of the bf16 values the weight-gradient GEMM will read) from shared memory, two
    // columns per lane; then the chunk is handed back.  This keeps stores and reductions off the epilogue's path.
    const int j = warp - (2 + kEpiWarps);
    float cs[n_hidden][2];
#pragma unroll
    for (int l = 0; l < n_hidden; ++l) { cs[l][0] = 0.f; cs[l][1] = 0.f; }
    uint32_t wr_cnt = 0, reg_uses[2] = {0, 0};
    auto help = [&](int l, int buf, int tile, float (&c2)[2]) {
      const uint32_t src = (buf ? ubuf : sbuf) + j * kChunkBytes;
      mbar_wait(written(j), wr_cnt & 1u, 11);
      ++wr_cnt;
      if (lane == 0 && !(p.debug & 32)) {
        tma_store_2d(&map_dy, src, 64 * j, (int)((int64_t)l * p.Bpad + (int64_t)tile * 128));
        tma_store_commit();
      }
      if (!(p.debug & 16)) {
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 128; r += 2) {
          const uint32_t w0 = ld_shared_u32(src + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
          const uint32_t w1 = ld_shared_u32(src + (r + 1) * 128 + ((((lane >> 2) ^ ((r + 1) & 7)) << 4) | ((lane & 3) << 2)));
          a0 += bf16_lo(w0); a1 += bf16_hi(w0);
          b0 += bf16_lo(w1); b1 += bf16_hi(w1);
        }
        c2[0] += a0 + b0; c2[1] += a1 + b1;
      }
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      if (lane == 0) mbar_arrive(chunk_free(buf, j));
    };
    int t = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      ++reg_uses[t & 1]; ++t;                         // s0
      help(2 * R, 0, tile, cs[2 * R]);
#pragma unroll
      for (int r = R - 1; r >= 0; --r) {
        ++reg_uses[t & 1]; ++t;                       // dU step
        help(2 * r + 1, 1, tile, cs[2 * r + 1]);
        const int region = t & 1;                     // s step: its accumulator being full means every MMA that read
        const uint32_t n_before = reg_uses[region];   // the dU buffer has retired
        ++reg_uses[region]; ++t;
        if (r == 0) {
          mbar_wait(acc_full(region), n_before & 1u, 12);
          if (lane == 0) mbar_arrive(ubuf_free);      // the producer may load the next tile's dHead rows into it
        }
        help(2 * r, 0, tile, cs[2 * r]);
      }
      if (DIN) { ++reg_uses[t & 1]; ++t; }
    }
    if (lane == 0) tma_store_wait_all();
    // bias gradients: lane c of helper j holds columns 64 j + 2 c, + 1 summed over every tile of this CTA
#pragma unroll
    for (int l = 0; l < n_hidden; ++l)
      if (p.db[l] && !(p.debug & 16)) {
        atomicAdd(p.db[l] + 64 * j + 2 * lane, cs[l][0]);
        atomicAdd(p.db[l] + 64 * j + 2 * lane + 1, cs[l][1]);
      }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- backward chain, LayerNorm nets
// Input-gradient chain of a ResidualMLP with hk.LayerNorm(-1, False, False) after every Linear (networks.py:117-129, the
// bsds config) + head, the VJP of net_fwd_kernel<SAVE, ., LN>.  With s = dL/dh (residual stream), X_l / rstd_l the
// normalised pre-activation and 1/sigma of LayerNorm l, M_l the relu bits, and
//     LNbwd(v; X, rstd) = rstd . (v - mean(v) - X . mean(v . X))            (means over the 256 columns of a row)
// a 128-row tile runs
//     s    = M_2R . (dHead @ W_head^T);                 g_2R   = LNbwd(s; X_2R)
//     for r = R-1 .. 0:   t = M_{2r+1} . (g_{2r+2} @ W_{2r+2}^T);   g_{2r+1} = LNbwd(t; X_{2r+1})
//                         s += M_{2r} . (g_{2r+1} @ W_{2r+1}^T);    g_{2r}   = LNbwd(s; X_{2r})
//     dIn  = g_0 @ W_0^T                                 (decoder only)
// g_l is dY_l, the gradient with respect to the output of Linear l: the A operand of the next contraction (bf16, built
// 64 columns at a time in shared memory like every other chain), streamed to HBM by TMA for the weight-gradient GEMMs, and
// column-summed for the bias gradients.  s lives in TMEM columns [256, 512) (fp32), the contraction results in [0, 256).
// X_l tiles arrive by TMA (four 64-column chunks, refilled as soon as the second pass of the previous epilogue has read
// them).  Every epilogue makes two passes over its row: sums first (the accumulator is handed back after this pass: odd
// Linears stash the masked values, rounded to bf16, in the operand buffer they are about to overwrite), then the output.
// The head contraction streams dHead through the operand buffer as a 4-chunk ring, so any head width works (the TriL
// heads of bsds have 2144 columns).
constexpr int kLnOffX = kOpndBytes;
constexpr int kLnOffW = 2 * kOpndBytes;
constexpr int kLnOffExch = kLnOffW + kWStages * kWStageBytes;
constexpr int kLnOffBar = kLnOffExch + 128 * 2 * 2 * 4;
constexpr int kLnSmemBytes = kLnOffBar + 512;
constexpr int kLnHelpers = 2;                               // helper warps, two operand chunks each
constexpr int kLnThreads = kThreads + kLnHelpers * 32;      // 384: leaves 168 registers per thread (no spills in the epilogue)
static_assert(kLnSmemBytes <= 232448, "shared-memory plan exceeds 227 KB");

struct BwdLnArgs {
  int64_t B; int num_tiles; int R;
  int k16_h;                  // K = 16 steps of the head contraction (ceil(head_N / 16))
  int din_N, din_cols;        // dIn: MMA N (multiple of 16, 0 = none) and valid columns
  float* dIn;                 // [B, din_cols] fp32
  const uint32_t* masks; const float* rstd; int64_t Bpad;
  float* db[kMaxLayers];      // bias-gradient destinations (atomicAdd), Linear 0..2R
  int debug;                  // PMVAE_FUSED_DEBUG bits (profiling only): 1024 no column sums, 2048 no dY stores, 4096 no xhat loads
};

__device__ long long g_ln_trace[1024];      // PMVAE_FUSED_DEBUG bit 8192: clock64 stamps of block 0 (profiling only)

// R (residual blocks) is a run-time argument and every per-Linear loop is a real loop (one copy of the epilogue and of
// the helper code): with the chain unrolled over its 2R + 1 Linears the kernel was 400 KB of SASS, far beyond the
// instruction caches, and half of all warp stalls were instruction fetches (profiles/r02_net_bwd_ln_ncu.md).
template <bool DIN>
__global__ void __launch_bounds__(kLnThreads, 1)
net_bwd_ln_kernel(const __grid_constant__ CUtensorMap map_dh, const __grid_constant__ CUtensorMap map_wh,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w0,
                  const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, BwdLnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("pmvae net_bwd_ln_kernel: shared memory base is not 1024-byte aligned\n"); __trap(); }
  const uint32_t abuf = sbase, xbuf = sbase + kLnOffX, wring = sbase + kLnOffW, bar = sbase + kLnOffBar;
  float* exch = reinterpret_cast<float*>(smem_raw + kLnOffExch);           // [128 rows][2 halves][2]
  auto w_full = [&](int s) { return bar + 8u * s; };
  auto w_empty = [&](int s) { return bar + 8u * (3 + s); };
  auto opnd_ready = [&](int c) { return bar + 8u * (6 + c); };
  auto hd_full = [&](int c) { return bar + 8u * (10 + c); };
  auto hd_empty = [&](int c) { return bar + 8u * (14 + c); };
  auto x_full = [&](int c) { return bar + 8u * (18 + c); };
  auto x_free = [&](int c) { return bar + 8u * (22 + c); };
  auto written = [&](int c) { return bar + 8u * (26 + c); };
  auto chunk_free = [&](int c) { return bar + 8u * (30 + c); };
  const uint32_t acc_full = bar + 8u * 34, acc_empty = bar + 8u * 35, bufa_free = bar + 8u * 36;
  const uint32_t tmem_slot = bar + 8u * 37;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + kLnOffBar + 8 * 37);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int R = p.R;
  const int n_hidden = 2 * R + 1;
  const int nkb_h = (p.k16_h + 3) >> 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dh); tma_prefetch_desc(&map_wh); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    if (DIN) tma_prefetch_desc(&map_w0);
    for (int s = 0; s < 3; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int c = 0; c < 4; ++c) {
      mbar_init(opnd_ready(c), kEpiWarps); mbar_init(hd_full(c), 1); mbar_init(hd_empty(c), 1);
      mbar_init(x_full(c), 1); mbar_init(x_free(c), kEpiWarps);
      mbar_init(written(c), kEpiWarps); mbar_init(chunk_free(c), 1);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, kEpiWarps);
    mbar_init(bufa_free, 2 * kLnHelpers + (DIN ? 1 : 0));
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== producer: dHead chunks, weight K-blocks, xhat tiles =====================
    int stage = 0; uint32_t ph = 0;
    uint32_t hd_cnt = 0, x_cnt = 0, t_cnt = 0;
    auto load_w = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes) {
      mbar_wait(w_empty(stage), ph ^ 1u, 1);
      __syncwarp();
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full(stage), bytes);
        tma_load_2d(wring + stage * kWStageBytes, m, w_full(stage), c0, c1);
      }
      __syncwarp();
      if (++stage == kWStages) { stage = 0; ph ^= 1u; }
    };
    auto load_x = [&](int l, int tile) {
      for (int j = 0; j < 4; ++j) {
        mbar_wait(x_free(j), (x_cnt & 1u) ^ 1u, 20);
        __syncwarp();
        if (elect_one()) {
          if (p.debug & 4096) {
            mbar_arrive(x_full(j));
          } else {
            mbar_arrive_expect_tx(x_full(j), (uint32_t)kChunkBytes);
            tma_load_2d(xbuf + j * kChunkBytes, &map_x, x_full(j), 64 * j, (int)((int64_t)l * p.Bpad + (int64_t)tile * 128));
          }
        }
        __syncwarp();
      }
      ++x_cnt;
    };
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t_cnt) {
      // the operand buffer is the dHead ring now: the previous tile's last gradient tile has been stored (and multiplied)
      mbar_wait(bufa_free, (t_cnt & 1u) ^ 1u, 21);
      load_x(2 * R, tile);
      for (int kb = 0; kb < nkb_h; ++kb, ++hd_cnt) {
        const int slot = (int)(hd_cnt & 3u);
        mbar_wait(hd_empty(slot), ((hd_cnt >> 2) & 1u) ^ 1u, 22);
        __syncwarp();
        if (elect_one()) {
          mbar_arrive_expect_tx(hd_full(slot), (uint32_t)kChunkBytes);
          tma_load_2d(abuf + slot * kChunkBytes, &map_dh, hd_full(slot), kb * 64, tile * 128);
        }
        __syncwarp();
        load_w(&map_wh, kb * 64, 0, kWStageBytes);
      }
#pragma unroll 1
      for (int l = 2 * R; l >= 1; --l) {
        for (int kb = 0; kb < 4; ++kb) load_w(&map_w, kb * 64, (l - 1) * 256, kWStageBytes);
        load_x(l - 1, tile);
      }
      if (DIN)
        for (int kb = 0; kb < 4; ++kb) load_w(&map_w0, kb * 64, 0, (uint32_t)p.din_N * 128u);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    int stage = 0; uint32_t ph = 0;
    uint32_t ready_par = 0, acc_uses = 0, hd_cnt = 0;
    constexpr uint32_t kDescHi = 0x40004040u;      // SBO 1024 | version 1 | SWIZZLE_128B (see smem_desc)
    auto desc_lo = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
    auto mk = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
    const uint32_t d_tmem = tmem_base;
    auto acquire_acc = [&]() {
      mbar_wait(acc_empty, (acc_uses & 1u) ^ 1u, 2);
      ++acc_uses;
      tc_fence_after();
    };
    // one 64-wide K-block: A chunk at `a_addr`, B = the ring stage; `ks` K = 16 slices
    auto kblock = [&](uint32_t a_addr, int ks, uint32_t idesc, bool first, uint32_t extra_commit, bool last) {
      const uint32_t b_w = w_full(stage);
      mbar_wait(b_w, ph, 4);
      tc_fence_after();
      const uint32_t a_lo = desc_lo(a_addr), b_lo = desc_lo(wring + stage * kWStageBytes);
      __syncwarp();
      if (elect_one()) {
        for (int k = 0; k < ks; ++k) umma_f16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, (!first || k > 0) ? 1u : 0u);
        umma_commit(w_empty(stage));
        if (extra_commit) umma_commit(extra_commit);
        if (last) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == kWStages) { stage = 0; ph ^= 1u; }
    };
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      // head: s_raw = dHead @ W_head^T, dHead streamed through the 4-chunk ring
      acquire_acc();
      for (int kb = 0; kb < nkb_h; ++kb, ++hd_cnt) {
        const int slot = (int)(hd_cnt & 3u);
        mbar_wait(hd_full(slot), (hd_cnt >> 2) & 1u, 8);
        kblock(abuf + slot * kChunkBytes, min(4, p.k16_h - 4 * kb), instr_desc(128, 256, 0, 0), kb == 0, hd_empty(slot), kb == nkb_h - 1);
      }
#pragma unroll 1
      for (int st = 0; st < 2 * R; ++st) {
        acquire_acc();
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(opnd_ready(kb), (ready_par >> kb) & 1u, 3);
          ready_par ^= 1u << kb;
          kblock(abuf + kb * kChunkBytes, 4, instr_desc(128, 256, 0, 0), kb == 0, 0u, kb == 3);
        }
      }
      if (DIN) {
        acquire_acc();
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(opnd_ready(kb), (ready_par >> kb) & 1u, 3);
          ready_par ^= 1u << kb;
          kblock(abuf + kb * kChunkBytes, 4, instr_desc(128, p.din_N, 0, 0), kb == 0, kb == 3 ? bufa_free : 0u, kb == 3);
        }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (8 warps): thread = row, two threads (half 0 / 1) per row =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_acc = t_lane + (uint32_t)(32 * half);
    const uint32_t t_s = t_lane + (uint32_t)(256 + 32 * half);
    uint32_t full_par = 0, e_cnt = 0;
    float* mine = exch + (row * 2 + half) * 2;
    const float* other = exch + (row * 2 + (half ^ 1)) * 2;

    // One epilogue: v = M_l . acc (+ s), row sums, g_l = LNbwd(v; X_l) -> operand buffer chunk by chunk.
    auto epi = [&](int l, bool first, int64_t g, bool has_consumer) {
      const bool even = (l & 1) == 0;
      const uint4 mq = *reinterpret_cast<const uint4*>(p.masks + (((int64_t)l * p.Bpad + g) * 8 + half * 4));
      const float rstd = p.rstd[(int64_t)l * p.Bpad + g];
      const uint32_t cf_par = (e_cnt & 1u) ^ 1u;      // chunk_free: the previous gradient tile in this chunk has been stored
      const uint32_t x_par = e_cnt & 1u;
      ++e_cnt;
      const bool tr = (p.debug & 8192) && blockIdx.x == 0 && ew == 0 && lane == 0 && e_cnt <= 24;
      if (tr) g_ln_trace[8 * e_cnt + 0] = clock64();
      mbar_wait(acc_full, full_par, 5);
      full_par ^= 1u;
      tc_fence_after();
      if (tr) g_ln_trace[8 * e_cnt + 1] = clock64();
      float S1 = 0.f, S2 = 0.f;
      // ---- pass 1: masked values, row sums; s -> TMEM (even) or bf16 stash in the operand buffer (odd)
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t r[32], so[32];
        tmem_ld32(t_acc + 64 * j, r);
        if (even && !first) tmem_ld32(t_s + 64 * j, so);
        if (tr && e_cnt == 6) g_ln_trace[512 + 8 * j + 0] = clock64();
        mbar_wait(x_full(j), x_par, 23);
        if (tr && e_cnt == 6) g_ln_trace[512 + 8 * j + 1] = clock64();
        uint32_t xq[16];
        const uint32_t xrow = xbuf + j * kChunkBytes + row * 128;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int slot = (half * 4 + i4) ^ (row & 7);
          ld_shared_v4(xrow + slot * 16, xq[4 * i4], xq[4 * i4 + 1], xq[4 * i4 + 2], xq[4 * i4 + 3]);
        }
        tmem_ld_wait();
        if (tr && e_cnt == 6) g_ln_trace[512 + 8 * j + 2] = clock64();
        const uint32_t m = j == 0 ? mq.x : (j == 1 ? mq.y : (j == 2 ? mq.z : mq.w));
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = ((int32_t)(m << (2 * i)) < 0) ? __uint_as_float(r[2 * i]) : 0.f;
          float a1 = ((int32_t)(m << (2 * i + 1)) < 0) ? __uint_as_float(r[2 * i + 1]) : 0.f;
          if (even && !first) { a0 += __uint_as_float(so[2 * i]); a1 += __uint_as_float(so[2 * i + 1]); }
          S1 += a0 + a1;
          S2 = fmaf(a0, bf16_lo(xq[i]), S2);
          S2 = fmaf(a1, bf16_hi(xq[i]), S2);
          r[2 * i] = __float_as_uint(a0); r[2 * i + 1] = __float_as_uint(a1);
        }
        if (tr && e_cnt == 6) g_ln_trace[512 + 8 * j + 3] = clock64();
        if (even) {
          tmem_st32(t_s + 64 * j, r);
        } else {
          mbar_wait(chunk_free(j), cf_par, 10);
          if (tr && e_cnt == 6) g_ln_trace[512 + 8 * j + 4] = clock64();
          const uint32_t rowaddr = abuf + j * kChunkBytes + row * 128;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int slot = (half * 4 + i4) ^ (row & 7);
            st_shared_v4(rowaddr + slot * 16, pack2(__uint_as_float(r[8 * i4]), __uint_as_float(r[8 * i4 + 1])),
                         pack2(__uint_as_float(r[8 * i4 + 2]), __uint_as_float(r[8 * i4 + 3])),
                         pack2(__uint_as_float(r[8 * i4 + 4]), __uint_as_float(r[8 * i4 + 5])),
                         pack2(__uint_as_float(r[8 * i4 + 6]), __uint_as_float(r[8 * i4 + 7])));
          }
        }
      }
      if (even) tmem_st_wait();
      // the accumulator has been read for the last time
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      if (tr) g_ln_trace[8 * e_cnt + 2] = clock64();
      mine[0] = S1; mine[1] = S2;
      named_bar_sync(1 + q, 64);
      if (tr) g_ln_trace[8 * e_cnt + 3] = clock64();
      const float c1 = (S1 + other[0]) * (1.0f / 256.0f), c2 = (S2 + other[1]) * (1.0f / 256.0f);
      // ---- pass 2: g = rstd (v - c1 - x c2)
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t r[32];
        uint32_t vq[16];
        const uint32_t rowaddr = abuf + j * kChunkBytes + row * 128;
        if (even) {
          tmem_ld32(t_s + 64 * j, r);
        } else {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int slot = (half * 4 + i4) ^ (row & 7);
            ld_shared_v4(rowaddr + slot * 16, vq[4 * i4], vq[4 * i4 + 1], vq[4 * i4 + 2], vq[4 * i4 + 3]);
          }
        }
        uint32_t xq[16];
        const uint32_t xrow = xbuf + j * kChunkBytes + row * 128;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int slot = (half * 4 + i4) ^ (row & 7);
          ld_shared_v4(xrow + slot * 16, xq[4 * i4], xq[4 * i4 + 1], xq[4 * i4 + 2], xq[4 * i4 + 3]);
        }
        if (even) {
          tmem_ld_wait();
          mbar_wait(chunk_free(j), cf_par, 10);
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float v0 = even ? __uint_as_float(r[2 * i]) : bf16_lo(vq[i]);
          const float v1 = even ? __uint_as_float(r[2 * i + 1]) : bf16_hi(vq[i]);
          const float g0 = rstd * (v0 - c1 - bf16_lo(xq[i]) * c2);
          const float g1 = rstd * (v1 - c1 - bf16_hi(xq[i]) * c2);
          pk[i] = pack2(g0, g1);
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int slot = (half * 4 + i4) ^ (row & 7);
          st_shared_v4(rowaddr + slot * 16, pk[4 * i4], pk[4 * i4 + 1], pk[4 * i4 + 2], pk[4 * i4 + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (has_consumer) mbar_arrive(opnd_ready(j));
          mbar_arrive(written(j));
          mbar_arrive(x_free(j));
        }
      }
      if (tr) g_ln_trace[8 * e_cnt + 4] = clock64();
      named_bar_sync(1 + q, 64);        // the exchange slots are rewritten by the next epilogue
    };

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t g = (int64_t)tile * 128 + row;
#pragma unroll 1
      for (int l = 2 * R; l >= 0; --l) epi(l, l == 2 * R, g, DIN || l > 0);
      if (DIN) {
        mbar_wait(acc_full, full_par, 6);
        full_par ^= 1u;
        tc_fence_after();
        for (int pc = half; pc * 32 < p.din_N; pc += 2) {
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(pc * 32), r);
          tmem_ld_wait();
          if (g < p.B) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (pc * 32 + i < p.din_cols) p.dIn[g * p.din_cols + pc * 32 + i] = __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    }
  } else {
    // ===================== helpers (2 warps): warp h owns chunks 2h, 2h + 1 of every gradient tile =====================
    // one TMA store of the chunk as dY_l, and the bias gradient (column sums of the bf16 values the weight-gradient
    // GEMM will read), two columns per lane, added to the gradient arena per tile (one atomicAdd per column)
    const int h = warp - (2 + kEpiWarps);
    uint32_t wr_cnt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int l = 2 * R; l >= 0; --l) {
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          const int j = 2 * h + u;
          const uint32_t src = abuf + j * kChunkBytes;
          mbar_wait(written(j), wr_cnt & 1u, 11);
          if (lane == 0) {
            if (!(p.debug & 2048)) tma_store_2d(&map_dy, src, 64 * j, (int)((int64_t)l * p.Bpad + (int64_t)tile * 128));
            tma_store_commit();
          }
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
          for (int r = 0; r < ((p.debug & 1024) ? 0 : 128); r += 2) {
            const uint32_t w0 = ld_shared_u32(src + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
            const uint32_t w1 = ld_shared_u32(src + (r + 1) * 128 + ((((lane >> 2) ^ ((r + 1) & 7)) << 4) | ((lane & 3) << 2)));
            a0 += bf16_lo(w0); a1 += bf16_hi(w0);
            b0 += bf16_lo(w1); b1 += bf16_hi(w1);
          }
          if (p.db[l]) {
            atomicAdd(p.db[l] + 64 * j + 2 * lane, a0 + b0);
            atomicAdd(p.db[l] + 64 * j + 2 * lane + 1, a1 + b1);
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          if (lane == 0) mbar_arrive(chunk_free(j));
        }
        ++wr_cnt;
      }
      if (lane == 0) { mbar_arrive(bufa_free); mbar_arrive(bufa_free); }   // both chunks of the tile's last gradient tile stored
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight images
struct PackSlab { uint64_t dst, w; int rows, cols, kind, src_rows, src_cols, D, tile0; };
struct PackTable { int n; PackSlab s[2 * kMaxLayers + 4]; };

// kind 0: dst[n][k] = W[k][n];  kind 1/2: first-Linear image, dst[n][kk] = W[srow(kk)][n] for the hi/lo
// operand layout (in_kind 0/1);  kind 3: dst[r][c] = W[r][c].  Everything outside the source is zero.
__global__ void __launch_bounds__(256) pack_fused_kernel(const float* __restrict__ params, bf16* __restrict__ img,
                                                         PackTable tb) {
  __shared__ float tile[32][33];
  int si = 0;
  while (si + 1 < tb.n && (int)blockIdx.x >= tb.s[si + 1].tile0) ++si;
  const PackSlab sl = tb.s[si];
  const int t = blockIdx.x - sl.tile0;
  const int tiles_c = (sl.cols + 31) / 32;
  const int tr = t / tiles_c, tcn = t % tiles_c;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* W = params + sl.w;
  bf16* dst = img + sl.dst;
  if (sl.kind == 3) {
    for (int i = ty; i < 32; i += 8) {
      const int r = tr * 32 + i, c = tcn * 32 + tx;
      if (r < sl.rows && c < sl.cols)
        dst[(uint64_t)r * sl.cols + c] = __float2bfloat16((r < sl.src_rows && c < sl.src_cols) ? W[(uint64_t)r * sl.src_cols + c] : 0.f);
    }
    return;
  }
  for (int i = ty; i < 32; i += 8) {
    const int kk = tcn * 32 + i, n = tr * 32 + tx;
    int sr = -1;
    if (sl.kind == 0) sr = kk < sl.src_rows ? kk : -1;
    else if (sl.kind == 4) {                                      // wide masked input: K-block 0 = W[:D], K-block 1 = W[D:2D]
      if (kk < sl.D) sr = kk;
      else if (kk >= 64 && kk < 64 + sl.D) sr = sl.D + (kk - 64);
    }
    else if (kk < sl.D) sr = kk;
    else if (kk < 2 * sl.D) sr = kk - sl.D;
    else if (sl.kind == 2 && kk < 3 * sl.D) sr = kk - sl.D;      // rows D..2D-1 of W multiply b
    tile[i][tx] = (sr >= 0 && n < sl.src_cols) ? W[(uint64_t)sr * sl.src_cols + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = tr * 32 + i, kk = tcn * 32 + tx;
    if (n < sl.rows && kk < sl.cols) dst[(uint64_t)n * sl.cols + kk] = __float2bfloat16(tile[tx][i]);
  }
}

// expanded first-Linear operand: plain input [hi | lo] when both fit one 64-wide K-block, else [hi] only (fan-in up
// to 64: operand rounding like every other Linear); masked input [hi(x*b) | lo(x*b) | b]
static bool wide_masked(const Net& n, int in_kind) { return in_kind == 1 && 3 * (n.in_dim / 2) > 64 && n.in_dim / 2 <= 64; }
static int first_kext(const Net& n, int in_kind, int* in_lo) {
  if (wide_masked(n, in_kind)) { *in_lo = 0; return n.in_dim / 2; }   // per half: [x*b] then [b], one K-block each
  if (in_kind == 1) { *in_lo = 1; return 3 * (n.in_dim / 2); }
  *in_lo = 2 * n.in_dim <= 64 ? 1 : 0;
  return *in_lo ? 2 * n.in_dim : n.in_dim;
}
bool forward_supported(const Net& n, int H, int in_kind) {
  if (H != 256) return false;
  int lo;
  const int kext = first_kext(n, in_kind, &lo);
  // one K-block for the first Linear; a row's fp32 inputs (x | mask) fit its staging slot; LayerNorm nets keep the
  // exchange buffer in the tail of the bias table
  return kext <= 64 && (n.in_dim <= kInPitch - 1 || wide_masked(n, in_kind)) && n.R < (n.ln ? 6 : kMaxBlocks);
}
bool supported(const Net& n, int H, int in_kind) { return !n.ln && forward_supported(n, H, in_kind); }
// training-mode forward (saved activations) of LayerNorm nets; their backward chain is fused_ln.cu
bool ln_train_supported(const Net& n, int H, int in_kind) { return n.ln && forward_supported(n, H, in_kind) && n.R >= 1; }

NetImages plan_images(const Net& n, const Leaf& head, int in_kind, bf16* base) {
  NetImages im{};
  im.R = n.R; im.in_kind = in_kind;
  im.D_in = (in_kind == 1) ? n.in_dim / 2 : n.in_dim;
  im.in_wide = wide_masked(n, in_kind) ? 1 : 0;
  const int kext = first_kext(n, in_kind, &im.in_lo);
  im.k16_0 = (kext + 15) / 16;
  im.head_N = head.cols;
  const int np16 = (head.cols + 15) / 16 * 16;
  im.head_tiles = (np16 + 255) / 256;
  im.head_NT = ((np16 + im.head_tiles - 1) / im.head_tiles + 15) / 16 * 16;
  // Several head tiles: the epilogue stores 32-column pieces, so a tile width that is not a multiple of 32 lets the last
  // piece of tile t spill 16 never-written accumulator columns over the first 16 columns of tile t + 1, and the two TMA
  // stores (issued by different warps) are not ordered: whenever the TMA unit was busy enough for the earlier store to
  // land last (the training-mode LayerNorm forward: ~1 tile in 500), those columns came out stale.
  if (im.head_tiles > 1) im.head_NT = (im.head_NT + 31) / 32 * 32;
  im.head_Kp = (head.cols + 63) / 64 * 64;
  uint64_t off = 0;
  auto take = [&](uint64_t elems) { bf16* p = base ? base + off : nullptr; off += align_up(elems, 512); return p; };
  im.stack_t = take((uint64_t)(1 + 2 * n.R) * 256 * 256);
  im.head_t = take((uint64_t)im.head_tiles * im.head_NT * 256);
  im.stack_n = take((uint64_t)(2 * n.R > 0 ? 2 * n.R : 1) * 256 * 256);
  im.head_n = take((uint64_t)256 * im.head_Kp);
  im.din_N = (n.in_dim + 15) / 16 * 16;
  im.w0_n = (in_kind == 0 && im.din_N <= 256) ? take((uint64_t)im.din_N * 256) : nullptr;
  if (!base && in_kind == 0 && im.din_N <= 256) im.w0_n = nullptr;
  im.has_w0_n = (in_kind == 0 && im.din_N <= 256);
  im.elems = off;
  return im;
}

int pack_images(const float* params, const Net& n, const Leaf& head, const NetImages& im, cudaStream_t s) {
  PackTable tb{};
  int tiles = 0;
  const bf16* base = im.stack_t;
  auto add = [&](const bf16* dst, uint64_t w, int rows, int cols, int kind, int src_rows, int src_cols) {
    PackSlab& sl = tb.s[tb.n++];
    sl.dst = (uint64_t)(dst - base); sl.w = w; sl.rows = rows; sl.cols = cols; sl.kind = kind;
    sl.src_rows = src_rows; sl.src_cols = src_cols; sl.D = im.D_in; sl.tile0 = tiles;
    tiles += ((rows + 31) / 32) * ((cols + 31) / 32);
  };
  add(im.stack_t, n.lin[0].w, 256, 256, im.in_wide ? 4 : (im.in_kind == 1 ? 2 : (im.in_lo ? 1 : 0)), n.lin[0].rows, 256);
  for (int l = 1; l <= 2 * n.R; ++l) add(im.stack_t + (uint64_t)l * 65536, n.lin[l].w, 256, 256, 0, 256, 256);
  add(im.head_t, head.w, im.head_tiles * im.head_NT, 256, 0, 256, head.cols);
  for (int l = 1; l <= 2 * n.R; ++l) add(im.stack_n + (uint64_t)(l - 1) * 65536, n.lin[l].w, 256, 256, 3, 256, 256);
  add(im.head_n, head.w, 256, im.head_Kp, 3, 256, head.cols);
  if (im.has_w0_n) add(im.w0_n, n.lin[0].w, im.din_N, 256, 3, n.lin[0].rows, 256);
  pack_fused_kernel<<<tiles, 256, 0, s>>>(params, const_cast<bf16*>(base), tb);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// PMVAE_FUSED_MAXGRID (tests): cap on the persistent grids, so that small batches exercise several tiles per CTA
static int grid_cap(int grid) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PMVAE_FUSED_MAXGRID"); v = e ? atoi(e) : 0; }
  return (v > 0 && grid > v) ? v : grid;
}

static bool cta2_enabled() {
  static int v = -1;
  // default on: a CTA pair halves the weight bytes written into, and the B-operand bytes read from, each SM's shared
  // memory, whose port is close to saturated in the forward chain (MMA operand reads + weight TMA + epilogue stores +
  // activation stores): 96 -> 92 us for the encoder chain, 1.31 -> 1.24 ms per train step (PMVAE_FUSED_CTA2=0 disables)
  if (v < 0) { const char* e = getenv("PMVAE_FUSED_CTA2"); v = e ? atoi(e) : 1; }
  return v != 0;
}

template <bool SAVE, bool CTA2, bool LN>
static int launch_fwd_t(int grid, const CUtensorMap& mw, const CUtensorMap& mh, const CUtensorMap& ms,
                        const CUtensorMap& mo, const CUtensorMap& mx, const FwdArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    PMVAE_CUDA(cudaFuncSetAttribute(net_fwd_kernel<SAVE, CTA2, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((SAVE && LN) ? kFwdThreadsLN : kFwdThreads);
  cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTA2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  PMVAE_CUDA(cudaLaunchKernelEx(&cfg, net_fwd_kernel<SAVE, CTA2, LN>, mw, mh, ms, mo, mx, a));
  return 0;
}
static int launch_fwd(bool save, bool cta2, bool ln, int grid, const CUtensorMap& mw, const CUtensorMap& mh,
                      const CUtensorMap& ms, const CUtensorMap& mo, const CUtensorMap& mx, const FwdArgs& a, cudaStream_t s) {
  if (ln && save) return cta2 ? launch_fwd_t<true, true, true>(grid, mw, mh, ms, mo, mx, a, s) : launch_fwd_t<true, false, true>(grid, mw, mh, ms, mo, mx, a, s);
  if (ln) return cta2 ? launch_fwd_t<false, true, true>(grid, mw, mh, ms, mo, mx, a, s) : launch_fwd_t<false, false, true>(grid, mw, mh, ms, mo, mx, a, s);
  if (save) return cta2 ? launch_fwd_t<true, true, false>(grid, mw, mh, ms, mo, mx, a, s) : launch_fwd_t<true, false, false>(grid, mw, mh, ms, mo, mx, a, s);
  return cta2 ? launch_fwd_t<false, true, false>(grid, mw, mh, ms, mo, mx, a, s) : launch_fwd_t<false, false, false>(grid, mw, mh, ms, mo, mx, a, s);
}

int net_forward(const float* params, const Net& n, const Leaf& head, const NetImages& im, const float* in,
                const float* msk, int64_t B, bf16* saved, uint32_t* masks, int64_t Bpad, float* out, int64_t ld_out,
                cudaStream_t s, bf16* xhat, float* rstd) {
  if (B <= 0) return 0;
  PMVAE_CHECK(forward_supported(n, 256, im.in_kind), "net not covered by the fused kernels");
  PMVAE_CHECK(saved == nullptr || !n.ln || (xhat != nullptr && rstd != nullptr),
              "LayerNorm nets in training mode also save xhat and 1/sigma");
  PMVAE_CHECK((im.in_kind == 1) == (msk != nullptr), "mask pointer does not match the first-layer layout");
  PMVAE_CHECK(B < (1ll << 30), "too many rows");
  FwdArgs a{};
  a.in = in; a.msk = msk; a.D_in = im.D_in; a.in_kind = im.in_kind; a.in_lo = im.in_lo; a.k16_0 = im.k16_0; a.R = n.R;
  a.B = B; a.num_tiles = (int)ceil_div(B, 128);
  for (int l = 0; l <= 2 * n.R; ++l) a.bias[l] = params + n.lin[l].b;
  a.head_bias = params + head.b; a.head_N = im.head_N; a.head_NT = im.head_NT; a.head_tiles = im.head_tiles;
  a.out = out; a.ld_out = ld_out;
  a.Bpad = Bpad;
  a.masks = nullptr;
  a.rstd = nullptr;
  a.in_wide = im.in_wide;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("PMVAE_FUSED_DEBUG"); dbg = e ? atoi(e) : 0; } a.debug = dbg; }
  CUtensorMap mw, mh, ms, mo, mx;
  const bool cta2 = cta2_enabled();
  PMVAE_TRY(make_map_2d(&mw, im.stack_t, 2, (uint64_t)(1 + 2 * n.R) * 256, 256, 256, 64, cta2 ? 128 : 256));
  PMVAE_TRY(make_map_2d(&mh, im.head_t, 2, (uint64_t)im.head_tiles * im.head_NT, 256, 256, 64,
                        (uint32_t)(cta2 ? im.head_NT / 2 : im.head_NT)));
  PMVAE_CHECK(im.k16_0 <= 4 && ld_out >= im.head_N, "first-Linear operand must fit one K-block");
  if (saved) {
    PMVAE_CHECK(Bpad % 128 == 0 && Bpad >= B, "saved activations need a 128-row padded slab pitch");
    PMVAE_CHECK((int64_t)(2 * n.R + 1) * Bpad < (1ll << 31), "saved activation stack too large");
    PMVAE_TRY(make_map_2d(&ms, saved, 2, (uint64_t)(2 * n.R + 1) * Bpad, 256, 256, 64, 128));
    a.masks = masks;
    if (n.ln) {
      PMVAE_TRY(make_map_2d(&mx, xhat, 2, (uint64_t)(2 * n.R + 1) * Bpad, 256, 256, 64, 128));
      a.rstd = rstd;
    } else {
      mx = mw;
    }
  } else {
    ms = mw; mx = mw;
  }
  a.head_tma = ((ld_out * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) ? 1 : 0;
  if (a.head_tma) PMVAE_TRY(make_map_2d(&mo, out, 4, (uint64_t)B, (uint64_t)im.head_N, (uint64_t)ld_out, 32, 32));
  else mo = mw;
  int grid = grid_cap(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
  if (cta2) {
    const int pairs = (a.num_tiles + 1) / 2;
    grid = 2 * grid_cap(pairs < num_sms() / 2 ? pairs : num_sms() / 2);
    PMVAE_CHECK(saved == nullptr || Bpad % 256 == 0, "CTA pairs need a 256-row padded slab pitch");
  }
  static long long* trace_buf = nullptr;
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("PMVAE_FUSED_TRACE"); trace_on = e ? atoi(e) : 0; }
  if (trace_on) {
    if (!trace_buf) cudaMalloc(&trace_buf, 2 * 2048 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 2 * 2048 * sizeof(long long), s);
    a.trace = trace_buf;
  }
  PMVAE_TRY(launch_fwd(saved != nullptr, cta2, n.ln != 0, grid, mw, mh, ms, mo, mx, a, s));
  PMVAE_LAUNCH_CHECK();
  if (trace_on) {
    static long long host[2 * 2048];
    cudaStreamSynchronize(s);
    cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
    if (trace_on == 1) {   // print once
      trace_on = 2;
      const long long t0 = host[2048 + 2];
      for (int role = 0; role < 2; ++role) {
        const long long* t = host + role * 2048;
        printf("trace role %d (%lld stamps)\n", role, t[0]);
        for (int i = 0; i < t[0] && i < 400; ++i) printf("  %4lld @ %8lld\n", t[1 + 2 * i], t[2 + 2 * i] - t0);
      }
    }
  }
  return 0;
}

template <int R, bool DIN>
static int launch_bwd(const CUtensorMap& mdh, const CUtensorMap& mwh, const CUtensorMap& mw, const CUtensorMap& mw0,
                      const CUtensorMap& mdy, const BwdArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    PMVAE_CUDA(cudaFuncSetAttribute(net_bwd_kernel<R, DIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    attr_set = true;
  }
  const int grid = grid_cap(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
  net_bwd_kernel<R, DIN><<<grid, kBwdThreads, kBwdSmemBytes, s>>>(mdh, mwh, mw, mw0, mdy, a);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

bool backward_supported(const Net& n, int H, int in_kind) { return supported(n, H, in_kind) && n.R >= 1 && n.R <= 4; }

int net_backward(const Net& n, const Leaf& head, const NetImages& im, const bf16* dHead, int64_t ld_dhead, int64_t B,
                 const uint32_t* masks, int64_t Bpad, bf16* dY, float* grads, float* dIn, cudaStream_t s) {
  if (B <= 0) return 0;
  PMVAE_CHECK(backward_supported(n, 256, im.in_kind), "net not covered by the fused backward kernel");
  PMVAE_CHECK(dIn == nullptr || im.has_w0_n, "no first-layer image for the input gradient");
  PMVAE_CHECK(Bpad % 128 == 0 && Bpad >= B && (int64_t)(2 * n.R + 1) * Bpad < (1ll << 31), "bad slab pitch");
  BwdArgs a{};
  a.B = B; a.num_tiles = (int)ceil_div(B, 128);
  a.k16_h = (im.head_N + 15) / 16;
  a.din_N = dIn ? im.din_N : 0; a.din_cols = n.in_dim; a.dIn = dIn;
  a.masks = masks; a.Bpad = Bpad;
  for (int l = 0; l <= 2 * n.R; ++l) a.db[l] = grads + n.lin[l].b;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("PMVAE_FUSED_DEBUG"); dbg = e ? atoi(e) : 0; } a.debug = dbg; }
  CUtensorMap mdh, mwh, mw, mw0, mdy;
  PMVAE_TRY(make_map_2d(&mdh, dHead, 2, (uint64_t)B, (uint64_t)im.head_N, (uint64_t)ld_dhead, 64, 128));
  PMVAE_TRY(make_map_2d(&mwh, im.head_n, 2, 256, (uint64_t)im.head_Kp, (uint64_t)im.head_Kp, 64, 256));
  PMVAE_TRY(make_map_2d(&mw, im.stack_n, 2, (uint64_t)(2 * n.R) * 256, 256, 256, 64, 256));
  PMVAE_TRY(make_map_2d(&mdy, dY, 2, (uint64_t)(2 * n.R + 1) * Bpad, 256, 256, 64, 128));
  if (dIn) PMVAE_TRY(make_map_2d(&mw0, im.w0_n, 2, (uint64_t)im.din_N, 256, 256, 64, (uint32_t)im.din_N));
  else mw0 = mw;
#define PMVAE_BWD_CASE(RR)                                                             \
  case RR: return dIn ? launch_bwd<RR, true>(mdh, mwh, mw, mw0, mdy, a, s) : launch_bwd<RR, false>(mdh, mwh, mw, mw0, mdy, a, s)
  switch (n.R) {
    PMVAE_BWD_CASE(1);
    PMVAE_BWD_CASE(2);
    PMVAE_BWD_CASE(3);
    PMVAE_BWD_CASE(4);
  }
#undef PMVAE_BWD_CASE
  PMVAE_CHECK(false, "unsupported number of residual blocks");
  return 1;
}

template <bool DIN>
static int launch_bwd_ln(const CUtensorMap& mdh, const CUtensorMap& mwh, const CUtensorMap& mw, const CUtensorMap& mw0,
                         const CUtensorMap& mdy, const CUtensorMap& mx, const BwdLnArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    PMVAE_CUDA(cudaFuncSetAttribute(net_bwd_ln_kernel<DIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmemBytes));
    attr_set = true;
  }
  const int grid = grid_cap(a.num_tiles < num_sms() ? a.num_tiles : num_sms());
  net_bwd_ln_kernel<DIN><<<grid, kLnThreads, kLnSmemBytes, s>>>(mdh, mwh, mw, mw0, mdy, mx, a);
  PMVAE_LAUNCH_CHECK();
  if (a.debug & 8192) {
    static int printed = 0;
    if (printed < 2) {
      ++printed;
      static long long host[1024];
      cudaStreamSynchronize(s);
      cudaMemcpyFromSymbol(host, g_ln_trace, sizeof(host));
      const long long t0 = host[8 + 0];
      printf("pass 1 of epilogue 5 (odd), per chunk: start, x_full, loads, math, chunk_free\n");
      for (int j = 0; j < 4; ++j)
        printf("  j %d: %8lld %8lld %8lld %8lld %8lld\n", j, host[512 + 8 * j] - t0, host[512 + 8 * j + 1] - t0,
               host[512 + 8 * j + 2] - t0, host[512 + 8 * j + 3] - t0, host[512 + 8 * j + 4] - t0);
      printf("net_bwd_ln trace (block 0, epilogue warp 0): epilogue e: start, acc_full, pass1 end, exchange, pass2 end (cycles)\n");
      for (int e = 1; e <= 24; ++e)
        printf("  e %2d: %8lld %8lld %8lld %8lld %8lld\n", e, host[8 * e] - t0, host[8 * e + 1] - t0, host[8 * e + 2] - t0,
               host[8 * e + 3] - t0, host[8 * e + 4] - t0);
      fflush(stdout);
    }
  }
  return 0;
}

bool backward_ln_supported(const Net& n, int H, int in_kind) { return ln_train_supported(n, H, in_kind); }

int net_backward_ln(const Net& n, const Leaf& head, const NetImages& im, const bf16* dHead, int64_t ld_dhead, int64_t B,
                    const uint32_t* masks, const bf16* xhat, const float* rstd, int64_t Bpad, bf16* dY, float* grads,
                    float* dIn, cudaStream_t s) {
  if (B <= 0) return 0;
  PMVAE_CHECK(backward_ln_supported(n, 256, im.in_kind), "net not covered by the fused LayerNorm backward kernel");
  PMVAE_CHECK(dIn == nullptr || im.has_w0_n, "no first-layer image for the input gradient");
  PMVAE_CHECK(Bpad % 128 == 0 && Bpad >= B && (int64_t)(2 * n.R + 1) * Bpad < (1ll << 31), "bad slab pitch");
  PMVAE_CHECK((ld_dhead * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(dHead) & 15u) == 0, "dHead must be 16-byte aligned with a 16-byte pitch");
  BwdLnArgs a{};
  a.B = B; a.num_tiles = (int)ceil_div(B, 128);
  a.k16_h = (im.head_N + 15) / 16;
  a.din_N = dIn ? im.din_N : 0; a.din_cols = n.in_dim; a.dIn = dIn;
  a.masks = masks; a.rstd = rstd; a.Bpad = Bpad;
  for (int l = 0; l <= 2 * n.R; ++l) a.db[l] = grads + n.lin[l].b;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("PMVAE_FUSED_DEBUG"); dbg = e ? atoi(e) : 0; } a.debug = dbg; }
  CUtensorMap mdh, mwh, mw, mw0, mdy, mx;
  PMVAE_TRY(make_map_2d(&mdh, dHead, 2, (uint64_t)B, (uint64_t)im.head_N, (uint64_t)ld_dhead, 64, 128));
  PMVAE_TRY(make_map_2d(&mwh, im.head_n, 2, 256, (uint64_t)im.head_Kp, (uint64_t)im.head_Kp, 64, 256));
  PMVAE_TRY(make_map_2d(&mw, im.stack_n, 2, (uint64_t)(2 * n.R) * 256, 256, 256, 64, 256));
  PMVAE_TRY(make_map_2d(&mdy, dY, 2, (uint64_t)(2 * n.R + 1) * Bpad, 256, 256, 64, 128));
  PMVAE_TRY(make_map_2d(&mx, xhat, 2, (uint64_t)(2 * n.R + 1) * Bpad, 256, 256, 64, 128));
  if (dIn) PMVAE_TRY(make_map_2d(&mw0, im.w0_n, 2, (uint64_t)im.din_N, 256, 256, 64, (uint32_t)im.din_N));
  else mw0 = mw;
  a.R = n.R;
  return dIn ? launch_bwd_ln<true>(mdh, mwh, mw, mw0, mdy, mx, a, s) : launch_bwd_ln<false>(mdh, mwh, mw, mw0, mdy, mx, a, s);
}

}  // namespace fused
}  // namespace pmvae
