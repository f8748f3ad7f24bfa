// Fused ResidualMLP kernels for sm_100a: one persistent CTA per SM walks 128-row tiles through every
// Linear of a net (networks.py:111-135) and its distribution-head Linear (distributions.py:44,104)
// without the activations leaving the SM.
//
//   warp 0      : TMA producer, streams the bf16 weight K-blocks ([256 n] x [64 k], SWIZZLE_128B) of the
//                 current Linear through a 3-stage mbarrier ring (the weights come from L2)
//   warp 1      : MMA issuer (one lane): tcgen05.mma kind::f16, M = 128, N = 256 (head: N <= 256), K = 16
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
#include "fused_mlp.cuh"

#include "kernels.h"
#include "tc_ptx.cuh"

namespace pmvae {
namespace fused {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kChunkBytes = 16384;             // [128 rows] x [64 k] bf16
constexpr int kOpndBytes = 4 * kChunkBytes;
constexpr int kWStageBytes = 32768;            // [256 n] x [64 k] bf16
constexpr int kWStages = 3;
constexpr int kStagBytes = kEpiWarps * 4096;   // per-warp [32 rows] x [32 cols] fp32 head staging
constexpr int kMaxLayers = 2 * kMaxBlocks + 1;
constexpr int kOffW = kOpndBytes;
constexpr int kOffStag = kOffW + kWStages * kWStageBytes;
constexpr int kOffBias = kOffStag + kStagBytes;
constexpr int kOffBar = kOffBias + kMaxLayers * 1024;
constexpr int kSmemBytes = kOffBar + 256 + 1024;

struct FwdArgs {
  const float* in; const float* msk;
  int D_in, in_kind, k16_0, R;
  int64_t B; int num_tiles;
  const float* bias[kMaxLayers];
  const float* head_bias; int head_N, head_NT, head_tiles;
  int64_t Bpad; uint32_t* masks;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
net_fwd_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h,
               const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_o, FwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t opnd = sbase, wring = sbase + kOffW, stag = sbase + kOffStag, bar = sbase + kOffBar;
  float* bias_tbl = reinterpret_cast<float*>(sgen + kOffBias);
  auto w_full = [&](int s) { return bar + 8u * s; };
  auto w_empty = [&](int s) { return bar + 8u * (kWStages + s); };
  auto opnd_ready = [&](int c) { return bar + 8u * (2 * kWStages + c); };
  auto acc_full = [&](int r) { return bar + 8u * (2 * kWStages + 4 + r); };
  auto acc_empty = [&](int r) { return bar + 8u * (2 * kWStages + 6 + r); };
  const uint32_t tmem_slot = bar + 8u * (2 * kWStages + 8);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + kOffBar + 8 * (2 * kWStages + 8));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R;
  const int n_hidden = 2 * R + 1;                 // Linears that feed the operand buffer
  const int nkb0 = (p.k16_0 + 3) >> 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_h); tma_prefetch_desc(&map_o);
    if (SAVE) tma_prefetch_desc(&map_s);
    for (int s = 0; s < kWStages; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int c = 0; c < 4; ++c) mbar_init(opnd_ready(c), kEpiWarps);
    for (int r = 0; r < 2; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // bias table: row l = what the epilogue of Linear l adds (running sum for the residual stream)
  for (int c = threadIdx.x; c < 256; c += kThreads) {
    float run = p.bias[0][c];
    bias_tbl[c] = run;
    for (int r = 0; r < R; ++r) {
      bias_tbl[(2 * r + 1) * 256 + c] = p.bias[2 * r + 1][c];
      run += p.bias[2 * r + 2][c];
      bias_tbl[(2 * r + 2) * 256 + c] = run;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes) {
        mbar_wait(w_empty(stage), ph ^ 1u, 1);
        mbar_arrive_expect_tx(w_full(stage), bytes);
        tma_load_2d(wring + stage * kWStageBytes, m, w_full(stage), c0, c1);
        if (++stage == kWStages) { stage = 0; ph ^= 1u; }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb0; ++kb) load(&map_w, kb * 64, 0, kWStageBytes);
        for (int l = 1; l < n_hidden; ++l)
          for (int kb = 0; kb < 4; ++kb) load(&map_w, kb * 64, l * 256, kWStageBytes);
        for (int t = 0; t < p.head_tiles; ++t)
          for (int kb = 0; kb < 4; ++kb) load(&map_h, kb * 64, t * p.head_NT, (uint32_t)p.head_NT * 128u);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      uint32_t ready_par = 0, use_cnt0 = 0, use_cnt1 = 0;
      auto step = [&](int s_idx, int nk16, int N, bool accum, bool wait_opnd) {
        const int region = (s_idx & 1) ? 0 : 1;
        uint32_t& uc = region ? use_cnt1 : use_cnt0;
        mbar_wait(acc_empty(region), (uc & 1u) ^ 1u, 2);
        ++uc;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(region * 256);
        const uint32_t idesc = instr_desc(128, N, 0, 0);
        const int nkb = (nk16 + 3) >> 2;
        for (int kb = 0; kb < nkb; ++kb) {
          if (wait_opnd) {
            mbar_wait(opnd_ready(kb), (ready_par >> kb) & 1u, 3);
            ready_par ^= 1u << kb;
          }
          mbar_wait(w_full(stage), ph, 4);
          tc_fence_after();
          const uint32_t sa = opnd + kb * kChunkBytes;
          const uint32_t sb = wring + stage * kWStageBytes;
          const int ks = min(4, nk16 - 4 * kb);
          for (int k = 0; k < ks; ++k)
            umma_f16(d_tmem, smem_desc(sa + k * 32, 16, 1024), smem_desc(sb + k * 32, 16, 1024), idesc,
                     (accum || kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(w_empty(stage));
          if (++stage == kWStages) { stage = 0; ph ^= 1u; }
        }
        umma_commit(acc_full(region));
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        step(0, p.k16_0, 256, false, true);
        for (int l = 1; l < n_hidden; ++l) step(l, 16, 256, (l & 1) == 0, true);
        for (int t = 0; t < p.head_tiles; ++t) step(n_hidden + t, 16, p.head_NT, false, t == 0);
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = ew >> 2;          // which 32 of the 64 columns of a chunk
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t my_stag = stag + ew * 4096;
    uint32_t full_par = 0;
    const int D = p.D_in;

    auto drain_sync = [&]() {
      // the TMA stores of the previous operand tile must have finished reading shared memory
      if (half == 0 && lane == 0) tma_store_wait_read0();
      named_bar_sync(1 + q, 64);
    };

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t g = (int64_t)tile * 128 + row;
      const bool row_ok = g < p.B;
      // ---- first-layer operand: hi/lo bf16 split of the fp32 input, [x*b, b] built here (vae.py:132-133)
      if (SAVE) drain_sync();
      {
        const float* xin = p.in + g * D;
        const float* xm = p.msk ? p.msk + g * D : nullptr;
        auto ext = [&](int kk) -> float {
          if (!row_ok) return 0.f;
          if (kk < D) {
            float v = __ldg(xin + kk);
            if (xm) v *= __ldg(xm + kk);
            return v;
          }
          if (kk < 2 * D) {
            float v = __ldg(xin + kk - D);
            if (xm) v *= __ldg(xm + kk - D);
            return v - bf16_round(v);
          }
          if (xm && kk < 3 * D) return __ldg(xm + kk - 2 * D);
          return 0.f;
        };
        for (int c = 0; c < nkb0; ++c) {
          const int kk0 = 64 * c + 32 * half;
          const uint32_t rowaddr = opnd + c * kChunkBytes + row * 128;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = pack2(ext(kk0 + 8 * i4 + 2 * i), ext(kk0 + 8 * i4 + 2 * i + 1));
            const int slot = (half * 4 + i4) ^ (row & 7);
            st_shared_v4(rowaddr + slot * 16, w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(opnd_ready(c));
        }
      }
      // ---- hidden Linears: accumulator -> bf16 operand of the next Linear
      for (int l = 0; l < n_hidden; ++l) {
        const int region = (l & 1) ? 0 : 1;
        mbar_wait(acc_full(region), (full_par >> region) & 1u, 5);
        full_par ^= 1u << region;
        tc_fence_after();
        if (SAVE) drain_sync();
        uint32_t mw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(region * 256 + 64 * j + 32 * half), r);
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(bias_tbl + l * 256 + 64 * j + 32 * half);
          uint32_t pk[16];
          uint32_t bits = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = bp[i];
            const float v0 = __uint_as_float(r[4 * i]) + bv.x, v1 = __uint_as_float(r[4 * i + 1]) + bv.y;
            const float v2 = __uint_as_float(r[4 * i + 2]) + bv.z, v3 = __uint_as_float(r[4 * i + 3]) + bv.w;
            bits |= (v0 > 0.f ? 1u : 0u) << (4 * i) | (v1 > 0.f ? 1u : 0u) << (4 * i + 1) |
                    (v2 > 0.f ? 1u : 0u) << (4 * i + 2) | (v3 > 0.f ? 1u : 0u) << (4 * i + 3);
            pk[2 * i] = pack2(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
            pk[2 * i + 1] = pack2(fmaxf(v2, 0.f), fmaxf(v3, 0.f));
          }
          mw[j] = bits;
          const uint32_t rowaddr = opnd + j * kChunkBytes + row * 128;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int slot = (half * 4 + i4) ^ (row & 7);
            st_shared_v4(rowaddr + slot * 16, pk[4 * i4], pk[4 * i4 + 1], pk[4 * i4 + 2], pk[4 * i4 + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(opnd_ready(j));
          if (SAVE) {
            named_bar_sync(1 + q, 64);
            if (half == 0 && lane == 0) {
              tma_store_2d(&map_s, opnd + j * kChunkBytes + q * 4096, 64 * j,
                           (int)((int64_t)l * p.Bpad + (int64_t)tile * 128 + q * 32));
              tma_store_commit();
            }
          }
        }
        if (SAVE && p.masks)
          *reinterpret_cast<uint4*>(p.masks + (((int64_t)l * p.Bpad + g) * 8 + half * 4)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(region));
      }
      // ---- head Linear: accumulator + bias -> fp32 rows, staged per warp and stored by TMA
      for (int t = 0; t < p.head_tiles; ++t) {
        const int region = ((n_hidden + t) & 1) ? 0 : 1;
        mbar_wait(acc_full(region), (full_par >> region) & 1u, 6);
        full_par ^= 1u << region;
        tc_fence_after();
        for (int pc = half; pc * 32 < p.head_NT; pc += 2) {
          const int nb = t * p.head_NT + pc * 32;
          if (nb >= p.head_N) break;
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)(region * 256 + pc * 32), r);
          tmem_ld_wait();
          if (lane == 0) tma_store_wait_read0();      // the staging tile is free again
          __syncwarp();
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = nb + 4 * i4 + i;
              v[i] = __uint_as_float(r[4 * i4 + i]) + (n < p.head_N ? __ldg(p.head_bias + n) : 0.f);
            }
            const int slot = i4 ^ (lane & 7);
            st_shared_v4(my_stag + lane * 128 + slot * 16, __float_as_uint(v[0]), __float_as_uint(v[1]),
                         __float_as_uint(v[2]), __float_as_uint(v[3]));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_o, my_stag, nb, (int)((int64_t)tile * 128 + q * 32));
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(region));
      }
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

bool supported(const Net& n, int H, int in_kind) {
  if (n.ln || H != 256) return false;
  const int kext = (in_kind == 1) ? 3 * (n.in_dim / 2) : 2 * n.in_dim;
  return kext <= 256;
}

NetImages plan_images(const Net& n, const Leaf& head, int in_kind, bf16* base) {
  NetImages im{};
  im.R = n.R; im.in_kind = in_kind;
  im.D_in = (in_kind == 1) ? n.in_dim / 2 : n.in_dim;
  const int kext = (in_kind == 1) ? 3 * im.D_in : 2 * im.D_in;
  im.k16_0 = (kext + 15) / 16;
  im.head_N = head.cols;
  const int np16 = (head.cols + 15) / 16 * 16;
  im.head_tiles = (np16 + 255) / 256;
  im.head_NT = ((np16 + im.head_tiles - 1) / im.head_tiles + 15) / 16 * 16;
  im.head_Kp = (head.cols + 63) / 64 * 64;
  uint64_t off = 0;
  auto take = [&](uint64_t elems) { bf16* p = base ? base + off : nullptr; off += align_up(elems, 512); return p; };
  im.stack_t = take((uint64_t)(1 + 2 * n.R) * 256 * 256);
  im.head_t = take((uint64_t)im.head_tiles * im.head_NT * 256);
  im.stack_n = take((uint64_t)(2 * n.R > 0 ? 2 * n.R : 1) * 256 * 256);
  im.head_n = take((uint64_t)256 * im.head_Kp);
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
  add(im.stack_t, n.lin[0].w, 256, 256, im.in_kind == 1 ? 2 : 1, n.lin[0].rows, 256);
  for (int l = 1; l <= 2 * n.R; ++l) add(im.stack_t + (uint64_t)l * 65536, n.lin[l].w, 256, 256, 0, 256, 256);
  add(im.head_t, head.w, im.head_tiles * im.head_NT, 256, 0, 256, head.cols);
  for (int l = 1; l <= 2 * n.R; ++l) add(im.stack_n + (uint64_t)(l - 1) * 65536, n.lin[l].w, 256, 256, 3, 256, 256);
  add(im.head_n, head.w, 256, im.head_Kp, 3, 256, head.cols);
  pack_fused_kernel<<<tiles, 256, 0, s>>>(params, const_cast<bf16*>(base), tb);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int net_forward(const float* params, const Net& n, const Leaf& head, const NetImages& im, const float* in,
                const float* msk, int64_t B, bf16* saved, uint32_t* masks, int64_t Bpad, float* out, int64_t ld_out,
                cudaStream_t s) {
  if (B <= 0) return 0;
  PMVAE_CHECK(supported(n, 256, im.in_kind), "net not covered by the fused kernels");
  PMVAE_CHECK((im.in_kind == 1) == (msk != nullptr), "mask pointer does not match the first-layer layout");
  PMVAE_CHECK(B < (1ll << 30), "too many rows");
  FwdArgs a{};
  a.in = in; a.msk = msk; a.D_in = im.D_in; a.in_kind = im.in_kind; a.k16_0 = im.k16_0; a.R = n.R;
  a.B = B; a.num_tiles = (int)ceil_div(B, 128);
  for (int l = 0; l <= 2 * n.R; ++l) a.bias[l] = params + n.lin[l].b;
  a.head_bias = params + head.b; a.head_N = im.head_N; a.head_NT = im.head_NT; a.head_tiles = im.head_tiles;
  a.Bpad = Bpad;
  a.masks = nullptr;
  CUtensorMap mw, mh, ms, mo;
  PMVAE_TRY(make_map_2d(&mw, im.stack_t, 2, (uint64_t)(1 + 2 * n.R) * 256, 256, 256, 64, 256));
  PMVAE_TRY(make_map_2d(&mh, im.head_t, 2, (uint64_t)im.head_tiles * im.head_NT, 256, 256, 64, (uint32_t)im.head_NT));
  PMVAE_TRY(make_map_2d(&mo, out, 4, (uint64_t)B, (uint64_t)im.head_N, (uint64_t)ld_out, 32, 32));
  if (saved) {
    PMVAE_CHECK(Bpad % 128 == 0 && Bpad >= B, "saved activations need a 128-row padded slab pitch");
    PMVAE_CHECK((int64_t)(2 * n.R + 1) * Bpad < (1ll << 31), "saved activation stack too large");
    PMVAE_TRY(make_map_2d(&ms, saved, 2, (uint64_t)(2 * n.R + 1) * Bpad, 256, 256, 64, 32));
    a.masks = masks;
  } else {
    ms = mw;
  }
  const int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  static bool attr_set[2] = {false, false};
  if (saved) {
    if (!attr_set[1]) {
      PMVAE_CUDA(cudaFuncSetAttribute(net_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      attr_set[1] = true;
    }
    net_fwd_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(mw, mh, ms, mo, a);
  } else {
    if (!attr_set[0]) {
      PMVAE_CUDA(cudaFuncSetAttribute(net_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      attr_set[0] = true;
    }
    net_fwd_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(mw, mh, ms, mo, a);
  }
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace fused
}  // namespace pmvae
