// tcgen05 / TMEM / TMA GEMM for the dense contractions of the PM-VAE path (sm_100a).
//
//   kind 0 ("NT", both operands contraction-major): C[M,N] = A[M,K] . Bt[N,K]^T
//        hk.Linear forward  (A = activations, Bt = W^T image)      networks.py:116,122,127
//        input gradient     (A = dY,          Bt = W image)        dX = dY @ W^T
//   kind 1 ("TN", both operands MN-major):     C[M,N] = A[K,M]^T . B[K,N]
//        weight gradient    (A = saved activations [rows,M], B = dY [rows,N]), the
//        contraction runs over the batch rows and is split across CTAs.
//
// Structure (one CTA per SM, persistent over tiles):
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2D, SWIZZLE_128B, 3-stage mbarrier ring)
//   warp 1      : MMA issuer    (one lane: tcgen05.mma kind::f16, M=128, N<=256, K=16, fp32 accum in TMEM)
//   warps 2..9  : epilogue      (tcgen05.ld 32x32b.x32 -> smem transpose -> bias / mask / residual / relu
//                                -> coalesced global rows; optional fused column sums for bias gradients)
// Two 256-column TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_gemm.h"
#include "tc_ptx.cuh"

namespace pmvae {
namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int kMaxN = 256;
constexpr int kStages = 3;
constexpr int kAccStages = 2;
constexpr int kThreads = 320;          // 10 warps: TMA, MMA, 8 epilogue
constexpr int kEpiWarps = 8;
constexpr int kMaxGroup = 8;
constexpr int kRowBatch = 16;          // rows whose epilogue loads are in flight together
constexpr int kStagePitch = 68;        // floats per staged row: 64 columns + 4 pad (conflict-free both ways)
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytes = kMaxN * kBlockK * 2;        // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + kEpiWarps * 32 * kStagePitch * 4 + 1024 /*align slack*/;

// ---------------------------------------------------------------- kernel
// EPI selects the auxiliary input streams of the NT epilogue: 0 none, 1 fp32 residual, 3 bf16 mask then fp32 residual,
// 2 bf16 relu mask (+ optional bf16 residual).
template <int KIND, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kAccStages + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * kAccStages);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + 8 * (2 * kStages + 2 * kAccStages));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
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

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;
  const int n_tile = p.n_tile;
  const uint32_t stage_tx = (uint32_t)(kABytes + n_tile * kBlockK * 2);

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % p.num_m_tiles;
        const int rest = tile / p.num_m_tiles;
        const int nt = rest % p.num_n_tiles;
        const int sp = rest / p.num_n_tiles;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          __syncwarp();
          if (elect_one()) {
            mbar_arrive_expect_tx(full_bar(stage), stage_tx);
            if (KIND == 0) {
              // A box: [64 k] x [128 rows]; B box: [64 k] x [n_tile rows]
              tma_load_2d(sa, &map_a, full_bar(stage), kb * kBlockK, mt * kBlockM);
              tma_load_2d(sb, &map_b, full_bar(stage), kb * kBlockK, nt * n_tile);
            } else {
              // MN-major: boxes of [64 m-or-n] x [64 k rows], one per 64 columns of the tile
              for (int j = 0; j < kBlockM / 64; ++j)
                tma_load_2d(sa + j * 8192, &map_a, full_bar(stage), mt * kBlockM + j * 64, kb * kBlockK);
              for (int j = 0; j < n_tile / 64; ++j)
                tma_load_2d(sb + j * 8192, &map_b, full_bar(stage), nt * n_tile + j * 64, kb * kBlockK);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues; see elect_one) =====================
    {
      const uint32_t idesc = instr_desc(kBlockM, n_tile, KIND, KIND);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int rest = tile / p.num_m_tiles;
        const int sp = rest / p.num_n_tiles;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kMaxN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          __syncwarp();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              uint64_t da, db;
              if (KIND == 0) {
                // K-major SW128: 8-row groups 1024 B apart; a K=16 slice is 32 B inside the swizzle row
                da = smem_desc(sa + k * 32, 16, 1024);
                db = smem_desc(sb + k * 32, 16, 1024);
              } else {
                // MN-major SW128: 64-element column groups 8192 B apart (LBO), 8-k groups 1024 B apart (SBO)
                da = smem_desc(sa + k * 2048, 8192, 1024);
                db = smem_desc(sb + k * 2048, 8192, 1024);
              }
              umma_f16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));          // frees the smem slot when these MMAs retire
            if (kb == kb1 - 1) umma_commit(tfull_bar(acc));   // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) {                          // empty K range: still hand the (stale) accumulator over
          if (elect_one()) umma_commit(tfull_bar(acc));
          __syncwarp();
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    // Warp w may touch TMEM lanes 32*(w%4)..+31 (accumulator rows); the two warps of a quadrant
    // take alternate 64-column chunks.  A chunk goes TMEM -> registers (thread = row) -> padded
    // shared-memory stage -> registers (lane = column pair), so every global access below is a
    // contiguous row segment (256 B fp32 / 128 B bf16 per warp instruction).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stage = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + 256) + (warp - 2) * (32 * kStagePitch);
    float2 csum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % p.num_m_tiles;
      const int rest = tile / p.num_m_tiles;
      const int nt = rest % p.num_n_tiles;
      const int sp = rest / p.num_n_tiles;
      mbar_wait(tfull_bar(acc), acc_phase, 4);
      tc_fence_after();
      const int64_t row0 = (int64_t)mt * kBlockM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kMaxN);
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = half * 64 + ci * 128;
        if (c0 >= n_tile || (p.debug & 2)) break;
        // ---- TMEM -> stage (thread = row)
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + (uint32_t)c0, r0);            // both halves in flight before the wait
          tmem_ld32(t_row + (uint32_t)(c0 + 32), r1);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(stage + lane * kStagePitch);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]), __uint_as_float(r0[4 * j + 2]),
                                 __uint_as_float(r0[4 * j + 3]));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[8 + j] = make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]), __uint_as_float(r1[4 * j + 2]),
                                     __uint_as_float(r1[4 * j + 3]));
        }
        __syncwarp();
        // ---- stage -> global (lane = column pair)
        const int n = nt * n_tile + c0 + 2 * lane;
        const bool col_ok = n < p.N && !(p.debug & 1);
        if (KIND == 1) {
          float* base = p.out_f32 + (p.atomic ? 0 : (int64_t)sp * p.M * p.ld_out_f32);
          for (int rr = 0; rr < 32; ++rr) {
            const int64_t row = row0 + rr;
            if (row >= p.M) break;
            const float2 v = *reinterpret_cast<const float2*>(stage + rr * kStagePitch + 2 * lane);
            if (col_ok) {
              float2* dst = reinterpret_cast<float2*>(base + row * p.ld_out_f32 + n);
              if (p.atomic) atomicAdd(dst, v);
              else *dst = v;
            }
          }
        } else {
          float2 bias2 = make_float2(0.f, 0.f);
          if (p.bias && col_ok) bias2 = __ldg(reinterpret_cast<const float2*>(p.bias + n));
          // Rows go in batches of kRowBatch: all global loads of a batch are issued before any is
          // consumed, so a warp keeps kRowBatch row segments in flight instead of one.
#pragma unroll 1
          for (int rb = 0; rb < 32; rb += kRowBatch) {
            uint32_t mk[kRowBatch], rbv[kRowBatch];
            float2 rf[kRowBatch];
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) {
              const int64_t row = row0 + rb + i;
              const bool ok = col_ok && row < p.M;
              if (EPI == 2 || EPI == 3)
                mk[i] = ok ? __ldg(reinterpret_cast<const uint32_t*>(p.mask_bf16 + row * p.ld_mask + n)) : 0u;
              if (EPI == 2)
                rbv[i] = (ok && p.resid_bf16) ? *reinterpret_cast<const uint32_t*>(p.resid_bf16 + row * p.ld_resid_bf16 + n) : 0u;
              if (EPI == 1 || EPI == 3)
                rf[i] = ok ? *reinterpret_cast<const float2*>(p.resid_f32 + row * p.ld_resid_f32 + n) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) {
              const int64_t row = row0 + rb + i;
              if (!(col_ok && row < p.M)) continue;
              float2 v = *reinterpret_cast<const float2*>(stage + (rb + i) * kStagePitch + 2 * lane);
              v.x += bias2.x; v.y += bias2.y;
              if (EPI == 2 || EPI == 3) {
                if (!(bf16_lo(mk[i]) > 0.f)) v.x = 0.f;
                if (!(bf16_hi(mk[i]) > 0.f)) v.y = 0.f;
              }
              if (EPI == 2) { v.x += bf16_lo(rbv[i]); v.y += bf16_hi(rbv[i]); }
              if (EPI == 1 || EPI == 3) { v.x += rf[i].x; v.y += rf[i].y; }   // (EPI 3: mask, then the fp32 residual)
              if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + row * p.ld_out_f32 + n) = v;
              if (p.out_bf16) {
                if (p.relu_out) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                const uint32_t o = pack_bf16(v.x, v.y);
                *reinterpret_cast<uint32_t*>(p.out_bf16 + row * p.ld_out_bf16 + n) = o;
                if (p.colsum_out) { csum[ci].x += bf16_lo(o); csum[ci].y += bf16_hi(o); }   // sums what dW will read
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
    if (KIND == 0 && p.colsum_out) {
      // bias gradient: column sums of the bf16 output over every tile this CTA produced
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int n = half * 64 + ci * 128 + 2 * lane;
        if (n < n_tile && n < p.N) { atomicAdd(p.colsum_out + n, csum[ci].x); atomicAdd(p.colsum_out + n + 1, csum[ci].y); }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- grouped weight gradients
// Up to kMaxGroup independent TN problems C_g[M_g, N_g] += A_g[rows, M_g]^T . B_g[rows, N_g] (all the Linears of one
// net) in ONE persistent launch: the tile list is the concatenation of every problem's (m-tile, split) pairs, so a
// net's weight gradients fill the machine once instead of once per Linear (at small batches the per-launch latency of
// eighteen 5 us GEMMs is most of the step).  Same pipeline as tc_gemm_kernel<1, 0>; fp32 atomics into C.
struct TnProblem {
  CUtensorMap ma, mb;
  float* out; int64_t ld_out;
  int M, N, n_tile, num_m_tiles, num_k_blocks, split_k, kb_per_split, tile0;
};
struct TnGroup { int n; int total_tiles; TnProblem p[kMaxGroup]; };

__global__ void __launch_bounds__(kThreads, 1) tn_grouped_kernel(const __grid_constant__ TnGroup g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kAccStages + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * kAccStages);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * kStageBytes + 8 * (2 * kStages + 2 * kAccStages));
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.n; ++i) { tma_prefetch_desc(&g.p[i].ma); tma_prefetch_desc(&g.p[i].mb); }
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // tile -> (problem, m-tile, split)
  auto locate = [&](int tile, int& pi, int& mt, int& sp) {
    pi = 0;
    while (pi + 1 < g.n && tile >= g.p[pi + 1].tile0) ++pi;
    const int t = tile - g.p[pi].tile0;
    mt = t % g.p[pi].num_m_tiles;
    sp = t / g.p[pi].num_m_tiles;
  };

  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int pi, mt, sp;
        locate(tile, pi, mt, sp);
        const TnProblem& P = g.p[pi];
        const int kb0 = sp * P.kb_per_split, kb1 = min(P.num_k_blocks, kb0 + P.kb_per_split);
        const uint32_t tx = (uint32_t)(kABytes + P.n_tile * kBlockK * 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
          __syncwarp();
          if (elect_one()) {
            mbar_arrive_expect_tx(full_bar(stage), tx);
            for (int j = 0; j < kBlockM / 64; ++j) tma_load_2d(sa + j * 8192, &P.ma, full_bar(stage), mt * kBlockM + j * 64, kb * kBlockK);
            for (int j = 0; j < P.n_tile / 64; ++j) tma_load_2d(sb + j * 8192, &P.mb, full_bar(stage), j * 64, kb * kBlockK);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int pi, mt, sp;
        locate(tile, pi, mt, sp);
        const TnProblem& P = g.p[pi];
        const int kb0 = sp * P.kb_per_split, kb1 = min(P.num_k_blocks, kb0 + P.kb_per_split);
        const uint32_t idesc = instr_desc(kBlockM, P.n_tile, 1, 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kMaxN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
          __syncwarp();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(d_tmem, smem_desc(sa + k * 2048, 8192, 1024), smem_desc(sb + k * 2048, 8192, 1024), idesc,
                       (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(empty_bar(stage));
            if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) {                          // empty K range: still hand the (stale) accumulator over
          if (elect_one()) umma_commit(tfull_bar(acc));
          __syncwarp();
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stage = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + 256) + (warp - 2) * (32 * kStagePitch);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      int pi, mt, sp;
      locate(tile, pi, mt, sp);
      const TnProblem& P = g.p[pi];
      const bool empty_split = sp * P.kb_per_split >= P.num_k_blocks;
      mbar_wait(tfull_bar(acc), acc_phase, 4);
      tc_fence_after();
      const int64_t row0 = (int64_t)mt * kBlockM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kMaxN);
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = half * 64 + ci * 128;
        if (c0 >= P.n_tile || empty_split) break;
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + (uint32_t)c0, r0);
          tmem_ld32(t_row + (uint32_t)(c0 + 32), r1);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(stage + lane * kStagePitch);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]), __uint_as_float(r0[4 * j + 2]),
                                 __uint_as_float(r0[4 * j + 3]));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[8 + j] = make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]), __uint_as_float(r1[4 * j + 2]),
                                     __uint_as_float(r1[4 * j + 3]));
        }
        __syncwarp();
        const int n = c0 + 2 * lane;
        if (n < P.N) {
          for (int rr = 0; rr < 32; ++rr) {
            const int64_t row = row0 + rr;
            if (row >= P.M) break;
            const float2 v = *reinterpret_cast<const float2*>(stage + rr * kStagePitch + 2 * lane);
            atomicAdd(reinterpret_cast<float2*>(P.out + row * P.ld_out + n), v);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_map_2d(CUtensorMap* m, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  PMVAE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  PMVAE_CHECK(elem_bytes == 2 || elem_bytes == 4, "tensor maps are built for bf16 or fp32 elements");
  PMVAE_CHECK((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (ld * elem_bytes) % 16 == 0,
              "TMA operand must be 16-byte aligned");
  PMVAE_CHECK(box_cols * elem_bytes <= 128 && box_rows <= 256, "TMA box too large for SWIZZLE_128B");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PMVAE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return 0;
}
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                    uint32_t box_rows) {
  return make_map_2d(m, base, 2, rows, cols, ld, box_cols, box_rows);
}

static int debug_bits() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PMVAE_TC_DEBUG"); v = e ? atoi(e) : 0; }
  return v;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int KIND, int EPI>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const TcGemmArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    PMVAE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<KIND, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int tiles = a.num_m_tiles * a.num_n_tiles * a.split_k;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  tc_gemm_kernel<KIND, EPI><<<grid, kThreads, kSmemBytes, s>>>(ma, mb, a);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

static int pick_n_tile(int N, int granule) {
  int nt = (N + granule - 1) / granule * granule;
  return nt > kMaxN ? kMaxN : nt;
}

// C[M,N] = A[M,K] . Bt[N,K]^T  (+ epilogue)
int gemm_nt(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* Bt, int64_t ldb, int64_t M, int N, int K,
            TcGemmArgs ep, cudaStream_t s) {
  PMVAE_CHECK(M >= 0 && N > 0 && K > 0, "bad gemm shape");
  if (M == 0) return 0;
  PMVAE_CHECK(N % 8 == 0, "tensor path needs N % 8 == 0");
  PMVAE_CHECK(M < (1ll << 31), "M too large");
  ep.M = (int)M; ep.N = N; ep.K = K;
  ep.n_tile = pick_n_tile(N, 32);
  ep.num_m_tiles = (int)ceil_div(M, kBlockM);
  ep.num_n_tiles = (int)ceil_div(N, ep.n_tile);
  ep.num_k_blocks = (int)ceil_div(K, kBlockK);
  ep.split_k = 1; ep.kb_per_split = ep.num_k_blocks; ep.atomic = 0; ep.debug = debug_bits();
  PMVAE_CHECK(ep.colsum_out == nullptr || (ep.num_n_tiles == 1 && ep.out_bf16 != nullptr),
              "fused column sums need a single N tile and a bf16 output");
  CUtensorMap ma, mb;
  PMVAE_TRY(make_map(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBlockK, kBlockM));
  PMVAE_TRY(make_map(&mb, Bt, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, kBlockK, (uint32_t)ep.n_tile));
  PMVAE_CHECK(!(ep.resid_f32 && ep.resid_bf16), "unsupported epilogue combination");
  PMVAE_CHECK(!(ep.resid_bf16 && !ep.mask_bf16), "bf16 residual needs a mask");
  if (ep.resid_f32 && ep.mask_bf16) return launch<0, 3>(ma, mb, ep, s);
  if (ep.resid_f32) return launch<0, 1>(ma, mb, ep, s);
  if (ep.mask_bf16) return launch<0, 2>(ma, mb, ep, s);
  return launch<0, 0>(ma, mb, ep, s);
}

// C[M,N] (fp32) = A[rows,M]^T . B[rows,N], contraction over rows split across CTAs.
// atomic = 1: atomicAdd into out (must be zeroed by the caller);
// atomic = 0: `out` is a [split, M, ld] partial buffer; *split_out returns the split count.
int gemm_tn(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B, int64_t ldb, int M, int N, int64_t rows,
            float* out, int64_t ld_out, int atomic, int max_split, int* split_out, cudaStream_t s) {
  PMVAE_CHECK(M > 0 && N > 0 && rows >= 0, "bad gemm shape");
  PMVAE_CHECK(N % 4 == 0, "tensor path needs N % 4 == 0");
  TcGemmArgs ep{};
  ep.M = M; ep.N = N; ep.K = (int)rows;
  ep.n_tile = pick_n_tile(N, 64);
  ep.num_m_tiles = (int)ceil_div(M, kBlockM);
  ep.num_n_tiles = (int)ceil_div(N, ep.n_tile);
  ep.num_k_blocks = (int)ceil_div(rows, kBlockK);
  int split = num_sms() / (ep.num_m_tiles * ep.num_n_tiles);
  if (split < 1) split = 1;
  if (split > ep.num_k_blocks) split = ep.num_k_blocks > 0 ? ep.num_k_blocks : 1;
  if (max_split > 0 && split > max_split) split = max_split;
  ep.kb_per_split = (int)ceil_div(ep.num_k_blocks, split);
  split = (int)ceil_div(ep.num_k_blocks, ep.kb_per_split);
  if (split < 1) split = 1;
  ep.split_k = split;
  ep.atomic = atomic;
  ep.out_f32 = out; ep.ld_out_f32 = ld_out;
  if (split_out) *split_out = split;
  if (rows == 0) return 0;
  CUtensorMap ma, mb;
  PMVAE_TRY(make_map(&ma, A, (uint64_t)rows, (uint64_t)M, (uint64_t)lda, 64, kBlockK));
  PMVAE_TRY(make_map(&mb, B, (uint64_t)rows, (uint64_t)N, (uint64_t)ldb, 64, kBlockK));
  return launch<1, 0>(ma, mb, ep, s);
}

int gemm_tn_grouped(const TnDesc* d, int count, cudaStream_t s) {
  PMVAE_CHECK(count >= 1 && count <= kMaxGroup, "bad group size");
  TnGroup g{};
  int out_tiles = 0;
  for (int i = 0; i < count; ++i) {
    PMVAE_CHECK(d[i].M > 0 && d[i].N > 0 && d[i].N % 4 == 0 && d[i].N <= kMaxN && d[i].rows >= 0, "bad grouped gemm shape");
    out_tiles += (int)ceil_div(d[i].M, kBlockM);
  }
  int split = num_sms() / out_tiles;
  if (split < 1) split = 1;
  int tiles = 0, n = 0;
  for (int i = 0; i < count; ++i) {
    if (d[i].rows == 0) continue;
    TnProblem& P = g.p[n];
    P.out = d[i].out; P.ld_out = d[i].ld_out; P.M = d[i].M; P.N = d[i].N;
    P.n_tile = pick_n_tile(d[i].N, 64);
    P.num_m_tiles = (int)ceil_div(d[i].M, kBlockM);
    P.num_k_blocks = (int)ceil_div(d[i].rows, kBlockK);
    int sp = split > P.num_k_blocks ? P.num_k_blocks : split;
    P.kb_per_split = (int)ceil_div(P.num_k_blocks, sp);
    P.split_k = (int)ceil_div(P.num_k_blocks, P.kb_per_split);
    P.tile0 = tiles;
    tiles += P.num_m_tiles * P.split_k;
    PMVAE_TRY(make_map(&P.ma, d[i].A, (uint64_t)d[i].rows, (uint64_t)d[i].M, (uint64_t)d[i].lda, 64, kBlockK));
    PMVAE_TRY(make_map(&P.mb, d[i].B, (uint64_t)d[i].rows, (uint64_t)d[i].N, (uint64_t)d[i].ldb, 64, kBlockK));
    ++n;
  }
  if (n == 0) return 0;
  g.n = n; g.total_tiles = tiles;
  static bool attr_set = false;
  if (!attr_set) {
    PMVAE_CUDA(cudaFuncSetAttribute(tn_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int grid = tiles < num_sms() ? tiles : num_sms();
  tn_grouped_kernel<<<grid, kThreads, kSmemBytes, s>>>(g);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace pmvae
