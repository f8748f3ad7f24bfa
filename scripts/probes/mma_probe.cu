// Micro-probe: how many cycles one tcgen05.mma (M=128 per CTA, K=16, bf16) occupies when issued back to back
// with nothing else going on in the SM: operands from shared memory (SS), A from tensor memory (TS), CTA pairs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../posterior_matching_b200/csrc -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace pmvae::tc;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// mode 0: SS one CTA; 1: TS one CTA (A in TMEM columns 256..); 2: SS CTA pair (M = 256, each CTA holds N/2 rows of B)
// traffic: extra warps hammering shared memory with v4 stores while the MMAs run (0 = none)
template <int MODE>
__global__ void __launch_bounds__(288, 1) probe(int N, int nkb, int traffic, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t abuf = sbase, bbuf = sbase + 65536, bar = sbase + 65536 + 3 * 32768 + 32768, slot = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pseudo-random bf16 patterns (finite, small) everywhere the MMAs read
  for (uint32_t i = threadIdx.x; i < (65536 + 3 * 32768) / 4; i += blockDim.x) {
    uint32_t h = i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 13;
    reinterpret_cast<uint32_t*>(sgen)[i] = (h & 0x007F007Fu) | 0x3C003C00u | ((h >> 3) & 0x80008000u);
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); }
  if (warp == 0) {
    if (MODE == 2) { tmem_alloc_cta2(slot, 512); tmem_relinquish_cta2(); }
    else { tmem_alloc(slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (MODE == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + (slot - sbase));
  if (MODE == 1 && warp < 4) {
    // some finite bf16 pairs in the A columns
    uint32_t r[32];
    for (int c = 0; c < 4; ++c) {
      for (int i = 0; i < 32; ++i) r[i] = 0x3C003C00u | ((threadIdx.x * 131u + i * 7u + c) & 0x007F007Fu);
      tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + 32 * c, r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t rank = MODE == 2 ? cluster_ctarank() : 0u;
  if (warp == 0 && lane == 0 && rank == 0) {
    const uint32_t idesc = instr_desc(MODE == 2 ? 256 : 128, N, 0, 0);
    const long long t0 = clock64();
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t sa = abuf + (kb & 3) * 16384, sb = bbuf + (kb % 3) * 32768;
      for (int k = 0; k < 4; ++k) {
        const uint64_t db = smem_desc(sb + k * 32, 16, 1024);
        if (MODE == 1) umma_f16_ts(tmem, tmem + 256 + (kb & 3) * 32 + k * 8, db, idesc, 1u);
        else if (MODE == 2) umma_f16_cta2(tmem, smem_desc(sa + k * 32, 16, 1024), db, idesc, 1u);
        else umma_f16(tmem, smem_desc(sa + k * 32, 16, 1024), db, idesc, 1u);
      }
      if (MODE == 2) umma_commit_cta2(bar + 8, 3); else umma_commit(bar + 8);     // like freeing a ring stage
    }
    if (MODE == 2) umma_commit_cta2(bar, 3); else umma_commit(bar);
    const long long t1 = clock64();
    mbar_wait(bar, 0, 1);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  } else if (warp >= 1 && traffic > 0 && warp <= traffic) {
    // shared-memory store traffic into a region nobody reads, until the MMAs are done
    const uint32_t dst = sbase + 65536 + 3 * 32768 + (uint32_t)(warp - 1) * 4096 + lane * 16;
    uint32_t it = 0;
    while (!mbar_try_wait(bar, 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) st_shared_v4(dst + ((it + j) & 7) * 512, it, it, it, it);
      ++it;
    }
  }
  if (MODE == 2 && rank == 1 && warp == 0) mbar_wait(bar, 0, 2);
  tc_fence_before();
  __syncthreads();
  if (MODE == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (MODE == 2) tmem_dealloc_cta2(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

// Ring hand-shake cost: an issuer thread and a producer thread exchange `stages` buffers through full / empty barriers.
// free_mode 0: the issuer frees a stage with tcgen05.commit, 1: with a plain mbarrier.arrive.  mmas: MMAs per K-block.
__global__ void __launch_bounds__(128, 1) ring_probe(int stages, int free_mode, int mmas, int peek, int nkb, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t abuf = sbase, bbuf = sbase + 65536, bar = sbase + 65536 + 4 * 32768, slot = bar + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (65536 + 3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sgen)[i] = 0x3C003C00u;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (8 + s); };
  const uint32_t done = bar + 8u * 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + (slot - sbase));
  if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(empty(stage), ph ^ 1u, 1);
      mbar_arrive(full(stage));
      if (++stage == stages) { stage = 0; ph ^= 1u; }
    }
  } else if (warp == 0 && lane == 0) {
    const uint32_t idesc = instr_desc(128, 256, 0, 0);
    int stage = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    bool ready = false;
    for (int kb = 0; kb < nkb; ++kb) {
      if (!ready) mbar_wait(full(stage), ph, 2);
      const uint32_t sa = abuf + (kb & 3) * 16384, sb = bbuf + (stage % 3) * 32768;
      int nstage = stage + 1; uint32_t nph = ph;
      if (nstage == stages) { nstage = 0; nph ^= 1u; }
      // peek: poll the next stage's barrier before the MMAs are issued, look at the answer after
      bool nready = false;
      if (peek) nready = mbar_try_wait(full(nstage), nph);
      for (int k = 0; k < mmas; ++k)
        umma_f16(tmem, smem_desc(sa + k * 32, 16, 1024), smem_desc(sb + k * 32, 16, 1024), idesc, 1u);
      if (free_mode == 0) umma_commit(empty(stage)); else mbar_arrive(empty(stage));
      stage = nstage; ph = nph; ready = nready;
    }
    umma_commit(done);
    mbar_wait(done, 0, 3);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
// Issue cost of the synchronisation primitives, one thread, nothing else running in the SM.
__global__ void __launch_bounds__(64, 1) prim_probe(long long* out) {
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t slot_s;
  const uint32_t b_done = smem_u32(&bars[0]), b_pend = smem_u32(&bars[1]), b_big = smem_u32(&bars[2]), b_cm = smem_u32(&bars[3]);
  if (threadIdx.x == 0) {
    mbar_init(b_done, 1); mbar_init(b_pend, 1); mbar_init(b_big, 1 << 19); mbar_init(b_cm, 1 << 19);
    fence_barrier_init();
    mbar_arrive(b_done);                       // phase 0 of b_done is complete
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot_s), 32); tmem_relinquish(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n = 1000;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < n; ++i) acc += mbar_try_wait(b_done, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
    for (int i = 0; i < 50; ++i) acc += mbar_try_wait(b_pend, 0);
    long long t2 = clock64();
    out[1] = (t2 - t1) * 20;
    for (int i = 0; i < n; ++i) mbar_arrive(b_big);
    long long t3 = clock64();
    out[2] = t3 - t2;
    for (int i = 0; i < n; ++i) umma_commit(b_cm);
    long long t4 = clock64();
    out[3] = t4 - t3;
    for (int i = 0; i < n; ++i) tc_fence_after();
    long long t5 = clock64();
    out[4] = t5 - t4;
    for (int i = 0; i < n; ++i) fence_proxy_async();
    long long t6 = clock64();
    out[5] = t6 - t5;
    for (int i = 0; i < n; ++i) { acc += mbar_try_wait(b_done, 0); mbar_arrive(b_big); }
    long long t7 = clock64();
    out[6] = t7 - t6;
    for (int i = 0; i < n; ++i) tc_fence_before();
    long long t8 = clock64();
    out[7] = t8 - t7;
    long long t9 = clock64();
    for (int i = 0; i < n; ++i) { if (!mbar_try_wait(b_done, 0)) break; asm volatile("" ::: "memory"); }
    out[8] = clock64() - t9;
    out[15] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(slot_s, 32);
}
// Latency of the epilogue's per-chunk steps for one warp on an otherwise idle SM (dependent chain, clock64 between).
__global__ void __launch_bounds__(128, 1) epi_probe(long long* out) {
  __shared__ __align__(1024) uint8_t buf[16384];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t slot_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1 << 19); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&slot_s), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot_s;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t sb = smem_u32(buf);
  long long acc[6] = {0, 0, 0, 0, 0, 0};
  uint32_t sink = 0;
  for (int it = 0; it < 200; ++it) {
    uint32_t r[32];
    long long t0 = clock64();
    tmem_ld32(t_lane + 32 * (it & 7), r);
    tmem_ld_wait();
    long long t1 = clock64();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = __uint_as_float(r[2 * i]) + 1.0f, b = __uint_as_float(r[2 * i + 1]) + 2.0f;
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(b), "f"(a));
    }
    long long t2 = clock64();
    const uint32_t rowaddr = sb + (lane + 32 * (warp & 1)) * 128;
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4)
      st_shared_v4(rowaddr + (((warp >> 1) * 4 + i4) ^ (lane & 7)) * 16, pk[4 * i4], pk[4 * i4 + 1], pk[4 * i4 + 2], pk[4 * i4 + 3]);
    long long t3 = clock64();
    fence_proxy_async();
    long long t4 = clock64();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[0]));
    long long t5 = clock64();
    tc_fence_before();
    long long t6 = clock64();
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; acc[5] += t6 - t5;
    sink ^= pk[3];
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 6; ++i) out[i] = acc[i]; out[7] = sink; }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
static void run_epi(long long* d_out) {
  epi_probe<<<1, 128>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("epi: %s\n", cudaGetErrorString(e)); return; }
  long long h[8];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[6] = {"tcgen05.ld x32 + wait::ld", "bias add + relu + bf16 pack (32 values)", "4 x st.shared.v4", "fence.proxy.async",
                          "syncwarp + mbarrier.arrive", "tcgen05.fence::before_thread_sync"};
  for (int i = 0; i < 6; ++i) printf("epilogue step %-42s %.1f cycles\n", names[i], h[i] / 200.0);
}
static void run_prims(long long* d_out) {
  prim_probe<<<1, 64>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("prims: %s\n", cudaGetErrorString(e)); return; }
  long long h[16];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[9] = {"try_wait (phase complete)", "try_wait (pending, returns false)", "mbarrier.arrive", "tcgen05.commit",
                          "tcgen05.fence::after_thread_sync", "fence.proxy.async", "try_wait + arrive", "tcgen05.fence::before_thread_sync", "try_wait + dependent branch"};
  for (int i = 0; i < 9; ++i) printf("prim %-36s %.1f cycles each\n", names[i], h[i] / 1000.0);
}
// Same ring, but the whole issuer warp runs the loop converged and one elected lane issues (the CUTLASS way): the
// operands stay in uniform registers and the per-instruction R2UR / ELECT waterfall of a lane-0 branch disappears.
// flags: 1 = the producer streams 32 KB per K-block from global memory (bulk copy, L2 resident), 2 = eight more warps
// spin on a pending mbarrier, 4 = eight warps read TMEM (tcgen05.ld x32) in a loop, 8 = eight warps write shared
// memory + fence.proxy.async in a loop, 16 = a second (already complete) barrier wait + tcgen05 fence per K-block
// VAR strips pieces of the issuing loop at compile time (bisecting what makes the K-block slower than 4 x 128 cycles):
// bit 0: no layered wait block, bit 1: D address fixed, bit 2: accumulate flag fixed, bit 3: no per-Linear commit
template <int VAR>
__global__ void __launch_bounds__(320, 1) ring_probe_w(int stages, int mmas, int flags, const uint8_t* wsrc, int nkb, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t abuf = sbase, bbuf = sbase + 65536, bar = sbase + 65536 + 4 * 32768, slot = bar + 256;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (65536 + 3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sgen)[i] = 0x3C003C00u;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (8 + s); };
  const uint32_t done = bar + 8u * 16, never = bar + 8u * 17, always = bar + 8u * 18;
  auto acc_full = [&](int r) { return bar + 8u * (19 + r); };
  auto acc_empty = [&](int r) { return bar + 8u * (21 + r); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(done, 1); mbar_init(never, 1); mbar_init(always, 1);
    for (int r = 0; r < 2; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), 8); }
    fence_barrier_init();
    mbar_arrive(always);
  }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + (slot - sbase));
  if (warp == 1) {
    int stage = 0; uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(empty(stage), ph ^ 1u, 1);
      __syncwarp();
      if (elect_one()) {
        if (flags & 1) {
          mbar_arrive_expect_tx(full(stage), 32768u);
          const uint8_t* src = wsrc + (size_t)((kb * 7 + blockIdx.x) % 20) * 32768;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bbuf + (stage % 3) * 32768), "l"(src), "r"(32768u), "r"(full(stage)) : "memory");
        } else {
          mbar_arrive(full(stage));
        }
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; ph ^= 1u; }
    }
  } else if (warp == 0) {
    const uint32_t idesc = instr_desc(128, 256, 0, 0);
    int stage = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    uint32_t uc[2] = {0, 0};
    for (int kb = 0; kb < nkb; ++kb) {
      const int layer = kb >> 2, region = (flags & 64) ? 0 : (layer & 1);
      if (!(VAR & 1) && (flags & 32) && (kb & 3) == 0) {
        // layered mode: four K-blocks make one Linear whose accumulator region must have been drained
        mbar_wait(acc_empty(region), (uc[region] & 1u) ^ 1u, 5);
        ++uc[region];
        tc_fence_after();
      }
      if (flags & 16) { mbar_wait(always, 0, 4); }
      mbar_wait(full(stage), ph, 2);
      if (flags & 16) tc_fence_after();
      const uint32_t sa = abuf + (kb & 3) * 16384, sb = bbuf + (stage % 3) * 32768;
      const uint32_t d = (!(VAR & 2) && (flags & 32)) ? tmem + 256u * region : tmem;
      __syncwarp();
      if (elect_one()) {
        for (int k = 0; k < mmas; ++k)
          umma_f16(d, smem_desc(sa + k * 32, 16, 1024), smem_desc(sb + k * 32, 16, 1024), idesc,
                   (!(VAR & 4) && (flags & 32) && (kb & 3) == 0 && k == 0 && !(flags & 128)) ? 0u : 1u);
        umma_commit(empty(stage));
        if (!(VAR & 8) && (flags & 32) && (kb & 3) == 3) umma_commit(acc_full(region));
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0, 3);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  } else {
    const int q = warp & 3;
    uint32_t sink = 0;
    if (flags & 32) {
      // the drain side of layered mode: wait for the Linear's accumulator, read it, hand the region back
      uint32_t fp[2] = {0, 0};
      for (int layer = 0; layer < nkb / 4; ++layer) {
        const int region = (flags & 64) ? 0 : (layer & 1);
        mbar_wait(acc_full(region), fp[region] & 1u, 6);
        fp[region] ^= 1u;
        tc_fence_after();
        for (int j = 0; j < 4; ++j) {
          uint32_t r[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256u * region + 32 * ((warp - 2) >> 2) + 64 * j, r);
          tmem_ld_wait();
          for (int i = 0; i < 32; ++i) sink ^= r[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(region));
      }
    } else if (flags & 2) {
      while (!mbar_try_wait(done, 0)) { sink += mbar_try_wait(never, 0); }
    } else if (flags & 4) {
      while (!mbar_try_wait(done, 0)) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + 32 * (warp & 7), r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) sink ^= r[i];
      }
    } else if (flags & 8) {
      const uint32_t dst = sbase + 65536 + 3 * 32768 + (uint32_t)(warp - 2) * 4096 + lane * 16;
      uint32_t it = 0;
      while (!mbar_try_wait(done, 0)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(dst + j * 512, it, it, it, it);
        fence_proxy_async();
        ++it;
      }
    }
    if (sink == 0x12345u) out[200] = sink;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int VAR>
static void run_ring_w(int stages, int mmas, int flags, const uint8_t* wsrc, long long* d_out, int sms) {
  const int nkb = 2000;
  const size_t smem = 65536 + 4 * 32768 + 1024 + 512;
  cudaFuncSetAttribute(ring_probe_w<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    ring_probe_w<VAR><<<sms, 320, smem>>>(stages, mmas, flags, wsrc, nkb, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ring_w: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[160];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double tot = 0;
  for (int b = 0; b < sms; ++b) tot += h[b];
  printf("ring (converged warp, elected lane) var=%2d stages=%d mmas/kb=%d flags=%2d: %.1f cycles per K-block\n", VAR, stages, mmas, flags,
         tot / sms / nkb);
}
static void run_ring(int stages, int free_mode, int mmas, int peek, long long* d_out, int sms) {
  const int nkb = 2000;
  const size_t smem = 65536 + 4 * 32768 + 1024 + 512;
  cudaFuncSetAttribute(ring_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    ring_probe<<<sms, 128, smem>>>(stages, free_mode, mmas, peek, nkb, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ring: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[160];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double tot = 0;
  for (int b = 0; b < sms; ++b) tot += h[b];
  printf("ring stages=%d free=%s mmas/kb=%d peek=%d: %.1f cycles per K-block\n", stages, free_mode ? "arrive" : "commit", mmas, peek,
         tot / sms / nkb);
}

template <int MODE>
static void run(const char* name, int N, int traffic, long long* d_out, int sms) {
  const int nkb = 2000;
  const size_t smem = 65536 + 4 * 32768 + 1024 + 256;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(MODE == 2 ? (sms / 2) * 2 : sms);
  cfg.blockDim = dim3(288);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = MODE == 2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_out, 0, sizeof(long long) * 2 * 160);
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<MODE>, N, nkb, traffic, d_out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[320];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0; int n = 0;
  for (int b = 0; b < (int)cfg.gridDim.x; ++b)
    if (h[2 * b + 1] > 0) { issue += h[2 * b]; total += h[2 * b + 1]; ++n; }
  printf("%-28s N=%3d traffic=%d: %.1f cycles per MMA retired (%.1f to issue), %d issuers\n", name, N, traffic,
         total / n / (4.0 * nkb), issue / n / (4.0 * nkb), n);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 2 * 160);
  run_prims(d_out);
  run_epi(d_out);
  uint8_t* wsrc;
  cudaMalloc(&wsrc, 20 * 32768);
  cudaMemset(wsrc, 0x3C, 20 * 32768);
  run_ring_w<15>(3, 4, 0, wsrc, d_out, sms);
  run_ring_w<0>(3, 4, 32, wsrc, d_out, sms);
  for (int st : {3})
    for (int mm : {0, 2, 4, 8})
      for (int pk : {0, 1}) run_ring(st, 0, mm, pk, d_out, sms);
  for (int N : {256, 128, 64}) run<0>("SS one CTA", N, 0, d_out, sms);
  for (int N : {256, 128}) run<1>("TS (A in TMEM) one CTA", N, 0, d_out, sms);
  for (int N : {256, 128}) run<2>("SS CTA pair (M=256)", N, 0, d_out, sms);
  for (int t : {2, 4, 8}) run<0>("SS one CTA + st.shared", 256, t, d_out, sms);
  for (int t : {4, 8}) run<1>("TS one CTA + st.shared", 256, t, d_out, sms);
  for (int t : {4, 8}) run<2>("SS CTA pair + st.shared", 256, t, d_out, sms);
  return 0;
}
