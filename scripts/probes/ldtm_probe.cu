// Micro-probe: TMEM read throughput of tcgen05.ld 32x32b.xN with W warps per SM (one CTA per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_probe ldtm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);

template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  return s;
}

// mode 0: ld + wait each; mode 1: two lds in flight before a wait (x32 only)
template <int X, int MODE>
__global__ void probe(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const int nw = blockDim.x >> 5;
  const int share = nw / 4;                 // warps per lane quadrant
  const int part = warp >> 2;               // which share of the 256 columns
  const int cols = 256 / share;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols; c += X) acc ^= ld_cols<X>(base + (uint32_t)((it & 1) * 256 + part * cols + c));
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int X>
void run(int warps, int iters) {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 1024 * 4);
  probe<X, 0><<<148, warps * 32>>>(iters, d, sink);
  cudaDeviceSynchronize();
  probe<X, 0><<<148, warps * 32>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = (double)h[0] / iters;
  printf("x%-3d warps %2d: %8.1f cycles per [128 x 256] fp32 tile read  (%.1f B/cycle/SM)  %s\n", X, warps, cyc, 131072.0 / cyc,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) { run<32>(w, 2000); run<16>(w, 2000); }
  return 0;
}
