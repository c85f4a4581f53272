// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (sm_100a) with 4, 8 and 16 warps, 32 columns per
// instruction.  Decides whether the scoring epilogue can afford extra TMEM traffic (re-initialising accumulators).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o benchmarks/bin/ubench_tmem benchmarks/ubench_tmem.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// mode 0: loads only, 1: stores only, 2: load + store per iteration
template <int kMode>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col0 = ((warp >> 2) * 64) & 511;  // warps of one quadrant work on different columns
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  st32(base + lane_addr + col0, r);
  st32(base + lane_addr + col0 + 32, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + lane_addr + col0 + (it & 1) * 32;
    if (kMode == 0 || kMode == 2) {
      ld32(a, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[31];
    }
    if (kMode == 1 || kMode == 2) {
      st32(a, r);
      if (kMode == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  if (kMode == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[3];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

template <int kMode>
static void run(const char* name, int threads) {
  const int iters = 20000;
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 512 * 4);
  tmem_kernel<kMode><<<148, threads>>>(iters, cyc, sink);
  cudaDeviceSynchronize();
  tmem_kernel<kMode><<<148, threads>>>(iters, cyc, sink);
  cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto c : h) avg += (double)c / 148;
  const double bytes_per_iter = (double)threads * 32 * 4 * (kMode == 2 ? 2 : 1);
  printf("%-28s threads=%3d  %8.1f cyc/iter  -> %7.1f B/cyc/SM   err=%s\n", name, threads, avg / iters,
         bytes_per_iter / (avg / iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0>("tcgen05.ld 32x32b.x32 + wait", th);
    run<1>("tcgen05.st 32x32b.x32", th);
    run<2>("ld + st", th);
  }
  return 0;
}
