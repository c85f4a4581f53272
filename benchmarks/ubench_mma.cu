// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M = 128) by kind (tf32 K = 8, bf16 K = 16), N (64 / 128 / 256),
// A operand location (TMEM "TS" or shared memory "SS") and accumulate-chain length -- how much of an MMA's time is a
// per-instruction cost that a wider N would amortise.  One CTA per SM, one issuing thread, operands are zeros.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ispeech_signal_processing_b200/csrc -Iinclude \
//        -o benchmarks/bin/ubench_mma benchmarks/ubench_mma.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "tc_common.cuh"

using namespace ssp::tc;

__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

// kKind 0: tf32, 1: bf16.  kTs: A in TMEM.  chain MMAs accumulate into one accumulator, then the next accumulator.
template <int kKind, bool kTs, int N, int CHAIN>
__global__ void __launch_bounds__(128, 1) mma_kernel(int chains, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 48 * 4 + N * 48 * 4) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = kKind == 0 ? make_idesc_tf32(128, N, 0, 0) : make_idesc_bf16(128, N);
    const uint64_t a_desc = make_desc(smem_u32(smem), 128 * 16u, 128u);
    const uint64_t b_desc = make_desc(smem_u32(smem) + 128 * 48 * 4, N * 16u, 128u);
    const uint32_t ks_a = (2u * 128 * 16u) >> 4, ks_b = (2u * N * 16u) >> 4;
    constexpr int n_acc = N == 256 ? 1 : N == 128 ? 2 : 4;  // accumulators after the 64 columns kept for the A operand
    const long long t0 = clock64();
#pragma unroll 1
    for (int c = 0; c < chains; ++c) {
      const uint32_t d = base + 64 + (uint32_t)(c & (n_acc - 1)) * N;
#pragma unroll
      for (int k = 0; k < CHAIN; ++k) {
        const int kk = k % 6;
        if (kTs) {
          if (kKind == 0) mma_tf32_ts(d, base + 8u * kk, b_desc + (uint64_t)(kk * ks_b), idesc, k > 0);
          else mma_bf16_ts(d, base + 8u * kk, b_desc + (uint64_t)(kk * ks_b), idesc, k > 0);
        } else {
          if (kKind == 0) tc_mma_tf32(d, a_desc + (uint64_t)(kk * ks_a), b_desc + (uint64_t)(kk * ks_b), idesc, k > 0);
          else mma_bf16_ss(d, a_desc + (uint64_t)(kk * ks_a), b_desc + (uint64_t)(kk * ks_b), idesc, k > 0);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

template <int kKind, bool kTs, int N, int CHAIN>
static void run(int grid) {
  const int chain = CHAIN, chains = 60000 / chain;
  long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const size_t smem = 128 * 48 * 4 + (size_t)N * 48 * 4 + 1024;
  cudaFuncSetAttribute(mma_kernel<kKind, kTs, N, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    mma_kernel<kKind, kTs, N, CHAIN><<<grid, 128, smem>>>(chains, cyc);
    cudaDeviceSynchronize();
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto c : h) avg += (double)c / grid;
  const double per = avg / ((double)chains * chain);
  const double flop = 2.0 * 128 * N * (kKind == 0 ? 8 : 16);
  printf("%s %s N=%3d chain=%2d grid=%3d  %7.1f cyc/mma  %7.0f FLOP/cyc/SM  err=%s\n", kKind == 0 ? "tf32" : "bf16", kTs ? "TS" : "SS", N,
         chain, grid, per, flop / per, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
}

template <int CHAIN>
static void sweep(int grid) {
  run<0, true, 64, CHAIN>(grid);
  run<0, true, 128, CHAIN>(grid);
  run<0, true, 256, CHAIN>(grid);
  run<0, false, 64, CHAIN>(grid);
  run<0, false, 128, CHAIN>(grid);
  run<1, true, 64, CHAIN>(grid);
  run<1, false, 64, CHAIN>(grid);
  run<1, false, 128, CHAIN>(grid);
  run<1, false, 256, CHAIN>(grid);
}

int main() {
  for (int grid : {1, 148}) {
    sweep<1>(grid);
    sweep<6>(grid);
    sweep<24>(grid);
  }
  return 0;
}
