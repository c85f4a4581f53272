// Micro-benchmark of the scoring epilogue's arithmetic (sm_100a): how many SM cycles does one 32-column chunk of
// "sum of 2^(v - m)" cost per warp when a share of the exponentials goes to the FMA pipe?  Register-only work, the
// shape of the real kernel (16 warps per SM, 1 CTA per SM), so the pipe model behind the choice of the polynomial
// share is measured, not guessed.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench benchmarks/ubench_epilogue.cu && /tmp/ubench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// mode bits: kPoly = pairs (of 16) on the FMA pipe, kDeg = polynomial degree (3 or 4), kMax = track the maximum,
// kClamp = clamp the polynomial argument at -126, kPackAcc = accumulate MUFU results with FADD2 (else two FADD)
template <int kPoly, int kDeg, bool kMax, bool kClamp, bool kPackAcc>
__device__ __forceinline__ void chunk(const float (&v)[32], float m, float2& accp, float2& accm0, float2& accm1, float& cmax) {
  if (kMax) {
    float cm = max3(v[0], v[1], v[2]);
#pragma unroll
    for (int i = 3; i < 31; i += 2) cm = max3(cm, v[i], v[i + 1]);
    cmax = max3(cmax, cm, v[31]);
  }
  const float2 nm = make_float2(-m, -m);
  const float MAGIC = 12582912.f;
  const float2 mg = make_float2(MAGIC, MAGIC), nmg = make_float2(-MAGIC, -MAGIC), neg1 = make_float2(-1.f, -1.f);
  const float2 c0 = make_float2(0.9999992847442627f, 0.9999992847442627f);
  const float2 c1 = make_float2(0.6931218504905701f, 0.6931218504905701f);
  const float2 c2 = make_float2(0.240247443318367f, 0.240247443318367f);
  const float2 c3 = make_float2(0.05591766536235809f, 0.05591766536235809f);
  const float2 c4 = make_float2(0.009570018388330936f, 0.009570018388330936f);
#pragma unroll
  for (int i = 0; i < kPoly; ++i) {
    float2 d = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), nm);
    if (kClamp) {
      d.x = fmaxf(d.x, -126.f);
      d.y = fmaxf(d.y, -126.f);
    }
    const float2 t = __fadd2_rn(d, mg);
    const float2 nn = __fadd2_rn(t, nmg);
    const float2 f = __ffma2_rn(nn, neg1, d);
    float2 p;
    if (kDeg == 4) {
      p = __ffma2_rn(c4, f, c3);
      p = __ffma2_rn(p, f, c2);
    } else {
      p = __ffma2_rn(c3, f, c2);
    }
    p = __ffma2_rn(p, f, c1);
    p = __ffma2_rn(p, f, c0);
    float2 e;
    e.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    e.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    accp = __fadd2_rn(accp, e);
  }
#pragma unroll
  for (int i = kPoly; i < 16; i += 2) {
    const float2 d0 = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), nm);
    const float2 d1 = __fadd2_rn(make_float2(v[2 * i + 2], v[2 * i + 3]), nm);
    if (kPackAcc) {
      accm0 = __fadd2_rn(accm0, make_float2(ex2(d0.x), ex2(d0.y)));
      accm1 = __fadd2_rn(accm1, make_float2(ex2(d1.x), ex2(d1.y)));
    } else {
      accm0.x += ex2(d0.x);
      accm0.y += ex2(d0.y);
      accm1.x += ex2(d1.x);
      accm1.y += ex2(d1.y);
    }
  }
}

template <int kPoly, int kDeg, bool kMax, bool kClamp, bool kPackAcc>
__global__ void __launch_bounds__(512, 1) epi_kernel(const float* __restrict__ in, float* __restrict__ out, int iters,
                                                     long long* __restrict__ cycles) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 4095];
  float2 accp = make_float2(0.f, 0.f), accm0 = accp, accm1 = accp;
  float cmax = -3e38f, m = 1.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    chunk<kPoly, kDeg, kMax, kClamp, kPackAcc>(v, m, accp, accm0, accm1, cmax);
    m += 0.0009765625f;  // the stabiliser changes every chunk: nothing can be hoisted
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = accp.x + accp.y + accm0.x + accm0.y + accm1.x + accm1.y + cmax;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- half-precision exponentials: ex2.approx.ftz.f16x2 produces two results per MUFU operation
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
  uint32_t y;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b));
  return y;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t x) {
  float lo, hi;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(x));
  return make_float2(lo, hi);
}
// kHalf of the 16 pairs through f16x2 (accumulated in f16x2 four pairs at a time), the rest through MUFU f32;
// no add of a stabiliser (the accumulator already holds it), like the shared-variance kernel's main loop
template <int kHalf>
__global__ void __launch_bounds__(512, 1) epi_half_kernel(const float* __restrict__ in, float* __restrict__ out, int iters,
                                                          long long* __restrict__ cycles) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 4095];
  float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0, acch = acc0;
  float m = 1.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += m;  // stands in for the fresh accumulator values of the next tile
#pragma unroll
    for (int g4 = 0; g4 < kHalf; g4 += 4) {
      uint32_t h = ex2_f16x2(cvt_f16x2(v[2 * g4], v[2 * g4 + 1]));
#pragma unroll
      for (int i = 1; i < 4; ++i) h = hadd2(h, ex2_f16x2(cvt_f16x2(v[2 * (g4 + i)], v[2 * (g4 + i) + 1])));
      acch = __fadd2_rn(acch, unpack_f16x2(h));
    }
#pragma unroll
    for (int i = kHalf; i < 16; i += 2) {
      acc0 = __fadd2_rn(acc0, make_float2(ex2(v[2 * i]), ex2(v[2 * i + 1])));
      acc1 = __fadd2_rn(acc1, make_float2(ex2(v[2 * i + 2]), ex2(v[2 * i + 3])));
    }
    m = -m;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0.x + acc0.y + acc1.x + acc1.y + acch.x + acch.y;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// plain instruction streams: what one warp-instruction of each kind costs per SM sub-partition
template <int kKind>
__global__ void __launch_bounds__(512, 1) pipe_kernel(const float* __restrict__ in, float* __restrict__ out, int iters,
                                                      long long* __restrict__ cycles) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 8 + i]);
  const float2 k = make_float2(in[5], in[6]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (kKind == 0) a[i] = __ffma2_rn(a[i], k, k);                                             // FFMA2
        if (kKind == 1) a[i] = __fadd2_rn(a[i], k);                                                // FADD2
        if (kKind == 2) { a[i].x = fmaf(a[i].x, k.x, k.y); a[i].y = fmaf(a[i].y, k.x, k.y); }      // 2 x FFMA
        if (kKind == 3) { a[i].x = ex2(a[i].x); a[i].y = ex2(a[i].y); }                            // 2 x MUFU
        if (kKind == 4) { a[i].x = fmaxf(a[i].x, k.x); a[i].y = fmaxf(a[i].y, k.y); }              // 2 x FMNMX
        if (kKind == 5) { a[i].x += k.x; a[i].y += k.y; }                                          // 2 x FADD
        if (kKind == 6) { a[i].x = __uint_as_float(cvt_f16x2(a[i].x, a[i].y)); }                   // cvt.rn.f16x2.f32
        if (kKind == 7) { a[i].x = __uint_as_float(ex2_f16x2(__float_as_uint(a[i].x))); }          // ex2.f16x2
        if (kKind == 8) { a[i] = unpack_f16x2(__float_as_uint(a[i].x)); }                          // 2 x cvt.f32.f16
        if (kKind == 9) { a[i].x = __uint_as_float(hadd2(__float_as_uint(a[i].x), __float_as_uint(k.x))); }  // add.f16x2
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int iters, double per_iter_elems) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  launch(iters, cyc);  // warm-up
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto c : h) avg += (double)c / 148;
  // 16 warps per SM -> 4 per sub-partition; cycles per (warp, iteration) on one sub-partition's pipes
  printf("%-44s %8.3f ms  %9.1f cyc/iter/SM  %7.2f cyc per warp-iter per SMSP  (%.2f cyc/elem/SMSP)  err=%s\n", name, ms,
         avg / iters, avg / iters / 4.0, avg / iters / 4.0 / per_iter_elems, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
}

int main() {
  float *in, *out;
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&out, 148 * 512 * 4);
  std::vector<float> h(4096);
  for (int i = 0; i < 4096; ++i) h[i] = -(float)((i * 37) % 97) * 0.21f;
  cudaMemcpy(in, h.data(), 4096 * 4, cudaMemcpyHostToDevice);
  const int iters = 20000;
#define EPI(P, D, M, C, A)                                                                                                 \
  run("epi poly=" #P " deg=" #D " max=" #M " clamp=" #C " packacc=" #A,                                                    \
      [&](int it, long long* c) { epi_kernel<P, D, M, C, A><<<148, 512>>>(in, out, it, c); }, iters, 32.0)
  EPI(0, 4, true, true, true);
  EPI(2, 4, true, true, true);
  EPI(4, 4, true, true, true);
  EPI(8, 4, true, true, true);
  EPI(0, 4, false, true, true);
  EPI(2, 4, false, true, true);
  EPI(4, 4, false, true, true);
  EPI(6, 4, false, true, true);
  EPI(8, 4, false, true, true);
  EPI(4, 3, false, true, true);
  EPI(6, 3, false, true, true);
  EPI(8, 3, false, true, true);
  EPI(10, 3, false, true, true);
  EPI(4, 3, false, false, true);
  EPI(6, 3, false, false, true);
  EPI(8, 3, false, false, true);
  EPI(0, 3, false, true, false);
  EPI(4, 3, false, true, false);
  EPI(6, 3, false, true, false);
  EPI(8, 3, false, true, false);
  EPI(16, 3, false, true, true);
  EPI(16, 3, false, false, true);
#define EPIH(H) \
  run("epi-half f16x2 pairs=" #H " (no stabiliser add)", [&](int it, long long* c) { epi_half_kernel<H><<<148, 512>>>(in, out, it, c); }, iters, 32.0)
  EPIH(0);
  EPIH(4);
  EPIH(8);
  EPIH(12);
  EPIH(16);
#define PIPE(K, NAME) \
  run(NAME, [&](int it, long long* c) { pipe_kernel<K><<<148, 512>>>(in, out, it, c); }, iters, 64.0)
  PIPE(0, "pipe: 32 x FFMA2 per iter");
  PIPE(1, "pipe: 32 x FADD2 per iter");
  PIPE(2, "pipe: 64 x FFMA per iter");
  PIPE(3, "pipe: 64 x MUFU.EX2 per iter");
  PIPE(4, "pipe: 64 x FMNMX per iter");
  PIPE(5, "pipe: 64 x FADD per iter");
  PIPE(6, "pipe: 32 x cvt.rn.f16x2.f32 per iter");
  PIPE(7, "pipe: 32 x ex2.approx.ftz.f16x2 per iter");
  PIPE(8, "pipe: 32 x (2 x cvt.f32.f16) per iter");
  PIPE(9, "pipe: 32 x add.rn.f16x2 per iter");
  return 0;
}
