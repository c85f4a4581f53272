// GMM-UBM scoring on the 5th-generation tensor cores (sm_100a, tcgen05 + TMEM + TMA bulk copies).
//
// The diagonal-Gaussian log-likelihood (sklearn _gaussian_mixture.py:536-553) is the dense contraction
//     L2[t, c] = [x_t, x_t^2, 1, 1] . [mu_c/var_c, -1/(2 var_c), c_hi, c_lo] * log2(e)
// with frames as M, (model, component) as N and 2D+2 as the K dimension.  One persistent CTA per SM owns
// 256 frames at a time ("unit"): the frames' [x, x^2, 1, 1] rows are rounded to TF32 (cvt.rna) and laid
// out ONCE in shared memory as two 128-row UMMA operands, then every model's component tiles (pre-packed
// as shared-memory images by gmm_pack_kernel) stream through a 3-stage cp.async.bulk ring.  All CTAs walk
// the same tile sequence, so each tile is read from HBM once per wave and served to the other SMs by L2.
//
//   warp 0      : producer   - one lane issues cp.async.bulk (global -> smem) + mbarrier expect_tx
//   warp 1      : MMA issuer - one lane issues tcgen05.mma kind::tf32 (M128 x N128 x K8) into 4 TMEM slots,
//                 tcgen05.commit signals "slot full" / "stage free"
//   warps 2..9  : epilogue   - tcgen05.ld 32 lanes x 32 columns; thread == frame row; online (max, sum 2^x)
//                 across a model's tiles -> per-frame log-likelihood -> warp-segmented per-utterance sum
//                 -> one double atomic per (warp, utterance, model).  The T x K logits never leave the SM.
//
// The log-constant rides through the MMA as two TF32-exact pieces (c_hi + c_lo) against A columns of 1.0,
// so the epilogue is max / ex2 / add only.
#include "common.cuh"

namespace ssp {

namespace tc {
constexpr int BM = 128;      // rows per accumulator (TMEM lanes)
constexpr int MB = 2;        // row blocks per unit
constexpr int UNIT = BM * MB;
constexpr int BN = kTileN;   // components per tile
constexpr int NSTAGE = 3;
constexpr int NSLOT = 4;     // TMEM accumulator slots (BN fp32 columns each) = all 512 columns
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int MAX_KD = 80;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU (2^32 SM cycles ~ 2-3 s).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 32)) {
      printf("ssp gmm_score_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
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
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, no-swizzle UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor, sm_100 "version 1"):
// core matrix = 8 rows x 16 bytes stored contiguously (128 B); SBO = stride between 8-row groups,
// LBO = stride between the two 16-byte K chunks of one K=8 TF32 step.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for Blackwell
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, K-major A and B,
// n_dim = N>>3 @17, m_dim = M>>4 @24.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Args {
  const float* feats;
  const int64_t* offsets;
  int64_t n_utts, total_frames;
  const float* tiles;  // [n_models][Kp/BN][KD/4][BN] float4 images
  int n_models, tiles_per_model, D, KD;
  int normalize;
  double* scores;
  float* frame_lse;
};

__global__ void __launch_bounds__(THREADS, 1) gmm_score_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int KD = a.KD;
  const uint32_t tile_bytes = (uint32_t)BN * KD * 4u;
  float* sA = reinterpret_cast<float*>(smem);                     // [MB][KC][BM][4]
  unsigned char* sB = smem + (size_t)MB * tile_bytes;             // [NSTAGE][KC][BN][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NSTAGE * tile_bytes);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + NSTAGE;
  uint64_t* t_full = bars + 2 * NSTAGE;
  uint64_t* t_empty = t_full + NSLOT;
  uint64_t* a_full = t_empty + NSLOT;
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < NSLOT; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, BM); }
    mbar_init(a_full, EPI_WARPS * 32);
    mbar_init(a_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(NSLOT * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_units = (a.total_frames + UNIT - 1) / UNIT;
  const int tiles_per_unit = a.n_models * a.tiles_per_model;
  const size_t tile_floats = (size_t)BN * KD;

  if (warp == 0) {
    // ===================== producer: stream every model's tiles, once per unit =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        for (int t = 0; t < tiles_per_unit; ++t, ++it) {
          const uint32_t stage = it % NSTAGE, ph = (it / NSTAGE) & 1u;
          mbar_wait(b_empty + stage, ph ^ 1u);
          mbar_arrive_expect_tx(b_full + stage, tile_bytes);
          bulk_g2s(sB + (size_t)stage * tile_bytes, a.tiles + (size_t)t * tile_floats, tile_bytes, b_full + stage);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t lbo = BM * 16u, sbo = 128u;  // chunk stride = 128 rows x 16 B; 8-row groups are contiguous
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      uint32_t it = 0, q = 0, unit_idx = 0;
      for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_idx) {
        mbar_wait(a_full, unit_idx & 1u);
        tc_fence_after();
        for (int t = 0; t < tiles_per_unit; ++t, ++it) {
          const uint32_t stage = it % NSTAGE, ph = (it / NSTAGE) & 1u;
          mbar_wait(b_full + stage, ph);
          tc_fence_after();
#pragma unroll
          for (int mb = 0; mb < MB; ++mb, ++q) {
            const uint32_t slot = q % NSLOT, sph = (q / NSLOT) & 1u;
            mbar_wait(t_empty + slot, sph ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + slot * BN;
            for (int k = 0; k < (KD >> 3); ++k) {
              const uint64_t ad = make_desc(a_base + mb * tile_bytes + k * 2u * lbo, lbo, sbo);
              const uint64_t bd = make_desc(b_base + stage * tile_bytes + k * 2u * lbo, lbo, sbo);
              tc_mma_tf32(d_tmem, ad, bd, kIdesc, k > 0 ? 1u : 0u);
            }
            tc_commit(t_full + slot);
          }
          tc_commit(b_empty + stage);
        }
        tc_commit(a_empty);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (also build the A operand) =====================
    const int etid = tid - 64;                 // 0..255: row of the unit this thread BUILDS
    const int g = (warp - 2) >> 2;             // row block this thread READS accumulators of
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // accumulator row within the row block
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float LN2 = 0.69314718055994530942f;
    uint32_t n = 0, unit_idx = 0;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_idx) {
      const int64_t frame0 = u * UNIT;
      // ---- A operand: row r of block mb at float index ((mb*KC + j/4)*BM + r)*4 + j%4
      mbar_wait(a_empty, (unit_idx & 1u) ^ 1u);
      {
        const int64_t fr = frame0 + etid;
        const bool live = fr < a.total_frames;
        const float* xr = a.feats + fr * a.D;
        float* dst = sA + (size_t)(etid >> 7) * tile_floats + (size_t)(etid & (BM - 1)) * 4;
        for (int j = 0; j < KD; ++j) {
          float v = 0.f;
          if (j < a.D) v = live ? rna_tf32(xr[j]) : 0.f;
          else if (j < 2 * a.D) { float x = live ? xr[j - a.D] : 0.f; v = rna_tf32(x * x); }
          else if (j < 2 * a.D + 2) v = 1.f;
          dst[(size_t)(j >> 2) * (BM * 4) + (j & 3)] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(a_full);
      }
      // ---- the frame this thread scores
      const int64_t fe = frame0 + (int64_t)g * BM + row;
      const bool live = fe < a.total_frames;
      const int utt = live ? find_segment(a.offsets, a.n_utts, fe) : -1;
      float wgt = 1.f;
      if (utt >= 0 && a.normalize) wgt = 1.f / (float)(a.offsets[utt + 1] - a.offsets[utt]);

      for (int model = 0; model < a.n_models; ++model) {
        float m_run = -3.0e38f, s_run = 0.f;
        for (int t = 0; t < a.tiles_per_model; ++t, ++n) {
          const uint32_t slot = 2u * (n & 1u) + (uint32_t)g, ph = (n >> 1) & 1u;
          mbar_wait(t_full + slot, ph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + slot * BN;
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            tc_ld32(taddr + c * 32, v);
            float cmax = v[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) cmax = fmaxf(cmax, v[i]);
            const float m_new = fmaxf(m_run, cmax);
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              acc0 += ex2(v[i] - m_new);
              acc1 += ex2(v[i + 1] - m_new);
            }
            s_run = fmaf(s_run, ex2(m_run - m_new), acc0 + acc1);
            m_run = m_new;
          }
          tc_fence_before();
          mbar_arrive(t_empty + slot);
        }
        const float lse = (m_run + lg2(s_run)) * LN2;
        if (live && a.frame_lse) a.frame_lse[(int64_t)model * a.total_frames + fe] = lse;
        warp_segmented_atomic_add(a.scores, utt, a.n_models, model, lse * wgt, lane);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(NSLOT * BN))
                 : "memory");
  }
}

}  // namespace tc

int launch_score_tc(const float* feats, const int64_t* offsets, int64_t n_utts, int64_t total_frames, const void* pack,
                    const PackLayout& L, bool normalize, double* scores, float* frame_lse, cudaStream_t st) {
  using namespace tc;
  if (L.KD > MAX_KD) {
    set_error("ssp_gmm_score(tf32): feature dim %d needs a contraction of %d > %d; use SSP_PREC_FP32", L.D, L.KD, MAX_KD);
    return SSP_EUNSUP;
  }
  SSP_CUDA_OK(cudaMemsetAsync(scores, 0, sizeof(double) * n_utts * L.n_models, st));
  if (total_frames == 0) return SSP_OK;
  Args a;
  a.feats = feats;
  a.offsets = offsets;
  a.n_utts = n_utts;
  a.total_frames = total_frames;
  a.tiles = (const float*)((const char*)pack + L.off_tile);
  a.n_models = L.n_models;
  a.tiles_per_model = L.Kp / BN;
  a.D = L.D;
  a.KD = L.KD;
  a.normalize = normalize ? 1 : 0;
  a.scores = scores;
  a.frame_lse = frame_lse;
  const size_t tile_bytes = (size_t)BN * L.KD * 4;
  const size_t smem = (MB + NSTAGE) * tile_bytes + (2 * NSTAGE + 2 * NSLOT + 2) * sizeof(uint64_t) + 16;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SSP_CUDA_OK(cudaGetDevice(&dev));
    SSP_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  SSP_CUDA_OK(cudaFuncSetAttribute(gmm_score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_units = (total_frames + UNIT - 1) / UNIT;
  const unsigned grid = (unsigned)(n_units < num_sms ? n_units : num_sms);
  gmm_score_tc_kernel<<<grid, THREADS, smem, st>>>(a);
  SSP_LAUNCH_CHECK("gmm_score_tc_kernel");
  return SSP_OK;
}

}  // namespace ssp
