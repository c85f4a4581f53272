// EM / MAP sufficient statistics on the tensor cores (sm_100a): posteriors and N/F/S as two chained GEMMs.
//
// Posteriors need FP32-grade logits (a 4e-3 logit error is a 0.4 % responsibility error), so the logit GEMM runs
// as 3xTF32: hi.hi + lo.hi + hi.lo with the frame operand [x, x^2, 1, 1] and the model operand
// log2(e) [mu/var, -1/(2var), c...] both split into TF32 hi + lo pieces (the constant rides as three exact
// pieces), FP32 accumulation in TMEM.  kind::tf32 only takes K-major shared-memory operands (an MN-major
// descriptor silently yields zeros on sm_100a), so every operand is laid out with its contraction index
// contiguous.
//
//   grid = (frame chunks, component tiles of 128); frame blocks stream through each CTA.  Two passes (the per-frame
//   normaliser needs ALL component tiles):
//
//   pass LSE   (gmm_em_lse_kernel, 128-frame blocks): logits[frame, comp]; the frame operand lives double-buffered in
//              TENSOR MEMORY (tcgen05.st, thread == frame row), the model tile (hi + lo) in shared memory; thread ==
//              frame row -> per-tile (max, sum 2^x) partials to the workspace.
//   pass STATS (gmm_em_stats_kernel, 64-frame blocks): TRANSPOSED logits[comp, frame] = B . F^T with the model tile
//              (hi + lo) resident in TENSOR MEMORY, so that thread == component row: gamma = 2^(L2 - lse2[frame]) (lse2
//              from the partials) is written by tcgen05.st IN PLACE over the logits and is the TMEM-side operand of
//              GEMM 2:  stats[comp, :] += gamma . [x, x^2, 1, 1] over the frames, accumulated in TMEM across all blocks
//              of a segment, hi + lo passes of the frame-contiguous feature operand Xt.  At a segment end the
//              128 x (2D+2) accumulator is added to the double-precision N / F / S outputs.
//   Both passes keep two blocks in flight (double / triple-buffered operands): the 16 builder / epilogue warps split
//   block k and post-process block k - 1 while the tensor core multiplies.
// Blocks never straddle a segment, so one kernel pair serves UBM EM (one segment) and batched MAP enrolment
// (one segment per speaker).
#include <cstdlib>

#include "tc_common.cuh"

namespace ssp {
namespace em {

using namespace tc;

constexpr int BN = 128;   // components per CTA
constexpr int BM1 = 128;  // frames per block, pass LSE
constexpr int BM2 = 64;   // frames per block, pass STATS
constexpr int EPI = 512;  // 16 builder / epilogue warps, four per TMEM lane quadrant: the thread work around the MMAs (operand
                          // split, exponentials) is latency-bound, more warps hide more of it (8 warps: 7 % slower)
constexpr int THREADS = 64 + EPI;
constexpr int MAX_KD = 80;

struct Args {
  const float* feats;
  const int64_t* seg;
  int64_t n_segs, total_frames, chunk;
  const float* tiles_hi;  // [Kp/128][KD/4][128] float4
  const float* tiles_lo;
  int K, D, KD, n_tiles;
  float2* partial;        // [2 * n_tiles][total_frames]: (max, sum 2^(x-max)) per 64-component half tile, log2 domain
  float* frame_lse;       // natural-log per-frame likelihood (written by tile 0 in the STATS pass)
  double* out_n;
  double* out_f;
  double* out_s;
  double* out_loglik;
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Every role walks the same block sequence: [t0, t0 + nt) inside one segment and inside this CTA's chunk.
template <int BM>
struct Walk {
  const int64_t* seg;
  int64_t n_segs, end, t0, seg_end;
  int cur;
  __device__ bool start(const int64_t* s, int64_t n, int64_t b, int64_t e) {
    seg = s; n_segs = n; end = e; t0 = b;
    if (b >= e) return false;
    cur = find_segment(seg, n_segs, t0);
    if (cur < 0) return false;
    seg_end = seg[cur + 1];
    return true;
  }
  __device__ int nt() const { return (int)min((int64_t)BM, min(end, seg_end) - t0); }
  // advance; returns false when the chunk is exhausted.  `flush`: the block just finished was the last one of its
  // segment inside this chunk.
  __device__ bool next(int n, bool& flush) {
    t0 += n;
    flush = (t0 >= seg_end) || (t0 >= end);
    if (t0 >= end) return false;
    if (t0 >= seg_end) {
      cur = find_segment(seg, n_segs, t0);
      if (cur < 0) return false;
      seg_end = seg[cur + 1];
    }
    return true;
  }
};

// [x, x^2, 1, 1, 0..] element j of a frame row held in shared memory, split into TF32 hi / lo
__device__ __forceinline__ void feat_split(const float* xr, int j, int D, bool live, float& hi, float& lo) {
  float v = 0.f;
  if (live) {
    if (j < D) v = xr[j];
    else if (j < 2 * D) { const float x = xr[j - D]; v = x * x; }
    else if (j < 2 * D + 2) v = 1.f;
  }
  hi = rna_tf32(v);
  lo = rna_tf32(v - hi);
}

// The four elements of K chunk jc of a frame row, split into TF32 hi / lo.  A chunk is almost always all-x or all-x^2
// (only the chunks that contain column D, 2D or the ones are mixed), and jc is warp-uniform, so the common case is
// straight-line: load, (square), round, subtract, round.
__device__ __forceinline__ void chunk_split(const float* xr, int jc, int D, bool live, float (&h)[4], float (&l)[4]) {
  const int j0 = 4 * jc;
  if (!live) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { h[e] = 0.f; l[e] = 0.f; }
  } else if (j0 + 3 < D) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = xr[j0 + e];
      h[e] = rna_tf32(v);
      l[e] = rna_tf32(v - h[e]);
    }
  } else if (j0 >= D && j0 + 3 < 2 * D) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x = xr[j0 + e - D];
      const float v = x * x;
      h[e] = rna_tf32(v);
      l[e] = rna_tf32(v - h[e]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) feat_split(xr, j0 + e, D, true, h[e], l[e]);
  }
}

// Each of the 128 builder threads keeps its share of the NEXT block's features in registers: the global loads are
// issued right after the current block's operands are handed to the MMA warp and land while the tensor core and
// the epilogue work, instead of being waited for at the top of every block.
template <int R>
__device__ __forceinline__ void prefetch_block(const float* __restrict__ src, int n, int et, float (&pf)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int idx = et + EPI * r;
    pf[r] = idx < n ? __ldg(src + idx) : 0.f;
  }
}
template <int R>
__device__ __forceinline__ void store_block(float* dst, int n, int et, const float (&pf)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int idx = et + EPI * r;
    if (idx < n) dst[idx] = pf[r];
  }
}

// ================================================================================================ pass STATS
// Two frame blocks in flight, so that the tensor core works on one block while the 16 builder / epilogue warps work
// on its neighbours:
//   * the CTA's model tile (TF32 hi + lo) is copied ONCE into tensor memory (160 columns) and is the TMEM-side (A)
//     operand of the logit GEMM (component == TMEM lane): 80 KB of shared memory become operand buffers;
//   * gamma is written by tcgen05.st IN PLACE over the logits it came from and is the TMEM-side operand of the
//     statistics GEMM: it never touches shared memory;
//   * two F buffers (frames as rows, GEMM 1), three Xt buffers (frames contiguous, GEMM 2) and two 64-column logit
//     buffers; thread program per block k: build F / Xt of block k, then turn the logits of block k - 1 into gamma;
//     MMA program: GEMM1(k), GEMM2(k - 1).  The builders never wait for a GEMM that was issued less than a block
//     ago, so in steady state the period is max(thread work, 1600 tensor cycles) per 64 frames.
namespace p3 {
constexpr int NF = 2, NX = 3;
constexpr uint32_t COL_BHI = 0, COL_BLO = 80;   // model tile, hi and lo (MAX_KD columns each)
constexpr uint32_t COL_LOGIT = 160;             // + 64 * (k & 1)
constexpr uint32_t COL_STAT = 288;              // statistics accumulator, n2 <= 80 columns
struct Carve {
  int n2, xrows;
  size_t f_bytes, x_bytes, o_x, o_stage, o_lse, o_bar, bytes;  // F buffers at 0 (hi | lo each), Xt buffers at o_x
};
__host__ __device__ inline Carve carve(int KD) {
  Carve c;
  c.n2 = (KD + 15) & ~15;
  c.xrows = c.n2 + 1;
  c.f_bytes = (size_t)BM2 * KD * 4;
  c.x_bytes = (size_t)(BM2 / 4) * c.xrows * 16;
  c.o_x = NF * 2 * c.f_bytes;
  c.o_stage = c.o_x + NX * 2 * c.x_bytes;
  c.o_lse = c.o_stage + (size_t)BM2 * (MAX_KD / 2) * 4;
  c.o_bar = c.o_lse + NX * BM2 * 4;
  c.bytes = c.o_bar + 128;
  return c;
}
}  // namespace p3

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) gmm_em_stats_kernel(const Args a) {
  using namespace p3;
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int BM = BM2;
  const int KD = a.KD, KC = KD >> 2, D = a.D;
  const Carve cv = carve(KD);
  const uint32_t tile_bytes = (uint32_t)BN * KD * 4u;
  float* sStage = reinterpret_cast<float*>(smem + cv.o_stage);
  float* sLse = reinterpret_cast<float*>(smem + cv.o_lse);  // [3][BM]: written while building block k, read by its epilogue one block later
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cv.o_bar);
  uint64_t* b_loaded = bars;       // model tile images landed in shared memory (start-up only)
  uint64_t* a_full = bars + 1;     // [2] F / Xt of block k are built               (k & 1)
  uint64_t* l_full = bars + 3;     // [2] logits of block k are in TMEM, F[k & 1] is free again
  uint64_t* g_full = bars + 5;     // [2] gamma of block k is in TMEM
  uint64_t* x_free = bars + 7;     // [3] GEMM 2 of block k has completed: Xt[k % 3] free, statistics include block k
  uint64_t* drained = bars + 10;   // the statistics accumulator has been read out after a segment end
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.y;
  if (tid == 0) {
    mbar_init(b_loaded, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(a_full + g, EPI);
      mbar_init(l_full + g, 1);
      mbar_init(g_full + g, EPI);
    }
    for (int g = 0; g < NX; ++g) mbar_init(x_free + g, 1);
    mbar_init(drained, EPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t begin = (int64_t)blockIdx.x * a.chunk;
  const int64_t end = min(begin + a.chunk, a.total_frames);
  if (begin >= end) {  // (uniform) nothing to do for this CTA
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    return;
  }

  // ---- start-up: model tile hi / lo -> shared memory (the F buffers as scratch) -> tensor memory
  float* scratch = reinterpret_cast<float*>(smem);  // 2 x tile_bytes == NF x 2 x f_bytes
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(b_loaded, 2u * tile_bytes);
    bulk_g2s(scratch, a.tiles_hi + (size_t)tile * BN * KD, tile_bytes, b_loaded);
    bulk_g2s(scratch + BN * KD, a.tiles_lo + (size_t)tile * BN * KD, tile_bytes, b_loaded);
  }
  if (warp >= 2 && warp < 10) {
    const int which = (warp - 2) >> 2;           // warps 2..5 copy hi, 6..9 copy lo
    const int row = ((warp & 3) << 5) | lane;    // component row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    mbar_wait(b_loaded, 0);
    const float4* img = reinterpret_cast<const float4*>(scratch + (size_t)which * BN * KD) + row;
    for (int k = 0; k < (KD >> 3); ++k) {
      const float4 c0 = img[(2 * k) * BN], c1 = img[(2 * k + 1) * BN];
      const float v[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      tc_st8(tmem_base + lane_addr + (which ? COL_BLO : COL_BHI) + 8u * k, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  __syncthreads();  // scratch is free again; the tile is visible to the MMA warp
  tc_fence_after();

  if (warp == 1) {
    // ===================== MMA issuer =====================
    Walk<BM> w;
    if (w.start(a.seg, a.n_segs, begin, end)) {
      constexpr uint32_t lbo_f = BM * 16u, sbo = 128u;
      constexpr uint32_t ks_f = (2u * lbo_f) >> 4;
      const uint32_t lbo_x = (uint32_t)cv.xrows * 16u, ks_x = (2u * lbo_x) >> 4;
      const uint32_t base = smem_u32(smem);
      const int ksteps = KD >> 3;
      const uint32_t idesc1 = make_idesc_tf32(BN, BM, 0, 0);
      const uint32_t idesc2 = make_idesc_tf32(BN, cv.n2, 0, 0);
      const uint32_t t_bhi = tmem_base + COL_BHI, t_blo = tmem_base + COL_BLO, t_stat = tmem_base + COL_STAT;
      uint32_t k = 0, n_drains = 0;
      bool prev_first = true, prev_flush = false, pending_drain = false, first_in_seg = true, more = true;
      // GEMM 2 of block p: statistics += gamma (TMEM, in the logit columns of p) . Xt[p % 3] (shared memory)
      auto gemm2 = [&](uint32_t p, bool first, bool flush) {
        mbar_wait(g_full + (p & 1u), (p >> 1) & 1u);
        if (pending_drain) {
          mbar_wait(drained, n_drains & 1u);
          ++n_drains;
          pending_drain = false;
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t xb = base + (uint32_t)(cv.o_x + (size_t)(p % NX) * 2 * cv.x_bytes);
          const uint64_t xhi = make_desc(xb, lbo_x, sbo), xlo = make_desc(xb + (uint32_t)cv.x_bytes, lbo_x, sbo);
          const uint32_t t_gam = tmem_base + COL_LOGIT + 64u * (p & 1u);
          for (int q = 0; q < BM / 8; ++q) tc_mma_tf32_ts(t_stat, t_gam + 8u * q, xhi + (uint64_t)(q * ks_x), idesc2, (first && q == 0) ? 0u : 1u);
          for (int q = 0; q < BM / 8; ++q) tc_mma_tf32_ts(t_stat, t_gam + 8u * q, xlo + (uint64_t)(q * ks_x), idesc2, 1u);
          tc_commit(x_free + (p % NX));
        }
        __syncwarp();
        if (flush) pending_drain = true;
      };
      while (more) {
        const int nt = w.nt();
        mbar_wait(a_full + (k & 1u), (k >> 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t fb = base + (uint32_t)((size_t)(k & 1u) * 2 * cv.f_bytes);
          const uint64_t fhi = make_desc(fb, lbo_f, sbo), flo = make_desc(fb + (uint32_t)cv.f_bytes, lbo_f, sbo);
          const uint32_t t_log = tmem_base + COL_LOGIT + 64u * (k & 1u);
          for (int q = 0; q < ksteps; ++q) tc_mma_tf32_ts(t_log, t_bhi + 8u * q, fhi + (uint64_t)(q * ks_f), idesc1, q > 0 ? 1u : 0u);
          for (int q = 0; q < ksteps; ++q) tc_mma_tf32_ts(t_log, t_bhi + 8u * q, flo + (uint64_t)(q * ks_f), idesc1, 1u);
          for (int q = 0; q < ksteps; ++q) tc_mma_tf32_ts(t_log, t_blo + 8u * q, fhi + (uint64_t)(q * ks_f), idesc1, 1u);
          tc_commit(l_full + (k & 1u));
        }
        __syncwarp();
        if (k > 0) gemm2(k - 1, prev_first, prev_flush);
        bool flush;
        more = w.next(nt, flush);
        prev_first = first_in_seg;
        prev_flush = flush;
        first_in_seg = flush;
        ++k;
      }
      gemm2(k - 1, prev_first, prev_flush);
    }
  } else if (warp >= 2) {
    // ===================== operand builders (thread == frame) and epilogue (thread == component) =====================
    const int row = ((warp & 3) << 5) | lane;  // TMEM lane == component row within the tile
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int et = tid - 64;                   // 0..511
    const int cq = (warp - 2) >> 2;            // which 16 of a block's 64 frame columns this warp turns into gamma
    const int fr = et & (BM - 1), part = et >> 6;  // frame row this thread builds, and which eighth of its K chunks
    const bool norm_thread = part == 7;        // these 64 threads also own the per-frame normaliser
    const float LN2 = 0.69314718055994530942f;
    const int n_part = 2 * a.n_tiles;
    Walk<BM> w;
    if (w.start(a.seg, a.n_segs, begin, end)) {
      constexpr int R = (BM * (MAX_KD / 2 - 1) + EPI - 1) / EPI;  // D <= 39
      constexpr int PT = 8;
      float pf[R];
      float2 pp[PT];
      auto prefetch = [&](const Walk<BM>& wb) {
        prefetch_block<R>(a.feats + wb.t0 * D, wb.nt() * D, et, pf);
        if (norm_thread && fr < wb.nt()) {
#pragma unroll
          for (int y = 0; y < PT; ++y)
            if (y < n_part) pp[y] = a.partial[(size_t)y * a.total_frames + wb.t0 + fr];
        }
      };
      prefetch(w);
      uint32_t k = 0, n_drained = 0;
      bool more = true, prev_flush = false;
      int prev_seg = -1, ll_seg = -1;
      float ll_acc = 0.f;
      auto flush_ll = [&]() {
        if (tile == 0 && ll_seg >= 0) {
          const float tot = warp_sum(ll_acc);
          if (lane == 0 && tot != 0.f) atomicAdd(a.out_loglik + ll_seg, (double)tot);
        }
        ll_acc = 0.f;
      };
      // logits of block p -> gamma in place; at a segment end also read out the statistics
      auto epilogue = [&](uint32_t p, int seg_id, bool flush) {
        mbar_wait(l_full + (p & 1u), (p >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_log = tmem_base + lane_addr + COL_LOGIT + 64u * (p & 1u) + 16u * cq;
        const float* lse = sLse + (p % NX) * BM + 16 * cq;
        uint32_t r[16];
        tc_ld16(t_log, r);
        float gam[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ls = *reinterpret_cast<const float4*>(lse + 4 * q);
          gam[4 * q + 0] = rna_tf32(ex2(__uint_as_float(r[4 * q + 0]) - ls.x));
          gam[4 * q + 1] = rna_tf32(ex2(__uint_as_float(r[4 * q + 1]) - ls.y));
          gam[4 * q + 2] = rna_tf32(ex2(__uint_as_float(r[4 * q + 2]) - ls.z));
          gam[4 * q + 3] = rna_tf32(ex2(__uint_as_float(r[4 * q + 3]) - ls.w));
        }
        tc_st16(t_log, gam);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(g_full + (p & 1u));
        if (flush) {
          mbar_wait(x_free + (p % NX), (p / NX) & 1u);  // GEMM 2 of block p (and of everything before it) is complete
          tc_fence_after();
          const int comp = tile * BN + row;
          const uint32_t saddr = tmem_base + lane_addr + COL_STAT;
#pragma unroll 1
          for (int c0 = 16 * cq; c0 < KD; c0 += 64) {
            uint32_t s16[16];
            tc_ld16(saddr + c0, s16);
            if (comp < a.K) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int jj = c0 + e;
                const double v = (double)__uint_as_float(s16[e]);
                if (jj < D) atomicAdd(a.out_f + ((int64_t)seg_id * a.K + comp) * D + jj, v);
                else if (jj < 2 * D) atomicAdd(a.out_s + ((int64_t)seg_id * a.K + comp) * D + (jj - D), v);
                else if (jj == 2 * D) atomicAdd(a.out_n + (int64_t)seg_id * a.K + comp, v);
              }
            }
          }
          tc_fence_before();
          mbar_arrive(drained);
          ++n_drained;
        }
      };
      while (more) {
        const int nt = w.nt();
        const int64_t t0 = w.t0;
        const int seg_id = w.cur;
        if (seg_id != ll_seg) {
          flush_ll();
          ll_seg = seg_id;
        }
        // ---- build block k: F[k & 1] was released by GEMM1(k - 2) (waited for in the epilogue of k - 2), Xt[k % 3] by
        // GEMM2(k - 3)
        mbar_wait(x_free + (k % NX), ((k / NX) & 1u) ^ 1u);
        store_block<R>(sStage, nt * D, et, pf);
        if (norm_thread) {
          float lse2 = 3.0e38f;  // dead frames: gamma = 2^(x - huge) = 0
          if (fr < nt) {
            float m = -3.0e38f;
#pragma unroll
            for (int y = 0; y < PT; ++y)
              if (y < n_part) m = fmaxf(m, pp[y].x);
            for (int y = PT; y < n_part; ++y) m = fmaxf(m, a.partial[(size_t)y * a.total_frames + t0 + fr].x);
            float ssum = 0.f;
#pragma unroll
            for (int y = 0; y < PT; ++y)
              if (y < n_part) ssum += pp[y].y * ex2(pp[y].x - m);
            for (int y = PT; y < n_part; ++y) {
              const float2 p = a.partial[(size_t)y * a.total_frames + t0 + fr];
              ssum += p.y * ex2(p.x - m);
            }
            lse2 = m + lg2(ssum);
            if (tile == 0) {
              const float lse = lse2 * LN2;
              a.frame_lse[t0 + fr] = lse;
              ll_acc += lse;
            }
          }
          sLse[(k % NX) * BM + fr] = lse2;
        }
        named_bar_sync(1, EPI);
        {
          const bool live = fr < nt;
          const float* xr = sStage + fr * D;
          unsigned char* fb = smem + (size_t)(k & 1u) * 2 * cv.f_bytes;
          unsigned char* xb = smem + cv.o_x + (size_t)(k % NX) * 2 * cv.x_bytes;
          float4* dhi = reinterpret_cast<float4*>(fb) + fr;
          float4* dlo = reinterpret_cast<float4*>(fb + cv.f_bytes) + fr;
          float* xth = reinterpret_cast<float*>(xb) + ((size_t)(fr >> 2) * cv.xrows) * 4 + (fr & 3);
          float* xtl = reinterpret_cast<float*>(xb + cv.x_bytes) + ((size_t)(fr >> 2) * cv.xrows) * 4 + (fr & 3);
          for (int jc = part; jc < KC; jc += 8) {
            float h[4], l[4];
            chunk_split(xr, jc, D, live, h, l);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              xth[(4 * jc + e) * 4] = h[e];
              xtl[(4 * jc + e) * 4] = l[e];
            }
            dhi[jc * BM] = make_float4(h[0], h[1], h[2], h[3]);
            dlo[jc * BM] = make_float4(l[0], l[1], l[2], l[3]);
          }
          if (part == 0)
            for (int jx = KD; jx < cv.n2; ++jx) { xth[jx * 4] = 0.f; xtl[jx * 4] = 0.f; }
        }
        named_bar_sync(1, EPI);  // staging consumed before the next block's store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        mbar_arrive(a_full + (k & 1u));
        {
          Walk<BM> wn = w;
          bool fl;
          if (wn.next(nt, fl)) prefetch(wn);
        }
        // ---- gamma of the previous block while the tensor core works on this one
        if (k > 0) epilogue(k - 1, prev_seg, prev_flush);
        bool flush;
        more = w.next(nt, flush);
        prev_seg = seg_id;
        prev_flush = flush;
        ++k;
      }
      epilogue(k - 1, prev_seg, prev_flush);
      flush_ll();
      (void)n_drained;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}


// ================================================================================================ pass LSE
// The frame operand [x, x^2, 1, 1] (TF32 hi + lo) is written straight into
// tensor memory (tcgen05.st, thread == frame row == TMEM lane) and double-buffered there (2 x 160 columns), the model
// tile stays in shared memory as the N-side operand, one 128-column accumulator.  Thread program per block k: build
// A[k & 1] (needs GEMM(k - 2) done), then the epilogue of block k - 1 (tcgen05.ld, release the accumulator at once,
// max / exp / sum from registers); MMA program: GEMM(k) once A[k & 1] is built and the accumulator has been read.
// The split of block k + 1 overlaps GEMM(k); in steady state the period is ~1920 tensor cycles + one TMEM read.
namespace l2 {
constexpr uint32_t COL_A = 0;      // + 160 * (k & 1): hi at +0, lo at +80
constexpr uint32_t COL_ACC = 320;  // 128 columns
}

__global__ void __launch_bounds__(THREADS, 1) gmm_em_lse_kernel(const Args a) {
  using namespace l2;
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int BM = BM1;
  const int KD = a.KD, D = a.D;
  const uint32_t tile_bytes = (uint32_t)BN * KD * 4u;
  float* sBhi = reinterpret_cast<float*>(smem);
  float* sBlo = sBhi + BN * KD;
  float* sX = sBlo + BN * KD;  // feature staging: BM x D floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + BM * MAX_KD / 2);
  uint64_t* b_full = bars;
  uint64_t* a_full = bars + 1;    // [2] A[k & 1] is built
  uint64_t* l_full = bars + 3;    // [2] GEMM(k) complete: accumulator holds block k, A[k & 1] is free
  uint64_t* acc_free = bars + 5;  // the epilogue warps hold the accumulator's contents in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.y;
  if (tid == 0) {
    mbar_init(b_full, 1);
    mbar_init(a_full, EPI);
    mbar_init(a_full + 1, EPI);
    mbar_init(l_full, 1);
    mbar_init(l_full + 1, 1);
    mbar_init(acc_free, EPI / 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t begin = (int64_t)blockIdx.x * a.chunk;
  const int64_t end = min(begin + a.chunk, a.total_frames);

  if (warp == 0) {
    if (begin < end && elect_one()) {
      mbar_arrive_expect_tx(b_full, 2u * tile_bytes);
      bulk_g2s(sBhi, a.tiles_hi + (size_t)tile * BN * KD, tile_bytes, b_full);
      bulk_g2s(sBlo, a.tiles_lo + (size_t)tile * BN * KD, tile_bytes, b_full);
    }
    __syncwarp();
  } else if (warp == 1) {
    Walk<BM> w;
    if (w.start(a.seg, a.n_segs, begin, end)) {
      constexpr uint32_t lbo = BN * 16u, sbo = 128u;
      constexpr uint32_t kstep = (2u * lbo) >> 4;
      const uint64_t bhi = make_desc(smem_u32(sBhi), lbo, sbo), blo = make_desc(smem_u32(sBlo), lbo, sbo);
      const int ksteps = KD >> 3;
      const uint32_t idesc = make_idesc_tf32(BM, BN, 0, 0);
      const uint32_t t_acc = tmem_base + COL_ACC;
      mbar_wait(b_full, 0);
      uint32_t k = 0;
      bool more = true;
      while (more) {
        const int nt = w.nt();
        mbar_wait(a_full + (k & 1u), (k >> 1) & 1u);
        if (k > 0) mbar_wait(acc_free, (k - 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ahi = tmem_base + COL_A + 160u * (k & 1u), alo = ahi + 80u;
          for (int q = 0; q < ksteps; ++q) tc_mma_tf32_ts(t_acc, ahi + 8u * q, bhi + (uint64_t)(q * kstep), idesc, q > 0 ? 1u : 0u);
          for (int q = 0; q < ksteps; ++q) tc_mma_tf32_ts(t_acc, alo + 8u * q, bhi + (uint64_t)(q * kstep), idesc, 1u);
          for (int q = 0; q < ksteps; ++q) tc_mma_tf32_ts(t_acc, ahi + 8u * q, blo + (uint64_t)(q * kstep), idesc, 1u);
          tc_commit(l_full + (k & 1u));
        }
        __syncwarp();
        bool flush;
        more = w.next(nt, flush);
        ++k;
      }
    }
  } else {
    const int row = ((warp & 3) << 5) | lane;  // TMEM lane == frame row (built AND reduced by this thread's warp quadrant)
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int et = tid - 64;                   // 0..511
    const int part = (warp - 2) >> 2;          // 0..3: which K columns this thread builds; epilogue: parts 0, 1 take 64 columns each
    const bool epi_warp = part < 2;
    Walk<BM> w;
    if (w.start(a.seg, a.n_segs, begin, end)) {
      constexpr int R = (BM * (MAX_KD / 2 - 1) + EPI - 1) / EPI;  // D <= 39
      float pf[R];
      prefetch_block<R>(a.feats + w.t0 * D, w.nt() * D, et, pf);
      uint32_t k = 0;
      bool more = true;
      int prev_nt = 0;
      int64_t prev_t0 = 0;
      // block p: accumulator -> registers -> (max, sum 2^(x - max)) over this thread's 64 columns
      auto epilogue = [&](uint32_t p, int nt_p, int64_t t0_p) {
        mbar_wait(l_full + (p & 1u), (p >> 1) & 1u);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        const uint32_t taddr = tmem_base + lane_addr + COL_ACC + part * 64;
        tc_ld32_issue(taddr, ra);
        tc_ld32_issue(taddr + 32, rb);
        tc_ld_wait2(ra, rb);
        tc_fence_before();
        mbar_arrive(acc_free);
        float cm = max3(__uint_as_float(ra[0]), __uint_as_float(ra[1]), __uint_as_float(rb[0]));
#pragma unroll
        for (int e = 2; e < 32; e += 2) cm = max3(cm, __uint_as_float(ra[e]), __uint_as_float(ra[e + 1]));
#pragma unroll
        for (int e = 1; e < 31; e += 2) cm = max3(cm, __uint_as_float(rb[e]), __uint_as_float(rb[e + 1]));
        cm = fmaxf(cm, __uint_as_float(rb[31]));
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          acc0 += ex2(__uint_as_float(ra[e]) - cm);
          acc1 += ex2(__uint_as_float(ra[e + 1]) - cm);
          acc2 += ex2(__uint_as_float(rb[e]) - cm);
          acc3 += ex2(__uint_as_float(rb[e + 1]) - cm);
        }
        if (row < nt_p) a.partial[(size_t)(2 * tile + part) * a.total_frames + t0_p + row] = make_float2(cm, (acc0 + acc1) + (acc2 + acc3));
      };
      while (more) {
        const int nt = w.nt();
        const int64_t t0 = w.t0;
        // ---- build A[k & 1]: GEMM(k - 2), its last reader, must be complete
        if (k >= 2) {
          mbar_wait(l_full + (k & 1u), ((k - 2) >> 1) & 1u);
          tc_fence_after();
        }
        store_block<R>(sX, nt * D, et, pf);
        named_bar_sync(1, EPI);
        {
          Walk<BM> wn = w;
          bool fl;
          if (wn.next(nt, fl)) prefetch_block<R>(a.feats + wn.t0 * D, wn.nt() * D, et, pf);
        }
        {
          const bool live = row < nt;
          const float* xr = sX + row * D;
          const uint32_t t_hi = tmem_base + lane_addr + COL_A + 160u * (k & 1u), t_lo = t_hi + 80u;
          for (int c = part; c < (KD >> 3); c += 4) {  // 8 contraction columns at a time
            float h0[4], l0[4], h1[4], l1[4];
            chunk_split(xr, 2 * c, D, live, h0, l0);
            chunk_split(xr, 2 * c + 1, D, live, h1, l1);
            const float hv[8] = {h0[0], h0[1], h0[2], h0[3], h1[0], h1[1], h1[2], h1[3]};
            const float lv[8] = {l0[0], l0[1], l0[2], l0[3], l1[0], l1[1], l1[2], l1[3]};
            tc_st8(t_hi + 8u * c, hv);
            tc_st8(t_lo + 8u * c, lv);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        named_bar_sync(1, EPI);  // staged rows consumed before the next block overwrites them
        tc_fence_before();
        mbar_arrive(a_full + (k & 1u));
        // ---- epilogue of the previous block while the tensor core works on this one
        if (k > 0 && epi_warp) epilogue(k - 1, prev_nt, prev_t0);
        prev_nt = nt;
        prev_t0 = t0;
        bool flush;
        more = w.next(nt, flush);
        ++k;
      }
      if (epi_warp) epilogue(k - 1, prev_nt, prev_t0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace em

bool stats_tc_supported(const PackLayout& L) { return L.KD <= em::MAX_KD && L.off_tile_lo != 0 && L.n_models == 1; }

int64_t stats_tc_workspace_bytes(const PackLayout& L, int64_t total_frames) {
  return stats_tc_supported(L) ? (int64_t)sizeof(float2) * 2 * (L.Kp / em::BN) * total_frames : 0;
}

int launch_stats_tc(const float* feats, const int64_t* seg_offsets, int64_t n_segs, int64_t total_frames, const void* pack,
                    const PackLayout& L, float* frame_lse, double* out_n, double* out_f, double* out_s, double* out_loglik,
                    void* workspace, cudaStream_t st) {
  using namespace em;
  SSP_CUDA_OK(cudaMemsetAsync(out_loglik, 0, sizeof(double) * n_segs, st));
  if (total_frames == 0) return SSP_OK;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SSP_CUDA_OK(cudaGetDevice(&dev));
    SSP_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  Args a;
  a.feats = feats;
  a.seg = seg_offsets;
  a.n_segs = n_segs;
  a.total_frames = total_frames;
  a.n_tiles = L.Kp / BN;
  int64_t gx = num_sms / a.n_tiles;
  if (gx < 1) gx = 1;
  int64_t chunk = (total_frames + gx - 1) / gx;
  chunk = (chunk + BM1 - 1) / BM1 * BM1;
  gx = (total_frames + chunk - 1) / chunk;
  a.chunk = chunk;
  a.tiles_hi = (const float*)((const char*)pack + L.off_tile);
  a.tiles_lo = (const float*)((const char*)pack + L.off_tile_lo);
  a.K = L.K;
  a.D = L.D;
  a.KD = L.KD;
  a.partial = (float2*)workspace;
  a.frame_lse = frame_lse;
  a.out_n = out_n;
  a.out_f = out_f;
  a.out_s = out_s;
  a.out_loglik = out_loglik;
  const size_t smem_lse = (size_t)(2 * BN * L.KD + BM1 * MAX_KD / 2) * sizeof(float) + 128;
  const size_t smem_stats = p3::carve(L.KD).bytes;
  dim3 grid((unsigned)gx, (unsigned)a.n_tiles);
  SSP_CUDA_OK(cudaFuncSetAttribute(gmm_em_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lse));
  SSP_CUDA_OK(cudaFuncSetAttribute(gmm_em_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stats));
  gmm_em_lse_kernel<<<grid, THREADS, smem_lse, st>>>(a);
  SSP_LAUNCH_CHECK("gmm_em_lse_kernel");
  gmm_em_stats_kernel<<<grid, THREADS, smem_stats, st>>>(a);
  SSP_LAUNCH_CHECK("gmm_em_stats_kernel");
  return SSP_OK;
}

}  // namespace ssp
