// Back half of the librosa-convention MFCC (librosa.feature.mfcc as called at MFCC_DTW.py:27-30): see ssp_mel_db_post in
// include/ssp_b200.h.  The fused front-end kernel leaves 10 log10(max(amin, S)) per mel band; power_to_db then clips
// every value of the utterance to (utterance maximum - top_db), which no frame can do on its own, and the DCT follows.
// One CTA per utterance: block-wide maximum, then one thread per (frame, cepstrum) with the DCT rows in shared memory.
#include "common.cuh"

namespace ssp {

__global__ void __launch_bounds__(256) mel_db_dct_kernel(const float* __restrict__ mel_db, const int64_t* __restrict__ frame_offsets,
                                                         int nm, int nc, const float* __restrict__ dct, float top_db,
                                                         float* __restrict__ out) {
  extern __shared__ float s_dct[];  // nc * nm
  __shared__ float s_max[8];
  const int u = blockIdx.x;
  const int64_t f0 = frame_offsets[u];
  const int64_t T = frame_offsets[u + 1] - f0;
  if (T <= 0) return;
  for (int i = threadIdx.x; i < nc * nm; i += blockDim.x) s_dct[i] = dct[i];
  const float* x = mel_db + f0 * nm;
  const int64_t total = T * nm;
  float mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) mx = fmaxf(mx, x[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_max[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_max[w]);
  const float floor_db = top_db >= 0.f ? mx - top_db : -INFINITY;  // top_db < 0: no clipping (librosa top_db=None)
  for (int64_t idx = threadIdx.x; idx < T * nc; idx += blockDim.x) {
    const int64_t t = idx / nc;
    const int j = (int)(idx - t * nc);
    const float* row = s_dct + j * nm;
    const float* v = x + t * nm;
    float a0 = 0.f, a1 = 0.f;
    int m = 0;
    for (; m + 1 < nm; m += 2) {
      a0 = fmaf(row[m], fmaxf(v[m], floor_db), a0);
      a1 = fmaf(row[m + 1], fmaxf(v[m + 1], floor_db), a1);
    }
    if (m < nm) a0 = fmaf(row[m], fmaxf(v[m], floor_db), a0);
    out[(f0 + t) * nc + j] = a0 + a1;
  }
}

}  // namespace ssp

extern "C" int ssp_mel_db_post(const float* mel_db, const int64_t* frame_offsets, int64_t n_utts, int32_t n_mels, int32_t n_ceps,
                               const float* dct, float top_db, float* out_ceps, void* stream) {
  SSP_REQUIRE(frame_offsets && dct && out_ceps, "ssp_mel_db_post: null pointer");
  SSP_REQUIRE(n_mels >= 1 && n_ceps >= 1 && n_ceps <= n_mels && (int64_t)n_mels * n_ceps <= 8192,
              "ssp_mel_db_post: n_ceps %d x n_mels %d outside the supported range (n_ceps <= n_mels, product <= 8192)", n_ceps,
              n_mels);
  if (n_utts <= 0) return SSP_OK;
  SSP_REQUIRE(mel_db, "ssp_mel_db_post: null mel_db");
  SSP_REQUIRE(n_utts < (1ll << 31), "ssp_mel_db_post: bad n_utts");
  const size_t smem = sizeof(float) * (size_t)n_mels * n_ceps;
  ssp::mel_db_dct_kernel<<<(unsigned)n_utts, 256, smem, (cudaStream_t)stream>>>(mel_db, frame_offsets, n_mels, n_ceps, dct, top_db,
                                                                               out_ceps);
  SSP_LAUNCH_CHECK("mel_db_dct_kernel");
  return SSP_OK;
}
