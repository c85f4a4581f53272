// Pieces shared by the generic and the nfft = 512 fast front-end kernels.
#pragma once
#include "common.cuh"

namespace ssp {

struct FrontendArgs {
  ssp_frontend_cfg cfg;
  const void* pcm;
  const int64_t* sample_offsets;
  const float* window;
  const int32_t* fb_start;
  const int32_t* fb_len;
  const int32_t* fb_offset;
  const float* fb_weights;
  const float* dct;
  const int64_t* frame_offsets;
  float* out_feats;
  float* out_log_energy;
  int max_frames;  // shared-memory rows reserved for cepstra
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

template <typename PcmT>
__device__ __forceinline__ float load_pcm(const void* p, int64_t i) {
  return (float)reinterpret_cast<const PcmT*>(p)[i];
}

// value of output feature j (0..OD-1) at frame t from the cepstra kept in shared memory;
// GMM_UBM.py:53-69 (delta, edge padding) applied once or twice.
__device__ __forceinline__ float delta_at(const float* __restrict__ ceps, int NC, int T, int t, int jj, int N, float inv_den) {
  float acc = 0.f;
  for (int n = 1; n <= N; ++n) {
    const int hi = min(t + n, T - 1), lo = max(t - n, 0);
    acc = fmaf((float)n, ceps[hi * NC + jj] - ceps[lo * NC + jj], acc);
  }
  return acc * inv_den;
}
__device__ __forceinline__ float feat_at(const float* __restrict__ ceps, int NC, int T, int t, int j, int N, float inv_den) {
  const int order = j / NC, jj = j - order * NC;
  if (order == 0) return ceps[t * NC + jj];
  if (order == 1) return delta_at(ceps, NC, T, t, jj, N, inv_den);
  float acc = 0.f;
  for (int n = 1; n <= N; ++n) {
    const int hi = min(t + n, T - 1), lo = max(t - n, 0);
    acc = fmaf((float)n, delta_at(ceps, NC, T, hi, jj, N, inv_den) - delta_at(ceps, NC, T, lo, jj, N, inv_den), acc);
  }
  return acc * inv_den;
}


int launch_frontend_fast(const FrontendArgs& a, int64_t n_utts, size_t* smem_out, cudaStream_t st);
bool frontend_fast_supported(const ssp_frontend_cfg& c);
size_t frontend_fast_smem(const ssp_frontend_cfg& c, int max_frames, bool materialize_deltas);

}  // namespace ssp
