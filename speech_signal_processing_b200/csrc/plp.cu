// Back half of the PLP front-end (sidekit.frontend.features.plp, GMM_UBM.py:20,94-99; rastamat rastaplp): see
// ssp_plp_post in include/ssp_b200.h.  One CTA per utterance.  Phase 1: one thread per critical band runs the RASTA
// filter (FIR [0.2 0.1 0 -0.1 -0.2], pole 0.94, direct form II transposed; the first four frames only load the FIR
// state and output zero) over the log energies, in place.  Phase 2: one thread per frame does the rest in double
// precision: equal loudness, ^0.33, band replication, autocorrelation by a small matrix, Levinson-Durbin, the
// LPC -> cepstrum recursion and the lifter.
#include "common.cuh"

namespace ssp {
namespace plp {

constexpr int MAXB = 32;   // critical bands
constexpr int MAXC = 16;   // cepstra (model order + 1)

__global__ void __launch_bounds__(128) plp_post_kernel(float* __restrict__ bands, const int64_t* __restrict__ frame_offsets,
                                                       int nb, int nc, const double* __restrict__ eql,
                                                       const double* __restrict__ idft, const double* __restrict__ lift,
                                                       int rasta, float* __restrict__ out) {
  __shared__ double s_eql[MAXB], s_idft[MAXC * MAXB], s_lift[MAXC];
  const int u = blockIdx.x;
  const int64_t f0 = frame_offsets[u];
  const int T = (int)(frame_offsets[u + 1] - f0);
  if (T <= 0) return;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s_eql[i] = eql[i];
  for (int i = threadIdx.x; i < nc * nb; i += blockDim.x) s_idft[i] = idft[i];
  for (int i = threadIdx.x; i < nc; i += blockDim.x) s_lift[i] = lift[i];
  float* x = bands + f0 * nb;
  if (rasta && threadIdx.x < nb) {
    const int b = threadIdx.x;
    const double b0 = 0.2, b1 = 0.1, b2 = 0.0, b3 = -0.1, b4 = -0.2, a1 = -0.94;
    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
    for (int t = 0; t < T; ++t) {
      const double v = log((double)x[(int64_t)t * nb + b]);
      const bool iir = t >= 4;
      const double y = b0 * v + z0;
      z0 = b1 * v + z1 - (iir ? a1 * y : 0.0);
      z1 = b2 * v + z2;
      z2 = b3 * v + z3;
      z3 = b4 * v;
      x[(int64_t)t * nb + b] = (float)exp(iir ? y : 0.0);
    }
  }
  __syncthreads();
  const int order = nc - 1;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double post[MAXB];
    for (int i = 0; i < nb; ++i) post[i] = pow(s_eql[i] * (double)x[(int64_t)t * nb + i], 0.33);
    post[0] = post[1];
    post[nb - 1] = post[nb - 2];
    double r[MAXC], a[MAXC], prev[MAXC];
    for (int k = 0; k < nc; ++k) {
      double acc = 0.0;
      for (int i = 0; i < nb; ++i) acc = fma(s_idft[k * nb + i], post[i], acc);
      r[k] = acc;
    }
    a[0] = 1.0;
    double e = r[0];
    for (int i = 1; i <= order; ++i) {
      double acc = r[i];
      for (int j = 1; j < i; ++j) acc += a[j] * r[i - j];
      const double k = -acc / e;
      for (int j = 1; j < i; ++j) prev[j] = a[j];
      for (int j = 1; j < i; ++j) a[j] = prev[j] + k * prev[i - j];
      a[i] = k;
      e *= (1.0 - k * k);
    }
    // lpc = a / e; cep[0] = -log(lpc[0]) = log(e); normalised coefficients = a
    double cep[MAXC];
    cep[0] = log(e);
    for (int n = 1; n < nc; ++n) {
      double s = 0.0;
      for (int m = 1; m < n; ++m) s += (double)(n - m) * a[m] * cep[n - m];
      cep[n] = -(a[n] + s / (double)n);
    }
    float* o = out + (f0 + t) * nc;
    for (int n = 0; n < nc; ++n) o[n] = (float)(cep[n] * s_lift[n]);
  }
}

}  // namespace plp
}  // namespace ssp

extern "C" int ssp_plp_post(float* bands, const int64_t* frame_offsets, int64_t n_utts, int32_t n_bands, int32_t n_ceps,
                            const double* eql, const double* idft, const double* lift, int32_t rasta, float* out_ceps,
                            void* stream) {
  SSP_REQUIRE(frame_offsets && eql && idft && lift && out_ceps, "ssp_plp_post: null pointer");
  SSP_REQUIRE(n_bands >= 3 && n_bands <= ssp::plp::MAXB, "ssp_plp_post: n_bands %d outside [3, %d]", n_bands, ssp::plp::MAXB);
  SSP_REQUIRE(n_ceps >= 2 && n_ceps <= ssp::plp::MAXC && n_ceps <= n_bands, "ssp_plp_post: n_ceps %d outside [2, min(%d, n_bands)]",
              n_ceps, ssp::plp::MAXC);
  if (n_utts <= 0) return SSP_OK;
  SSP_REQUIRE(bands, "ssp_plp_post: null bands");
  ssp::plp::plp_post_kernel<<<(unsigned)n_utts, 128, 0, (cudaStream_t)stream>>>(bands, frame_offsets, n_bands, n_ceps, eql, idft,
                                                                               lift, rasta, out_ceps);
  SSP_LAUNCH_CHECK("plp_post_kernel");
  return SSP_OK;
}
