// Fused front-end: raw PCM -> pre-emphasis -> framing -> window -> real FFT (half-size complex Stockham
// FFT in shared memory, one warp per frame) -> power/magnitude spectrum -> triangular filterbank -> log ->
// DCT -> cepstra kept in shared memory for the whole utterance -> delta / delta-delta (edge padded) ->
// per-utterance CMVN -> coalesced store.  One CTA per utterance; HBM traffic is exactly the PCM in and
// the features out.
//
// Conventions are data (tables + ssp_frontend_cfg), not code: the same kernel serves the sidekit
// recipe GMM_UBM.py:89 calls, python_speech_features' and utils/processing.py:110-144.
#include "frontend_common.cuh"

namespace ssp {

template <typename PcmT>
__global__ void __launch_bounds__(256) frontend_kernel(const FrontendArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ssp_frontend_cfg& cfg = a.cfg;
  const int nfft = cfg.nfft, NH = nfft >> 1, NC = cfg.n_ceps, NF = cfg.n_filt, FL = cfg.frame_len;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int OD = NC * (1 + cfg.delta_order);

  // ---- shared memory carve-up
  float2* tw_fft = reinterpret_cast<float2*>(smem_raw);          // e^{-2 pi i q / NH}, q < NH
  float2* tw_real = tw_fft + NH;                                 // e^{-2 pi i k / nfft}, k <= NH
  float* win = reinterpret_cast<float*>(tw_real + NH + 2);       // FL (+pad to even)
  float* red = win + ((FL + 3) & ~3);                            // 4*64*... reduction scratch: (blockDim/64)*64
  float* mean = red + blockDim.x;                                // 64
  float* istd = mean + 64;                                       // 64
  float* warp_base = istd + 64;
  const int per_warp = (4 * NH + (NH + 4) + ((NF + 3) & ~3) + 3) & ~3;  // bufA, bufB (float2 each), spectrum, mel; 16-byte multiple
  float* wb = warp_base + (size_t)warp * per_warp;
  float2* bufA = reinterpret_cast<float2*>(wb);
  float2* bufB = bufA + NH;
  float* pw = reinterpret_cast<float*>(bufB + NH);
  float* mel = pw + NH + 4;
  float* ceps = warp_base + (size_t)nwarps * per_warp;            // max_frames * NC

  const int u = blockIdx.x;
  const int64_t s_begin = a.sample_offsets[u];
  const int64_t n_samp = a.sample_offsets[u + 1] - s_begin;
  const int64_t f_begin = a.frame_offsets[u];
  const int T = (int)(a.frame_offsets[u + 1] - f_begin);
  if (T <= 0) return;

  // FFT lengths that are not a power of two (utils/processing.py:129 uses the frame length, e.g. 400) take a direct
  // DFT against one table e^{-2 pi i q / nfft}, q < nfft, laid over the two FFT tables (2 nfft <= 4 NH + 4 floats).
  // Even lengths whose half is 2^a 3^b 5^c (400 = 2 * 200) run the half-size complex FFT with radix-3 / radix-5 stages
  // next to the radix-4 / radix-2 ones; everything else (odd lengths, other prime factors) the direct DFT.
  const bool pow2 = (nfft & (nfft - 1)) == 0;
  bool smooth = (nfft & 1) == 0;
  {
    int m = NH;
    while ((m & 1) == 0) m >>= 1;
    while (m % 3 == 0) m /= 3;
    while (m % 5 == 0) m /= 5;
    smooth = smooth && m == 1;
  }
  if (smooth) {
    for (int q = threadIdx.x; q < NH; q += blockDim.x) {
      float s, c;
      sincospif(-2.0f * (float)q / (float)NH, &s, &c);
      tw_fft[q] = make_float2(c, s);
    }
    for (int k = threadIdx.x; k <= NH; k += blockDim.x) {
      float s, c;
      sincospif(-2.0f * (float)k / (float)nfft, &s, &c);
      tw_real[k] = make_float2(c, s);
    }
  } else {
    for (int q = threadIdx.x; q < nfft; q += blockDim.x) {
      double s, c;
      sincospi(-2.0 * (double)q / (double)nfft, &s, &c);
      tw_fft[q] = make_float2((float)c, (float)s);
    }
  }
  for (int i = threadIdx.x; i < FL; i += blockDim.x) win[i] = a.window[i];
  __syncthreads();

  const float pre = cfg.preemph;
  const int pmode = cfg.preemph_mode;
  const float LOG10_E = 0.43429448190325176f;

  // framing 3 / 4 (librosa center=True, MFCC_DTW.py:27-30): frame f is centred on sample f * shift; positions outside
  // the utterance are mirrored about the first / last sample (numpy 'reflect', any number of folds) or read as zero.
  const bool centred = cfg.framing >= 3;
  const int64_t refl_period = 2 * (n_samp - 1);
  for (int f = warp; f < T; f += nwarps) {
    const int64_t s0 = (int64_t)f * cfg.frame_shift - (centred ? (FL >> 1) : 0);
    // ---- load, pre-emphasis, time-domain energy, window; packed as NH complex values
    float energy = 0.f;
    float* zr = reinterpret_cast<float*>(bufA);
    for (int i = lane; i < nfft; i += 32) {
      float v = 0.f;
      if (i < FL) {
        int64_t gi = s0 + i;
        if (cfg.framing == 3 && (gi < 0 || gi >= n_samp)) {
          if (refl_period == 0) {
            gi = 0;
          } else {
            gi %= refl_period;
            if (gi < 0) gi += refl_period;
            if (gi >= n_samp) gi = refl_period - gi;
          }
        }
        float y = 0.f;
        if (gi >= 0 && gi < n_samp) {
          const float cur = load_pcm<PcmT>(a.pcm, s_begin + gi);
          if (pmode == 0) {
            y = cur;
          } else {
            float prev;
            if (i == 0) prev = (pmode == 1) ? cur : (gi > 0 ? load_pcm<PcmT>(a.pcm, s_begin + gi - 1) : 0.f);
            else prev = load_pcm<PcmT>(a.pcm, s_begin + gi - 1);
            y = fmaf(-pre, prev, cur);
          }
        }
        energy = fmaf(y, y, energy);
        v = y * win[i];
      }
      zr[i] = v;
    }
    energy = warp_sum(energy);
    __syncwarp();

    // ---- Stockham autosort FFT of NH complex points (radix 4, then radix 2 / 3 / 5 stages for what is left of NH)
    float2* src = bufA;
    float2* dst = bufB;
    for (int Ns = 1; smooth && Ns < NH;) {
      const int rem = NH / Ns;
      if ((rem & 3) == 0) {
        const int q4 = NH >> 2;
        const int tstep = NH / (Ns * 4);
        for (int j = lane; j < q4; j += 32) {
          const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
          float2 v0 = src[j], v1 = src[j + q4], v2 = src[j + 2 * q4], v3 = src[j + 3 * q4];
          if (Ns > 1) {
            v1 = cmul(v1, tw_fft[k * tstep]);
            v2 = cmul(v2, tw_fft[2 * k * tstep]);
            v3 = cmul(v3, tw_fft[3 * k * tstep]);
          }
          const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
          const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
          const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
          const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // (v1 - v3) * (-i)
          const int j0 = ((j - k) << 2) + k;                           // (j / Ns) * Ns * 4 + k
          dst[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
          dst[j0 + Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
          dst[j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
          dst[j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
        }
        Ns <<= 2;
      } else if ((rem & 1) == 0) {
        const int q2 = NH >> 1;
        const int tstep = NH / (Ns * 2);
        for (int j = lane; j < q2; j += 32) {
          const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
          float2 v0 = src[j], v1 = src[j + q2];
          if (Ns > 1) v1 = cmul(v1, tw_fft[k * tstep]);
          const int j0 = ((j - k) << 1) + k;
          dst[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
          dst[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        Ns <<= 1;
      } else if (rem % 3 == 0) {
        const int q3 = NH / 3;
        const int tstep = NH / (Ns * 3);
        const float S3 = 0.86602540378443865f;  // sin(2 pi / 3)
        for (int j = lane; j < q3; j += 32) {
          const int k = j % Ns;
          float2 v0 = src[j], v1 = src[j + q3], v2 = src[j + 2 * q3];
          if (Ns > 1) {
            v1 = cmul(v1, tw_fft[k * tstep]);
            v2 = cmul(v2, tw_fft[2 * k * tstep]);
          }
          const float2 t = make_float2(v1.x + v2.x, v1.y + v2.y);
          const float2 a = make_float2(fmaf(-0.5f, t.x, v0.x), fmaf(-0.5f, t.y, v0.y));
          const float2 b = make_float2(S3 * (v1.x - v2.x), S3 * (v1.y - v2.y));
          const int j0 = (j - k) * 3 + k;
          dst[j0] = make_float2(v0.x + t.x, v0.y + t.y);
          dst[j0 + Ns] = make_float2(a.x + b.y, a.y - b.x);      // a - i b
          dst[j0 + 2 * Ns] = make_float2(a.x - b.y, a.y + b.x);  // a + i b
        }
        Ns *= 3;
      } else {
        const int q5 = NH / 5;
        const int tstep = NH / (Ns * 5);
        const float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f;  // cos(2 pi / 5), cos(4 pi / 5)
        const float S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;   // sin(2 pi / 5), sin(4 pi / 5)
        for (int j = lane; j < q5; j += 32) {
          const int k = j % Ns;
          float2 v0 = src[j], v1 = src[j + q5], v2 = src[j + 2 * q5], v3 = src[j + 3 * q5], v4 = src[j + 4 * q5];
          if (Ns > 1) {
            v1 = cmul(v1, tw_fft[k * tstep]);
            v2 = cmul(v2, tw_fft[2 * k * tstep]);
            v3 = cmul(v3, tw_fft[3 * k * tstep]);
            v4 = cmul(v4, tw_fft[4 * k * tstep]);
          }
          const float2 t1 = make_float2(v1.x + v4.x, v1.y + v4.y), t2 = make_float2(v2.x + v3.x, v2.y + v3.y);
          const float2 t3 = make_float2(v1.x - v4.x, v1.y - v4.y), t4 = make_float2(v2.x - v3.x, v2.y - v3.y);
          const float2 a1 = make_float2(fmaf(C1, t1.x, fmaf(C2, t2.x, v0.x)), fmaf(C1, t1.y, fmaf(C2, t2.y, v0.y)));
          const float2 a2 = make_float2(fmaf(C2, t1.x, fmaf(C1, t2.x, v0.x)), fmaf(C2, t1.y, fmaf(C1, t2.y, v0.y)));
          const float2 b1 = make_float2(fmaf(S1, t3.x, S2 * t4.x), fmaf(S1, t3.y, S2 * t4.y));
          const float2 b2 = make_float2(fmaf(S2, t3.x, -S1 * t4.x), fmaf(S2, t3.y, -S1 * t4.y));
          const int j0 = (j - k) * 5 + k;
          dst[j0] = make_float2(v0.x + t1.x + t2.x, v0.y + t1.y + t2.y);
          dst[j0 + Ns] = make_float2(a1.x + b1.y, a1.y - b1.x);      // a1 - i b1
          dst[j0 + 2 * Ns] = make_float2(a2.x + b2.y, a2.y - b2.x);  // a2 - i b2
          dst[j0 + 3 * Ns] = make_float2(a2.x - b2.y, a2.y + b2.x);  // a2 + i b2
          dst[j0 + 4 * Ns] = make_float2(a1.x - b1.y, a1.y + b1.x);  // a1 + i b1
        }
        Ns *= 5;
      }
      __syncwarp();
      float2* tmp = src; src = dst; dst = tmp;
    }
    // src now holds Z = FFT_NH(z).  Real-input split: X[k] = Xe[k] + e^{-2 pi i k/nfft} Xo[k].
    float etot = 0.f;
    for (int k = lane; k <= NH; k += 32) {
      float re, im;
      if (smooth) {
        const float2 zk = src[k == NH ? 0 : k];
        const float2 zm = src[k == 0 ? 0 : NH - k];
        const float2 xe = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 xo = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        const float2 x = cmul(tw_real[k], xo);
        re = xe.x + x.x;
        im = xe.y + x.y;
      } else {
        // X[k] = sum_n z[n] e^{-2 pi i k n / nfft}; the table index k n mod nfft is carried exactly in integers and
        // two partial sums per component keep the FP32 accumulation error at the FFT path's level.
        float re0 = 0.f, im0 = 0.f, re1 = 0.f, im1 = 0.f;
        int idx = 0;
        int n = 0;
        for (; n + 1 < FL; n += 2) {
          const float2 w0 = tw_fft[idx];
          idx += k;
          if (idx >= nfft) idx -= nfft;
          const float2 w1 = tw_fft[idx];
          idx += k;
          if (idx >= nfft) idx -= nfft;
          const float v0 = zr[n], v1 = zr[n + 1];
          re0 = fmaf(v0, w0.x, re0);
          im0 = fmaf(v0, w0.y, im0);
          re1 = fmaf(v1, w1.x, re1);
          im1 = fmaf(v1, w1.y, im1);
        }
        if (n < FL) {
          const float2 w0 = tw_fft[idx];
          re0 = fmaf(zr[n], w0.x, re0);
          im0 = fmaf(zr[n], w0.y, im0);
        }
        re = re0 + re1;
        im = im0 + im1;
      }
      float p = fmaf(re, re, im * im);
      if (cfg.spec_type == 1) p = sqrtf(p);
      p *= cfg.spec_scale;
      pw[k] = p;
      etot += p;
    }
    if (cfg.energy_mode == 2) etot = warp_sum(etot);
    __syncwarp();
    // ---- triangular filterbank + log
    for (int m = lane; m < NF; m += 32) {
      const float* w = a.fb_weights + a.fb_offset[m];
      const float* p = pw + a.fb_start[m];
      const int len = a.fb_len[m];
      float acc = 0.f;
      for (int i = 0; i < len; ++i) acc = fmaf(__ldg(w + i), p[i], acc);
      if (cfg.log_zero_floor > 0.f && acc == 0.f) acc = cfg.log_zero_floor;
      acc += cfg.log_add;
      if (cfg.log_type == 3) {  // librosa power_to_db: 10 log10(max(amin, S)), ref = 1
        mel[m] = 10.f * log10f(fmaxf(acc, cfg.log_zero_floor));
      } else {
        mel[m] = cfg.log_type == 2 ? acc : (cfg.log_type == 1 ? logf(acc) * LOG10_E : logf(acc));
      }
    }
    __syncwarp();
    // ---- DCT (rows chosen by the host: c0 kept or dropped, lifter folded in)
    for (int j = lane; j < NC; j += 32) {
      const float* row = a.dct + j * NF;
      float acc = 0.f;
      for (int m = 0; m < NF; ++m) acc = fmaf(__ldg(row + m), mel[m], acc);
      if (cfg.energy_mode == 2 && j == 0) {
        float e = etot;
        if (cfg.log_zero_floor > 0.f && e == 0.f) e = cfg.log_zero_floor;
        acc = logf(e);
      }
      ceps[f * NC + j] = acc;
    }
    if (lane == 0 && a.out_log_energy && cfg.energy_mode == 1) a.out_log_energy[f_begin + f] = logf(energy);
    __syncwarp();
  }
  __syncthreads();

  // ---- delta / delta-delta / CMVN epilogue straight out of shared memory
  const int N = cfg.delta_n;
  float den = 0.f;
  for (int n = 1; n <= N; ++n) den += 2.f * n * n;
  const float inv_den = den > 0.f ? 1.f / den : 0.f;
  if (cfg.cmvn) {
    const int j = threadIdx.x & 63, g = threadIdx.x >> 6, G = blockDim.x >> 6;
    float part = 0.f;
    if (j < OD)
      for (int t = g; t < T; t += G) part += feat_at(ceps, NC, T, t, j, N, inv_den);
    red[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
      for (int gg = 0; gg < G; ++gg) s += red[gg * 64 + threadIdx.x];
      mean[threadIdx.x] = s / (float)T;
    }
    __syncthreads();
    part = 0.f;
    if (j < OD) {
      const float mu = mean[j];
      for (int t = g; t < T; t += G) {
        const float d = feat_at(ceps, NC, T, t, j, N, inv_den) - mu;
        part = fmaf(d, d, part);
      }
    }
    red[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
      for (int gg = 0; gg < G; ++gg) s += red[gg * 64 + threadIdx.x];
      float sd = sqrtf(s / (float)T);
      if (sd < 10.f * 1.1920929e-7f) sd = 1.f;  // sklearn/preprocessing/_data.py:127 (_handle_zeros_in_scale)
      istd[threadIdx.x] = 1.f / sd;
    }
    __syncthreads();
  }
  float* out = a.out_feats + f_begin * OD;
  const int total = T * OD;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int t = idx / OD, j = idx - t * OD;
    float v = feat_at(ceps, NC, T, t, j, N, inv_den);
    if (cfg.cmvn) v = (v - mean[j]) * istd[j];
    out[idx] = v;
  }
}

static int frontend_warps(int nfft) { return nfft <= 512 ? 8 : (nfft <= 1024 ? 4 : 2); }

static size_t frontend_smem(const ssp_frontend_cfg& c, int max_frames) {
  const int NH = c.nfft / 2, nw = frontend_warps(c.nfft);
  size_t floats = 2 * (size_t)NH + 2 * ((size_t)NH + 2) + ((c.frame_len + 3) & ~3) + nw * 32 + 128;
  floats += (size_t)nw * ((4 * NH + (NH + 4) + ((c.n_filt + 3) & ~3) + 3) & ~3);
  floats += (size_t)max_frames * c.n_ceps;
  return floats * sizeof(float);
}

static bool frontend_cfg_ok(const ssp_frontend_cfg* c) {
  if (!c) return false;
  if (c->nfft < 64 || c->nfft > 4096) return false;  // even with a 2^a 3^b 5^c half: FFT; anything else: direct DFT
  if (c->frame_len < 1 || c->frame_len > c->nfft || c->frame_shift < 1) return false;
  if (c->n_filt < 1 || c->n_filt > 256 || c->n_ceps < 1 || c->n_ceps > c->n_filt) return false;
  if (c->delta_order < 0 || c->delta_order > 2 || (c->delta_order > 0 && c->delta_n < 1)) return false;
  // 64 mean / inverse-std slots for the fused CMVN; without CMVN and deltas a frame may keep every filter output
  // (PLP critical bands, librosa's 128 log-mel bands) for a post kernel
  if (c->n_ceps * (1 + c->delta_order) > 64 && (c->cmvn || c->delta_order)) return false;
  if (c->framing < 0 || c->framing > 4 || c->preemph_mode < 0 || c->preemph_mode > 2) return false;
  if (c->framing >= 3 && c->preemph_mode != 0) return false;
  if (c->log_type < 0 || c->log_type > 3 || (c->log_type == 3 && !(c->log_zero_floor > 0.f))) return false;
  if (c->pcm_dtype < 0 || c->pcm_dtype > 1) return false;
  return true;
}

constexpr size_t kFrontendSmemMax = 226 * 1024;  // 227 KB is the sm_100 per-CTA limit

// -------------------------------------------------------------------------------- standalone delta / CMVN
__global__ void delta_kernel(const float* __restrict__ feat, int64_t T, int F, int N, float inv_den, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * F) return;
  const int64_t t = idx / F;
  const int j = (int)(idx - t * F);
  float acc = 0.f;
  for (int n = 1; n <= N; ++n) {
    const int64_t hi = min(t + n, T - 1), lo = max(t - (int64_t)n, (int64_t)0);
    acc = fmaf((float)n, feat[hi * F + j] - feat[lo * F + j], acc);
  }
  out[idx] = acc * inv_den;
}

// one CTA per utterance, two-pass mean / variance like sklearn.preprocessing.scale
__global__ void __launch_bounds__(256) cmvn_kernel(const float* __restrict__ feat, const int64_t* __restrict__ offsets, int F,
                                                   float* __restrict__ out) {
  __shared__ float red[256];
  __shared__ float mean[256], istd[256];
  const int64_t b = offsets[blockIdx.x];
  const int64_t T = offsets[blockIdx.x + 1] - b;
  if (T <= 0) return;
  const int lanes = F <= 256 ? (256 / F) : 0;  // frame groups per pass
  for (int j0 = 0; j0 < F; j0 += 256) {
    const int FJ = min(F - j0, 256);
    const int G = max(256 / FJ, 1);
    const int j = threadIdx.x % FJ, g = threadIdx.x / FJ;
    float part = 0.f;
    if (g < G)
      for (int64_t t = g; t < T; t += G) part += feat[(b + t) * F + j0 + j];
    red[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x < FJ) {
      float s = 0.f;
      for (int gg = 0; gg < G; ++gg) s += red[gg * FJ + threadIdx.x];
      mean[threadIdx.x] = s / (float)T;
    }
    __syncthreads();
    part = 0.f;
    if (g < G) {
      const float mu = mean[j];
      for (int64_t t = g; t < T; t += G) {
        const float d = feat[(b + t) * F + j0 + j] - mu;
        part = fmaf(d, d, part);
      }
    }
    red[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x < FJ) {
      float s = 0.f;
      for (int gg = 0; gg < G; ++gg) s += red[gg * FJ + threadIdx.x];
      float sd = sqrtf(s / (float)T);
      if (sd < 10.f * 1.1920929e-7f) sd = 1.f;
      istd[threadIdx.x] = 1.f / sd;
    }
    __syncthreads();
    if (g < G)
      for (int64_t t = g; t < T; t += G) {
        const int64_t i = (b + t) * F + j0 + j;
        out[i] = (feat[i] - mean[j]) * istd[j];
      }
    __syncthreads();
  }
  (void)lanes;
}

}  // namespace ssp

extern "C" int64_t ssp_frontend_num_frames(const ssp_frontend_cfg* c, int64_t n) {
  if (!ssp::frontend_cfg_ok(c) || n <= 0) return 0;
  const int64_t len = c->frame_len, hop = c->frame_shift;
  switch (c->framing) {
    case 0: return n < len ? 0 : (n - len) / hop + 1;
    case 1: return n <= len ? 1 : 1 + (n - len + hop - 1) / hop;
    case 2: return (n + hop - 1) / hop;
    default: return 1 + n / hop;  // centred: the signal is padded by len / 2 on both sides
  }
}

extern "C" int64_t ssp_frontend_max_frames(const ssp_frontend_cfg* c) {
  if (!ssp::frontend_cfg_ok(c)) return 0;
  const size_t fixed = ssp::frontend_fast_supported(*c) ? ssp::frontend_fast_smem(*c, 0, false) : ssp::frontend_smem(*c, 0);
  if (fixed >= ssp::kFrontendSmemMax) return 0;
  return (int64_t)((ssp::kFrontendSmemMax - fixed) / (sizeof(float) * c->n_ceps));
}

extern "C" int ssp_frontend_batch(const void* pcm, const int64_t* sample_offsets, int64_t n_utts,
                                  const ssp_frontend_cfg* cfg, const float* window, const int32_t* fb_start,
                                  const int32_t* fb_len, const int32_t* fb_offset, const float* fb_weights,
                                  const float* dct, const int64_t* frame_offsets, int64_t max_frames_per_utt,
                                  float* out_feats, float* out_log_energy, void* stream) {
  using namespace ssp;
  SSP_REQUIRE(frontend_cfg_ok(cfg), "ssp_frontend_batch: unsupported front-end configuration");
  SSP_REQUIRE(pcm && sample_offsets && window && fb_start && fb_len && fb_offset && fb_weights && dct && frame_offsets &&
                  out_feats,
              "ssp_frontend_batch: null pointer");
  SSP_REQUIRE(n_utts >= 0 && n_utts < (1ll << 31), "ssp_frontend_batch: bad n_utts");
  if (n_utts == 0) return SSP_OK;
  if (max_frames_per_utt > ssp_frontend_max_frames(cfg)) {
    set_error("ssp_frontend_batch: utterance of %lld frames exceeds the fused kernel's %lld-frame shared-memory bound; "
              "split the utterance", (long long)max_frames_per_utt, (long long)ssp_frontend_max_frames(cfg));
    return SSP_EUNSUP;
  }
  FrontendArgs a;
  a.cfg = *cfg;
  a.pcm = pcm;
  a.sample_offsets = sample_offsets;
  a.window = window;
  a.fb_start = fb_start;
  a.fb_len = fb_len;
  a.fb_offset = fb_offset;
  a.fb_weights = fb_weights;
  a.dct = dct;
  a.frame_offsets = frame_offsets;
  a.out_feats = out_feats;
  a.out_log_energy = out_log_energy;
  a.max_frames = (int)max_frames_per_utt;
  if (frontend_fast_supported(*cfg)) return launch_frontend_fast(a, n_utts, nullptr, (cudaStream_t)stream);
  const size_t smem = frontend_smem(*cfg, (int)max_frames_per_utt);
  const int threads = 32 * frontend_warps(cfg->nfft);
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->pcm_dtype == 0) {
    SSP_CUDA_OK(cudaFuncSetAttribute(frontend_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    frontend_kernel<int16_t><<<(unsigned)n_utts, threads, smem, st>>>(a);
  } else {
    SSP_CUDA_OK(cudaFuncSetAttribute(frontend_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    frontend_kernel<float><<<(unsigned)n_utts, threads, smem, st>>>(a);
  }
  SSP_LAUNCH_CHECK("frontend_kernel");
  return SSP_OK;
}

extern "C" int ssp_delta(const float* feat, int64_t n_frames, int32_t n_feat, int32_t delta_n, float* out, void* stream) {
  SSP_REQUIRE(feat && out, "ssp_delta: null pointer");
  SSP_REQUIRE(delta_n >= 1, "N must be an integer >= 1");  // GMM_UBM.py:59-60
  SSP_REQUIRE(n_frames >= 0 && n_feat >= 1, "ssp_delta: bad shape");
  if (n_frames == 0) return SSP_OK;
  float den = 0.f;
  for (int n = 1; n <= delta_n; ++n) den += 2.f * n * n;
  const int64_t total = n_frames * n_feat;
  ssp::delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(feat, n_frames, n_feat, delta_n,
                                                                                      1.f / den, out);
  SSP_LAUNCH_CHECK("delta_kernel");
  return SSP_OK;
}

extern "C" int ssp_cmvn(const float* feat, const int64_t* frame_offsets, int64_t n_utts, int32_t n_feat, float* out,
                        void* stream) {
  SSP_REQUIRE(feat && out && frame_offsets, "ssp_cmvn: null pointer");
  SSP_REQUIRE(n_utts >= 0 && n_feat >= 1, "ssp_cmvn: bad shape");
  if (n_utts == 0) return SSP_OK;
  ssp::cmvn_kernel<<<(unsigned)n_utts, 256, 0, (cudaStream_t)stream>>>(feat, frame_offsets, n_feat, out);
  SSP_LAUNCH_CHECK("cmvn_kernel");
  return SSP_OK;
}
