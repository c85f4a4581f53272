// Voice-activity detection front filter (reference VAD.py) on the GPU.
//
//   vad_features_kernel : one CTA per utterance.  Peak of the utterance (VAD.py:133 normalises by max |x|), then one
//                         warp per frame of 256 samples (hop 128, ceil(N / hop) frames, zero-padded tail, no window --
//                         VAD.py:28-49): zero-crossing count = number of strictly negative neighbour products
//                         (VAD.py:52-63), energy = sum x^2 in double (VAD.py:66-76), spectral entropy over ten 12-bin
//                         sub-bands of |FFT|^2 of the first 128 bins (VAD.py:79-108) from a 256-point radix-4 Stockham
//                         FFT in shared memory.
//   vad_detect_kernel   : one thread per utterance walks the double-threshold state machine of VAD.py:137-182 over its
//                         frames, quirks included (a voiced run is only closed by a later quiet frame; the backward
//                         search uses Python's wrapped negative indices; runs are never merged because `last_end` is
//                         only assigned inside the merge branch).
#include "common.cuh"

namespace ssp {
namespace vad {

constexpr int FL = 256;     // frame length == FFT length (VAD.py:22)
constexpr int WARPS = 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

struct FeatArgs {
  const void* pcm;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  int frame_shift, n_blocks, normalize_peak;
  float eps;
  float* out_zcr;
  double* out_power;
  float* out_entropy;
};

template <typename PcmT>
__global__ void __launch_bounds__(WARPS * 32) vad_features_kernel(const FeatArgs a) {
  __shared__ float2 tw[FL];                 // e^{-2 pi i q / 256}
  __shared__ float2 buf[WARPS][2][FL];
  __shared__ float red[WARPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u = blockIdx.x;
  const int64_t s_begin = a.sample_offsets[u];
  const int64_t n_samp = a.sample_offsets[u + 1] - s_begin;
  const int64_t f_begin = a.frame_offsets[u];
  const int T = (int)(a.frame_offsets[u + 1] - f_begin);
  if (T <= 0) return;
  const PcmT* x = reinterpret_cast<const PcmT*>(a.pcm) + s_begin;

  for (int q = tid; q < FL; q += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)q / (float)FL, &s, &c);
    tw[q] = make_float2(c, s);
  }
  // ---- peak of the utterance
  double inv_peak = 1.0;
  if (a.normalize_peak) {
    float m = 0.f;
    for (int64_t i = tid; i < n_samp; i += blockDim.x) m = fmaxf(m, fabsf((float)x[i]));
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < WARPS; ++w) m = fmaxf(m, red[w]);
    inv_peak = 1.0 / (double)m;  // an all-zero utterance divides by zero in the reference too (NaN features)
  }
  __syncthreads();

  const int n_blocks = a.n_blocks, sub = (FL / 2) / n_blocks;
  for (int f = warp; f < T; f += WARPS) {
    const int64_t s0 = (int64_t)f * a.frame_shift;
    // lane owns samples [8 lane, 8 lane + 8) of the frame
    float v[9];
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t g = s0 + 8 * lane + i;
      const double xd = g < n_samp ? (double)x[g] * inv_peak : 0.0;
      e = fma(xd, xd, e);
      v[i] = (float)xd;
    }
    v[8] = __shfl_down_sync(0xffffffffu, v[0], 1);
    int z = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((i < 7 || lane < 31) && v[i] * v[i + 1] < 0.f) ++z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      z += __shfl_xor_sync(0xffffffffu, z, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    // ---- 256-point FFT of the (real) frame: four radix-4 Stockham passes
    float2* src = buf[warp][0];
    float2* dst = buf[warp][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) src[8 * lane + i] = make_float2(v[i], 0.f);
    __syncwarp();
    for (int Ns = 1; Ns < FL; Ns <<= 2) {
      const int q4 = FL >> 2, tstep = FL / (Ns * 4);
      for (int j = lane; j < q4; j += 32) {
        const int k = j & (Ns - 1);
        float2 v0 = src[j], v1 = src[j + q4], v2 = src[j + 2 * q4], v3 = src[j + 3 * q4];
        if (Ns > 1) {
          v1 = cmul(v1, tw[k * tstep]);
          v2 = cmul(v2, tw[2 * k * tstep]);
          v3 = cmul(v3, tw[3 * k * tstep]);
        }
        const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y), t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y), t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));
        const int j0 = ((j - k) << 2) + k;
        dst[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
        dst[j0 + Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
        dst[j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
        dst[j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
      }
      __syncwarp();
      float2* t = src; src = dst; dst = t;
    }
    // ---- spectral entropy: lane b < n_blocks sums its band; the total runs over all 128 bins
    float tot = 0.f;
    for (int k = lane; k < FL / 2; k += 32) tot = fmaf(src[k].x, src[k].x, fmaf(src[k].y, src[k].y, tot));
    tot = warp_sum(tot);
    float s_b = 0.f;
    if (lane < n_blocks) {
      float acc = 0.f;
      for (int k = lane * sub; k < (lane + 1) * sub; ++k) acc = fmaf(src[k].x, src[k].x, fmaf(src[k].y, src[k].y, acc));
      s_b = acc / (tot + a.eps);
      s_b = -s_b * log2f(s_b + a.eps);
    }
    const float ent = warp_sum(s_b);
    if (lane == 0) {
      a.out_zcr[f_begin + f] = (float)z;
      a.out_power[f_begin + f] = e;
      a.out_entropy[f_begin + f] = ent;
    }
    __syncwarp();
  }
}

// Python indexing of a length-n array with an index in [-n, n): negative indices wrap (VAD.py:161 relies on it)
__device__ __forceinline__ int64_t py_index(int64_t n, int64_t i) { return i < 0 ? i + n : i; }

__global__ void vad_detect_kernel(const float* __restrict__ zcr, const double* __restrict__ power,
                                  const int64_t* __restrict__ frame_offsets, int64_t n_utts, float zcr_gate, double ampl,
                                  double amph, int min_len, unsigned char* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_utts) return;
  const int64_t f0 = frame_offsets[u], n = frame_offsets[u + 1] - f0;
  const float* z = zcr + f0;
  const double* p = power + f0;
  unsigned char* res = out + f0;
  for (int64_t i = 0; i < n; ++i) res[i] = 0;
  int status = 0;
  int64_t start = 0, end = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (p[i] > amph) {
      if (status != 1) start = i;
      end = i;
      status = 1;
    } else if (end - start + 1 > min_len) {
      // the reference raises IndexError below -n; stop there instead
      while (start >= -n && (p[py_index(n, start)] > ampl || z[py_index(n, start)] > zcr_gate)) --start;
      ++start;
      while (p[end] > ampl || z[end] > zcr_gate) {
        ++end;
        if (end == n) break;
      }
      --end;
      // res[start : end + 1] = 1 with numpy slice semantics for a negative start
      const int64_t lo = start < 0 ? (start + n > 0 ? start + n : 0) : start;
      for (int64_t j = lo; j <= end && j < n; ++j) res[j] = 1;
      start = 0;
      end = 0;
      status = 0;
    }
  }
}

}  // namespace vad
}  // namespace ssp

extern "C" int64_t ssp_vad_num_frames(const ssp_vad_cfg* cfg, int64_t n_samples) {
  if (!cfg || cfg->frame_shift < 1 || n_samples <= 0) return 0;
  return (n_samples + cfg->frame_shift - 1) / cfg->frame_shift;  // VAD.py:36
}

extern "C" int ssp_vad_features(const void* pcm, const int64_t* sample_offsets, int64_t n_utts, const ssp_vad_cfg* cfg,
                                const int64_t* frame_offsets, float* out_zcr, double* out_power, float* out_entropy,
                                void* stream) {
  using namespace ssp::vad;
  SSP_REQUIRE(cfg && sample_offsets && frame_offsets && out_zcr && out_power && out_entropy, "ssp_vad_features: null pointer");
  SSP_REQUIRE(cfg->frame_len == FL, "ssp_vad_features: frame_len must be %d (VAD.py:22), got %d", FL, cfg->frame_len);
  SSP_REQUIRE(cfg->frame_shift >= 1 && cfg->frame_shift <= FL, "ssp_vad_features: frame_shift %d outside [1, %d]", cfg->frame_shift, FL);
  SSP_REQUIRE(cfg->n_blocks >= 1 && cfg->n_blocks <= 32, "ssp_vad_features: n_blocks %d outside [1, 32]", cfg->n_blocks);
  SSP_REQUIRE(cfg->pcm_dtype == 0 || cfg->pcm_dtype == 1, "ssp_vad_features: pcm_dtype must be 0 (int16) or 1 (float32)");
  if (n_utts <= 0) return SSP_OK;
  SSP_REQUIRE(pcm, "ssp_vad_features: null pcm");
  FeatArgs a;
  a.pcm = pcm;
  a.sample_offsets = sample_offsets;
  a.frame_offsets = frame_offsets;
  a.frame_shift = cfg->frame_shift;
  a.n_blocks = cfg->n_blocks;
  a.normalize_peak = cfg->normalize_peak;
  a.eps = cfg->eps;
  a.out_zcr = out_zcr;
  a.out_power = out_power;
  a.out_entropy = out_entropy;
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->pcm_dtype == 0) vad_features_kernel<int16_t><<<(unsigned)n_utts, WARPS * 32, 0, st>>>(a);
  else vad_features_kernel<float><<<(unsigned)n_utts, WARPS * 32, 0, st>>>(a);
  SSP_LAUNCH_CHECK("vad_features_kernel");
  return SSP_OK;
}

extern "C" int ssp_vad_detect(const float* zcr, const double* power, const int64_t* frame_offsets, int64_t n_utts,
                              float zcr_gate, double ampl, double amph, int32_t min_len, uint8_t* out_speech, void* stream) {
  SSP_REQUIRE(zcr && power && frame_offsets && out_speech, "ssp_vad_detect: null pointer");
  if (n_utts <= 0) return SSP_OK;
  const int threads = 64;
  ssp::vad::vad_detect_kernel<<<(unsigned)((n_utts + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      zcr, power, frame_offsets, n_utts, zcr_gate, ampl, amph, min_len, out_speech);
  SSP_LAUNCH_CHECK("vad_detect_kernel");
  return SSP_OK;
}
