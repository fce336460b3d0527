// Post-processing of the decoded waveform on the GPU: the steps AFTER the hot path that
// the reference runs with torchaudio and CPU pydub (tts_service.py:251-256,287-306; watermarking.py:35-39):
// sinc resampling around the watermarker, peak normalisation, 16-bit PCM conversion, silence padding, fades.
// All of it is HBM-bound element-wise / short-FIR work: coalesced loads, one pass per step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace post {

// torchaudio.functional.resample, default arguments (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99):
// kernel[p][j] for output phase p < nf and tap j < 2 width + of, evaluated in fp64 and rounded to fp32 like
// torchaudio's ``_get_sinc_resample_kernel`` (of / nf = the rates divided by their gcd).
__global__ void k_resample_table(int of, int nf, int width, double rolloff, int lpw, float* __restrict__ tab) {
  const int taps = 2 * width + of;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf * taps) return;
  const int p = i / taps, j = i - p * taps;
  const double base = (double)(of < nf ? of : nf) * rolloff;
  double t = ((double)(-p) / (double)nf + (double)(j - width) / (double)of) * base;
  t = t < -(double)lpw ? -(double)lpw : (t > (double)lpw ? (double)lpw : t);
  const double PI = 3.14159265358979323846;
  const double c = cos(t * PI / (double)lpw / 2.0);
  const double window = c * c;
  t *= PI;
  const double scale = base / (double)of;
  const double sinc = t == 0.0 ? 1.0 : sin(t) / t;
  tab[i] = (float)(sinc * window * scale);
}

// y[n * nf + p] = sum_j tab[p][j] * xpad[n * of + j], xpad = x with ``width`` zeros in front (and zeros behind):
// torchaudio's conv1d(stride = of) over the padded waveform, transposed and trimmed to ``n_out`` samples.
// One CTA per block of input strides; the x window of the block is staged in shared memory.
constexpr int RS_STRIDES = 8;  // input strides (n values) per CTA
__global__ void __launch_bounds__(256) k_resample(const float* __restrict__ x, long long n_in, const float* __restrict__ tab, int of,
                                                  int nf, int width, float* __restrict__ y, long long n_out) {
  extern __shared__ float xs[];  // [(RS_STRIDES - 1) * of + taps]
  const int taps = 2 * width + of;
  const long long n0 = (long long)blockIdx.x * RS_STRIDES;
  const int span = (RS_STRIDES - 1) * of + taps;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const long long src = n0 * of + i - width;
    xs[i] = (src >= 0 && src < n_in) ? x[src] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < RS_STRIDES * nf; o += blockDim.x) {
    const int dn = o / nf, p = o - dn * nf;
    const long long out = (n0 + dn) * nf + p;
    if (out >= n_out) continue;
    const float* tp = tab + (size_t)p * taps;
    const float* xp = xs + dn * of;
    float acc = 0.f;
    for (int j = 0; j < taps; ++j) acc = fmaf(tp[j], xp[j], acc);
    y[out] = acc;
  }
}

// max |x| (fp32) into *out via ordered-int atomicMax (|x| >= 0, so the bit pattern orders like the value)
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, long long n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    atomicMax(out, __float_as_uint(m));
  }
}

// tts_service.generate_audio_segment (tts_service.py:287-306) on the device:
//   a = audio / max(|audio|.max(), 1e-6)   (fp32);  pcm = int16(a * 32767)   (numpy astype: truncation);
//   [s0 zeros] + pcm + [s1 zeros];  pydub fade_in / fade_out of f_in / f_out samples (precise per-sample form,
//   fades <= 100 ms): sample * (from + step * i) in double, floor, clipped -- audioop.mul.
__global__ void __launch_bounds__(256) k_pcm16_segment(const float* __restrict__ audio, long long n, const unsigned int* __restrict__ absmax,
                                                       long long s0, long long s1, long long f_in, long long f_out,
                                                       int16_t* __restrict__ out) {
  const long long total = s0 + n + s1;
  const float peak = fmaxf(__uint_as_float(*absmax), 1e-6f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int v = 0;
    if (i >= s0 && i < s0 + n) {
      const float a = __fdiv_rn(audio[i - s0], peak);
      v = (int)(__fmul_rn(a, 32767.f));  // C cast = truncation toward zero, like ndarray.astype("int16")
    }
    const double lo = 1e-6;  // db_to_float(-120)
    if (f_in > 0 && i < f_in) {
      const double g = lo + ((1.0 - lo) / (double)f_in) * (double)i;
      double f = floor((double)v * g);
      v = (int)(f < -32768.0 ? -32768.0 : (f > 32767.0 ? 32767.0 : f));
    }
    if (f_out > 0 && i >= total - f_out) {
      const double g = 1.0 + ((lo - 1.0) / (double)f_out) * (double)(i - (total - f_out));
      double f = floor((double)v * g);
      v = (int)(f < -32768.0 ? -32768.0 : (f > 32767.0 ? 32767.0 : f));
    }
    out[i] = (int16_t)v;
  }
}

}  // namespace post
