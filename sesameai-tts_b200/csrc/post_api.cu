// C ABI of the waveform post-processing (include/csm_b200.h, csm_post_* entry points).
#include <stdio.h>
#include <stdint.h>

#include "../../include/csm_b200.h"
#include "nvtx_ranges.h"
#include "post_kernels.cuh"

int csm_set_error(int code, const char* msg);  // api.cu
void csm_count_launches(unsigned long long n);

namespace {
long long gcd_ll(long long a, long long b) {
  while (b) {
    long long t = a % b;
    a = b;
    b = t;
  }
  return a;
}
struct Rs {
  int of, nf, width, taps;
};
bool rs_params(int orig, int neu, Rs* r) {
  if (orig < 1 || neu < 1) return false;
  const long long g = gcd_ll(orig, neu);
  r->of = (int)(orig / g);
  r->nf = (int)(neu / g);
  const double base = (double)(r->of < r->nf ? r->of : r->nf) * 0.99;
  const double w = 6.0 * (double)r->of / base;
  r->width = (int)w;
  if ((double)r->width < w) ++r->width;  // ceil
  r->taps = 2 * r->width + r->of;
  return r->of <= 4096 && r->nf <= 4096;
}
}  // namespace

extern "C" int64_t csm_post_resample_len(int64_t n, int32_t orig_freq, int32_t new_freq) {
  Rs r;
  if (n < 0 || !rs_params(orig_freq, new_freq, &r)) return -1;
  return ((long long)r.nf * n + r.of - 1) / r.of;  // ceil(new * n / orig)
}
extern "C" size_t csm_post_resample_workspace_bytes(int32_t orig_freq, int32_t new_freq) {
  Rs r;
  if (!rs_params(orig_freq, new_freq, &r)) return 0;
  return (size_t)r.nf * r.taps * sizeof(float);
}

extern "C" int32_t csm_post_resample(const float* x, int64_t n, int32_t orig_freq, int32_t new_freq, float* y, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  Rs r;
  if (!x || !y || n < 1 || !rs_params(orig_freq, new_freq, &r)) return csm_set_error(CSM_ERR_ARG, "csm_post_resample: bad arguments");
  NvtxRange nvtx_rs("csm.post.resample");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_out = csm_post_resample_len(n, orig_freq, new_freq);
  if (r.of == r.nf) {
    cudaError_t e = cudaMemcpyAsync(y, x, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? CSM_OK : csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e));
  }
  if (!workspace || workspace_bytes < csm_post_resample_workspace_bytes(orig_freq, new_freq))
    return csm_set_error(CSM_ERR_WORKSPACE, "csm_post_resample: workspace too small");
  float* tab = (float*)workspace;
  post::k_resample_table<<<(r.nf * r.taps + 255) / 256, 256, 0, st>>>(r.of, r.nf, r.width, 0.99, 6, tab);
  const long long strides = (n_out + r.nf - 1) / r.nf;
  const size_t smem = (size_t)((post::RS_STRIDES - 1) * r.of + r.taps) * sizeof(float);
  if (smem > 48 * 1024) return csm_set_error(CSM_ERR_ARG, "csm_post_resample: rate ratio too large");
  post::k_resample<<<(unsigned)((strides + post::RS_STRIDES - 1) / post::RS_STRIDES), 256, smem, st>>>(x, n, tab, r.of, r.nf, r.width, y, n_out);
  csm_count_launches(2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? CSM_OK : csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e));
}

extern "C" int32_t csm_post_pcm16_segment(const float* audio, int64_t n, int64_t start_silence, int64_t end_silence, int64_t fade_in,
                                          int64_t fade_out, int16_t* out, void* scratch4, void* stream) {
  if (!audio || !out || !scratch4 || n < 1 || start_silence < 0 || end_silence < 0 || fade_in < 0 || fade_out < 0)
    return csm_set_error(CSM_ERR_ARG, "csm_post_pcm16_segment: bad arguments");
  const long long total = start_silence + n + end_silence;
  if (fade_in > total || fade_out > total) return csm_set_error(CSM_ERR_ARG, "csm_post_pcm16_segment: fade longer than the segment");
  NvtxRange nvtx_pcm("csm.post.pcm16");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch4, 0, 4, st);
  if (e != cudaSuccess) return csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e));
  post::k_absmax<<<148 * 4, 256, 0, st>>>(audio, n, (unsigned int*)scratch4);
  post::k_pcm16_segment<<<148 * 4, 256, 0, st>>>(audio, n, (const unsigned int*)scratch4, start_silence, end_silence, fade_in, fade_out, out);
  csm_count_launches(2);
  e = cudaGetLastError();
  return e == cudaSuccess ? CSM_OK : csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e));
}
