// C ABI of the Mimi decode path (include/csm_b200.h, mimi_* entry points).
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/csm_b200.h"
#include "mimi_kernels.cuh"

int csm_set_error(int code, const char* msg);  // api.cu
void csm_count_launches(unsigned long long n);

namespace {

#define MCU_TRY(expr)                                                   \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) {                                            \
      char buf[400];                                                    \
      snprintf(buf, sizeof(buf), "%s: %s", #expr, cudaGetErrorString(_e)); \
      return csm_set_error(CSM_ERR_CUDA, buf);                          \
    }                                                                   \
  } while (0)

const int PAD = 8;  // zero rows in front of every conv input
const int RATIOS[4] = {8, 6, 5, 4};

struct Carver {
  char* base;
  size_t off;
  float* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += n * sizeof(float);
    return p;
  }
};

}  // namespace

struct mimi_ctx {
  int max_frames;
  const float* w[MIMI_W_COUNT];
  // packed weights
  float *emb, *wproj, *conv0, *convtr[4], *res1[4], *finalw;
  float final_bias;
  // activations (per utterance)
  float *q512, *e, *xs, *xn, *qkv, *att, *ff, *c0, *u[4], *r[4];
  size_t pad_rows_bytes;
};

static size_t mimi_carve(mimi_ctx* x, char* base) {
  Carver cv{base, 0};
  const size_t T = x->max_frames, L = 2 * T;
  x->emb = cv.take((size_t)32 * 2048 * 256);
  x->wproj = cv.take(512 * 512);
  x->conv0 = cv.take((size_t)1024 * 7 * 512);
  int ch = 1024;
  for (int s = 0; s < 4; ++s) {
    x->convtr[s] = cv.take((size_t)RATIOS[s] * (ch / 2) * 2 * ch);
    x->res1[s] = cv.take((size_t)(ch / 4) * 3 * (ch / 2));
    ch /= 2;
  }
  x->finalw = cv.take(3 * 64);
  x->q512 = cv.take(T * 512);
  x->e = cv.take(T * 512);
  x->xs = cv.take((L + PAD) * 512);
  x->xn = cv.take(L * 512);
  x->qkv = cv.take(L * 1536);
  x->att = cv.take(L * 512);
  x->ff = cv.take(L * 2048);
  x->c0 = cv.take((L + PAD) * 1024);
  size_t rows = L;
  ch = 1024;
  for (int s = 0; s < 4; ++s) {
    rows *= RATIOS[s];
    x->u[s] = cv.take((rows + PAD) * (ch / 2));
    x->r[s] = cv.take(rows * (ch / 4));
    ch /= 2;
  }
  return (cv.off + 255) & ~(size_t)255;
}

extern "C" size_t mimi_workspace_bytes(int32_t max_frames) {
  if (max_frames < 1 || max_frames > 8192) return 0;
  mimi_ctx tmp;
  tmp.max_frames = max_frames;
  return mimi_carve(&tmp, nullptr);
}

static cudaError_t gemm(cudaStream_t st, const float* A, long long lda, const float* B, float* C, long long ldc, long long M,
                        int N, int K, const float* bias, int bias_period, int flags, const float* R = nullptr,
                        long long ldr = 0, const float* scale = nullptr) {
  mimi::GemmArgs g;
  g.A = A; g.lda = lda; g.B = B; g.C = C; g.ldc = ldc; g.M = (int)M; g.N = N; g.K = K;
  g.bias = bias; g.bias_period = bias_period > 0 ? bias_period : 1; g.R = R; g.ldr = ldr; g.scale = scale; g.flags = flags;
  dim3 grid((N + mimi::BN - 1) / mimi::BN, (unsigned)((M + mimi::BM - 1) / mimi::BM));
  mimi::k_sgemm<<<grid, 256, 0, st>>>(g);
  csm_count_launches(1);
  return cudaGetLastError();
}

extern "C" int32_t mimi_create(const void* const* weights, int32_t n_weights, int32_t max_frames, void* workspace,
                               size_t workspace_bytes, void* stream, mimi_ctx** out) {
  if (!out) return csm_set_error(CSM_ERR_ARG, "out is null");
  *out = nullptr;
  if (!weights || n_weights != MIMI_W_COUNT) return csm_set_error(CSM_ERR_ARG, "mimi_create: expected MIMI_W_COUNT weight pointers");
  for (int i = 0; i < MIMI_W_COUNT; ++i)
    if (!weights[i]) return csm_set_error(CSM_ERR_ARG, "mimi_create: null weight pointer");
  int ndev = 0;
  MCU_TRY(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return csm_set_error(CSM_ERR_CUDA, "no CUDA device (libcsm_b200 has no CPU fallback)");
  const size_t need = mimi_workspace_bytes(max_frames);
  if (!need) return csm_set_error(CSM_ERR_ARG, "mimi_create: bad max_frames");
  if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < need)
    return csm_set_error(CSM_ERR_WORKSPACE, "mimi_create: workspace missing, misaligned or too small");
  mimi_ctx* x = new (std::nothrow) mimi_ctx();
  if (!x) return csm_set_error(CSM_ERR_ARG, "out of host memory");
  x->max_frames = max_frames;
  for (int i = 0; i < MIMI_W_COUNT; ++i) x->w[i] = (const float*)weights[i];
  mimi_carve(x, (char*)workspace);
  cudaStream_t st = (cudaStream_t)stream;
  // zero the pad rows (and everything else once, cheaply enough) then pack weights
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
  if (e != cudaSuccess) { delete x; return csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e)); }
  for (int k = 0; k < 32; ++k) {
    mimi::k_pack_embedding<<<2048, 256, 0, st>>>(x->w[MIMI_W_CODEBOOK0 + 2 * k], x->w[MIMI_W_CODEBOOK0 + 2 * k + 1],
                                                 x->emb + (size_t)k * 2048 * 256);
  }
  mimi::k_pack_rvq_proj<<<1024, 256, 0, st>>>(x->w[MIMI_W_RVQ_FIRST_PROJ], x->w[MIMI_W_RVQ_REST_PROJ], x->wproj);
  mimi::k_pack_conv<<<(1024 * 512 * 7 + 255) / 256, 256, 0, st>>>(x->w[MIMI_W_CONV0], 1024, 512, 7, x->conv0);
  int ch = 1024;
  for (int s = 0; s < 4; ++s) {
    const float* const* sw = &x->w[MIMI_W_STAGE0 + 6 * s];
    const long long nt = (long long)RATIOS[s] * (ch / 2) * 2 * ch;
    mimi::k_pack_convtr<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(sw[0], ch, ch / 2, RATIOS[s], x->convtr[s]);
    const long long nr = (long long)(ch / 4) * (ch / 2) * 3;
    mimi::k_pack_conv<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(sw[2], ch / 4, ch / 2, 3, x->res1[s]);
    ch /= 2;
  }
  mimi::k_pack_conv<<<1, 256, 0, st>>>(x->w[MIMI_W_FINAL], 1, 64, 3, x->finalw);
  csm_count_launches(32 + 3 + 8 + 1);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&x->final_bias, x->w[MIMI_W_FINAL + 1], sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { delete x; return csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e)); }
  *out = x;
  return CSM_OK;
}

extern "C" void mimi_destroy(mimi_ctx* x) { delete x; }

extern "C" int32_t mimi_decode(mimi_ctx* x, const int64_t* codes, int32_t B, int32_t K, int32_t T, float* out, void* stream) {
  if (!x) return csm_set_error(CSM_ERR_STATE, "mimi_decode: null context");
  if (!codes || !out || B < 1 || K < 1 || K > 32 || T < 1) return csm_set_error(CSM_ERR_ARG, "mimi_decode: bad arguments");
  if (T > x->max_frames) return csm_set_error(CSM_ERR_OVERFLOW, "mimi_decode: more frames than the codec was created for");
  cudaStream_t st = (cudaStream_t)stream;
  const long long L = 2LL * T;
  using namespace mimi;
  for (int b = 0; b < B; ++b) {
    const int64_t* cb = codes + (size_t)b * K * T;
    float* wav = out + (size_t)b * 1920 * T;
    k_rvq_gather<<<T, 256, 0, st>>>(cb, K, T, x->emb, x->q512);
    MCU_TRY(gemm(st, x->q512, 512, x->wproj, x->e, 512, T, 512, 512, nullptr, 0, 0));
    float* xs = x->xs + (size_t)PAD * 512;
    k_upsample2<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(x->e, x->w[MIMI_W_UPSAMPLE], T, 512, xs);
    csm_count_launches(2);
    for (int l = 0; l < 8; ++l) {
      const float* const* lw = &x->w[MIMI_W_LAYER0 + 10 * l];
      k_layernorm512<<<(unsigned)((L + 7) / 8), 256, 0, st>>>(xs, lw[2], lw[3], (int)L, 1e-5f, x->xn);
      MCU_TRY(gemm(st, x->xn, 512, lw[0], x->qkv, 1536, L, 1536, 512, nullptr, 0, 0));
      k_rope_qk<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(x->qkv, (int)L);
      k_attn_window<<<dim3((unsigned)((L + 3) / 4), 8), 128, 0, st>>>(x->qkv, (int)L, 250, x->att);
      MCU_TRY(gemm(st, x->att, 512, lw[1], xs, 512, L, 512, 512, nullptr, 0, F_LAYERSCALE | F_RESID, xs, 512, lw[8]));
      k_layernorm512<<<(unsigned)((L + 7) / 8), 256, 0, st>>>(xs, lw[4], lw[5], (int)L, 1e-5f, x->xn);
      MCU_TRY(gemm(st, x->xn, 512, lw[6], x->ff, 2048, L, 2048, 512, nullptr, 0, F_GELU));
      MCU_TRY(gemm(st, x->ff, 2048, lw[7], xs, 512, L, 512, 2048, nullptr, 0, F_LAYERSCALE | F_RESID, xs, 512, lw[9]));
      csm_count_launches(4);
    }
    // SEANet decoder
    float* c0 = x->c0 + (size_t)PAD * 1024;
    MCU_TRY(gemm(st, xs - 6 * 512, 512, x->conv0, c0, 1024, L, 1024, 7 * 512, x->w[MIMI_W_CONV0 + 1], 1024, 0));
    const float* in = c0;
    long long rows = L;
    int ch = 1024;
    for (int s = 0; s < 4; ++s) {
      const float* const* sw = &x->w[MIMI_W_STAGE0 + 6 * s];
      const int r = RATIOS[s], co = ch / 2, hid = ch / 4;
      float* u = x->u[s] + (size_t)PAD * co;
      // ELU -> ConvTranspose1d(ch -> ch/2, kernel 2r, stride r): rows x[q-1], x[q]
      MCU_TRY(gemm(st, in - ch, ch, x->convtr[s], u, (long long)r * co, rows, r * co, 2 * ch, sw[1], co, F_A_ELU));
      rows *= r;
      // residual block: u + conv1(ELU(conv3(ELU(u))))
      MCU_TRY(gemm(st, u - 2 * co, co, x->res1[s], x->r[s], hid, rows, hid, 3 * co, sw[3], hid, F_A_ELU));
      MCU_TRY(gemm(st, x->r[s], hid, sw[4], u, co, rows, co, hid, sw[5], co, F_A_ELU | F_RESID, u, co));
      in = u;
      ch = co;
    }
    k_final_conv<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(in, x->finalw, x->final_bias, rows, wav);
    csm_count_launches(1);
    MCU_TRY(cudaGetLastError());
  }
  return CSM_OK;
}
