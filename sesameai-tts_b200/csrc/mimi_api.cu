// C ABI of the Mimi decode path (include/csm_b200.h, mimi_* entry points).
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
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
  // encode side
  float *enc_res1[4], *enc_down[4], *enc_final, *dsw, *enorm, *wavbuf, *p1, *p2, *dots;
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
  {
    int ech = 64;
    const int eratio[4] = {4, 5, 6, 8};
    for (int s = 0; s < 4; ++s) {
      x->enc_res1[s] = cv.take((size_t)(ech / 2) * 3 * ech);
      x->enc_down[s] = cv.take((size_t)(2 * ech) * 2 * eratio[s] * ech);
      ech *= 2;
    }
    x->enc_final = cv.take((size_t)512 * 3 * 1024);
    x->dsw = cv.take((size_t)512 * 4 * 512);
    x->enorm = cv.take((size_t)32 * 2048);
    x->wavbuf = cv.take(T * 1920 + 64);
    x->p1 = cv.take(T * 256);
    x->p2 = cv.take(T * 256);
    x->dots = cv.take(T * 2048);
  }
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

// GEMM dispatch: TF32 tensor cores when K is a whole number of 32-wide tiles and rows are 16-byte
// aligned (every Mimi shape is); precise = 3xTF32 (encode side), else single-pass TF32.
// MIMI_GEMM env (debug): "fp32" forces the CUDA-core SGEMM, "tf32x3" forces the split everywhere.
static int g_gemm_mode = -1;  // 0 auto, 1 fp32, 2 tf32x3 (read-only after the first call)
static thread_local bool g_precise = false;  // per calling thread: two threads may drive two codecs
static cudaError_t gemm(cudaStream_t st, const float* A, long long lda, const float* B, float* C, long long ldc, long long M,
                        int N, int K, const float* bias, int bias_period, int flags, const float* R = nullptr,
                        long long ldr = 0, const float* scale = nullptr) {
  if (g_gemm_mode < 0) {
    const char* e = getenv("MIMI_GEMM");
    g_gemm_mode = !e ? 0 : (!strcmp(e, "fp32") ? 1 : (!strcmp(e, "tf32x3") ? 2 : 0));
  }
  mimi::GemmArgs g;
  g.A = A; g.lda = lda; g.B = B; g.C = C; g.ldc = ldc; g.M = (int)M; g.N = N; g.K = K;
  g.bias = bias; g.bias_period = bias_period > 0 ? bias_period : 1; g.R = R; g.ldr = ldr; g.scale = scale; g.flags = flags;
  const bool tc_ok = g_gemm_mode != 1 && K % mimi::TBK == 0 && lda % 4 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0;
  if (tc_ok) {
    dim3 grid((N + mimi::TBN - 1) / mimi::TBN, (unsigned)((M + mimi::TBM - 1) / mimi::TBM));
    if (g_precise || g_gemm_mode == 2) mimi::k_tgemm<3><<<grid, 256, 0, st>>>(g);
    else mimi::k_tgemm<1><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid((N + mimi::BN - 1) / mimi::BN, (unsigned)((M + mimi::BM - 1) / mimi::BM));
    mimi::k_sgemm<<<grid, 256, 0, st>>>(g);
  }
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
  {
    int ech = 64;
    const int eratio[4] = {4, 5, 6, 8};
    for (int s = 0; s < 4; ++s) {
      const float* const* sw = &x->w[MIMI_W_ENC_STAGE0 + 6 * s];
      const long long n1 = (long long)(ech / 2) * ech * 3, n2 = (long long)(2 * ech) * ech * 2 * eratio[s];
      mimi::k_pack_conv<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(sw[0], ech / 2, ech, 3, x->enc_res1[s]);
      mimi::k_pack_conv<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(sw[4], 2 * ech, ech, 2 * eratio[s], x->enc_down[s]);
      ech *= 2;
    }
    mimi::k_pack_conv<<<(512 * 1024 * 3 + 255) / 256, 256, 0, st>>>(x->w[MIMI_W_ENC_FINAL], 512, 1024, 3, x->enc_final);
    mimi::k_pack_conv<<<(512 * 512 * 4 + 255) / 256, 256, 0, st>>>(x->w[MIMI_W_DOWNSAMPLE], 512, 512, 4, x->dsw);
    mimi::k_row_sumsq256<<<32 * 2048 / 8, 256, 0, st>>>(x->emb, 32LL * 2048, x->enorm);
  }
  csm_count_launches(32 + 3 + 8 + 1 + 11);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&x->final_bias, x->w[MIMI_W_FINAL + 1], sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { delete x; return csm_set_error(CSM_ERR_CUDA, cudaGetErrorString(e)); }
  *out = x;
  return CSM_OK;
}

extern "C" void mimi_destroy(mimi_ctx* x) { delete x; }

// 8-layer causal transformer (context 250) in place on xs [L, 512]; lw0 = first of the 80 layer tensors
static int mimi_transformer(mimi_ctx* x, float* xs, long long L, int w_layer0, cudaStream_t st) {
  using namespace mimi;
  for (int l = 0; l < 8; ++l) {
    const float* const* lw = &x->w[w_layer0 + 10 * l];
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
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}

extern "C" int32_t mimi_encode(mimi_ctx* x, const float* wav, int32_t B, int64_t Lin, int32_t K, int64_t* codes, void* stream) {
  if (!x) return csm_set_error(CSM_ERR_STATE, "mimi_encode: null context");
  if (!wav || !codes || B < 1 || Lin < 1 || K < 1 || K > 32) return csm_set_error(CSM_ERR_ARG, "mimi_encode: bad arguments");
  const long long T = (Lin + 1919) / 1920, Lp = T * 1920;
  if (T > x->max_frames) return csm_set_error(CSM_ERR_OVERFLOW, "mimi_encode: more frames than the codec was created for");
  cudaStream_t st = (cudaStream_t)stream;
  using namespace mimi;
  struct Precise {  // the nearest-centroid search amplifies rounding: 3xTF32 products on the encode side
    Precise() { g_precise = true; }
    ~Precise() { g_precise = false; }
  } precise_scope;
  const int eratio[4] = {4, 5, 6, 8};
  for (int b = 0; b < B; ++b) {
    // waveform, zero-padded to whole frames, 8 zero samples in front (causal k = 7)
    float* wv = x->wavbuf + 8;
    MCU_TRY(cudaMemsetAsync(x->wavbuf, 0, (size_t)(Lp + 8) * sizeof(float), st));
    MCU_TRY(cudaMemcpyAsync(wv, wav + (size_t)b * Lin, (size_t)Lin * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // SEANet encoder, mirrored onto the decoder's activation buffers (same shapes in reverse order)
    float* cur = x->u[3] + (size_t)PAD * 64;  // [Lp, 64]
    k_enc_conv0<<<(unsigned)((Lp * 64 + 255) / 256), 256, 0, st>>>(wv, x->w[MIMI_W_ENC_CONV0], x->w[MIMI_W_ENC_CONV0 + 1], Lp, cur);
    csm_count_launches(1);
    long long rows = Lp;
    int ch = 64;
    for (int s = 0; s < 4; ++s) {
      const float* const* sw = &x->w[MIMI_W_ENC_STAGE0 + 6 * s];
      float* hid = x->r[3 - s];  // [rows, ch/2]
      MCU_TRY(gemm(st, cur - 2 * ch, ch, x->enc_res1[s], hid, ch / 2, rows, ch / 2, 3 * ch, sw[1], ch / 2, F_A_ELU));
      MCU_TRY(gemm(st, hid, ch / 2, sw[2], cur, ch, rows, ch, ch / 2, sw[3], ch, F_A_ELU | F_RESID, cur, ch));
      const int r = eratio[s];
      float* nxt = (s < 3 ? x->u[2 - s] + (size_t)PAD * 2 * ch : x->c0 + (size_t)PAD * 1024);  // [rows/r, 2ch]
      MCU_TRY(gemm(st, cur - (size_t)r * ch, (long long)r * ch, x->enc_down[s], nxt, 2 * ch, rows / r, 2 * ch, 2 * r * ch, sw[5],
                   2 * ch, F_A_ELU));
      rows /= r;
      cur = nxt;
      ch *= 2;
    }
    const long long L2 = rows;  // = 2T
    float* xs = x->xs + (size_t)PAD * 512;
    MCU_TRY(gemm(st, cur - 2 * 1024, 1024, x->enc_final, xs, 512, L2, 512, 3 * 1024, x->w[MIMI_W_ENC_FINAL + 1], 512, F_A_ELU));
    int rc = mimi_transformer(x, xs, L2, MIMI_W_ENC_LAYER0, st);
    if (rc != CSM_OK) return rc;
    // stride-2 downsample with replicate padding
    k_pad_rows<<<2, 256, 0, st>>>(xs, 512, 1);
    MCU_TRY(gemm(st, xs - 2 * 512, 1024, x->dsw, x->e, 512, T, 512, 4 * 512, nullptr, 0, 0));
    k_pad_rows<<<2, 256, 0, st>>>(xs, 512, 0);
    // split RVQ: semantic codebook on its own projection, acoustic codebooks on theirs
    MCU_TRY(gemm(st, x->e, 512, x->w[MIMI_W_RVQ_FIRST_INPROJ], x->p1, 256, T, 256, 512, nullptr, 0, 0));
    MCU_TRY(gemm(st, x->e, 512, x->w[MIMI_W_RVQ_REST_INPROJ], x->p2, 256, T, 256, 512, nullptr, 0, 0));
    for (int k = 0; k < K; ++k) {
      float* res = k == 0 ? x->p1 : x->p2;
      const float* emb = x->emb + (size_t)k * 2048 * 256;
      MCU_TRY(gemm(st, res, 256, emb, x->dots, 2048, T, 2048, 256, nullptr, 0, 0));
      k_rvq_argmin<<<(unsigned)T, 256, 0, st>>>(x->dots, x->enorm + (size_t)k * 2048, emb, res, codes + ((size_t)b * K + k) * T);
    }
    csm_count_launches(2 + K);
    MCU_TRY(cudaGetLastError());
  }
  return CSM_OK;
}

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
    {
      int rc = mimi_transformer(x, xs, L, MIMI_W_LAYER0, st);
      if (rc != CSM_OK) return rc;
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
