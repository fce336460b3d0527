// C ABI of the Mimi decode path (include/csm_b200.h, mimi_* entry points).
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "../../include/csm_b200.h"
#include "nvtx_ranges.h"
#include "mimi_kernels.cuh"
#include "mimi_tc.cuh"

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
const int BT_GROUP = 8;  // utterances whose transformer passes run as one batch (mimi_decode, B > 1)
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
  float* own_state;  // stream state of the one-shot decode (windows of max_frames carry their left context in it)
  // tensor-core decode path (mimi_tc.cuh): TF32-rounded copies of the matrices used straight from the caller's
  // tensors, the raw ConvTranspose1d outputs (residual operand) and a rounded copy of the transformer output
  bool tc;                 // decode on tcgen05 (default) or on the mma.sync kernels (MIMI_DECODE=mma)
  float* tr_r[8][4];       // in_proj, out_proj, linear1, linear2 per layer
  float* res2_r[4];        // residual block's 1x1 conv per stage
  float* u_raw[4];
  float* xs_r;
  float* bt_xr;            // batched one-shot decode: BT_GROUP rounded transformer outputs, (L + PAD) rows each
};

// ---- streamed decode state ------------------------------------------------------------------------
// Everything a causal chunk needs from the frames before it (device floats, caller-owned buffer):
//   e_prev [512]            last row of the RVQ output (depthwise ConvTranspose1d k=4 s=2 reads x[q-1])
//   kv[l]  [249][1024]      RoPE'd K | V of the last <= 249 transformer positions, per layer (window 250)
//   xs_tail [6][512]        last 6 transformer outputs (SEANet Conv1d k=7)
//   c0_tail [1024]          last row of conv0's output (stage-0 ConvTranspose1d reads x[q-1])
//   per stage s: pre[s] [2][C_s] last 2 rows of the ConvTranspose1d output BEFORE the residual block adds to it
//                           (its Conv1d k=3), post[s] [n][C_s] last row(s) after it (next stage's x[q-1]; the final
//                           Conv1d k=3 needs n = 2)
namespace {
const int HIST = 249;
struct StateLayout {
  size_t e_prev, kv[8], xs_tail, c0_tail, pre[4], post[4], total;
};
StateLayout state_layout() {
  StateLayout L;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) & ~(size_t)63; return r; };
  L.e_prev = take(512);
  for (int l = 0; l < 8; ++l) L.kv[l] = take((size_t)HIST * 1024);
  L.xs_tail = take(6 * 512);
  L.c0_tail = take(1024);
  int co = 512;
  for (int s = 0; s < 4; ++s) {
    L.pre[s] = take(2 * (size_t)co);
    L.post[s] = take((s == 3 ? 2 : 1) * (size_t)co);
    co /= 2;
  }
  L.total = o;
  return L;
}
}  // namespace

struct mimi_stream {
  mimi_ctx* x;
  float* st;
  long long frames;  // frames decoded so far
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
  x->qkv = cv.take((L + HIST + 7) * 1536);  // + carried K/V rows of a streamed decode in front
  x->own_state = cv.take(state_layout().total);
  for (int l = 0; l < 8; ++l) {
    x->tr_r[l][0] = cv.take(1536 * 512);
    x->tr_r[l][1] = cv.take(512 * 512);
    x->tr_r[l][2] = cv.take(2048 * 512);
    x->tr_r[l][3] = cv.take(512 * 2048);
  }
  x->xs_r = cv.take((L + PAD) * 512);
  x->bt_xr = cv.take((size_t)BT_GROUP * (L + PAD) * 512);
  {
    size_t rr = L;
    int c2 = 1024;
    for (int s = 0; s < 4; ++s) {
      rr *= RATIOS[s];
      x->res2_r[s] = cv.take((size_t)(c2 / 2) * (c2 / 4));
      x->u_raw[s] = cv.take(rr * (c2 / 2));
      c2 /= 2;
    }
  }
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
                        long long ldr = 0, const float* scale = nullptr, bool fp32_fma = false) {
  if (g_gemm_mode < 0) {
    const char* e = getenv("MIMI_GEMM");
    g_gemm_mode = !e ? 0 : (!strcmp(e, "fp32") ? 1 : (!strcmp(e, "tf32x3") ? 2 : 0));
  }
  mimi::GemmArgs g;
  g.A = A; g.lda = lda; g.B = B; g.C = C; g.ldc = ldc; g.M = (int)M; g.N = N; g.K = K;
  g.bias = bias; g.bias_period = bias_period > 0 ? bias_period : 1; g.R = R; g.ldr = ldr; g.scale = scale; g.flags = flags;
  // fp32_fma: true fp32 multiply-add on the CUDA cores (the split-RVQ search: an argmin over 2048 distances whose
  // top-2 gap can be a few fp32 ulps must not see tensor-core product rounding at all)
  const bool tc_ok = !fp32_fma && g_gemm_mode != 1 && K % mimi::TBK == 0 && lda % 4 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0;
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

// ---- tcgen05 GEMM launcher (mimi_tc.cuh) ------------------------------------------------------------
typedef CUresult (*mimi_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static mimi_encode_tiled_fn mimi_encode_tiled() {
  static mimi_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (mimi_encode_tiled_fn)p;
  }
  return fn;
}
// C[M, N] = A . B^T (+ epilogue), A row m = the contiguous slice of ``taps`` activation rows of ``cin`` channels
// starting at A + m * lda (lda == cin for every Mimi conv: overlapping rows, no im2col), B [N, taps * cin] row-major.
static int gemm_tc(cudaStream_t st, const float* A, long long lda, int cin, int taps, const float* B, long long M, int N,
                   const mtc::Args& ep) {
  mimi_encode_tiled_fn enc = mimi_encode_tiled();
  if (!enc) return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  const int K = cin * taps;
  // output width: whole 128-column tiles, or one tile of 32 / 64 columns (every Mimi shape); bias periods are multiples of 4
  const bool n_ok = N % 128 == 0 || N == 32 || N == 64;
  if (cin % mtc::BK || !n_ok || M < 1 || (((uintptr_t)A | (uintptr_t)B) & 15) || (lda * 4) % 16 || (ep.bias && ep.bias_period % 4))
    return csm_set_error(CSM_ERR_ARG, "mimi gemm_tc: unsupported shape");
  static bool attr[64] = {false};
  int dev = 0;
  MCU_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    MCU_TRY(cudaFuncSetAttribute(mtc::k_gemm_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mtc::SMEM_BYTES));
    MCU_TRY(cudaFuncSetAttribute(mimi::k_attn_window_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(3 * 64 * mimi::AT_LD * sizeof(float))));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  CUtensorMap ma, mb;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)M};
    cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)lda * 4};
    cuuint32_t box[3] = {(cuuint32_t)mtc::BK, 1, (cuuint32_t)mtc::BM};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)A, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled (A, 3-d) failed");
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)mtc::BK, (cuuint32_t)mtc::BN};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)B, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled (B) failed");
  }
  mtc::Args a = ep;
  a.M = (int)M; a.N = N; a.K = K;
  if (a.bias_period < 1) a.bias_period = 1;
  static int sms[64] = {0};
  if (dev >= 0 && dev < 64 && !sms[dev]) MCU_TRY(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (long long)((N + mtc::BN - 1) / mtc::BN) * ((M + mtc::BM - 1) / mtc::BM);
  const int nsm = (dev >= 0 && dev < 64 && sms[dev] > 0) ? sms[dev] : 148;
  const unsigned grid = (unsigned)(tiles < nsm ? tiles : nsm);  // persistent: one CTA per SM walks the tiles
  mtc::k_gemm_tf32<<<grid, mtc::THREADS, mtc::SMEM_BYTES, st>>>(ma, mb, a, cin);
  csm_count_launches(1);
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}
// The weight-resident kernel (mimi_tc.cuh: k_gemm_tf32_r) for one-tile-wide GEMMs over a small weight matrix; falls
// back to gemm_tc when the shape does not qualify.  ``fin`` (optional) carries the fused final convolution.
static bool g_mimi_resident = true;  // MIMI_RESIDENT=0: measurement aid
static int gemm_tc_r(cudaStream_t st, const float* A, long long lda, int cin, int taps, const float* B, long long M, int N,
                     const mtc::Args& ep, const mtc::RArgs* fin = nullptr) {
  const int K = cin * taps;
  const size_t budget = 220 * 1024;
  const bool resid = (ep.flags & mtc::F_RESID) != 0;
  // shared memory: weights + staging tile, then residual tiles (up to 3 in flight) and A stages (up to 8) as they fit
  int stages = 0, nres = 0;
  // tap-shift mode: every 32-channel block of the 128 + taps - 1 activation rows of a tile is loaded once
  static const bool shift_on = !(getenv("MIMI_TAPSHIFT") && getenv("MIMI_TAPSHIFT")[0] == '0');
  const bool shift = shift_on && taps > 1 && taps <= 8 && lda == cin;
  if ((N == 32 || N == 64 || N == 128) && !ep.C2 && !(ep.flags & (mtc::F_GELU | mtc::F_LAYERSCALE)) &&
      mtc::r_smem_bytes(K, N, 2, resid ? 1 : 0, shift) <= budget && (!fin || (N == 64 && resid && M % mtc::BM == 0))) {
    if (resid) {
      nres = mtc::R_MAX_RES;
      while (nres > 1 && mtc::r_smem_bytes(K, N, 3, nres, shift) > budget) --nres;
    }
    stages = mtc::R_MAX_STAGES;
    while (stages > 2 && mtc::r_smem_bytes(K, N, stages, nres, shift) > budget) --stages;
  }
  if (stages < 2 || (!g_mimi_resident && !fin)) {
    if (fin) return csm_set_error(CSM_ERR_ARG, "mimi gemm_tc_r: the fused final conv needs the weight-resident kernel");
    return gemm_tc(st, A, lda, cin, taps, B, M, N, ep);
  }
  mimi_encode_tiled_fn enc = mimi_encode_tiled();
  if (!enc) return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  if (cin % mtc::BK || M < 1 || (((uintptr_t)A | (uintptr_t)B) & 15) || (lda * 4) % 16 || (ep.bias && ep.bias_period % 4))
    return csm_set_error(CSM_ERR_ARG, "mimi gemm_tc_r: unsupported shape");
  typedef void (*rkern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, mtc::RArgs, int);
  static const rkern_t kerns[7] = {mtc::k_gemm_tf32_r<32, false, false>,  mtc::k_gemm_tf32_r<64, false, false>,
                                   mtc::k_gemm_tf32_r<128, false, false>, mtc::k_gemm_tf32_r<32, true, false>,
                                   mtc::k_gemm_tf32_r<64, true, false>,   mtc::k_gemm_tf32_r<128, true, false>,
                                   mtc::k_gemm_tf32_r<64, true, true>};
  const int ki = fin ? 6 : (resid ? 3 : 0) + (N == 32 ? 0 : (N == 64 ? 1 : 2));
  static bool attr[64] = {false};
  static int sms[64] = {0};
  int dev = 0;
  MCU_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return csm_set_error(CSM_ERR_ARG, "mimi gemm_tc_r: device index out of range");
  if (!attr[dev]) {
    for (int i = 0; i < 7; ++i) MCU_TRY(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    MCU_TRY(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
    attr[dev] = true;
  }
  CUtensorMap ma, mb;
  if (shift) {
    // the activation itself, [M + taps - 1 rows, cin]: a stage = 136 rows x 32 channels starting at the tile's first row
    cuuint64_t gdim[2] = {(cuuint64_t)cin, (cuuint64_t)(M + taps - 1)};
    cuuint64_t gstr[1] = {(cuuint64_t)lda * 4};
    cuuint32_t box[2] = {(cuuint32_t)mtc::BK, 136};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)A, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled (A, tap-shift) failed");
  } else {
    cuuint64_t gdim[3] = {(cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)M};
    cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)lda * 4};
    cuuint32_t box[3] = {(cuuint32_t)mtc::BK, 1, (cuuint32_t)mtc::BM};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)A, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled (A, 3-d) failed");
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)mtc::BK, (cuuint32_t)N};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)B, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled (B) failed");
  }
  CUtensorMap mr = ma;  // (unused without a residual)
  if (resid) {
    if (((uintptr_t)ep.R & 15) || (ep.ldr * 4) % 16) return csm_set_error(CSM_ERR_ARG, "mimi gemm_tc_r: misaligned residual");
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)ep.ldr * 4};
    cuuint32_t box[2] = {(cuuint32_t)N, (cuuint32_t)mtc::BM};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ep.R, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return csm_set_error(CSM_ERR_CUDA, "cuTensorMapEncodeTiled (residual) failed");
  }
  mtc::RArgs ra;
  memset(&ra, 0, sizeof(ra));
  if (fin) ra = *fin;
  ra.e = ep;
  ra.e.M = (int)M; ra.e.N = N; ra.e.K = K;
  if (ra.e.bias_period < 1) ra.e.bias_period = 1;
  ra.stages = stages;
  ra.nres = nres;
  ra.taps = shift ? taps : 0;
  const long long tiles = (M + mtc::BM - 1) / mtc::BM;
  const int nsm = sms[dev] > 0 ? sms[dev] : 148;
  const unsigned grid = (unsigned)(tiles < nsm ? tiles : nsm);
  kerns[ki]<<<grid, mtc::R_THREADS, mtc::r_smem_bytes(K, N, stages, nres, shift), st>>>(ma, mb, mr, ra, cin);
  csm_count_launches(1);
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}
static mtc::Args ep_plain(float* C, long long ldc, int flags = 0, const float* bias = nullptr, int period = 1) {
  mtc::Args e;
  memset(&e, 0, sizeof(e));
  e.C = C; e.ldc = ldc; e.flags = flags; e.bias = bias; e.bias_period = period;
  return e;
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
    // TF32 (round-to-nearest) copies / in-place rounding of every matrix a decode GEMM reads: the mma.sync path
    // rounds operands when it loads fragments, the tcgen05 path truncates -- on rounded values both see the same bits
    const char* mode = getenv("MIMI_DECODE");
    x->tc = !(mode && !strcmp(mode, "mma"));
    const char* res = getenv("MIMI_RESIDENT");
    g_mimi_resident = !(res && res[0] == '0');
    auto round_to = [&](const float* src, float* dst, long long n) {
      mimi::k_round_copy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n);
      csm_count_launches(1);
    };
    for (int l = 0; l < 8; ++l) {
      const float* const* lw = &x->w[MIMI_W_LAYER0 + 10 * l];
      round_to(lw[0], x->tr_r[l][0], 1536LL * 512);
      round_to(lw[1], x->tr_r[l][1], 512LL * 512);
      round_to(lw[6], x->tr_r[l][2], 2048LL * 512);
      round_to(lw[7], x->tr_r[l][3], 512LL * 2048);
    }
    round_to(x->wproj, x->wproj, 512LL * 512);
    round_to(x->conv0, x->conv0, 1024LL * 7 * 512);
    int c2 = 1024;
    for (int s = 0; s < 4; ++s) {
      round_to(x->convtr[s], x->convtr[s], (long long)RATIOS[s] * (c2 / 2) * 2 * c2);
      round_to(x->res1[s], x->res1[s], (long long)(c2 / 4) * 3 * (c2 / 2));
      round_to(x->w[MIMI_W_STAGE0 + 6 * s + 4], x->res2_r[s], (long long)(c2 / 2) * (c2 / 4));
      c2 /= 2;
    }
  }
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

static void copy_rows(cudaStream_t st, const float* src, long long lds, float* dst, long long ldd, long long rows, int cols) {
  if (rows <= 0) return;
  const long long n = rows * cols;
  mimi::k_copy_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lds, dst, ldd, (int)rows, cols);
  csm_count_launches(1);
}

// 8-layer causal transformer (context 250) in place on xs [L, 512]; lw0 = first of the 80 layer tensors.
// ``kv`` (streamed decode, else null): per-layer history of the hist = min(pos0, 249) positions before this chunk
// (pos0 = absolute position of row 0); it is read into the rows in front of the chunk's q/k/v and updated.
// ``sc`` / ``nutt`` (tensor-core path, one-shot): xs holds nutt utterances of L / nutt rows each and the scratch
// buffers are the caller's; every row-wise kernel runs once over all rows, attention per utterance (grid z).
struct TrScratch {
  float *xn, *qkv, *att, *ff;
};
static int mimi_transformer(mimi_ctx* x, float* xs, long long L, int w_layer0, cudaStream_t st, float* const* kv = nullptr,
                            long long pos0 = 0, bool tc = false, const TrScratch* sc = nullptr, int nutt = 1) {
  using namespace mimi;
  const int hist = kv ? (int)(pos0 < HIST ? pos0 : HIST) : 0;
  float* const qkv0 = sc ? sc->qkv : x->qkv;
  float* const xn = sc ? sc->xn : x->xn;
  float* const att = sc ? sc->att : x->att;
  float* const ff = sc ? sc->ff : x->ff;
  const long long Lu = L / nutt;
  float* qkv = qkv0 + (size_t)hist * 1536;  // rows of this chunk
  for (int l = 0; l < 8; ++l) {
    const float* const* lw = &x->w[w_layer0 + 10 * l];
    if (tc) {
      // the same layer on the tcgen05 GEMM: operands rounded to TF32 where they are produced
      int rc;
      k_layernorm512<<<(unsigned)((L + 7) / 8), 256, 0, st>>>(xs, lw[2], lw[3], (int)L, 1e-5f, xn, 1);
      if ((rc = gemm_tc(st, xn, 512, 512, 1, x->tr_r[l][0], L, 1536, ep_plain(qkv, 1536))) != CSM_OK) return rc;
      k_rope_qk<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(qkv, (int)L, kv ? pos0 : 0, nutt > 1 ? (int)Lu : 0);
      if (kv) copy_rows(st, kv[l], 1024, qkv0 + 512, 1536, hist, 1024);
      {
        const size_t at_smem = (size_t)3 * 64 * mimi::AT_LD * sizeof(float);
        k_attn_window_tc<<<dim3((unsigned)((Lu + 63) / 64), 8, (unsigned)nutt), 128, at_smem, st>>>(qkv0, (int)Lu, 250, att, hist,
                                                                                                   (kv ? pos0 : 0) - hist, 1);
      }
      if (kv) {
        const long long keep = hist + L < HIST ? hist + L : HIST;
        copy_rows(st, qkv0 + (size_t)(hist + L - keep) * 1536 + 512, 1536, kv[l], 1024, keep, 1024);
      }
      mtc::Args e1 = ep_plain(xs, 512, mtc::F_LAYERSCALE | mtc::F_RESID);
      e1.R = xs; e1.ldr = 512; e1.scale = lw[8];
      if ((rc = gemm_tc(st, att, 512, 512, 1, x->tr_r[l][1], L, 512, e1)) != CSM_OK) return rc;
      k_layernorm512<<<(unsigned)((L + 7) / 8), 256, 0, st>>>(xs, lw[4], lw[5], (int)L, 1e-5f, xn, 1);
      if ((rc = gemm_tc(st, xn, 512, 512, 1, x->tr_r[l][2], L, 2048, ep_plain(ff, 2048, mtc::F_GELU | mtc::F_ROUND))) != CSM_OK) return rc;
      mtc::Args e2 = ep_plain(xs, 512, mtc::F_LAYERSCALE | mtc::F_RESID);
      e2.R = xs; e2.ldr = 512; e2.scale = lw[9];
      if ((rc = gemm_tc(st, ff, 2048, 2048, 1, x->tr_r[l][3], L, 512, e2)) != CSM_OK) return rc;
      csm_count_launches(4);
      continue;
    }
    k_layernorm512<<<(unsigned)((L + 7) / 8), 256, 0, st>>>(xs, lw[2], lw[3], (int)L, 1e-5f, x->xn);
    MCU_TRY(gemm(st, x->xn, 512, lw[0], qkv, 1536, L, 1536, 512, nullptr, 0, 0));
    k_rope_qk<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(qkv, (int)L, kv ? pos0 : 0);
    if (kv) copy_rows(st, kv[l], 1024, x->qkv + 512, 1536, hist, 1024);
    k_attn_window<<<dim3((unsigned)((L + 3) / 4), 8), 128, 0, st>>>(x->qkv, (int)L, 250, x->att, hist);
    if (kv) {
      const long long keep = hist + L < HIST ? hist + L : HIST;
      copy_rows(st, x->qkv + (size_t)(hist + L - keep) * 1536 + 512, 1536, kv[l], 1024, keep, 1024);
    }
    MCU_TRY(gemm(st, x->att, 512, lw[1], xs, 512, L, 512, 512, nullptr, 0, F_LAYERSCALE | F_RESID, xs, 512, lw[8]));
    k_layernorm512<<<(unsigned)((L + 7) / 8), 256, 0, st>>>(xs, lw[4], lw[5], (int)L, 1e-5f, x->xn);
    MCU_TRY(gemm(st, x->xn, 512, lw[6], x->ff, 2048, L, 2048, 512, nullptr, 0, F_GELU));
    MCU_TRY(gemm(st, x->ff, 2048, lw[7], xs, 512, L, 512, 2048, nullptr, 0, F_LAYERSCALE | F_RESID, xs, 512, lw[9]));
    csm_count_launches(4);
  }
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}

// the conv inputs' pad rows are zero for the encoder (it starts every utterance from silence)
static cudaError_t zero_pads(mimi_ctx* x, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(x->xs, 0, (size_t)PAD * 512 * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(x->c0, 0, (size_t)PAD * 1024 * sizeof(float), st);
  int co = 512;
  for (int s = 0; s < 4 && e == cudaSuccess; ++s) {
    e = cudaMemsetAsync(x->u[s], 0, (size_t)PAD * co * sizeof(float), st);
    co /= 2;
  }
  return e;
}

// split RVQ search on the latent x->e [T, 512]: the semantic codebook on its own projection, the acoustic
// codebooks on theirs; nearest centroid per layer on the running residual (true fp32 FMA distances)
static int rvq_search(mimi_ctx* x, int T, int K, int64_t* codes /*[K, T]*/, cudaStream_t st) {
  using namespace mimi;
  MCU_TRY(gemm(st, x->e, 512, x->w[MIMI_W_RVQ_FIRST_INPROJ], x->p1, 256, T, 256, 512, nullptr, 0, 0, nullptr, 0, nullptr, true));
  MCU_TRY(gemm(st, x->e, 512, x->w[MIMI_W_RVQ_REST_INPROJ], x->p2, 256, T, 256, 512, nullptr, 0, 0, nullptr, 0, nullptr, true));
  for (int k = 0; k < K; ++k) {
    float* res = k == 0 ? x->p1 : x->p2;
    const float* emb = x->emb + (size_t)k * 2048 * 256;
    MCU_TRY(gemm(st, res, 256, emb, x->dots, 2048, T, 2048, 256, nullptr, 0, 0, nullptr, 0, nullptr, true));
    k_rvq_argmin<<<(unsigned)T, 256, 0, st>>>(x->dots, x->enorm + (size_t)k * 2048, emb, res, codes + (size_t)k * T);
  }
  csm_count_launches(K);
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}

extern "C" int32_t mimi_encode(mimi_ctx* x, const float* wav, int32_t B, int64_t Lin, int32_t K, int64_t* codes, void* stream) {
  if (!x) return csm_set_error(CSM_ERR_STATE, "mimi_encode: null context");
  if (!wav || !codes || B < 1 || Lin < 1 || K < 1 || K > 32) return csm_set_error(CSM_ERR_ARG, "mimi_encode: bad arguments");
  NvtxRange nvtx_enc("mimi.encode");
  const long long T = (Lin + 1919) / 1920, Lp = T * 1920;
  if (T > x->max_frames) return csm_set_error(CSM_ERR_OVERFLOW, "mimi_encode: more frames than the codec was created for");
  cudaStream_t st = (cudaStream_t)stream;
  using namespace mimi;
  struct Precise {  // the nearest-centroid search amplifies rounding: 3xTF32 products on the encode side
    Precise() { g_precise = true; }
    ~Precise() { g_precise = false; }
  } precise_scope;
  const int eratio[4] = {4, 5, 6, 8};
  MCU_TRY(zero_pads(x, st));  // a streamed decode leaves carried rows there
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
    int rc2 = rvq_search(x, (int)T, K, codes + (size_t)b * K * T, st);
    if (rc2 != CSM_OK) return rc2;
    MCU_TRY(cudaGetLastError());
  }
  return CSM_OK;
}

// Test entry: the split-RVQ search alone on a given latent (time-major [T, 512] fp32) -> codes [K, T].
extern "C" int32_t mimi_k_rvq_encode(mimi_ctx* x, const float* latent, int32_t T, int32_t K, int64_t* codes, void* stream) {
  if (!x) return csm_set_error(CSM_ERR_STATE, "mimi_k_rvq_encode: null context");
  if (!latent || !codes || T < 1 || T > x->max_frames || K < 1 || K > 32) return csm_set_error(CSM_ERR_ARG, "mimi_k_rvq_encode: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  MCU_TRY(cudaMemcpyAsync(x->e, latent, (size_t)T * 512 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return rvq_search(x, T, K, codes, st);
}

// One causal chunk of one utterance: frames [0, T) of ``codes`` ([K, ldt] layout) continue the stream whose
// left context is ``sbuf`` (``frames_done`` frames so far; an all-zero state = the start of an utterance).
static int decode_chunk_tc(mimi_ctx* x, float* sbuf, long long frames_done, const int64_t* codes, int K, int T, long long ldt,
                           float* wav, cudaStream_t st);
static int decode_front_tc(mimi_ctx* x, float* sbuf, const int64_t* codes, int K, int T, long long ldt, float* xs, cudaStream_t st);
static int decode_seanet_tc(mimi_ctx* x, float* sbuf, float* xr, long long L, float* wav, cudaStream_t st);

static int decode_chunk(mimi_ctx* x, float* sbuf, long long frames_done, const int64_t* codes, int K, int T, long long ldt,
                        float* wav, cudaStream_t st) {
  if (x->tc) return decode_chunk_tc(x, sbuf, frames_done, codes, K, T, ldt, wav, st);
  using namespace mimi;
  const StateLayout SL = state_layout();
  const long long L = 2LL * T;
  k_rvq_gather<<<T, 256, 0, st>>>(codes, K, T, ldt, x->emb, x->q512);
  MCU_TRY(gemm(st, x->q512, 512, x->wproj, x->e, 512, T, 512, 512, nullptr, 0, 0));
  float* xs = x->xs + (size_t)PAD * 512;
  k_upsample2<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(x->e, x->w[MIMI_W_UPSAMPLE], T, 512, xs, sbuf + SL.e_prev);
  copy_rows(st, x->e + (size_t)(T - 1) * 512, 512, sbuf + SL.e_prev, 512, 1, 512);
  csm_count_launches(2);
  {
    float* kv[8];
    for (int l = 0; l < 8; ++l) kv[l] = sbuf + SL.kv[l];
    int rc = mimi_transformer(x, xs, L, MIMI_W_LAYER0, st, kv, 2 * frames_done);
    if (rc != CSM_OK) return rc;
  }
  // SEANet decoder.  Before a conv reads a buffer, the rows in front of it are loaded with the carried tail of
  // the previous chunk; the new tail is saved by reading THROUGH those rows (a chunk may be shorter than the tail).
  copy_rows(st, sbuf + SL.xs_tail, 512, xs - 6 * 512, 512, 6, 512);
  copy_rows(st, xs + (L - 6) * 512, 512, sbuf + SL.xs_tail, 512, 6, 512);
  float* c0 = x->c0 + (size_t)PAD * 1024;
  MCU_TRY(gemm(st, xs - 6 * 512, 512, x->conv0, c0, 1024, L, 1024, 7 * 512, x->w[MIMI_W_CONV0 + 1], 1024, 0));
  copy_rows(st, sbuf + SL.c0_tail, 1024, c0 - 1024, 1024, 1, 1024);
  copy_rows(st, c0 + (L - 1) * 1024, 1024, sbuf + SL.c0_tail, 1024, 1, 1024);
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
    // residual block: u + conv1(ELU(conv3(ELU(u)))); conv3 sees the previous chunk's last two PRE-residual rows
    copy_rows(st, sbuf + SL.pre[s], co, u - 2 * co, co, 2, co);
    copy_rows(st, u + (rows - 2) * co, co, sbuf + SL.pre[s], co, 2, co);
    MCU_TRY(gemm(st, u - 2 * co, co, x->res1[s], x->r[s], hid, rows, hid, 3 * co, sw[3], hid, F_A_ELU));
    MCU_TRY(gemm(st, x->r[s], hid, sw[4], u, co, rows, co, hid, sw[5], co, F_A_ELU | F_RESID, u, co));
    // what follows (next ConvTranspose1d: 1 row; final Conv1d k=3: 2 rows) sees the POST-residual tail
    const int np = s == 3 ? 2 : 1;
    copy_rows(st, sbuf + SL.post[s], co, u - (size_t)np * co, co, np, co);
    copy_rows(st, u + (rows - np) * co, co, sbuf + SL.post[s], co, np, co);
    in = u;
    ch = co;
  }
  k_final_conv<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(in, x->finalw, x->final_bias, rows, wav);
  csm_count_launches(1);
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}

// The same chunk on the tensor cores (mimi_tc.cuh).  SEANet activations are stored the way their consumers read
// them -- ELU applied, rounded to TF32 -- next to the raw ConvTranspose1d output the residual add needs; the
// carried tails (stream state) hold the same representation, so a stream must not switch paths mid-utterance
// (the path is a property of the context).
static int decode_chunk_tc(mimi_ctx* x, float* sbuf, long long frames_done, const int64_t* codes, int K, int T, long long ldt,
                           float* wav, cudaStream_t st) {
  const long long L = 2LL * T;
  int rc;
  float* xs = x->xs + (size_t)PAD * 512;
  if ((rc = decode_front_tc(x, sbuf, codes, K, T, ldt, xs, st)) != CSM_OK) return rc;
  {
    const StateLayout SL = state_layout();
    float* kv[8];
    for (int l = 0; l < 8; ++l) kv[l] = sbuf + SL.kv[l];
    if ((rc = mimi_transformer(x, xs, L, MIMI_W_LAYER0, st, kv, 2 * frames_done, true)) != CSM_OK) return rc;
  }
  // conv0 reads a TF32-rounded copy of the transformer output (the residual stream itself stays fp32)
  float* xr = x->xs_r + (size_t)PAD * 512;
  mtc::k_round_tf32<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(xs, xr, L * 512);
  csm_count_launches(1);
  return decode_seanet_tc(x, sbuf, xr, L, wav, st);
}

// codes -> RVQ sum -> projection -> depthwise ConvTranspose1d x2: the transformer's input rows xs [2T, 512]
static int decode_front_tc(mimi_ctx* x, float* sbuf, const int64_t* codes, int K, int T, long long ldt, float* xs, cudaStream_t st) {
  using namespace mimi;
  const StateLayout SL = state_layout();
  const long long L = 2LL * T;
  int rc;
  k_rvq_gather<<<T, 256, 0, st>>>(codes, K, T, ldt, x->emb, x->q512, 1);
  if ((rc = gemm_tc(st, x->q512, 512, 512, 1, x->wproj, T, 512, ep_plain(x->e, 512))) != CSM_OK) return rc;
  k_upsample2<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(x->e, x->w[MIMI_W_UPSAMPLE], T, 512, xs, sbuf + SL.e_prev);
  copy_rows(st, x->e + (size_t)(T - 1) * 512, 512, sbuf + SL.e_prev, 512, 1, 512);
  csm_count_launches(2);
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}

// SEANet decoder on the TF32-rounded transformer output xr [L, 512] (>= 6 writable rows in front of it)
static int decode_seanet_tc(mimi_ctx* x, float* sbuf, float* xr, long long L, float* wav, cudaStream_t st) {
  using namespace mimi;
  const StateLayout SL = state_layout();
  int rc;
  copy_rows(st, sbuf + SL.xs_tail, 512, xr - 6 * 512, 512, 6, 512);
  copy_rows(st, xr + (L - 6) * 512, 512, sbuf + SL.xs_tail, 512, 6, 512);
  float* c0 = x->c0 + (size_t)PAD * 1024;  // holds ELU(conv0(..))
  if ((rc = gemm_tc(st, xr - 6 * 512, 512, 512, 7, x->conv0, L, 1024, ep_plain(c0, 1024, mtc::F_OUT_ELU, x->w[MIMI_W_CONV0 + 1], 1024))) != CSM_OK)
    return rc;
  copy_rows(st, sbuf + SL.c0_tail, 1024, c0 - 1024, 1024, 1, 1024);
  copy_rows(st, c0 + (L - 1) * 1024, 1024, sbuf + SL.c0_tail, 1024, 1, 1024);
  const float* in = c0;
  long long rows = L;
  int ch = 1024;
  for (int s = 0; s < 4; ++s) {
    const float* const* sw = &x->w[MIMI_W_STAGE0 + 6 * s];
    const int r = RATIOS[s], co = ch / 2, hid = ch / 4;
    float* ue = x->u[s] + (size_t)PAD * co;  // ELU'd, rounded: what the convs read
    float* ur = x->u_raw[s];                 // raw ConvTranspose1d output: the residual operand
    // ConvTranspose1d(ch -> ch/2, kernel 2r, stride r) on the ELU'd input: rows x[q-1], x[q]
    mtc::Args ec = ep_plain(ur, (long long)r * co, 0, sw[1], co);
    ec.C2 = ue; ec.ldc2 = (long long)r * co;
    if ((rc = gemm_tc(st, in - ch, ch, ch, 2, x->convtr[s], rows, r * co, ec)) != CSM_OK) return rc;
    rows *= r;
    copy_rows(st, sbuf + SL.pre[s], co, ue - 2 * co, co, 2, co);
    copy_rows(st, ue + (rows - 2) * co, co, sbuf + SL.pre[s], co, 2, co);
    // residual block: u + conv1(ELU(conv3(ELU(u)))), stored as ELU(..) for what follows
    if ((rc = gemm_tc_r(st, ue - 2 * co, co, co, 3, x->res1[s], rows, hid, ep_plain(x->r[s], hid, mtc::F_OUT_ELU, sw[3], hid))) != CSM_OK) return rc;
    mtc::Args er = ep_plain(ue, co, mtc::F_OUT_ELU | mtc::F_RESID, sw[5], co);
    er.R = ur; er.ldr = co;
    const int np = s == 3 ? 2 : 1;
    if (s == 3 && g_mimi_resident) {
      // last stage: the final Conv1d(64 -> 1, k = 3) runs in this GEMM's epilogue; the stage's activation is not stored
      copy_rows(st, sbuf + SL.post[s], co, ue - (size_t)np * co, co, np, co);  // rows -2, -1: the previous chunk's tail
      MCU_TRY(cudaMemsetAsync(wav, 0, (size_t)rows * sizeof(float), st));
      mtc::RArgs fin;
      memset(&fin, 0, sizeof(fin));
      fin.fw = x->finalw; fin.fb = x->final_bias; fin.wav = wav; fin.halo = ue - (size_t)np * co; fin.tail = sbuf + SL.post[s];
      er.C = nullptr;
      if ((rc = gemm_tc_r(st, x->r[s], hid, hid, 1, x->res2_r[s], rows, co, er, &fin)) != CSM_OK) return rc;
      MCU_TRY(cudaGetLastError());
      return CSM_OK;
    }
    if ((rc = gemm_tc_r(st, x->r[s], hid, hid, 1, x->res2_r[s], rows, co, er)) != CSM_OK) return rc;
    copy_rows(st, sbuf + SL.post[s], co, ue - (size_t)np * co, co, np, co);
    copy_rows(st, ue + (rows - np) * co, co, sbuf + SL.post[s], co, np, co);
    in = ue;
    ch = co;
  }
  k_final_conv<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(in, x->finalw, x->final_bias, rows, wav, 1);
  csm_count_launches(1);
  MCU_TRY(cudaGetLastError());
  return CSM_OK;
}

// One-shot decode of several utterances (T <= max_frames): the transformer -- 8 layers of small row-wise kernels, a
// third of a 60 s decode when run per utterance on 1500 rows -- runs ONCE over the rows of up to BT_GROUP utterances
// (attention per utterance); front end and SEANet stay per utterance.  Row-wise kernels and per-utterance attention
// tiles compute every row exactly as the per-utterance path does: the result is bit-identical.  The batch's scratch
// (residual stream, xn, qkv, att, ff: 5120 floats per row) lives in the last stage's raw ConvTranspose1d buffer,
// which is dead until the SEANet of the group's first utterance runs; the rounded outputs wait in bt_xr.
static int decode_batch_tc(mimi_ctx* x, const int64_t* codes, int B, int K, int T, float* out, cudaStream_t st) {
  const StateLayout SL = state_layout();
  const long long L = 2LL * T;
  const size_t cap = (size_t)2 * x->max_frames * 960 * 64;  // floats in u_raw[3]
  int G = BT_GROUP;
  while (G > 1 && (size_t)G * L * 5120 > cap) --G;
  float* bxs = x->u_raw[3];
  TrScratch sc;
  sc.xn = bxs + (size_t)G * L * 512;
  sc.qkv = sc.xn + (size_t)G * L * 512;
  sc.att = sc.qkv + (size_t)G * L * 1536;
  sc.ff = sc.att + (size_t)G * L * 512;
  const size_t slot = (size_t)(2 * x->max_frames + PAD) * 512;
  int rc;
  for (int b0 = 0; b0 < B; b0 += G) {
    const int g = B - b0 < G ? B - b0 : G;
    for (int i = 0; i < g; ++i) {
      MCU_TRY(cudaMemsetAsync(x->own_state + SL.e_prev, 0, 512 * sizeof(float), st));
      if ((rc = decode_front_tc(x, x->own_state, codes + (size_t)(b0 + i) * K * T, K, T, T, bxs + (size_t)i * L * 512, st)) != CSM_OK) return rc;
    }
    if ((rc = mimi_transformer(x, bxs, (long long)g * L, MIMI_W_LAYER0, st, nullptr, 0, true, &sc, g)) != CSM_OK) return rc;
    for (int i = 0; i < g; ++i) {
      mtc::k_round_tf32<<<(unsigned)((L * 512 + 255) / 256), 256, 0, st>>>(bxs + (size_t)i * L * 512, x->bt_xr + i * slot + (size_t)PAD * 512, L * 512);
      csm_count_launches(1);
    }
    for (int i = 0; i < g; ++i) {
      MCU_TRY(cudaMemsetAsync(x->own_state, 0, SL.kv[0] * sizeof(float), st));  // e_prev (the K/V history is not used here)
      MCU_TRY(cudaMemsetAsync(x->own_state + SL.xs_tail, 0, (SL.total - SL.xs_tail) * sizeof(float), st));
      if ((rc = decode_seanet_tc(x, x->own_state, x->bt_xr + i * slot + (size_t)PAD * 512, L, out + (size_t)(b0 + i) * 1920 * T, st)) != CSM_OK) return rc;
    }
  }
  return CSM_OK;
}

extern "C" int32_t mimi_decode(mimi_ctx* x, const int64_t* codes, int32_t B, int32_t K, int32_t T, float* out, void* stream) {
  if (!x) return csm_set_error(CSM_ERR_STATE, "mimi_decode: null context");
  if (!codes || !out || B < 1 || K < 1 || K > 32 || T < 1) return csm_set_error(CSM_ERR_ARG, "mimi_decode: bad arguments");
  NvtxRange nvtx_dec("mimi.decode");
  cudaStream_t st = (cudaStream_t)stream;
  const StateLayout SL = state_layout();
  static const bool batch_on = !(getenv("MIMI_BATCH") && getenv("MIMI_BATCH")[0] == '0');  // measurement aid
  if (x->tc && B > 1 && T <= x->max_frames && batch_on) return decode_batch_tc(x, codes, B, K, T, out, st);
  // any length: windows of at most max_frames frames, the causal left context carried between them
  // (moshi's MimiModel.decode has no length limit either)
  for (int b = 0; b < B; ++b) {
    MCU_TRY(cudaMemsetAsync(x->own_state, 0, SL.total * sizeof(float), st));
    for (int t0 = 0; t0 < T; t0 += x->max_frames) {
      const int n = T - t0 < x->max_frames ? T - t0 : x->max_frames;
      int rc = decode_chunk(x, x->own_state, t0, codes + (size_t)b * K * T + t0, K, n, T, out + (size_t)b * 1920 * T + (size_t)t0 * 1920, st);
      if (rc != CSM_OK) return rc;
    }
  }
  return CSM_OK;
}

extern "C" size_t mimi_stream_state_bytes(void) { return state_layout().total * sizeof(float); }

extern "C" int32_t mimi_stream_create(mimi_ctx* x, void* state, size_t state_bytes, void* stream, mimi_stream** out) {
  if (!out) return csm_set_error(CSM_ERR_ARG, "out is null");
  *out = nullptr;
  if (!x) return csm_set_error(CSM_ERR_STATE, "mimi_stream_create: null context");
  if (!state || ((uintptr_t)state & 255) || state_bytes < mimi_stream_state_bytes())
    return csm_set_error(CSM_ERR_WORKSPACE, "mimi_stream_create: state buffer missing, misaligned or too small");
  mimi_stream* s = new (std::nothrow) mimi_stream();
  if (!s) return csm_set_error(CSM_ERR_ARG, "out of host memory");
  s->x = x; s->st = (float*)state; s->frames = 0;
  MCU_TRY(cudaMemsetAsync(state, 0, mimi_stream_state_bytes(), (cudaStream_t)stream));
  *out = s;
  return CSM_OK;
}
extern "C" int32_t mimi_stream_reset(mimi_stream* s, void* stream) {
  if (!s) return csm_set_error(CSM_ERR_STATE, "mimi_stream_reset: null stream");
  s->frames = 0;
  MCU_TRY(cudaMemsetAsync(s->st, 0, mimi_stream_state_bytes(), (cudaStream_t)stream));
  return CSM_OK;
}
extern "C" void mimi_stream_destroy(mimi_stream* s) { delete s; }

extern "C" int32_t mimi_decode_stream(mimi_stream* s, const int64_t* codes, int32_t K, int32_t T, float* out, void* stream) {
  if (!s) return csm_set_error(CSM_ERR_STATE, "mimi_decode_stream: null stream");
  if (!codes || !out || K < 1 || K > 32 || T < 1) return csm_set_error(CSM_ERR_ARG, "mimi_decode_stream: bad arguments");
  NvtxRange nvtx_decs("mimi.decode_stream");
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < T; t0 += s->x->max_frames) {
    const int n = T - t0 < s->x->max_frames ? T - t0 : s->x->max_frames;
    int rc = decode_chunk(s->x, s->st, s->frames, codes + t0, K, n, T, out + (size_t)t0 * 1920, st);
    if (rc != CSM_OK) return rc;
    s->frames += n;
  }
  return CSM_OK;
}
