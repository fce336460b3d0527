// Mimi codec decode kernels (fp32, moshi 0.2.2 semantics).
//
// Layout: every activation is TIME-MAJOR [rows, channels] fp32 with zeroed pad rows in front, so
//   * a causal Conv1d(k) is a plain GEMM: row t of the im2col matrix is the contiguous slice
//     x[(t-k+1) .. t] (k*Cin floats, row stride Cin -> overlapping rows), weights packed [Cout][k][Cin];
//   * a causal ConvTranspose1d(kernel 2s, stride s) is a plain GEMM with K = 2*Cin (rows x[q-1], x[q])
//     and N = s*Cout: output row q holds the s upsampled time steps back to back, which IS the
//     time-major layout of the upsampled sequence (the k-s trimmed samples are never computed);
//   * ELU is applied when the A operand is staged, bias / GELU / LayerScale / residual in the epilogue.
// GEMMs: k_tgemm (TF32 tensor cores, mma.sync, fp32 accumulate; 3xTF32 split on the encode side where a
// nearest-centroid search consumes the result) with k_sgemm (CUDA cores) for shapes it does not take.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mimi {

enum { F_A_ELU = 1, F_GELU = 2, F_RESID = 4, F_LAYERSCALE = 8 };

struct GemmArgs {
  const float* A;
  long long lda;
  const float* B;  // [N, K] row-major
  float* C;
  long long ldc;
  int M, N, K;
  const float* bias;  // bias[n % bias_period] or null
  int bias_period;
  const float* R;  // residual [M, ldr] (may alias C)
  long long ldr;
  const float* scale;  // LayerScale [N]
  int flags;
};

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
// round-to-nearest to TF32 (10-bit mantissa): the tensor-core decode path (mimi_tc.cuh) rounds every GEMM operand
// where it is produced, so that tcgen05's truncation of the low mantissa bits changes nothing
__device__ __forceinline__ float rtf32(float x, int on) {
  if (!on) return x;
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) k_sgemm(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lc = (tid & 3) * 4;  // loader: row 0..63, k offset 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;       // compute: 4x4 micro tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_ok = m0 + lr < g.M, b_ok = n0 + lr < g.N;
  const float* ap = g.A + (long long)(m0 + lr) * g.lda + lc;
  const float* bp = g.B + (long long)(n0 + lr) * g.K + lc;
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (a_ok) a = *reinterpret_cast<const float4*>(ap + k0);
    if (b_ok) b = *reinterpret_cast<const float4*>(bp + k0);
    if (g.flags & F_A_ELU) {
      a.x = elu1(a.x); a.y = elu1(a.y); a.z = elu1(a.z); a.w = elu1(a.w);
    }
    __syncthreads();
    As[lc + 0][lr] = a.x; As[lc + 1][lr] = a.y; As[lc + 2][lr] = a.z; As[lc + 3][lr] = a.w;
    Bs[lc + 0][lr] = b.x; Bs[lc + 1][lr] = b.y; Bs[lc + 2][lr] = b.z; Bs[lc + 3][lr] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n % g.bias_period];
      if (g.flags & F_GELU) v = gelu_erf(v);
      if (g.flags & F_LAYERSCALE) v *= g.scale[n];
      if (g.flags & F_RESID) v += g.R[(long long)m * g.ldr + n];
      g.C[(long long)m * g.ldc + n] = v;
    }
  }
}

// ---- tensor-core GEMM (TF32 mma.sync, fp32 accumulate) -------------------------------------------
// Same contract as k_sgemm.  CTA tile 128 x 64 x 32, 8 warps as 4 (M) x 2 (N), warp tile 32 x 32 =
// 2 x 4 mma.m16n8k8 tiles; the next k-tile is fetched into registers while the current one is
// multiplied (one shared-memory stage, padded rows: conflict-free fragment loads).
// PASSES == 1: operands rounded to TF32 (10-bit mantissa).  PASSES == 3: the 3xTF32 split
// a = hi + lo, a*b ~ hi*hi + hi*lo + lo*hi: fp32-grade products at a third of the tensor rate (used
// where a nearest-centroid search consumes the result).
constexpr int TBM = 128, TBN = 64, TBK = 32, TPAD = 4;

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int PASSES>
__global__ void __launch_bounds__(256) k_tgemm(GemmArgs g) {
  __shared__ __align__(16) float As[TBM][TBK + TPAD];
  __shared__ __align__(16) float Bs[TBN][TBK + TPAD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, q = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const long long m0 = (long long)blockIdx.y * TBM;
  const int n0 = blockIdx.x * TBN;
  // loaders: A tile 128 x 32 floats = 1024 float4 (4 per thread), B tile 64 x 32 = 512 float4 (2 per thread)
  const int lrow = tid >> 3, lk = (tid & 7) * 4;  // row 0..31 (+32 per step), k offset 0..28
  float4 pa[4], pb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + lrow + 32 * i;
      pa[i] = m < g.M ? *reinterpret_cast<const float4*>(g.A + m * g.lda + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int n = n0 + lrow + 32 * i;
      pb[i] = n < g.N ? *reinterpret_cast<const float4*>(g.B + (long long)n * g.K + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
  const bool elu = (g.flags & F_A_ELU) != 0;
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += TBK) {
    __syncthreads();  // the previous tile has been consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 a = pa[i];
      if (elu) {
        // single-pass TF32 keeps 10 mantissa bits: exp(x) - 1 from the fast exponential (absolute error ~1e-7)
        // is far inside that, and expm1f was the dominant cost of the skinny SEANet GEMMs
        if (PASSES == 1) {
          a.x = a.x > 0.f ? a.x : __expf(a.x) - 1.f; a.y = a.y > 0.f ? a.y : __expf(a.y) - 1.f;
          a.z = a.z > 0.f ? a.z : __expf(a.z) - 1.f; a.w = a.w > 0.f ? a.w : __expf(a.w) - 1.f;
        } else {
          a.x = elu1(a.x); a.y = elu1(a.y); a.z = elu1(a.z); a.w = elu1(a.w);
        }
      }
      *reinterpret_cast<float4*>(&As[lrow + 32 * i][lk]) = a;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&Bs[lrow + 32 * i][lk]) = pb[i];
    __syncthreads();
    if (k0 + TBK < g.K) fetch(k0 + TBK);
#pragma unroll
    for (int kk = 0; kk < TBK; kk += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float* ap = &As[wm + 16 * i + gq][kk + q];
        const float v[4] = {ap[0], ap[8 * (TBK + TPAD)], ap[4], ap[8 * (TBK + TPAD) + 4]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ah[i][e] = f2tf32(v[e]);
          if (PASSES == 3) al[i][e] = f2tf32(v[e] - __uint_as_float(ah[i][e]));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* bp = &Bs[wn + 8 * j + gq][kk + q];
        const float v[2] = {bp[0], bp[4]};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          bh[j][e] = f2tf32(v[e]);
          if (PASSES == 3) bl[j][e] = f2tf32(v[e] - __uint_as_float(bh[j][e]));
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (PASSES == 3) {
            mma_tf32(acc[i][j], al[i], bh[j]);
            mma_tf32(acc[i][j], ah[i], bl[j]);
          }
          mma_tf32(acc[i][j], ah[i], bh[j]);
        }
    }
  }
  // epilogue: accumulator (i, j): rows wm + 16 i + gq (+8), columns wn + 8 j + 2 q (+1)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long m = m0 + wm + 16 * i + gq + 8 * h;
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = n0 + wn + 8 * j + 2 * q + e;
          if (n >= g.N) continue;
          float v = acc[i][j][2 * h + e];
          if (g.bias) v += g.bias[n % g.bias_period];
          if (g.flags & F_GELU) v = gelu_erf(v);
          if (g.flags & F_LAYERSCALE) v *= g.scale[n];
          if (g.flags & F_RESID) v += g.R[m * g.ldr + n];
          g.C[m * g.ldc + n] = v;
        }
    }
}

// split-RVQ lookup: q[t] = [ sum over semantic codebooks | sum over acoustic codebooks ]  (2 x 256)
// embedding = embedding_sum / clamp(cluster_usage, 1e-5) is precomputed at create time.
__global__ void k_rvq_gather(const int64_t* __restrict__ codes /*[K, ldt] of this utterance, frames [0, T) of it*/, int K, int T,
                             long long ldt, const float* __restrict__ emb /*[32][2048][256]*/, float* __restrict__ out /*[T, 512]*/,
                             int round = 0) {
  const int t = blockIdx.x;
  const int d = threadIdx.x;  // 256 threads
  float first = 0.f, rest = 0.f;
  for (int k = 0; k < K; ++k) {
    long long code = codes[(long long)k * ldt + t];
    code = code < 0 ? 0 : (code > 2047 ? 2047 : code);  // memory safety: the reference would raise on these
    const float v = emb[((long long)k * 2048 + code) * 256 + d];
    if (k == 0) first += v;
    else rest += v;
  }
  out[(long long)t * 512 + d] = rtf32(first, round);
  out[(long long)t * 512 + 256 + d] = rtf32(rest, round);
}

// depthwise ConvTranspose1d(k=4, s=2, groups=C), causal (trim 2 on the right):
// y[2q + r][c] = x[q][c] * w[c][r] + x[q-1][c] * w[c][r+2]
// ``prev`` [C] = the row before x[0] (the last row of the previous chunk of a streamed decode; zeros at the start)
__global__ void k_upsample2(const float* __restrict__ x /*[T, C]*/, const float* __restrict__ w /*[C][4]*/, int T, int C,
                            float* __restrict__ y /*[2T, C]*/, const float* __restrict__ prev) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)2 * T * C) return;
  const int c = i % C;
  const long long o = i / C;
  const long long q = o >> 1;
  const int r = o & 1;
  float v = x[q * C + c] * w[c * 4 + r];
  v += (q > 0 ? x[(q - 1) * C + c] : prev[c]) * w[c * 4 + r + 2];
  y[i] = v;
}
// dst[r][c] = src[r][c] for r < rows: carries causal left context (conv pad rows, K/V history) between chunks
__global__ void k_copy_rows(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, int rows, int cols) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  dst[r * ldd + c] = src[r * lds + c];
}

// LayerNorm over 512 channels, one warp per row
__global__ void __launch_bounds__(256) k_layernorm512(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ b, int rows, float eps,
                                                      float* __restrict__ y, int round = 0) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * 512);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 512.0f);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    var += a * a + bb * bb + c * c + d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float inv = rsqrtf(var * (1.0f / 512.0f) + eps);
  float4* yr = reinterpret_cast<float4*>(y + (long long)row * 512);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 ww = reinterpret_cast<const float4*>(w)[lane + 32 * i];
    const float4 bb = reinterpret_cast<const float4*>(b)[lane + 32 * i];
    float4 o;
    o.x = rtf32((v[i].x - mean) * inv * ww.x + bb.x, round);
    o.y = rtf32((v[i].y - mean) * inv * ww.y + bb.y, round);
    o.z = rtf32((v[i].z - mean) * inv * ww.z + bb.z, round);
    o.w = rtf32((v[i].w - mean) * inv * ww.w + bb.w, round);
    yr[lane + 32 * i] = o;
  }
}

// interleaved RoPE (moshi apply_rope, max_period 10000) in place on the q and k thirds of qkv [L, 1536]
// rows [0, L) of ``qkv`` are absolute positions pos0 .. pos0 + L - 1
// period > 0: the rows are several utterances of ``period`` rows each, every one starting at position pos0
__global__ void k_rope_qk(float* __restrict__ qkv, int L, long long pos0, int period = 0) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // (t, which, head, pair)
  if (i >= (long long)L * 2 * 8 * 32) return;
  const int pair = i & 31;
  const int head = (i >> 5) & 7;
  const int which = (i >> 8) & 1;
  const long long t = i >> 9;
  const float freq = expf((float)pair * (-9.210340371976184f * 2.0f / 64.0f));  // ln(10000)
  float sn, cs;
  sincosf(freq * (float)(pos0 + (period > 0 ? t % period : t)), &sn, &cs);
  float* p = qkv + t * 1536 + which * 512 + head * 64 + pair * 2;
  const float xr = p[0], xi = p[1];
  p[0] = xr * cs - xi * sn;
  p[1] = xr * sn + xi * cs;
}

// causal attention with a `context`-key window; one warp per (query, head); qkv [L, 1536] -> out [L, 512]
// Queries are rows [hist, hist + L) of ``qkv``; rows [0, hist) hold the carried K/V of the positions before
// this chunk (streamed decode), so row index differences are position differences.  out row = t - hist.
__global__ void __launch_bounds__(128) k_attn_window(const float* __restrict__ qkv, int L, int context,
                                                     float* __restrict__ out, int hist, int round = 0) {
  const int tq = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int h = blockIdx.y;
  const int lane = threadIdx.x & 31;
  if (tq >= L) return;
  const int t = tq + hist;
  const float scale = 0.125f;  // 1/sqrt(64)
  float q[64];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (long long)t * 1536 + h * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = qp[i];
      q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
    }
  }
  const int j0 = t - context + 1 > 0 ? t - context + 1 : 0;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int c0 = j0; c0 <= t; c0 += 32) {
    const int key = c0 + lane;
    float s = -INFINITY;
    if (key <= t) {
      const float4* kp = reinterpret_cast<const float4*>(qkv + (long long)key * 1536 + 512 + h * 64);
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 v = kp[i];
        a = fmaf(q[4 * i], v.x, a); a = fmaf(q[4 * i + 1], v.y, a);
        a = fmaf(q[4 * i + 2], v.z, a); a = fmaf(q[4 * i + 3], v.w, a);
      }
      s = a * scale;
    }
    float cm = s;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, of));
    const float mn = fmaxf(m, cm);
    const float corr = expf(m - mn);
    const float e = key <= t ? expf(s - mn) : 0.f;
    float cs = e;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, of);
    l = l * corr + cs;
    o0 *= corr;
    o1 *= corr;
    const int nk = t - c0 + 1 < 32 ? t - c0 + 1 : 32;
    for (int jj = 0; jj < nk; ++jj) {
      const float p = __shfl_sync(0xffffffffu, e, jj);
      const float2 v = *reinterpret_cast<const float2*>(qkv + (long long)(c0 + jj) * 1536 + 1024 + h * 64 + lane * 2);
      o0 = fmaf(p, v.x, o0);
      o1 = fmaf(p, v.y, o1);
    }
    m = mn;
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<float2*>(out + (long long)tq * 512 + h * 64 + lane * 2) = make_float2(rtf32(o0 * inv, round), rtf32(o1 * inv, round));
}

// ---- windowed causal attention on the tensor cores (TF32 mma.sync, fp32 accumulate) -------------------------
// One CTA = 64 consecutive queries of one head, 4 warps x 16 query rows; key tiles of 64 rows are staged in shared
// memory.  Key tiles are aligned to ABSOLUTE positions (tile = absolute key position / 64; ``abs0`` = absolute
// position of buffer row 0), and every row folds the tiles it can see in ascending order with the online-softmax
// rule, so a row's result does not depend on which rows share its CTA or on where a streamed chunk starts: the
// chunked decode stays bit-identical to the one-shot decode.  S = Q K^T: m16n8k8, A = Q rows, B = K rows; the
// probabilities come out of the accumulator in exactly the A-fragment slots of the second product when its k
// slots are numbered (q, q+4) <-> keys (2q, 2q+1) of an 8-key group -- V's fragments use the same numbering.
constexpr int AT_LD = 68;  // padded row (floats) of the 64 x 64 tiles: conflict-free fragment loads
__device__ __forceinline__ uint32_t tf32u(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__global__ void __launch_bounds__(128) k_attn_window_tc(const float* __restrict__ qkv, int L, int context, float* __restrict__ out,
                                                        int hist, long long abs0, int round) {
  extern __shared__ __align__(16) float at_smem[];
  // blockIdx.z: utterance of a batched one-shot decode (hist == 0), L rows each
  qkv += (long long)blockIdx.z * L * 1536;
  out += (long long)blockIdx.z * L * 512;
  float* Qs = at_smem;                 // [64][AT_LD]
  float* Ks = Qs + 64 * AT_LD;
  float* Vs = Ks + 64 * AT_LD;
  const int h = blockIdx.y, tq0 = blockIdx.x * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int nq = min(64, L - tq0);
  // buffer rows: queries at hist + tq0 + i; a query at buffer row t sees keys [max(t - context + 1, 0), t]
  const int t_first = hist + tq0, t_last = t_first + nq - 1;
  const int key_lo = max(t_first - context + 1, 0);
  for (int u = tid; u < 64 * 16; u += 128) {
    const int r = u >> 4, c4 = (u & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nq) v = *reinterpret_cast<const float4*>(qkv + (long long)(t_first + r) * 1536 + h * 64 + c4);
    *reinterpret_cast<float4*>(Qs + r * AT_LD + c4) = v;
  }
  __syncthreads();
  uint32_t qa[8][4];  // A fragments of this warp's 16 rows, 8 k-steps of 8 dims
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const float* qp = Qs + (warp * 16 + g) * AT_LD + kk * 8 + q;
    qa[kk][0] = tf32u(qp[0]);
    qa[kk][1] = tf32u(qp[8 * AT_LD]);
    qa[kk][2] = tf32u(qp[4]);
    qa[kk][3] = tf32u(qp[8 * AT_LD + 4]);
  }
  const int t_lo = t_first + warp * 16 + g, t_hi = t_lo + 8;  // buffer rows of this lane's two query rows
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[d][e] = 0.f;
  const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64), scores in log2 units
  // absolute-aligned key tiles that intersect [key_lo, t_last]
  const long long kt_lo = (abs0 + key_lo) >> 6, kt_hi = (abs0 + t_last) >> 6;
  for (long long kt = kt_lo; kt <= kt_hi; ++kt) {
    const long long row0 = kt * 64 - abs0;  // buffer row of the tile's first key (may be negative)
    __syncthreads();                        // the previous tile has been consumed
    for (int u = tid; u < 64 * 16; u += 128) {
      const int r = u >> 4, c4 = (u & 15) * 4;
      const long long key = row0 + r;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (key >= key_lo && key <= t_last) {
        kv = *reinterpret_cast<const float4*>(qkv + key * 1536 + 512 + h * 64 + c4);
        vv = *reinterpret_cast<const float4*>(qkv + key * 1536 + 1024 + h * 64 + c4);
      }
      *reinterpret_cast<float4*>(Ks + r * AT_LD + c4) = kv;
      *reinterpret_cast<float4*>(Vs + r * AT_LD + c4) = vv;
    }
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float* kp = Ks + (j * 8 + g) * AT_LD + kk * 8 + q;
        const uint32_t b[2] = {tf32u(kp[0]), tf32u(kp[4])};
        mma_tf32(s[j], qa[kk], b);
      }
    }
    // window mask, running max
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const long long key = row0 + j * 8 + 2 * q + e;
        const bool ok_lo = key <= t_lo && key > (long long)t_lo - context && key >= 0 && t_lo <= t_last;
        const bool ok_hi = key <= t_hi && key > (long long)t_hi - context && key >= 0 && t_hi <= t_last;
        s[j][e] = ok_lo ? s[j][e] * sl2 : -INFINITY;
        s[j][2 + e] = ok_hi ? s[j][2 + e] * sl2 : -INFINITY;
        mx_lo = fmaxf(mx_lo, s[j][e]);
        mx_hi = fmaxf(mx_hi, s[j][2 + e]);
      }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    // a row that has not seen a key yet keeps max = -inf: its probabilities of this tile are all 0
    const float c_lo = mx_lo == -INFINITY ? 1.f : exp2f(m_lo - mx_lo), c_hi = mx_hi == -INFINITY ? 1.f : exp2f(m_hi - mx_hi);
    m_lo = mx_lo;
    m_hi = mx_hi;
    float ps_lo = 0.f, ps_hi = 0.f;
    uint32_t pa[8][4];  // P as A fragments of the 8-key groups: (a0,a1,a2,a3) = (c0,c2,c1,c3)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = m_lo == -INFINITY ? 0.f : exp2f(s[j][0] - m_lo), p1 = m_lo == -INFINITY ? 0.f : exp2f(s[j][1] - m_lo);
      const float p2 = m_hi == -INFINITY ? 0.f : exp2f(s[j][2] - m_hi), p3 = m_hi == -INFINITY ? 0.f : exp2f(s[j][3] - m_hi);
      ps_lo += p0 + p1;
      ps_hi += p2 + p3;
      pa[j][0] = tf32u(p0); pa[j][1] = tf32u(p2); pa[j][2] = tf32u(p1); pa[j][3] = tf32u(p3);
    }
    l_lo = l_lo * c_lo + ps_lo;
    l_hi = l_hi * c_hi + ps_hi;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      o[d][0] *= c_lo; o[d][1] *= c_lo;
      o[d][2] *= c_hi; o[d][3] *= c_hi;
    }
    // O += P V: k slots (q, q+4) of key group j are keys (8j + 2q, 8j + 2q + 1)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float* vp = Vs + (j * 8 + 2 * q) * AT_LD + d * 8 + g;
        const uint32_t b[2] = {tf32u(vp[0]), tf32u(vp[AT_LD])};
        mma_tf32(o[d], pa[j], b);
      }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    if (r_lo < nq)
      *reinterpret_cast<float2*>(out + (long long)(tq0 + r_lo) * 512 + h * 64 + d * 8 + 2 * q) =
          make_float2(rtf32(o[d][0] * i_lo, round), rtf32(o[d][1] * i_lo, round));
    if (r_hi < nq)
      *reinterpret_cast<float2*>(out + (long long)(tq0 + r_hi) * 512 + h * 64 + d * 8 + 2 * q) =
          make_float2(rtf32(o[d][2] * i_hi, round), rtf32(o[d][3] * i_hi, round));
  }
}

// final causal conv 64 -> 1, k = 3, with the ELU on its input: one thread per output sample
// pre_elu: x already holds ELU(u) (the tensor-core path stores activations that way)
__global__ void k_final_conv(const float* __restrict__ x /*[L, 64], 2 zero pad rows in front*/, const float* __restrict__ w
                             /*[3][64] tap-major*/, float bias, long long L, float* __restrict__ y, int pre_elu = 0) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float4* xp = reinterpret_cast<const float4*>(x + (t - 2) * 64);
  const float4* wp = reinterpret_cast<const float4*>(w);
  float a = bias;
#pragma unroll 8
  for (int i = 0; i < 48; ++i) {
    const float4 v = xp[i], ww = wp[i];
    if (pre_elu) {
      a = fmaf(v.x, ww.x, a); a = fmaf(v.y, ww.y, a);
      a = fmaf(v.z, ww.z, a); a = fmaf(v.w, ww.w, a);
    } else {
      a = fmaf(elu1(v.x), ww.x, a); a = fmaf(elu1(v.y), ww.y, a);
      a = fmaf(elu1(v.z), ww.z, a); a = fmaf(elu1(v.w), ww.w, a);
    }
  }
  y[t] = a;
}

// ---- encode side --------------------------------------------------------------------------------
// first encoder conv: 1 -> 64 channels, k = 7, causal; x is the waveform with >= 6 zero samples in front
__global__ void k_enc_conv0(const float* __restrict__ x, const float* __restrict__ w /*[64][7]*/, const float* __restrict__ b,
                            long long L, float* __restrict__ y /*[L, 64]*/) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= L * 64) return;
  const int co = i & 63;
  const long long t = i >> 6;
  float a = b[co];
#pragma unroll
  for (int k = 0; k < 7; ++k) a = fmaf(w[co * 7 + k], x[t - 6 + k], a);
  y[i] = a;
}
// replicate padding for the stride-2 downsample: rows -1 and -2 of xs become copies of row 0 (mode 1),
// or are cleared again afterwards (mode 0) because the decoder expects zero pad rows there
__global__ void k_pad_rows(float* __restrict__ xs, int C, int mode) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float v = mode ? xs[c] : 0.f;
  xs[c - C] = v;
  xs[c - 2 * C] = v;
}
// |e_j|^2 of every codebook entry
__global__ void k_row_sumsq256(const float* __restrict__ emb, long long rows, float* __restrict__ out) {
  const long long r = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  for (int d = lane; d < 256; d += 32) {
    const float v = emb[r * 256 + d];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = s;
}
// nearest centroid of one RVQ layer: argmin_j(|e_j|^2 - 2 x.e_j) (first index on ties, as torch.argmin),
// writes the code and subtracts the chosen centroid from the running residual
__global__ void __launch_bounds__(256) k_rvq_argmin(const float* __restrict__ dots /*[T, 2048]*/,
                                                    const float* __restrict__ enorm /*[2048]*/,
                                                    const float* __restrict__ emb /*[2048, 256]*/, float* __restrict__ res
                                                    /*[T, 256]*/, int64_t* __restrict__ codes /*[T]*/) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int j = tid; j < 2048; j += 256) {
    const float d = enorm[j] - 2.0f * dots[(long long)t * 2048 + j];
    if (d < best) {
      best = d;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  best = sv[0];
  bi = si[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (sv[w] < best || (sv[w] == best && si[w] < bi)) {
      best = sv[w];
      bi = si[w];
    }
  if (tid == 0) codes[t] = bi;
  res[(long long)t * 256 + tid] -= emb[(long long)bi * 256 + tid];
}

// ---- weight packing (create time) ---------------------------------------------------------------
__global__ void k_pack_embedding(const float* __restrict__ esum, const float* __restrict__ usage, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // [2048][256]
  if (i >= 2048LL * 256) return;
  out[i] = esum[i] / fmaxf(usage[i >> 8], 1e-5f);
}
// Conv1d weight [Cout][Cin][k] -> [Cout][k][Cin]
__global__ void k_round_copy(const float* w, float* out, long long n) {  // (in place when out == w)
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = rtf32(w[i], 1);
}
__global__ void k_pack_conv(const float* __restrict__ w, int Cout, int Cin, int k, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * k) return;
  const int ci = i % Cin;
  const int kk = (i / Cin) % k;
  const long long co = i / ((long long)Cin * k);
  out[i] = w[(co * Cin + ci) * k + kk];
}
// ConvTranspose1d weight [Cin][Cout][2s] -> B[n = r*Cout + co][K = (x[q-1] taps r+s | x[q] taps r)]
__global__ void k_pack_convtr(const float* __restrict__ w, int Cin, int Cout, int s, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)s * Cout * 2 * Cin;
  if (i >= total) return;
  const int kidx = i % (2 * Cin);
  const long long n = i / (2 * Cin);
  const int co = n % Cout, r = n / Cout;
  const int ci = kidx % Cin;
  const int tap = kidx < Cin ? r + s : r;
  out[i] = w[((long long)ci * Cout + co) * (2 * s) + tap];
}
// [Wf | Wr]: out[n][0..255] = Wf[n][:], out[n][256..511] = Wr[n][:]   (1x1 convs [512][256][1])
__global__ void k_pack_rvq_proj(const float* __restrict__ wf, const float* __restrict__ wr, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 512 * 512) return;
  const int n = i / 512, k = i % 512;
  out[i] = k < 256 ? wf[n * 256 + k] : wr[n * 256 + k - 256];
}

}  // namespace mimi
