// Skinny GEMM for batched decode (2..64 activation rows): Y[n, r] = sum_k X[n, k] W[r, k], bf16 in,
// fp32 accumulate, on the tensor cores, reading the FRAGMENT-MAJOR packed weights of the megakernel
// (mega.cuh / k_pack_frag).  A 128-row tcgen05 tile is the wrong shape here: a 1024-row projection is
// 8 CTAs and the step is occupancy bound; this kernel runs one CTA per row group of R = 8 / 16 rows
// (64..1024 CTAs per projection), whose 8 warps split K, stream their slice of the group with
// coalesced 16-byte loads (every weight byte is read once per launch), and feed mma.m16n8k16 with
// the streams as the other operand (8 per tile, up to 8 tiles).  Partial sums meet in shared memory
// in a fixed order; the epilogues are those of the tcgen05 GEMM (store / + residual / SwiGLU pairs).
#pragma once
#include "common.cuh"

namespace sk {

enum { EPI_STORE = 0, EPI_ADD_RESID = 1, EPI_SWIGLU_PAIRS = 2, EPI_ROPE_KV = 3 };

struct Args {
  const bf16* Wf;  // fragment-major packed [G][K/32][R*64 bytes]
  int R, K, n_out, G;
  const bf16* X;
  long long ldx;
  int N;
  bf16* out;
  long long ldo;
  const bf16* resid;
  int epi;
  const bf16* norm_scale;  // NORM: X rows are RMS-normalised on the fly, bf16(bf16(x * inv) * scale) like k_rmsnorm
  float eps;
  // EPI_ROPE_KV (the fused [q;k;v] projection): same arithmetic as k_rope_kv_rows -- rotate q and k row pairs,
  // q -> out [N, heads*hd], k / v -> the cache at the row's slot
  const bf16* rope;
  const int *row_stream, *row_pos, *row_slot;  // null: stream n % imp_B, position = slot = imp_pos + n / imp_B
  int imp_B, imp_pos, heads, kv_heads, hd, slots;
  bf16 *k_cache, *v_cache;
};

__device__ __forceinline__ float sk_sumsq8(const uint4& v) {
  float s = 0.f, a;
  a = bflo(v.x); s = fmaf(a, a, s); a = bfhi(v.x); s = fmaf(a, a, s);
  a = bflo(v.y); s = fmaf(a, a, s); a = bfhi(v.y); s = fmaf(a, a, s);
  a = bflo(v.z); s = fmaf(a, a, s); a = bfhi(v.z); s = fmaf(a, a, s);
  a = bflo(v.w); s = fmaf(a, a, s); a = bfhi(v.w); s = fmaf(a, a, s);
  return s;
}
__device__ __forceinline__ uint32_t sk_norm2(uint32_t x, float inv, uint32_t sc) {
  __nv_bfloat162 o = __floats2bfloat162_rn(rbf(bflo(x) * inv) * bflo(sc), rbf(bfhi(x) * inv) * bfhi(sc));
  return *reinterpret_cast<uint32_t*>(&o);
}
__device__ __forceinline__ uint4 sk_norm8(const uint4& v, float inv, const uint4& sc) {
  return make_uint4(sk_norm2(v.x, inv, sc.x), sk_norm2(v.y, inv, sc.y), sk_norm2(v.z, inv, sc.z), sk_norm2(v.w, inv, sc.w));
}

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NT, bool NORM>  // stream tiles of 8; fused RMSNorm of the activation rows
__global__ void __launch_bounds__(256) k_skinny(Args a) {
  __shared__ __align__(16) float psum[8][16][NT * 8 + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int grp = blockIdx.x, R = a.R, K = a.K;
  const bool r16 = R == 16;
  const int ks = K >> 3, nblk = ks >> 5;  // this warp's k slice, in 32-wide blocks
  const unsigned char* wp = reinterpret_cast<const unsigned char*>(a.Wf) + ((size_t)grp * R * K + (size_t)warp * R * ks) * 2 + lane * 16;
  const int blk = R * 64;
  // Programmatic dependent launch: this CTA's weight slice (R * K bf16, contiguous) does not depend on the
  // previous kernel -- pull it into L2 while that kernel drains, then wait for the activations.
  {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(a.Wf) + (size_t)grp * R * K * 2;
    for (int off = tid * 128; off < R * K * 2; off += 256 * 128) prefetch_l2(base + off);
  }
  pdl_wait();
  pdl_trigger();
  // x fragment rows of this lane: stream t*8 + g (clamped: rows >= N feed outputs nobody reads)
  const bf16* xp[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int n = t * 8 + g;
    xp[t] = a.X + (size_t)(n < a.N ? n : 0) * a.ldx + warp * ks + q * 8;
  }
  float inv[NT];
  if (NORM) {
    // sum of squares of every stream row: each warp over its k slice, the four lanes of a quad and then
    // the eight warps through shared memory (psum doubles as scratch before the main loop)
    float* ss = &psum[0][0][0];  // [8 warps][NT * 8]
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      float s = 0.f;
      for (int b = 0; b < nblk; ++b) s += sk_sumsq8(*reinterpret_cast<const uint4*>(xp[t] + b * 32));
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) ss[warp * (NT * 8) + t * 8 + g] = s;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += ss[w * (NT * 8) + t * 8 + g];
      inv[t] = 1.0f / sqrtf(s / (float)K + a.eps);
    }
    __syncthreads();
  }
  const bf16* scp = NORM ? a.norm_scale + warp * ks + q * 8 : nullptr;
  float acc[NT][2][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][h][e] = 0.f;
  // four blocks per round: the weight loads (HBM latency) all go out first, the activation loads
  // (L1 / L2 hits, shared by every CTA) follow block by block
#pragma unroll 1
  for (int b0 = 0; b0 < nblk; b0 += 4) {
    uint4 w0[4], w1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool on = b0 + i < nblk;
      w0[i] = on ? ld_stream(wp + (size_t)(b0 + i) * blk) : make_uint4(0, 0, 0, 0);
      w1[i] = (on && r16) ? ld_stream(wp + (size_t)(b0 + i) * blk + 512) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + i < nblk ? b0 + i : b0;
      uint4 sc = make_uint4(0, 0, 0, 0);
      if (NORM) sc = *reinterpret_cast<const uint4*>(scp + b * 32);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        uint4 xv = *reinterpret_cast<const uint4*>(xp[t] + b * 32);
        if (NORM) xv = sk_norm8(xv, inv[t], sc);
        if (r16) {
          // weights = A operand (pre-packed quads), streams = B columns
          mma(acc[t][0], w0[i].x, w0[i].y, w0[i].z, w0[i].w, xv.x, xv.y);
          mma(acc[t][1], w1[i].x, w1[i].y, w1[i].z, w1[i].w, xv.z, xv.w);
        } else {
          // weights = B operand (8 rows), streams = A rows 0-7; the quad (P0,P2,P1,P3) serves both halves
          mma(acc[t][0], xv.x, xv.z, xv.y, xv.w, w0[i].x, w0[i].y);
          mma(acc[t][1], xv.x, xv.z, xv.y, xv.w, w0[i].z, w0[i].w);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    if (r16) {  // lane (g, q): rows 2g, 2g+1 x streams t*8 + 2q, +1
      psum[warp][2 * g][t * 8 + 2 * q] = acc[t][0][0] + acc[t][1][0];
      psum[warp][2 * g][t * 8 + 2 * q + 1] = acc[t][0][1] + acc[t][1][1];
      psum[warp][2 * g + 1][t * 8 + 2 * q] = acc[t][0][2] + acc[t][1][2];
      psum[warp][2 * g + 1][t * 8 + 2 * q + 1] = acc[t][0][3] + acc[t][1][3];
    } else {    // lane (g, q): stream t*8 + g x rows 2q, 2q+1
      psum[warp][2 * q][t * 8 + g] = acc[t][0][0] + acc[t][1][2];
      psum[warp][2 * q + 1][t * 8 + g] = acc[t][0][1] + acc[t][1][3];
    }
  }
  __syncthreads();
  const int pairs = R >> 1;
  for (int i = tid; i < pairs * a.N; i += 256) {
    const int p = i / a.N, n = i - p * a.N;
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {  // fixed order
      y0 += psum[w][2 * p][n];
      y1 += psum[w][2 * p + 1][n];
    }
    const int r0 = grp * R + 2 * p;
    if (r0 >= a.n_out) continue;
    y0 = rbf(y0);
    y1 = rbf(y1);
    if (a.epi == EPI_ROPE_KV) {
      const int hd = a.hd, qrows = a.heads * hd, krows = a.kv_heads * hd;
      const int pos = a.row_stream ? a.row_pos[n] : a.imp_pos + n / a.imp_B;
      const int slot = a.row_stream ? a.row_slot[n] : a.imp_pos + n / a.imp_B;
      const int stream = a.row_stream ? a.row_stream[n] : n % a.imp_B;
      float o0 = y0, o1 = y1;
      if (r0 < qrows + krows) {
        const __nv_bfloat162 cs = *reinterpret_cast<const __nv_bfloat162*>(a.rope + ((size_t)pos * (hd / 2) + ((r0 % hd) >> 1)) * 2);
        const float c = __low2float(cs), sn = __high2float(cs);
        o0 = rbf(__fsub_rn(__fmul_rn(y0, c), __fmul_rn(y1, sn)));
        o1 = rbf(__fadd_rn(__fmul_rn(y1, c), __fmul_rn(y0, sn)));
      }
      const __nv_bfloat162 ov = __floats2bfloat162_rn(o0, o1);
      if (r0 < qrows) {
        *reinterpret_cast<__nv_bfloat162*>(a.out + (size_t)n * qrows + r0) = ov;
      } else {
        const bool isk = r0 < qrows + krows;
        const int rr = r0 - (isk ? qrows : qrows + krows);
        const int kvh = rr / hd, d = rr % hd;
        bf16* dst = (isk ? a.k_cache : a.v_cache) + (((size_t)stream * a.kv_heads + kvh) * a.slots + slot) * hd + d;
        *reinterpret_cast<__nv_bfloat162*>(dst) = ov;
      }
    } else if (a.epi == EPI_SWIGLU_PAIRS) {
      a.out[(size_t)n * a.ldo + (r0 >> 1)] = f2bf(silu_bf(y0) * y1);
    } else {
      bf16* o = a.out + (size_t)n * a.ldo + r0;
      if (a.epi == EPI_ADD_RESID) {
        y0 += bf2f(a.resid[(size_t)n * a.ldo + r0]);
        if (r0 + 1 < a.n_out) y1 += bf2f(a.resid[(size_t)n * a.ldo + r0 + 1]);
      }
      o[0] = f2bf(y0);
      if (r0 + 1 < a.n_out) o[1] = f2bf(y1);
    }
  }
}

}  // namespace sk
