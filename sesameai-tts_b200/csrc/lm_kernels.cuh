// CSM-1B language-model kernels, small-row (decode / short prefill chunk) path.
//
// Every weight matrix is streamed from HBM exactly once per launch with 16-byte coalesced loads
// (one warp per output-row pair, warp-shuffle reduction); the activation rows (<= 8 per launch
// tile) sit in shared memory.  Norm, RoPE, KV append, SwiGLU and the residual add are fused into
// the prologue / epilogue of the GEMV that produces or consumes them, at the reference's bf16
// rounding points.
#pragma once
#include "common.cuh"

// Per-call parameters, written by k_set_params immediately before the (captured) kernel chain.
struct FrameParams {
  const int64_t* tokens;   // [B, S, C+1]
  const uint8_t* mask;     // [B, S, C+1]
  const int64_t* pos;      // [B, S]
  int32_t* out;            // [B, C]
  const bf16* noise;       // [C, B, V] or null
  const int32_t* forced;   // [B, C] or null
  bf16* logits_out;        // [C, B, V] or null
  int32_t* sampled_out;    // [B, C] or null
  unsigned long long seed, offset;
  float temperature;
  int topk;
  int B, S;
  int cache_len;           // backbone positions already in the cache before this call (lock-step batch: every stream)
  int s0;                  // first prompt row handled by the current prefill chunk
  // Continuous batching: batch row b runs on cache lane lane_meta[b] which already holds lane_meta[B + b]
  // positions.  Null: row b = lane b, every lane holds cache_len positions (the reference's lock-step batch).
  const int* lane_meta;
};
__device__ __forceinline__ int row_lane(const FrameParams* P, int b) { return P->lane_meta ? P->lane_meta[b] : b; }
__device__ __forceinline__ int row_len(const FrameParams* P, int b) { return P->lane_meta ? P->lane_meta[P->B + b] : P->cache_len; }

// lane table of one call, passed in kernel-parameter space (no host buffer has to outlive the call)
struct LaneChunk {
  int n, off, B;
  int lane[256], len[256];
};
__global__ void k_set_lanes(int* meta, LaneChunk c) {
  const int i = threadIdx.x;
  if (i < c.n) {
    meta[c.off + i] = c.lane[i];
    meta[c.B + c.off + i] = c.len[i];
  }
}

struct DevStatus;
__global__ void k_set_params(FrameParams* dst, FrameParams v, DevStatus* st_reset = nullptr);

// Device-side status of a context: sticky error code of the current call, mirrored into a mapped host
// word so that the host can look at it without a CUDA call (csm_check_error).  Codes:
//   0x100-0x4ff  a wait of the decode megakernel gave up (trip cap): the launch drained, tokens are garbage
//   0x801        token id outside its embedding table        (the reference raises IndexError there)
//   0x802        teacher-forced token id outside the audio vocabulary
//   0x803        input_pos outside the RoPE table, or != cache position (the reference's mask row / cache slot
//                coincide only for sequential use from a reset state; anything else is rejected, not guessed)
struct DevStatus {
  unsigned int seq;    // frame counter of the megakernel (tag salt)
  unsigned int error;  // first error code of this call (0 = none)
  unsigned int* host_error;  // mapped host mirror, sticky until csm_check_error clears it
  unsigned int* att_part;    // megakernel: tagged partial results of the split long-context attention (mega.cuh: attn_split)
};
__global__ void k_set_params(FrameParams* dst, FrameParams v, DevStatus* st_reset) {
  pdl_wait();
  pdl_trigger();
  *dst = v;
  if (st_reset) st_reset->error = 0;  // per call; the host mirror stays sticky until csm_check_error reads it
}
__device__ __noinline__ void report_error(DevStatus* st, unsigned code) {
  if (st && atomicCAS(&st->error, 0u, code) == 0u && st->host_error) {
    *reinterpret_cast<volatile unsigned int*>(st->host_error) = code;
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------------
// K1: _embed_tokens + mask-mul + sum  (sesameai/models.py:155-157,193-203)
// h[n,:] = sum_{c<C} m_c * A[tok_c + V*c] + m_C * T[tok_C], fp32 accumulate in column order,
// one rounding to bf16.  One CTA per frame row, 16-byte gathers.
// ---------------------------------------------------------------------------------------------
// token id of column c, range-checked (ids the reference would raise on are reported and read as 0)
__device__ __forceinline__ size_t checked_token(const int64_t* tok, int c, int C, int V, int TV, DevStatus* st) {
  const int64_t t = tok[c];
  const int64_t lim = c < C ? V : TV;
  if (t < 0 || (lim > 0 && t >= lim)) {
    report_error(st, 0x801);
    return 0;
  }
  return (size_t)t;
}
__device__ __forceinline__ void embed_row(const int64_t* tok, const uint8_t* msk, const bf16* text_emb,
                                          const bf16* audio_emb, int C, int V, int D, bf16* out, int TV = 0,
                                          DevStatus* st = nullptr) {
  for (int d8 = threadIdx.x; d8 < D / 8; d8 += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int c = 0; c <= C; ++c) {
      if (!msk[c]) continue;  // masked slots contribute row*0 (exact for finite weights)
      const size_t t = checked_token(tok, c, C, V, TV, st);
      const bf16* row = (c < C) ? audio_emb + (t + (size_t)V * c) * D : text_emb + t * D;
      uint4 v = *reinterpret_cast<const uint4*>(row + d8 * 8);
      acc[0] += bflo(v.x); acc[1] += bfhi(v.x); acc[2] += bflo(v.y); acc[3] += bfhi(v.y);
      acc[4] += bflo(v.z); acc[5] += bfhi(v.z); acc[6] += bflo(v.w); acc[7] += bfhi(v.w);
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
    *reinterpret_cast<uint4*>(out + d8 * 8) = *reinterpret_cast<uint4*>(o);
  }
}

__global__ void k_embed_frames(const int64_t* tokens, const uint8_t* mask, const bf16* text_emb,
                               const bf16* audio_emb, int C, int V, int D, bf16* out) {
  const int n = blockIdx.x;
  embed_row(tokens + (size_t)n * (C + 1), mask + (size_t)n * (C + 1), text_emb, audio_emb, C, V, D,
            out + (size_t)n * D);
}

// Backbone input rows for one pass: rows n = b*chunk + t  <->  prompt frame s = s0 + t of stream b.
// Also emits the row metadata (stream, RoPE position, cache slot) the layer kernels use.
__global__ void k_embed_pass(const FrameParams* __restrict__ P, const bf16* text_emb, const bf16* audio_emb, int C,
                             int V, int D, int chunk, bf16* h, int* row_stream, int* row_pos, int* row_slot, int TV,
                             int rope_len, DevStatus* st) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x;
  const int b = n / chunk, t = n % chunk;
  const int s = P->s0 + t;
  const size_t fr = (size_t)b * P->S + s;
  embed_row(P->tokens + fr * (C + 1), P->mask + fr * (C + 1), text_emb, audio_emb, C, V, D, h + (size_t)n * D, TV, st);
  if (threadIdx.x == 0) {
    const int slot = row_len(P, b) + s;
    int64_t pos = P->pos[fr];
    // the key range of a row is its cache slot (key <= slot); the reference masks by input_pos: they agree
    // exactly when input_pos == cache position, the only use the reference makes of it
    if (pos != slot || pos < 0 || pos >= rope_len) {
      report_error(st, 0x803);
      pos = slot;
    }
    row_stream[n] = row_lane(P, b);
    row_pos[n] = (int)pos;
    row_slot[n] = slot;
  }
}

// ---------------------------------------------------------------------------------------------
// torchtune RMSNorm (Appendix A.3) on NB rows already resident in smem as raw bf16:
// xn = bf16( bf16(x * rsqrt(mean(x^2)+eps)) * scale ), written back in place.
// ---------------------------------------------------------------------------------------------
template <int NB, int NT, int BAR>
__device__ __forceinline__ void rmsnorm_smem(bf16* xs, int K, const bf16* __restrict__ scale, float eps,
                                             float* scratch, int tid) {
  // 8-element (16-byte) units per thread; one CTA-wide reduction for all NB rows
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int nw = NT / 32;
  float ss[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) ss[nb] = 0.f;
  for (int u = tid; u < K / 8; u += NT) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const uint4 v = *reinterpret_cast<const uint4*>(xs + nb * K + u * 8);
      float a;
      a = bflo(v.x); ss[nb] = fmaf(a, a, ss[nb]); a = bfhi(v.x); ss[nb] = fmaf(a, a, ss[nb]);
      a = bflo(v.y); ss[nb] = fmaf(a, a, ss[nb]); a = bfhi(v.y); ss[nb] = fmaf(a, a, ss[nb]);
      a = bflo(v.z); ss[nb] = fmaf(a, a, ss[nb]); a = bfhi(v.z); ss[nb] = fmaf(a, a, ss[nb]);
      a = bflo(v.w); ss[nb] = fmaf(a, a, ss[nb]); a = bfhi(v.w); ss[nb] = fmaf(a, a, ss[nb]);
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    ss[nb] = warp_sum(ss[nb]);
    if (lane == 0) scratch[nb * nw + warp] = ss[nb];
  }
  csync<NT, BAR>();
  float inv[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < nw; ++w) t += scratch[nb * nw + w];
    inv[nb] = 1.0f / sqrtf(t / (float)K + eps);
  }
  for (int u = tid; u < K / 8; u += NT) {
    const uint4 sc = *reinterpret_cast<const uint4*>(scale + u * 8);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      uint4 v = *reinterpret_cast<const uint4*>(xs + nb * K + u * 8);
      __nv_bfloat162 o[4];
      o[0] = __floats2bfloat162_rn(rbf(bflo(v.x) * inv[nb]) * bflo(sc.x), rbf(bfhi(v.x) * inv[nb]) * bfhi(sc.x));
      o[1] = __floats2bfloat162_rn(rbf(bflo(v.y) * inv[nb]) * bflo(sc.y), rbf(bfhi(v.y) * inv[nb]) * bfhi(sc.y));
      o[2] = __floats2bfloat162_rn(rbf(bflo(v.z) * inv[nb]) * bflo(sc.z), rbf(bfhi(v.z) * inv[nb]) * bfhi(sc.z));
      o[3] = __floats2bfloat162_rn(rbf(bflo(v.w) * inv[nb]) * bflo(sc.w), rbf(bfhi(v.w) * inv[nb]) * bfhi(sc.w));
      *reinterpret_cast<uint4*>(xs + nb * K + u * 8) = *reinterpret_cast<uint4*>(o);
    }
  }
  csync<NT, BAR>();
}

// ---------------------------------------------------------------------------------------------
// GEMV family: y[n, r] = sum_k W[r, k] * x[n, k] for NB activation rows per CTA tile.
// One warp owns an output-row PAIR (2i, 2i+1): RoPE rotates such pairs, and the gate/up matrices
// are interleaved so a pair is (gate_i, up_i).
// ---------------------------------------------------------------------------------------------
enum { EPI_PLAIN = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_ROPE_KV = 3 };

struct GemvArgs {
  const bf16* W;  // [rows, K] row-major
  int rows, K;
  const bf16* x;  // [N, ldx]
  int ldx, N;
  const bf16* norm_scale;  // NORM prologue
  float eps;
  bf16* out;  // PLAIN: [N, ldo] (rows cols); RESID: [N, ldo]; SWIGLU: [N, ldo] (rows/2 cols)
  int ldo;
  const bf16* resid;  // RESID: [N, ldr]
  int ldr;
  // EPI_ROPE_KV
  bf16* q_out;  // [N, heads*hd]
  bf16* k_cache;  // [streams, kv_heads, slots, hd]
  bf16* v_cache;
  const bf16* rope;  // [max_pos, hd/2, 2]
  const int* row_stream;  // null -> implicit: stream = n % imp_B, pos = slot = imp_pos + n / imp_B
  const int* row_pos;
  const int* row_slot;
  int imp_B, imp_pos;
  int heads, kv_heads, hd, slots;
};

template <int NB, int EPI, bool NORM>
__global__ void __launch_bounds__(256) k_gemv(GemvArgs a) {  // launched with exactly 256 threads
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float scratch[72];
  const int K = a.K;
  const int n0 = blockIdx.y * NB;
  pdl_wait();
  pdl_trigger();

  // stage the activation rows (zero-fill rows past N)
  for (int i = threadIdx.x; i < NB * (K / 8); i += blockDim.x) {
    const int nb = i / (K / 8), k8 = i % (K / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n0 + nb < a.N) v = *reinterpret_cast<const uint4*>(a.x + (size_t)(n0 + nb) * a.ldx + k8 * 8);
    *reinterpret_cast<uint4*>(xs + nb * K + k8 * 8) = v;
  }
  __syncthreads();
  if (NORM) rmsnorm_smem<NB, 256, 0>(xs, K, a.norm_scale, a.eps, scratch, threadIdx.x);

  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int npairs = (a.rows + 1) >> 1;
  for (int p = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); p < npairs; p += gridDim.x * warps_per_cta) {
    const int r0 = 2 * p, r1 = 2 * p + 1;
    const bool has1 = r1 < a.rows;
    const bf16* w0 = a.W + (size_t)r0 * K;
    const bf16* w1 = a.W + (size_t)(has1 ? r1 : r0) * K;
    float acc0[NB], acc1[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) acc0[nb] = acc1[nb] = 0.f;
#pragma unroll 4
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 a0 = ld_stream(w0 + k);
      const uint4 a1 = ld_stream(w1 + k);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + nb * K + k);
        acc0[nb] = dot8(a0, xv, acc0[nb]);
        acc1[nb] = dot8(a1, xv, acc1[nb]);
      }
    }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      acc0[nb] = warp_sum(acc0[nb]);
      acc1[nb] = warp_sum(acc1[nb]);
    }
    if (lane == 0) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int n = n0 + nb;
        if (n >= a.N) break;
        const float y0 = rbf(acc0[nb]), y1 = rbf(acc1[nb]);  // nn.Linear output is bf16
        if (EPI == EPI_PLAIN) {
          a.out[(size_t)n * a.ldo + r0] = f2bf(y0);
          if (has1) a.out[(size_t)n * a.ldo + r1] = f2bf(y1);
        } else if (EPI == EPI_RESID) {
          a.out[(size_t)n * a.ldo + r0] = f2bf(y0 + bf2f(a.resid[(size_t)n * a.ldr + r0]));
          if (has1) a.out[(size_t)n * a.ldo + r1] = f2bf(y1 + bf2f(a.resid[(size_t)n * a.ldr + r1]));
        } else if (EPI == EPI_SWIGLU) {
          // pair = (gate_p, up_p):  bf16( bf16(silu(gate)) * up )
          a.out[(size_t)n * a.ldo + p] = f2bf(silu_bf(y0) * y1);
        } else {  // EPI_ROPE_KV
          const int hd = a.hd, qrows = a.heads * hd, krows = a.kv_heads * hd;
          float o0 = y0, o1 = y1;
          const int m_stream = a.row_stream ? a.row_stream[n] : n % a.imp_B;
          const int m_pos = a.row_stream ? a.row_pos[n] : a.imp_pos + n / a.imp_B;
          const int m_slot = a.row_stream ? a.row_slot[n] : a.imp_pos + n / a.imp_B;
          if (r0 < qrows + krows) {
            const int j = (r0 % hd) >> 1;
            const bf16* cs = a.rope + ((size_t)m_pos * (hd / 2) + j) * 2;
            const float c = bf2f(cs[0]), s = bf2f(cs[1]);
            // fp32, un-fused, exactly as the reference evaluates it (Appendix A.5)
            o0 = rbf(__fsub_rn(__fmul_rn(y0, c), __fmul_rn(y1, s)));
            o1 = rbf(__fadd_rn(__fmul_rn(y1, c), __fmul_rn(y0, s)));
          }
          if (r0 < qrows) {
            a.q_out[(size_t)n * qrows + r0] = f2bf(o0);
            a.q_out[(size_t)n * qrows + r1] = f2bf(o1);
          } else {
            const bool isk = r0 < qrows + krows;
            const int rr = r0 - (isk ? qrows : qrows + krows);
            const int kvh = rr / hd, d = rr % hd;
            bf16* dst = (isk ? a.k_cache : a.v_cache) +
                        (((size_t)m_stream * a.kv_heads + kvh) * a.slots + m_slot) * hd + d;
            dst[0] = f2bf(o0);
            dst[1] = f2bf(o1);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3: attention over the GQA-compact KV cache, keys [0 .. slot] of the row's stream
// (== the reference's bool-mask row over the zero-padded cache).
// One CTA per (row, q-head); fp32 scores/softmax, bf16 output.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128) k_attn_rows(const bf16* __restrict__ q, const bf16* __restrict__ k_cache,
                                                   const bf16* __restrict__ v_cache, const int* __restrict__ row_stream,
                                                   const int* __restrict__ row_slot, int imp_B, int imp_pos, int heads,
                                                   int kv_heads, int slots, float scale, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);  // [slots]
  __shared__ float scratch[33];
  __shared__ float part[128];
  const int n = blockIdx.x, h = blockIdx.y;
  const int kvh = h / (heads / kv_heads);
  const int nkeys = (row_stream ? row_slot[n] : imp_pos + n / imp_B) + 1;
  const size_t base = ((size_t)(row_stream ? row_stream[n] : n % imp_B) * kv_heads + kvh) * slots * HD;
  const bf16* kp = k_cache + base;
  const bf16* vp = v_cache + base;

  // q in registers (fp32)
  float qf[HD];
  {
    const bf16* qr = q + ((size_t)n * heads + h) * HD;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      uint4 v = *reinterpret_cast<const uint4*>(qr + i * 8);
      qf[i * 8 + 0] = bflo(v.x); qf[i * 8 + 1] = bfhi(v.x); qf[i * 8 + 2] = bflo(v.y); qf[i * 8 + 3] = bfhi(v.y);
      qf[i * 8 + 4] = bflo(v.z); qf[i * 8 + 5] = bfhi(v.z); qf[i * 8 + 6] = bflo(v.w); qf[i * 8 + 7] = bfhi(v.w);
    }
  }
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < nkeys; j += blockDim.x) {
    const bf16* kr = kp + (size_t)j * HD;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      uint4 v = *reinterpret_cast<const uint4*>(kr + i * 8);
      s = fmaf(qf[i * 8 + 0], bflo(v.x), s); s = fmaf(qf[i * 8 + 1], bfhi(v.x), s);
      s = fmaf(qf[i * 8 + 2], bflo(v.y), s); s = fmaf(qf[i * 8 + 3], bfhi(v.y), s);
      s = fmaf(qf[i * 8 + 4], bflo(v.z), s); s = fmaf(qf[i * 8 + 5], bfhi(v.z), s);
      s = fmaf(qf[i * 8 + 6], bflo(v.w), s); s = fmaf(qf[i * 8 + 7], bfhi(v.w), s);
    }
    s *= scale;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_max<128, 0>(mx, scratch, threadIdx.x);
  float sum = 0.f;
  for (int j = threadIdx.x; j < nkeys; j += blockDim.x) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_sum<128, 0>(sum, scratch, threadIdx.x);  // (its barrier also publishes sc[])
  const float inv = 1.0f / sum;

  // P.V : thread -> (key group g, dim d); groups stride over keys
  constexpr int G = 128 / HD;  // 2 for hd 64, 1 for hd 128
  const int g = threadIdx.x / HD, d = threadIdx.x % HD;
  float acc = 0.f;
  for (int j = g; j < nkeys; j += G) acc = fmaf(sc[j], bf2f(vp[(size_t)j * HD + d]), acc);
  if (G > 1) {
    part[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0) {
#pragma unroll
      for (int gg = 1; gg < G; ++gg) acc += part[gg * HD + d];
    }
  }
  if (g == 0) out[((size_t)n * heads + h) * HD + d] = f2bf(acc * inv);
}

// RMSNorm of the row-batched path, one WARP per row (D = 1024 / 2048: the row lives in registers, one global read,
// no block barrier): k_rmsnorm -- a CTA per row, two passes, two block reductions -- was 6 us per launch whatever
// the row count, 312 launches per decode step (22 % of a 32-stream step).  Same arithmetic per element
// (bf16(bf16(x * inv) * scale)); the sum of squares is added in a different order than k_rmsnorm's.
template <int CH>  // D = CH * 256
__global__ void __launch_bounds__(256) k_rmsnorm_rows(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ scale,
                                                      float eps, bf16* __restrict__ y, int ldy, int N) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  constexpr int D = CH * 256;
  const bf16* xr = x + (size_t)n * ldx;
  uint4 v[CH];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    v[c] = *reinterpret_cast<const uint4*>(xr + c * 256 + lane * 8);
    const uint32_t w[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ss = fmaf(bflo(w[i]), bflo(w[i]), ss);
      ss = fmaf(bfhi(w[i]), bfhi(w[i]), ss);
    }
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / sqrtf(ss / (float)D + eps);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const uint4 sc = *reinterpret_cast<const uint4*>(scale + c * 256 + lane * 8);
    const uint32_t w[4] = {v[c].x, v[c].y, v[c].z, v[c].w}, g[4] = {sc.x, sc.y, sc.z, sc.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 p = __floats2bfloat162_rn(rbf(bflo(w[i]) * inv) * bflo(g[i]), rbf(bfhi(w[i]) * inv) * bfhi(g[i]));
      o[i] = *reinterpret_cast<const uint32_t*>(&p);
    }
    *reinterpret_cast<uint4*>(y + (size_t)n * ldy + c * 256 + lane * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// K3 (depth decoder, row-batched path): head_dim 128 over a cache of <= 32 positions.  k_attn_rows spends a CTA of 128
// threads on one (row, q-head): with <= 32 keys a quarter of them computes a 128-long dot product each from
// scattered 16-byte loads, and 2048 such CTAs per launch took 22 us at 256 streams (12 % of the decode step).
// Here one CTA serves a (row, KV head): the K / V rows are staged ONCE in shared memory (coalesced) for the
// heads / kv_heads q-heads that share them, one warp per q-head: lane j scores key j (K rows padded to 130
// bf16: conflict-free), the softmax is warp shuffles, lane l accumulates output dims 4l .. 4l+3.  Every sum runs
// in the order k_attn_rows uses (dims ascending, keys ascending, the same shuffle tree): bit-identical output.
// ---------------------------------------------------------------------------------------------
constexpr int AD_LD = 130;
__global__ void __launch_bounds__(256) k_attn_dec(const bf16* __restrict__ q, const bf16* __restrict__ k_cache,
                                                  const bf16* __restrict__ v_cache, const int* __restrict__ row_stream,
                                                  const int* __restrict__ row_slot, int imp_B, int imp_pos, int heads,
                                                  int kv_heads, int slots, float scale, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) bf16 Ks[32 * AD_LD];
  __shared__ __align__(16) bf16 Vs[32 * 128];
  __shared__ float qs[8][128];
  const int n = blockIdx.x, kvh = blockIdx.y, gq = heads / kv_heads;  // launched with 32 * gq threads, gq <= 8
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkeys = (row_stream ? row_slot[n] : imp_pos + n / imp_B) + 1;  // <= 32
  const size_t base = ((size_t)(row_stream ? row_stream[n] : n % imp_B) * kv_heads + kvh) * slots * 128;
  for (int u = threadIdx.x; u < nkeys * 16; u += blockDim.x) {
    const int r = u >> 4, c8 = (u & 15) * 8;
    const uint4 kv = *reinterpret_cast<const uint4*>(k_cache + base + (size_t)r * 128 + c8);
    uint32_t* kd = reinterpret_cast<uint32_t*>(Ks + r * AD_LD + c8);  // (4-byte aligned: AD_LD and c8 are even)
    kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
    *reinterpret_cast<uint4*>(Vs + r * 128 + c8) = *reinterpret_cast<const uint4*>(v_cache + base + (size_t)r * 128 + c8);
  }
  const int h = kvh * gq + warp;
  {
    const uint2 v = *reinterpret_cast<const uint2*>(q + ((size_t)n * heads + h) * 128 + lane * 4);
    qs[warp][lane * 4 + 0] = bflo(v.x); qs[warp][lane * 4 + 1] = bfhi(v.x);
    qs[warp][lane * 4 + 2] = bflo(v.y); qs[warp][lane * 4 + 3] = bfhi(v.y);
  }
  __syncthreads();
  float sc = -INFINITY;
  if (lane < nkeys) {
    const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + lane * AD_LD);
    const float* qr = qs[warp];
    float s = 0.f;
#pragma unroll 16
    for (int i = 0; i < 64; ++i) {
      const uint32_t kk = kr[i];
      s = fmaf(qr[2 * i], bflo(kk), s);
      s = fmaf(qr[2 * i + 1], bfhi(kk), s);
    }
    sc = s * scale;
  }
  const float mx = warp_max(sc);
  const float e = lane < nkeys ? expf(sc - mx) : 0.f;
  const float inv = 1.0f / warp_sum(e);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < nkeys; ++j) {
    const float p = __shfl_sync(0xffffffffu, e, j);
    const uint2 v = *reinterpret_cast<const uint2*>(Vs + j * 128 + lane * 4);
    acc[0] = fmaf(p, bflo(v.x), acc[0]); acc[1] = fmaf(p, bfhi(v.x), acc[1]);
    acc[2] = fmaf(p, bflo(v.y), acc[2]); acc[3] = fmaf(p, bfhi(v.y), acc[3]);
  }
  __nv_bfloat162 o[2] = {__floats2bfloat162_rn(acc[0] * inv, acc[1] * inv), __floats2bfloat162_rn(acc[2] * inv, acc[3] * inv)};
  *reinterpret_cast<uint2*>(out + ((size_t)n * heads + h) * 128 + lane * 4) = *reinterpret_cast<uint2*>(o);
}

// ---------------------------------------------------------------------------------------------
// K3 (prompt prefill): tiled causal attention on the tensor cores, head_dim 64.
// One CTA = 64 consecutive prompt rows of one stream x one q-head; 4 warps x 16 rows.  Key tiles of
// 64 cache rows are double-buffered in shared memory with cp.async; S = Q K^T and O += P V are
// mma.m16n8k16 (K fragments by ldmatrix, V fragments by ldmatrix.trans), the softmax is the online
// (running max / running sum) form in fp32 registers, P is rounded to bf16 for the second product.
// Rows: n = b * chunk + t, cache slot row_slot[n] (ascending in t), keys [0 .. slot] of stream b.
// ---------------------------------------------------------------------------------------------
constexpr int FA_LD = 72;  // padded row length (bf16) of the 64 x 64 tiles: conflict-free ldmatrix

__device__ __forceinline__ void fa_ldm4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void fa_ldm4t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void fa_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void fa_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t fa_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------
// K3 (backbone, row-batched decode at LONG context: a voice prompt).  k_attn_rows gives a CTA to one (row, q-head), a thread
// per key reading its 128-byte row by itself and a serial P.V loop over the keys -- 115 us per layer for 32 streams at
// 1568 keys, and 8 streams are only 256 CTAs.  Flash-decoding form: the keys of a (row, KV head) are cut into AS_SPLITS
// ranges, one CTA each (k_attn_split64_mma below), whose partial (o[64], m, l) go to ``part``; k_attn_combine64 adds
// the ranges of a (row, q-head) in range order.  Used from 256 keys on (host decision, a separate captured graph): below
// that the two-pass k_attn_rows stays, so short-context results are unchanged.
// ---------------------------------------------------------------------------------------------
#ifndef CSM_AS_SPLITS
#define CSM_AS_SPLITS 8
#endif
constexpr int AS_SPLITS = CSM_AS_SPLITS;
constexpr int AS_PW = 66;  // floats per partial: o[64], m (log2 units), l
// the key ranges of a (row, q-head) added in range order; one thread per output dim
__global__ void __launch_bounds__(64) k_attn_combine64(const float* __restrict__ part, int heads, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x, h = blockIdx.y, d = threadIdx.x;
  const float* pp = part + ((size_t)n * heads + h) * AS_SPLITS * AS_PW;
  float M = -INFINITY;
#pragma unroll
  for (int z = 0; z < AS_SPLITS; ++z) M = fmaxf(M, pp[z * AS_PW + 64]);
  float L = 0.f, O = 0.f;
#pragma unroll
  for (int z = 0; z < AS_SPLITS; ++z) {
    const float w = exp2f(pp[z * AS_PW + 64] - M);  // (an empty range: exp2(-inf) = 0)
    L = fmaf(pp[z * AS_PW + 65], w, L);
    O = fmaf(pp[z * AS_PW + d], w, O);
  }
  out[((size_t)n * heads + h) * 64 + d] = f2bf(O / L);
}

__global__ void __launch_bounds__(128) k_attn_flash64(const bf16* __restrict__ q, const bf16* __restrict__ k_cache,
                                                      const bf16* __restrict__ v_cache, const int* __restrict__ row_slot, int chunk,
                                                      int heads, int kv_heads, int slots, float scale, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) bf16 Qs[64 * FA_LD];
  __shared__ __align__(16) bf16 Ks[2][64 * FA_LD];
  __shared__ __align__(16) bf16 Vs[2][64 * FA_LD];
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, qd = lane & 3;
  const int kvh = h / (heads / kv_heads);
  const int nrows = min(64, chunk - t0);          // valid rows of this tile
  const size_t n0 = (size_t)b * chunk + t0;       // first row
  const bf16* kp = k_cache + ((size_t)b * kv_heads + kvh) * slots * 64;
  const bf16* vp = v_cache + ((size_t)b * kv_heads + kvh) * slots * 64;
  const int slot_last = row_slot[n0 + nrows - 1];
  const int nkt = slot_last / 64 + 1;

  auto load_kv = [&](int kt, int buf) {
    for (int u = tid; u < 64 * 8; u += 128) {  // 64 rows x 8 units of 16 bytes, for K and for V
      const int r = u >> 3, c8 = (u & 7) * 8;
      const int key = kt * 64 + r;
      if (key <= slot_last) {
        fa_cp16(&Ks[buf][r * FA_LD + c8], kp + (size_t)key * 64 + c8);
        fa_cp16(&Vs[buf][r * FA_LD + c8], vp + (size_t)key * 64 + c8);
      } else {  // never-written cache rows may hold Inf / NaN patterns: P is 0 there, but 0 * Inf is not
        *reinterpret_cast<uint4*>(&Ks[buf][r * FA_LD + c8]) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(&Vs[buf][r * FA_LD + c8]) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_kv(0, 0);
  for (int u = tid; u < 64 * 8; u += 128) {
    const int r = u >> 3, c8 = (u & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nrows) v = *reinterpret_cast<const uint4*>(q + ((n0 + r) * heads + h) * 64 + c8);
    *reinterpret_cast<uint4*>(&Qs[r * FA_LD + c8]) = v;
  }
  __syncthreads();
  uint32_t qa[4][4];  // A fragments of this warp's 16 rows, 4 k-steps of 16 dims
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    fa_ldm4(qa[kk], &Qs[(warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * FA_LD + kk * 16 + (lane >> 4) * 8]);
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
  const int slot_lo = r_lo < nrows ? row_slot[n0 + r_lo] : 0, slot_hi = r_hi < nrows ? row_slot[n0 + r_hi] : 0;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[d][e] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;  // scores in log2 units: exp(x) = exp2(x log2 e)

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) load_kv(kt + 1, buf ^ 1);
    if (kt + 1 < nkt) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // S = Q K^T for 64 keys: 8 n-tiles x 4 k-steps
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
      for (int kk2 = 0; kk2 < 2; ++kk2) {
        uint32_t kb[4];  // (b0, b1) of k-step 2 kk2 and of k-step 2 kk2 + 1
        fa_ldm4(kb, &Ks[buf][(j * 8 + (lane & 7)) * FA_LD + kk2 * 32 + (lane >> 3) * 8]);
        fa_mma(s[j], qa[2 * kk2], kb[0], kb[1]);
        fa_mma(s[j], qa[2 * kk2 + 1], kb[2], kb[3]);
      }
    }
    // causal mask + running max
    const int key0 = kt * 64 + 2 * qd;
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = key0 + j * 8 + e;
        s[j][e] = key <= slot_lo ? s[j][e] * sl2 : -INFINITY;
        s[j][2 + e] = key <= slot_hi ? s[j][2 + e] * sl2 : -INFINITY;
        mx_lo = fmaxf(mx_lo, s[j][e]);
        mx_hi = fmaxf(mx_hi, s[j][2 + e]);
      }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    // key 0 is visible to every row, so the running max is finite from the first tile on
    const float c_lo = exp2f(m_lo - mx_lo), c_hi = exp2f(m_hi - mx_hi);
    m_lo = mx_lo;
    m_hi = mx_hi;
    float ps_lo = 0.f, ps_hi = 0.f;
    uint32_t pa[4][4];  // P as A fragments: k-step t = keys 16 t .. 16 t + 15 = n-tiles 2t, 2t+1
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - m_lo), p1 = exp2f(s[j][1] - m_lo);
      const float p2 = exp2f(s[j][2] - m_hi), p3 = exp2f(s[j][3] - m_hi);
      ps_lo += p0 + p1;
      ps_hi += p2 + p3;
      pa[j >> 1][(j & 1) * 2] = fa_pack(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = fa_pack(p2, p3);
    }
    l_lo = l_lo * c_lo + ps_lo;
    l_hi = l_hi * c_hi + ps_hi;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      o[d][0] *= c_lo; o[d][1] *= c_lo;
      o[d][2] *= c_hi; o[d][3] *= c_hi;
    }
    // O += P V: 8 n-tiles of 8 dims x 4 k-steps of 16 keys
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t vb[4];  // (b0, b1) of dims 16 d2 .. +7 and of dims 16 d2 + 8 .. +15
        fa_ldm4t(vb, &Vs[buf][(t * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * FA_LD + d2 * 16 + (lane >> 4) * 8]);
        fa_mma(o[2 * d2], pa[t], vb[0], vb[1]);
        fa_mma(o[2 * d2 + 1], pa[t], vb[2], vb[3]);
      }
    __syncthreads();  // the buffer is free for the load after next
  }
  // the row sums live spread over the four lanes of a quad
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    if (r_lo < nrows)
      *reinterpret_cast<uint32_t*>(out + ((n0 + r_lo) * heads + h) * 64 + d * 8 + 2 * qd) = fa_pack(o[d][0] * i_lo, o[d][1] * i_lo);
    if (r_hi < nrows)
      *reinterpret_cast<uint32_t*>(out + ((n0 + r_hi) * heads + h) * 64 + d * 8 + 2 * qd) = fa_pack(o[d][2] * i_hi, o[d][3] * i_hi);
  }
}

// ---------------------------------------------------------------------------------------------
// One key range on the TENSOR CORES.  A CTA serves (row, KV head, key range): the q-heads of the group are the first
// rows of ONE 16-row mma tile (the rest zero), the range is walked in passes of 256 keys staged cooperatively
// (cp.async), warp w takes keys [64 w, 64 w + 64) of the pass with k_attn_flash64's inner loop -- S = Q K^T and
// O += P V as mma.m16n8k16, online softmax in log2 units -- and the four warps' (o, m, l) are combined in shared
// memory before the range's partial goes to ``part``.  (A CUDA-core version of the same decomposition -- a warp per
// q-head, a lane per key -- was bound by its shared-memory instruction stream, a load per two multiply-adds:
// 5.59 / 8.54 ms per step at 8 / 32 streams against 5.24 / 8.13 ms.)
// ---------------------------------------------------------------------------------------------
constexpr int AM_PASS = 256;  // keys per pass: 4 warps x 64
constexpr size_t AM_SMEM = (size_t)(16 + 2 * AM_PASS) * FA_LD * sizeof(bf16);
__global__ void __launch_bounds__(128) k_attn_split64_mma(const bf16* __restrict__ q, const bf16* __restrict__ k_cache,
                                                          const bf16* __restrict__ v_cache, const int* __restrict__ row_stream,
                                                          const int* __restrict__ row_slot, int imp_B, int imp_pos, int heads,
                                                          int kv_heads, int slots, float scale, float* __restrict__ part) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char am_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(am_raw);  // [16][FA_LD]: rows >= gq are zero
  bf16* Ks = Qs + 16 * FA_LD;                  // [AM_PASS][FA_LD]
  bf16* Vs = Ks + AM_PASS * FA_LD;             // [AM_PASS][FA_LD]
  const int n = blockIdx.x, kvh = blockIdx.y, z = blockIdx.z, gq = heads / kv_heads;  // gq <= 8
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, qd = lane & 3;
  const int nkeys = (row_stream ? row_slot[n] : imp_pos + n / imp_B) + 1;
  const int chunk = (((nkeys + AS_SPLITS - 1) / AS_SPLITS) + 63) & ~63;  // whole 64-key tiles per range
  const int k0 = z * chunk, k1 = k0 + chunk < nkeys ? k0 + chunk : nkeys;
  const int h0 = kvh * gq;
  if (k0 >= nkeys) {  // an empty range: weight 0 in the combine
    for (int i = tid; i < gq * AS_PW; i += 128) {
      const int r = i / AS_PW, d = i - r * AS_PW;
      part[(((size_t)n * heads + h0 + r) * AS_SPLITS + z) * AS_PW + d] = d == 64 ? -INFINITY : 0.f;
    }
    return;
  }
  const size_t base = ((size_t)(row_stream ? row_stream[n] : n % imp_B) * kv_heads + kvh) * slots * 64;
  const bf16* kp = k_cache + base;
  const bf16* vp = v_cache + base;
  for (int u = tid; u < 16 * 8; u += 128) {
    const int r = u >> 3, c8 = (u & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < gq) v = *reinterpret_cast<const uint4*>(q + ((size_t)n * heads + h0 + r) * 64 + c8);
    *reinterpret_cast<uint4*>(&Qs[r * FA_LD + c8]) = v;
  }
  float m_lo = -INFINITY, l_lo = 0.f;
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[d][e] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;  // scores in log2 units
  uint32_t qa[4][4];
  bool have_q = false;
  for (int kb = k0; kb < k1; kb += AM_PASS) {
    // stage the pass: 256 rows x 8 units of 16 bytes for K and for V; rows past the range are zero
    for (int u = tid; u < AM_PASS * 8; u += 128) {
      const int r = u >> 3, c8 = (u & 7) * 8, key = kb + r;
      if (key < k1) {
        fa_cp16(&Ks[r * FA_LD + c8], kp + (size_t)key * 64 + c8);
        fa_cp16(&Vs[r * FA_LD + c8], vp + (size_t)key * 64 + c8);
      } else {
        *reinterpret_cast<uint4*>(&Ks[r * FA_LD + c8]) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(&Vs[r * FA_LD + c8]) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (!have_q) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        fa_ldm4(qa[kk], &Qs[(((lane >> 3) & 1) * 8 + (lane & 7)) * FA_LD + kk * 16 + (lane >> 4) * 8]);
      have_q = true;
    }
    const int tk0 = kb + warp * 64;  // this warp's 64 keys of the pass
    if (tk0 < k1) {
      const bf16* Kt = Ks + warp * 64 * FA_LD;
      const bf16* Vt = Vs + warp * 64 * FA_LD;
      float sc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[j][e] = 0.f;
#pragma unroll
        for (int kk2 = 0; kk2 < 2; ++kk2) {
          uint32_t kf[4];
          fa_ldm4(kf, &Kt[(j * 8 + (lane & 7)) * FA_LD + kk2 * 32 + (lane >> 3) * 8]);
          fa_mma(sc[j], qa[2 * kk2], kf[0], kf[1]);
          fa_mma(sc[j], qa[2 * kk2 + 1], kf[2], kf[3]);
        }
      }
      const int key0 = tk0 + 2 * qd;
      float mx = m_lo;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[j][e] = (key0 + j * 8 + e < k1) ? sc[j][e] * sl2 : -INFINITY;
          mx = fmaxf(mx, sc[j][e]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float c_lo = exp2f(m_lo - mx);  // (the tile's first key is inside the range: mx is finite)
      m_lo = mx;
      float ps = 0.f;
      uint32_t pa[4][4];  // P as A fragments; the rows g + 8 of the tile are unused (zero q rows): their P is 0
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = exp2f(sc[j][0] - m_lo), p1 = exp2f(sc[j][1] - m_lo);
        ps += p0 + p1;
        pa[j >> 1][(j & 1) * 2] = fa_pack(p0, p1);
        pa[j >> 1][(j & 1) * 2 + 1] = 0u;
      }
      l_lo = l_lo * c_lo + ps;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        o[d][0] *= c_lo;
        o[d][1] *= c_lo;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int d2 = 0; d2 < 4; ++d2) {
          uint32_t vb[4];
          fa_ldm4t(vb, &Vt[(t * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * FA_LD + d2 * 16 + (lane >> 4) * 8]);
          fa_mma(o[2 * d2], pa[t], vb[0], vb[1]);
          fa_mma(o[2 * d2 + 1], pa[t], vb[2], vb[3]);
        }
    }
    __syncthreads();  // the pass buffers are free for the next pass (and, after the last one, for the warps' partials)
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  // the four warps' partials of row g (q-head h0 + g): [warp][g][AS_PW] floats over the K buffer
  float* wp = reinterpret_cast<float*>(Ks);
  if (g < gq) {
    float* dst = wp + ((size_t)warp * 8 + g) * AS_PW;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      dst[d * 8 + 2 * qd] = o[d][0];
      dst[d * 8 + 2 * qd + 1] = o[d][1];
    }
    if (qd == 0) {
      dst[64] = m_lo;
      dst[65] = l_lo;
    }
  }
  __syncthreads();
  for (int i = tid; i < gq * 64; i += 128) {
    const int r = i >> 6, d = i & 63;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, wp[((size_t)w * 8 + r) * AS_PW + 64]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {  // warp order: deterministic (a warp without keys: exp2(-inf) = 0)
      const float* src = wp + ((size_t)w * 8 + r) * AS_PW;
      const float wgt = exp2f(src[64] - M);
      L = fmaf(src[65], wgt, L);
      O = fmaf(src[d], wgt, O);
    }
    float* pp = part + (((size_t)n * heads + h0 + r) * AS_SPLITS + z) * AS_PW;
    pp[d] = O;
    if (d == 0) {
      pp[64] = M;
      pp[65] = L;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-batched RoPE + KV append for the tensor-core prefill path: qkv [N, (H+2KV)*hd] (GEMM output,
// already rounded to bf16) -> rotated q [N, H*hd], rotated k and v into the cache at the row's slot.
// Same arithmetic as the EPI_ROPE_KV epilogue.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rope_kv_rows(const bf16* __restrict__ qkv, const bf16* __restrict__ rope,
                                                      const int* __restrict__ row_stream, const int* __restrict__ row_pos,
                                                      const int* __restrict__ row_slot, int imp_B, int imp_pos, int heads,
                                                      int kv_heads, int hd, int slots, bf16* __restrict__ q_out,
                                                      bf16* __restrict__ k_cache, bf16* __restrict__ v_cache) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x;
  const int qrows = heads * hd, krows = kv_heads * hd, total = qrows + 2 * krows;
  const bf16* src = qkv + (size_t)n * total;
  const int pos = row_stream ? row_pos[n] : imp_pos + n / imp_B;
  const int slot = row_stream ? row_slot[n] : imp_pos + n / imp_B;
  const int stream = row_stream ? row_stream[n] : n % imp_B;
  for (int p = threadIdx.x; p < total / 2; p += blockDim.x) {
    const int r0 = 2 * p;
    const __nv_bfloat162 in = *reinterpret_cast<const __nv_bfloat162*>(src + r0);
    const float y0 = __low2float(in), y1 = __high2float(in);
    float o0 = y0, o1 = y1;
    if (r0 < qrows + krows) {
      const __nv_bfloat162 cs = *reinterpret_cast<const __nv_bfloat162*>(rope + ((size_t)pos * (hd / 2) + ((r0 % hd) >> 1)) * 2);
      const float c = __low2float(cs), s = __high2float(cs);
      o0 = rbf(__fsub_rn(__fmul_rn(y0, c), __fmul_rn(y1, s)));
      o1 = rbf(__fadd_rn(__fmul_rn(y1, c), __fmul_rn(y0, s)));
    }
    const __nv_bfloat162 out = __floats2bfloat162_rn(o0, o1);
    if (r0 < qrows) {
      *reinterpret_cast<__nv_bfloat162*>(q_out + (size_t)n * qrows + r0) = out;
    } else {
      const bool isk = r0 < qrows + krows;
      const int rr = r0 - (isk ? qrows : qrows + krows);
      const int kvh = rr / hd, d = rr % hd;
      bf16* dst = (isk ? k_cache : v_cache) + (((size_t)stream * kv_heads + kvh) * slots + slot) * hd + d;
      *reinterpret_cast<__nv_bfloat162*>(dst) = out;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone RMSNorm (final norm of a stack): one CTA per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rmsnorm(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ scale,
                                                 int D, float eps, bf16* __restrict__ y, int ldy) {
  pdl_wait();
  pdl_trigger();
  __shared__ float scratch[33];
  const int n = blockIdx.x;
  const bf16* xr = x + (size_t)n * ldx;
  float ss = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const float v = bf2f(xr[k]);
    ss = fmaf(v, v, ss);
  }
  ss = block_sum<256, 0>(ss, scratch, threadIdx.x);
  const float inv = 1.0f / sqrtf(ss / (float)D + eps);
  for (int k = threadIdx.x; k < D; k += blockDim.x)
    y[(size_t)n * ldy + k] = f2bf(rbf(bf2f(xr[k]) * inv) * bf2f(scale[k]));
}

// ---------------------------------------------------------------------------------------------
// K7/K8: sample_topk (sesameai/models.py:72-87) fused with the embedding gather that feeds the
// next depth-decoder step.  One CTA per stream.
//   x = bf16(logit / T); thr = k-th largest x (exact 16-bit radix select); x[x < thr] = -inf
//   (ties with thr kept); ls = bf16(log_softmax(x)); p = bf16(softmax(ls)); r = bf16(p / q);
//   token = first argmax(r).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf_key(bf16 v) {
  const uint32_t b = __bfloat16_as_ushort(v);
  return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  const uint32_t b = (k & 0x8000u) ? (k & 0x7fffu) : (~k & 0xffffu);
  return __uint_as_float(b << 16);
}

// counter-based Exp(1) draw for production mode (no shared noise tensor): -log(u), u in (0,1],
// rounded to bf16 like ``torch.empty_like(probs).exponential_(1)`` on a bf16 tensor.
__device__ __forceinline__ float exp1_draw(unsigned long long seed, unsigned long long ctr) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = ((float)(uint32_t)(z >> 40) + 1.0f) * (1.0f / 16777216.0f);
  return fmaxf(rbf(-logf(u)), 1.1754944e-38f);
}

#define SAMPLE_THREADS 256
#define SAMPLE_MAXV 4096

// k-th largest of a 256-bin histogram, counted from the top: warp 0 scans 8 bins per lane.
// Writes {bin, rank inside the bin} to out[0..1]; callers sync afterwards.
__device__ __forceinline__ void hist_select_from_top(const unsigned int* hist, int k, int* out, int tid) {
  if (tid >= 32) return;
  const int lane = tid;
  // lane l owns bins [255-8l-7 .. 255-8l] i.e. descending order across lanes
  unsigned int cnt[8], tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    cnt[j] = hist[255 - 8 * lane - j];
    tot += cnt[j];
  }
  unsigned int incl = tot;  // inclusive prefix over lanes (top bins first)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned int excl = incl - tot;
  const bool mine = excl < (unsigned)k && incl >= (unsigned)k;
  const unsigned int all = __shfl_sync(0xffffffffu, incl, 31);
  if (mine) {
    unsigned int cum = excl;
    int j = 0;
    for (; j < 7; ++j) {
      if (cum + cnt[j] >= (unsigned)k) break;
      cum += cnt[j];
    }
    out[0] = 255 - 8 * lane - j;
    out[1] = k - (int)cum;
  } else if (lane == 31 && all < (unsigned)k) {  // fewer than k entries: everything survives
    out[0] = 0;
    out[1] = 1;
  }
}

// Core: returns the sampled token for one logits row (all NT threads return the same value).
template <int NT, int BAR, bool CG>
__device__ int sample_row(const bf16* __restrict__ logits, const bf16* __restrict__ noise, int V, float temperature,
                          int topk, unsigned long long seed, unsigned long long ctr0, float* xs /*[V]*/,
                          unsigned int* hist /*[256]*/, float* scratch /*[33]*/, int* iscratch /*[36]*/, int tid) {
  // 1. temperature + radix histogram of the high key byte.  ``logits / temperature`` with a
  //    Python-float divisor is evaluated by torch's CUDA div kernel as x * (1/T) in fp32
  //    (the reference path on a GPU), then rounded to bf16.
  const int k = topk < 1 ? 1 : (topk > V ? V : topk);
  const bool greedy = k == 1;
  if (!greedy) {
    for (int i = tid; i < 256; i += NT) hist[i] = 0;
    csync<NT, BAR>();
  }
  float mx = -INFINITY;
  const float inv_t = 1.0f / temperature;
  {
    // all of this thread's logits are requested before any is used: one L2 round trip, not V/NT
    constexpr int MAXPT = SAMPLE_MAXV / NT;
    unsigned short raw[MAXPT];
#pragma unroll
    for (int t = 0; t < MAXPT; ++t) {
      const int i = tid + t * NT;
      raw[t] = 0;
      if (i < V)
        raw[t] = CG ? __ldcg(reinterpret_cast<const unsigned short*>(logits) + i)
                    : reinterpret_cast<const unsigned short*>(logits)[i];
    }
#pragma unroll
    for (int t = 0; t < MAXPT; ++t) {
      const int i = tid + t * NT;
      if (i < V) {
        const bf16 xb = f2bf(bf2f(__ushort_as_bfloat16(raw[t])) * inv_t);
        const float xv = bf2f(xb);
        xs[i] = xv;
        mx = fmaxf(mx, xv);
        if (!greedy) atomicAdd(&hist[bf_key(xb) >> 8], 1u);
      }
    }
  }
  mx = block_max<NT, BAR>(mx, scratch, tid);  // (its barriers also publish xs[] and hist[])
  // greedy short cut: topk == 1 with a UNIQUE maximum leaves one survivor, whose probability is
  // exactly 1 (log_softmax -> 0, softmax -> 1) while every masked entry scores 0: the race cannot
  // change the winner, so the token is the arg-max.  (Ties at the maximum take the general path:
  // the reference resolves them by the Exp(1) race)
  if (greedy) {
    float cnt = 0.f;
    int first = -1;
    for (int i = tid; i < V; i += NT)
      if (xs[i] == mx) {
        cnt += 1.f;
        if (first < 0) first = i;
      }
    cnt = block_sum<NT, BAR>(cnt, scratch, tid);
    if (cnt == 1.f) {
      if (first >= 0) iscratch[0] = first;
      csync<NT, BAR>();
      const int tok = iscratch[0];
      csync<NT, BAR>();
      return tok;
    }
    csync<NT, BAR>();
  }
  // 2. threshold = k-th largest value: exact 16-bit radix select (top-1 is just the max)
  float thr = mx;
  if (!greedy) {
    hist_select_from_top(hist, k, iscratch, tid);
    csync<NT, BAR>();
    const int hb = iscratch[0];
    const int krem = iscratch[1];
    csync<NT, BAR>();
    for (int i = tid; i < 256; i += NT) hist[i] = 0;
    csync<NT, BAR>();
    for (int i = tid; i < V; i += NT) {
      const uint32_t key = bf_key(f2bf(xs[i]));
      if ((int)(key >> 8) == hb) atomicAdd(&hist[key & 0xffu], 1u);
    }
    csync<NT, BAR>();
    hist_select_from_top(hist, krem, iscratch + 2, tid);
    csync<NT, BAR>();
    thr = key_to_float(((uint32_t)hb << 8) | (uint32_t)iscratch[2]);
  }
  // 3. mask (ties with the threshold are kept) + log_softmax, fp32 internals, bf16 result
  //    (non-survivors contribute exp(-inf) = 0 exactly and end with probability 0: skip their math)
  float sum = 0.f;
  for (int i = tid; i < V; i += NT) {
    const float v = xs[i];
    if (v < thr) xs[i] = -INFINITY;
    else sum += expf(v - mx);
  }
  sum = block_sum<NT, BAR>(sum, scratch, tid);
  const float lse = logf(sum);
  // 4. softmax of the bf16 log-probs: their max is the entry of the largest x
  const float mx2 = rbf((mx - mx) - lse);
  float sum2 = 0.f;
  for (int i = tid; i < V; i += NT) {
    const float v = xs[i];
    if (v == -INFINITY) continue;
    const float ls = rbf((v - mx) - lse);
    xs[i] = ls;
    sum2 += expf(ls - mx2);
  }
  sum2 = block_sum<NT, BAR>(sum2, scratch, tid);
  // 5. exponential race, first-index argmax
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = tid; i < V; i += NT) {
    const float v = xs[i];
    // masked entries have p = 0 -> r = 0 (q > 0); they can only win when every r rounds to 0
    float r = 0.f;
    if (v != -INFINITY) {
      const float p = rbf(expf(v - mx2) / sum2);
      const float q = noise ? bf2f(noise[i]) : exp1_draw(seed, ctr0 + i);
      r = rbf(p / q);
    }
    if (r > best) {  // strided ascending i per thread -> keeps the first index on ties
      best = r;
      besti = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) {
      best = ob;
      besti = oi;
    }
  }
  csync<NT, BAR>();
  if ((tid & 31) == 0) {
    scratch[tid >> 5] = best;
    iscratch[4 + (tid >> 5)] = besti;
  }
  csync<NT, BAR>();
  if (tid < 32) {
    constexpr int nw = NT / 32;
    best = tid < nw ? scratch[tid] : -INFINITY;
    besti = tid < nw ? iscratch[4 + tid] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
      }
    }
    if (tid == 0) iscratch[0] = besti;
  }
  csync<NT, BAR>();
  const int tok = iscratch[0];
  csync<NT, BAR>();
  return tok;
}

__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample_only(const bf16* logits, const bf16* noise, int V,
                                                                float temperature, int topk, int* out) {
  __shared__ float xs[SAMPLE_MAXV];
  __shared__ unsigned int hist[256];
  __shared__ float scratch[33];
  __shared__ int iscratch[36];
  const int b = blockIdx.x;
  const int tok = sample_row<SAMPLE_THREADS, 0, false>(logits + (size_t)b * V, noise ? noise + (size_t)b * V : nullptr, V,
                                                      temperature, topk, 0, 0, xs, hist, scratch, iscratch, threadIdx.x);
  if (threadIdx.x == 0) out[b] = tok;
}

// Sampling step ``cb`` of a frame: logits [B, ldl] -> token; records it; gathers the embedding of
// the (possibly teacher-forced) token into the depth decoder's next input row.
__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample_step(const FrameParams* __restrict__ P,
                                                                const bf16* __restrict__ logits, int ldl, int cb,
                                                                int V, int C, const bf16* __restrict__ audio_emb, int D,
                                                                bf16* __restrict__ next_in /*[B, D] or null*/,
                                                                DevStatus* st = nullptr) {
  pdl_wait();
  pdl_trigger();
  // ``audio_emb``/``D`` may also be the projection(embedding) table and its row length
  __shared__ float xs[SAMPLE_MAXV];
  __shared__ unsigned int hist[256];
  __shared__ float scratch[33];
  __shared__ int iscratch[36];
  const int b = blockIdx.x, B = P->B;
  const bf16* lrow = logits + (size_t)b * ldl;
  if (P->logits_out) {
    bf16* lo = P->logits_out + ((size_t)cb * B + b) * V;
    for (int i = threadIdx.x; i < V; i += blockDim.x) lo[i] = lrow[i];
  }
  const bf16* nz = P->noise ? P->noise + ((size_t)cb * B + b) * V : nullptr;
  const unsigned long long ctr = ((P->offset * (unsigned long long)C + cb) * (unsigned long long)B + b) * 4096ull;
  int tok = sample_row<SAMPLE_THREADS, 0, false>(lrow, nz, V, P->temperature, P->topk, P->seed, ctr, xs, hist, scratch,
                                                 iscratch, threadIdx.x);
  if (threadIdx.x == 0 && P->sampled_out) P->sampled_out[(size_t)b * C + cb] = tok;
  if (P->forced) tok = P->forced[(size_t)b * C + cb];
  if ((unsigned)tok >= (unsigned)V) {  // teacher-forced id out of range
    if (threadIdx.x == 0) report_error(st, 0x802);
    tok = 0;
  }
  if (threadIdx.x == 0) P->out[(size_t)b * C + cb] = tok;
  if (next_in) {
    const bf16* row = audio_emb + ((size_t)tok + (size_t)cb * V) * D;
    for (int d8 = threadIdx.x; d8 < D / 8; d8 += blockDim.x)
      *reinterpret_cast<uint4*>(next_in + (size_t)b * D + d8 * 8) = *reinterpret_cast<const uint4*>(row + d8 * 8);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight re-packing (csm_create): fused [q;k;v], interleaved gate/up, transposed audio heads.
// ---------------------------------------------------------------------------------------------
__global__ void k_copy_rows(const bf16* src, bf16* dst, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}
// dst[2r] = w1[r], dst[2r+1] = w3[r]
__global__ void k_interleave_rows(const bf16* w1, const bf16* w3, bf16* dst, int rows, int K8) {
  const size_t total = (size_t)rows * K8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K8, k = i % K8;
    reinterpret_cast<uint4*>(dst)[(2 * r) * K8 + k] = reinterpret_cast<const uint4*>(w1)[i];
    reinterpret_cast<uint4*>(dst)[(2 * r + 1) * K8 + k] = reinterpret_cast<const uint4*>(w3)[i];
  }
}
// src [S, K, V] -> dst [S, Vp, K] (rows v >= V zero)
__global__ void k_transpose_heads(const bf16* src, bf16* dst, int K, int V, int Vp) {
  __shared__ bf16 tile[32][33];
  const int s = blockIdx.z;
  const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const bf16* sp = src + (size_t)s * K * V;
  bf16* dp = dst + (size_t)s * Vp * K;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, v = v0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && v < V) ? sp[(size_t)k * V + v] : f2bf(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int v = v0 + i, k = k0 + threadIdx.x;
    if (v < Vp && k < K) dp[(size_t)v * K + k] = tile[threadIdx.x][i];
  }
}
