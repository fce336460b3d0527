// Persistent decode megakernel: one audio frame (backbone step + codebook-0 sample + 31 depth-
// decoder steps, ~670 dependent phases) in ONE launch, for one stream.
//
// Why: at batch 1 a frame is ~800 tiny dependent GEMVs (0.3-5 us of HBM time each); launched one
// by one they are latency bound (v1: 17 % of the HBM roofline).  Here one CTA per SM stays
// resident for the whole frame:
//   * a producer warp streams this CTA's share of every weight matrix, in consumption order,
//     through per-warp shared-memory rings with bulk async copies (cp.async.bulk + mbarrier
//     complete_tx).  Weight addresses are data independent, so the stream runs AHEAD of the
//     dependency chain: 148 SMs x 192 KB = 28 MB of weights (~4 us of HBM time) are in flight
//     while grid barriers, norms, attention and sampling resolve;
//   * 8 consumer warps each own output-row groups: dot products out of shared memory, warp
//     shuffle reduction, fused epilogue (RoPE + KV append / residual / SwiGLU / logits) at the
//     reference's bf16 rounding points;
//   * phases are separated by a monotonic grid barrier (release/acquire counter in L2).
// Every spin has a trip-count cap and traps instead of hanging the GPU.
#pragma once
#include "lm_kernels.cuh"

namespace mega {

constexpr int NW = 8;                 // consumer warps
constexpr int NCT = NW * 32;          // consumer threads
constexpr int NTHREADS = NCT + 32;    // + one producer warp
constexpr int SLOTS = 3;              // ring slots per consumer warp
constexpr int CHUNK_ELEMS = 4096;     // bf16 per slot (8 KB)
constexpr int MAXNB = 2;              // activation rows per phase (depth step 1 carries 2)
constexpr int XBUF_ELEMS = MAXNB * 8192;
constexpr int CBAR = 1;               // named barrier of the consumer warps
constexpr size_t SMEM_RING = (size_t)NW * SLOTS * CHUNK_ELEMS * 2;
constexpr size_t SMEM_X = (size_t)XBUF_ELEMS * 2;
constexpr size_t SMEM_MISC = 1024;
constexpr size_t SMEM_BYTES = SMEM_RING + SMEM_X + SMEM_MISC;

enum { PH_GEMV = 0, PH_EMBED = 1, PH_ATTN = 2, PH_SAMPLE = 3 };
enum { POS_FIXED = 0, POS_BACKBONE = 1 };

struct Phase {
  int type, epi, norm, nb;
  // GEMV
  const bf16* W;
  int rows, K, R, KC, G;
  const bf16* x;  // [nb, ldx] activations in global memory
  int ldx;
  const bf16* norm_scale;
  float eps;
  bf16* out;
  int ldo;
  const bf16* resid;
  bf16* x_copy_out;  // CTA 0 publishes the staged (normalised) rows here
  // RoPE / KV append / attention
  bf16 *q, *kc, *vc;
  const bf16* rope;
  int heads, kv_heads, hd, slots, pos_mode, pos0;
  int attn_prologue;  // GEMV: x = attention(q, cache) computed redundantly by every CTA (<= 32 keys)
  bf16* att_out;      // PH_ATTN: [heads*hd]
  // sample
  int cb, V, C, D, ldl;
  const bf16* logits;
  bf16* next_in;
  const bf16 *audio_emb, *text_emb;
  bf16* h_out;  // PH_EMBED
};

struct Sync {
  unsigned int counter;  // grid barrier arrivals, zeroed by k_mega_prepare before every frame
  unsigned int error;
};

__global__ void k_mega_prepare(FrameParams* dst, FrameParams v, Sync* sync) {
  *dst = v;
  sync->counter = 0;
}

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float ldcg_bf(const bf16* p) {
  return bf2f(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

__device__ __forceinline__ void die(Sync* sync, unsigned code) {
  atomicExch(&sync->error, code);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, Sync* sync, unsigned code) {
  for (unsigned spin = 0; !mbar_try(bar, parity); ++spin)
    if (spin > (1u << 22)) die(sync, code);
}

// work distribution: group index of the i-th group owned by (cta, warp)
__device__ __forceinline__ int group_of(int cta, int ncta, int warp, int i) { return cta + ncta * (warp + NW * i); }

__device__ __forceinline__ void phase_pos(const Phase& ph, const FrameParams* P, int n, int& pos, int& slot) {
  if (ph.pos_mode == POS_BACKBONE) {
    pos = (int)P->pos[(size_t)n * P->S + (P->S - 1)];  // row n == stream n (batch 1: n == 0)
    slot = P->cache_len + P->S - 1;
  } else {
    pos = slot = ph.pos0 + n;
  }
}

// ---- producer: stream this CTA's weight chunks through the per-warp rings -------------------------
__device__ void producer(const Phase* __restrict__ phases, int nphases, bf16* ring, uint64_t* full, uint64_t* empty,
                         Sync* sync) {
  const int w = threadIdx.x & 31;  // lane w feeds consumer warp w
  if (w >= NW) return;
  const int cta = blockIdx.x, ncta = gridDim.x;
  unsigned cnt = 0;
  for (int p = 0; p < nphases; ++p) {
    if (phases[p].type != PH_GEMV) continue;
    const bf16* W = phases[p].W;
    const int K = phases[p].K, R = phases[p].R, KC = phases[p].KC, G = phases[p].G, rows = phases[p].rows;
    const int nkc = K / KC;
    for (int i = 0;; ++i) {
      const int g = group_of(cta, ncta, w, i);
      if (g >= G) break;
      const int r0 = g * R;
      const int nr = min(R, rows - r0);
      for (int kc = 0; kc < nkc; ++kc, ++cnt) {
        const int slot = cnt % SLOTS;
        const uint32_t par = ((cnt / SLOTS) & 1) ^ 1;
        uint64_t* fb = &full[w * SLOTS + slot];
        uint64_t* eb = &empty[w * SLOTS + slot];
        for (unsigned spin = 0; !mbar_test(eb, par); ++spin)
          if (spin > (1u << 26)) die(sync, 0x100 + w);
        bf16* dst = ring + (size_t)(w * SLOTS + slot) * CHUNK_ELEMS;
        if (nkc == 1) {  // rows are contiguous: one copy
          mbar_expect_tx(fb, (uint32_t)nr * K * 2);
          bulk_g2s(dst, W + (size_t)r0 * K, (uint32_t)nr * K * 2, fb);
        } else {
          mbar_expect_tx(fb, (uint32_t)nr * KC * 2);
          for (int r = 0; r < nr; ++r)
            bulk_g2s(dst + r * KC, W + (size_t)(r0 + r) * K + (size_t)kc * KC, (uint32_t)KC * 2, fb);
        }
      }
    }
  }
}

// ---- consumer pieces ---------------------------------------------------------------------------------
struct Ctx {
  const FrameParams* P;
  bf16* ring;
  bf16* xs;
  uint64_t *full, *empty;
  float* scratch;  // 33 floats
  int* iscratch;   // 36 ints
  Sync* sync;
  unsigned cnt;  // chunks consumed by this warp so far
  int tid, warp, lane;
};

// fused epilogue of one output-row pair (row r0, r0+1) for activation row n
__device__ __forceinline__ void epilogue(const Phase& ph, const FrameParams* P, int r0, int n, float a0, float a1) {
  const bool has1 = r0 + 1 < ph.rows;
  const float y0 = rbf(a0), y1 = rbf(a1);
  if (ph.epi == EPI_PLAIN) {
    ph.out[(size_t)n * ph.ldo + r0] = f2bf(y0);
    if (has1) ph.out[(size_t)n * ph.ldo + r0 + 1] = f2bf(y1);
  } else if (ph.epi == EPI_RESID) {
    const float h0 = ldcg_bf(ph.resid + (size_t)n * ph.ldo + r0);
    const float h1 = has1 ? ldcg_bf(ph.resid + (size_t)n * ph.ldo + r0 + 1) : 0.f;
    ph.out[(size_t)n * ph.ldo + r0] = f2bf(y0 + h0);
    if (has1) ph.out[(size_t)n * ph.ldo + r0 + 1] = f2bf(y1 + h1);
  } else if (ph.epi == EPI_SWIGLU) {
    ph.out[(size_t)n * ph.ldo + (r0 >> 1)] = f2bf(silu_bf(y0) * y1);
  } else {  // EPI_ROPE_KV
    const int hd = ph.hd, qrows = ph.heads * hd, krows = ph.kv_heads * hd;
    int pos, slot;
    phase_pos(ph, P, n, pos, slot);
    float o0 = y0, o1 = y1;
    if (r0 < qrows + krows) {
      const int j = (r0 % hd) >> 1;
      const bf16* cs = ph.rope + ((size_t)pos * (hd / 2) + j) * 2;
      const float c = bf2f(cs[0]), s = bf2f(cs[1]);
      o0 = rbf(__fsub_rn(__fmul_rn(y0, c), __fmul_rn(y1, s)));
      o1 = rbf(__fadd_rn(__fmul_rn(y1, c), __fmul_rn(y0, s)));
    }
    if (r0 < qrows) {
      ph.q[(size_t)n * qrows + r0] = f2bf(o0);
      ph.q[(size_t)n * qrows + r0 + 1] = f2bf(o1);
    } else {
      const bool isk = r0 < qrows + krows;
      const int rr = r0 - (isk ? qrows : qrows + krows);
      const int kvh = rr / hd, d = rr % hd;
      bf16* dst = (isk ? ph.kc : ph.vc) + ((size_t)kvh * ph.slots + slot) * hd + d;  // stream 0
      dst[0] = f2bf(o0);
      dst[1] = f2bf(o1);
    }
  }
}

template <int R, int NB>
__device__ __forceinline__ void gemv_groups(const Phase& ph, Ctx& c) {
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int K = ph.K, KC = ph.KC, nkc = K / KC;
  for (int i = 0;; ++i) {
    const int g = group_of(cta, ncta, c.warp, i);
    if (g >= ph.G) break;
    float acc[R][NB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[r][n] = 0.f;
    for (int kc = 0; kc < nkc; ++kc, ++c.cnt) {
      const int slot = c.cnt % SLOTS;
      const uint32_t par = (c.cnt / SLOTS) & 1;
      mbar_wait(&c.full[c.warp * SLOTS + slot], par, c.sync, 0x200 + c.warp);
      const bf16* chunk = c.ring + (size_t)(c.warp * SLOTS + slot) * CHUNK_ELEMS;
      const bf16* xk = c.xs + (size_t)kc * KC;
#pragma unroll 2
      for (int k = c.lane * 8; k < KC; k += 256) {
        uint4 xv[NB];
#pragma unroll
        for (int n = 0; n < NB; ++n) xv[n] = *reinterpret_cast<const uint4*>(xk + (size_t)n * K + k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint4 wv = *reinterpret_cast<const uint4*>(chunk + r * KC + k);
#pragma unroll
          for (int n = 0; n < NB; ++n) acc[r][n] = dot8(wv, xv[n], acc[r][n]);
        }
      }
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&c.empty[c.warp * SLOTS + slot]);
    }
    // reduce, then lane (pair, n) runs that pair's epilogue
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int r = 0; r < R / 2; ++r)
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float s0 = warp_sum(acc[2 * r][n]);
        const float s1 = warp_sum(acc[2 * r + 1][n]);
        if (c.lane == r + n * (R / 2)) {
          y0 = s0;
          y1 = s1;
        }
      }
    if (c.lane < (R / 2) * NB) {
      const int pr = c.lane % (R / 2), n = c.lane / (R / 2);
      const int r0 = g * R + 2 * pr;
      if (r0 < ph.rows && n < ph.nb) epilogue(ph, c.P, r0, n, y0, y1);
    }
  }
}

// attention over <= 32 cached keys for every (row, head), redundantly in every CTA, straight into
// the activation buffer of the output projection (depth decoder: 32 slots per stream)
__device__ void attn_small_into_x(const Phase& ph, Ctx& c) {
  const int hd = ph.hd, heads = ph.heads, grp = heads / ph.kv_heads;
  const float scale = 1.0f / sqrtf((float)hd);
  const int dims = hd / 32;  // output dims per lane (2 or 4)
  for (int item = c.warp; item < ph.nb * heads; item += NW) {
    const int n = item / heads, h = item % heads, kvh = h / grp;
    int pos, slot;
    phase_pos(ph, c.P, n, pos, slot);
    const int nkeys = slot + 1;
    const bf16* kp = ph.kc + (size_t)kvh * ph.slots * hd;
    const bf16* vp = ph.vc + (size_t)kvh * ph.slots * hd;
    const bf16* qr = ph.q + ((size_t)n * heads + h) * hd;
    float s = -INFINITY;
    if (c.lane < nkeys) {
      const bf16* kr = kp + (size_t)c.lane * hd;
      float a = 0.f;
      for (int i = 0; i < hd / 8; ++i) {
        const uint4 kv = ldcg16(kr + i * 8);
        const uint4 qv = ldcg16(qr + i * 8);
        a = dot8(kv, qv, a);
      }
      s = a * scale;
    }
    const float mx = warp_max(s);
    const float e = (c.lane < nkeys) ? expf(s - mx) : 0.f;
    const float sum = warp_sum(e);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < nkeys; ++j) {
      const float pj = __shfl_sync(0xffffffffu, e, j);
      const bf16* vr = vp + (size_t)j * hd + c.lane * dims;
      if (dims == 4) {
        const uint2 v = __ldcg(reinterpret_cast<const uint2*>(vr));
        o[0] = fmaf(pj, bflo(v.x), o[0]); o[1] = fmaf(pj, bfhi(v.x), o[1]);
        o[2] = fmaf(pj, bflo(v.y), o[2]); o[3] = fmaf(pj, bfhi(v.y), o[3]);
      } else {
        const uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(vr));
        o[0] = fmaf(pj, bflo(v), o[0]); o[1] = fmaf(pj, bfhi(v), o[1]);
      }
    }
    const float inv = 1.0f / sum;
    bf16* dst = c.xs + (size_t)n * ph.K + h * hd + c.lane * dims;
    for (int d = 0; d < dims; ++d) dst[d] = f2bf(o[d] * inv);
  }
  csync<NCT, CBAR>();
}

// stage the phase's activation rows into shared memory (+ RMSNorm prologue)
__device__ void stage_x(const Phase& ph, Ctx& c) {
  if (ph.attn_prologue) {
    attn_small_into_x(ph, c);
    return;
  }
  const int K = ph.K, nb = ph.nb;
  for (int i = c.tid; i < nb * (K / 8); i += NCT) {
    const int n = i / (K / 8), k8 = i % (K / 8);
    *reinterpret_cast<uint4*>(c.xs + (size_t)n * K + k8 * 8) = ldcg16(ph.x + (size_t)n * ph.ldx + k8 * 8);
  }
  csync<NCT, CBAR>();
  if (ph.norm) {
    if (nb == 1) rmsnorm_smem<1, NCT, CBAR>(c.xs, K, ph.norm_scale, ph.eps, c.scratch, c.tid);
    else rmsnorm_smem<2, NCT, CBAR>(c.xs, K, ph.norm_scale, ph.eps, c.scratch, c.tid);
  }
  if (ph.x_copy_out && blockIdx.x == 0) {
    for (int i = c.tid; i < nb * (K / 8); i += NCT) {
      const int n = i / (K / 8), k8 = i % (K / 8);
      *reinterpret_cast<uint4*>(ph.x_copy_out + (size_t)n * K + k8 * 8) =
          *reinterpret_cast<const uint4*>(c.xs + (size_t)n * K + k8 * 8);
    }
  }
}

__device__ void gemv_phase(const Phase& ph, Ctx& c) {
  stage_x(ph, c);
  if (ph.nb == 1) {
    switch (ph.R) {
      case 2: gemv_groups<2, 1>(ph, c); break;
      case 4: gemv_groups<4, 1>(ph, c); break;
      case 8: gemv_groups<8, 1>(ph, c); break;
      default: gemv_groups<16, 1>(ph, c); break;
    }
  } else {
    switch (ph.R) {
      case 2: gemv_groups<2, 2>(ph, c); break;
      case 4: gemv_groups<4, 2>(ph, c); break;
      case 8: gemv_groups<8, 2>(ph, c); break;
      default: gemv_groups<16, 2>(ph, c); break;
    }
  }
}

// backbone attention: CTA h < heads owns q-head h over keys [0 .. slot]
__device__ void attn_phase(const Phase& ph, Ctx& c) {
  const int h = blockIdx.x;
  if (h >= ph.heads) return;
  const int hd = ph.hd;
  int pos, slot;
  phase_pos(ph, c.P, 0, pos, slot);
  const int nkeys = slot + 1;
  const int kvh = h / (ph.heads / ph.kv_heads);
  const bf16* kp = ph.kc + (size_t)kvh * ph.slots * hd;
  const bf16* vp = ph.vc + (size_t)kvh * ph.slots * hd;
  float* sc = reinterpret_cast<float*>(c.xs);      // [slots] scores (<= 8 KB)
  float* part = sc + ph.slots;                      // [NCT] partial outputs
  float* qs = part + NCT;                           // [hd]
  const float scale = 1.0f / sqrtf((float)hd);
  for (int d = c.tid; d < hd; d += NCT) qs[d] = ldcg_bf(ph.q + (size_t)h * hd + d);
  csync<NCT, CBAR>();
  float mx = -INFINITY;
  for (int j = c.tid; j < nkeys; j += NCT) {
    const bf16* kr = kp + (size_t)j * hd;
    float s = 0.f;
    for (int i = 0; i < hd / 8; ++i) {
      const uint4 v = ldcg16(kr + i * 8);
      const float* q8 = qs + i * 8;
      s = fmaf(q8[0], bflo(v.x), s); s = fmaf(q8[1], bfhi(v.x), s);
      s = fmaf(q8[2], bflo(v.y), s); s = fmaf(q8[3], bfhi(v.y), s);
      s = fmaf(q8[4], bflo(v.z), s); s = fmaf(q8[5], bfhi(v.z), s);
      s = fmaf(q8[6], bflo(v.w), s); s = fmaf(q8[7], bfhi(v.w), s);
    }
    s *= scale;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_max<NCT, CBAR>(mx, c.scratch, c.tid);
  float sum = 0.f;
  for (int j = c.tid; j < nkeys; j += NCT) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_sum<NCT, CBAR>(sum, c.scratch, c.tid);
  const int G = NCT / hd;  // key groups
  const int g = c.tid / hd, d = c.tid % hd;
  float acc = 0.f;
  for (int j = g; j < nkeys; j += G) acc = fmaf(sc[j], ldcg_bf(vp + (size_t)j * hd + d), acc);
  part[c.tid] = acc;
  csync<NCT, CBAR>();
  if (g == 0) {
    for (int gg = 1; gg < G; ++gg) acc += part[gg * hd + d];
    ph.att_out[(size_t)h * hd + d] = f2bf(acc * (1.0f / sum));
  }
}

__device__ void embed_phase(const Phase& ph, Ctx& c) {
  // one 16-byte unit of h per thread: unit u = cta + ncta * tid
  const FrameParams* P = c.P;
  const int u = blockIdx.x + gridDim.x * c.tid;
  if (u >= ph.D / 8) return;
  const size_t fr = (size_t)(P->S - 1);  // last prompt row of stream 0
  const int64_t* tok = P->tokens + fr * (ph.C + 1);
  const uint8_t* msk = P->mask + fr * (ph.C + 1);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int cb = 0; cb <= ph.C; ++cb) {
    if (!msk[cb]) continue;
    const bf16* row = (cb < ph.C) ? ph.audio_emb + ((size_t)tok[cb] + (size_t)ph.V * cb) * ph.D
                                  : ph.text_emb + (size_t)tok[cb] * ph.D;
    const uint4 v = *reinterpret_cast<const uint4*>(row + u * 8);
    acc[0] += bflo(v.x); acc[1] += bfhi(v.x); acc[2] += bflo(v.y); acc[3] += bfhi(v.y);
    acc[4] += bflo(v.z); acc[5] += bfhi(v.z); acc[6] += bflo(v.w); acc[7] += bfhi(v.w);
  }
  __nv_bfloat162 o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
  *reinterpret_cast<uint4*>(ph.h_out + u * 8) = *reinterpret_cast<uint4*>(o);
}

__device__ void sample_phase(const Phase& ph, Ctx& c) {
  if (blockIdx.x != 0) return;
  const FrameParams* P = c.P;
  float* xs = reinterpret_cast<float*>(c.xs);                       // [4096]
  unsigned int* hist = reinterpret_cast<unsigned int*>(xs + SAMPLE_MAXV);  // [256]
  const int V = ph.V, C = ph.C, cb = ph.cb;
  if (P->logits_out)
    for (int i = c.tid; i < V; i += NCT) P->logits_out[(size_t)cb * V + i] = __ushort_as_bfloat16(__ldcg(
        reinterpret_cast<const unsigned short*>(ph.logits) + i));
  const bf16* nz = P->noise ? P->noise + (size_t)cb * V : nullptr;
  const unsigned long long ctr = ((P->offset * (unsigned long long)C + cb)) * 4096ull;
  int tok = sample_row<NCT, CBAR, true>(ph.logits, nz, V, P->temperature, P->topk, P->seed, ctr, xs, hist, c.scratch,
                                        c.iscratch, c.tid);
  if (c.tid == 0 && P->sampled_out) P->sampled_out[cb] = tok;
  if (P->forced) tok = P->forced[cb];
  if (c.tid == 0) P->out[cb] = tok;
  if (ph.next_in) {
    const bf16* row = ph.audio_emb + ((size_t)tok + (size_t)cb * V) * ph.D;
    for (int d8 = c.tid; d8 < ph.D / 8; d8 += NCT)
      *reinterpret_cast<uint4*>(ph.next_in + d8 * 8) = *reinterpret_cast<const uint4*>(row + d8 * 8);
  }
}

// ---- the kernel ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
k_frame_mega(const Phase* __restrict__ phases, int nphases, const FrameParams* __restrict__ P, Sync* sync) {
  extern __shared__ __align__(128) unsigned char smem[];
  bf16* ring = reinterpret_cast<bf16*>(smem);
  bf16* xs = reinterpret_cast<bf16*>(smem + SMEM_RING);
  unsigned char* misc = smem + SMEM_RING + SMEM_X;
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);           // [NW*SLOTS]
  uint64_t* empty = full + NW * SLOTS;                           // [NW*SLOTS]
  float* scratch = reinterpret_cast<float*>(empty + NW * SLOTS);  // [33]
  int* iscratch = reinterpret_cast<int*>(scratch + 34);           // [36]

  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * SLOTS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x >= NCT) {
    producer(phases, nphases, ring, full, empty, sync);
    return;
  }
  Ctx c;
  c.P = P; c.ring = ring; c.xs = xs; c.full = full; c.empty = empty; c.scratch = scratch; c.iscratch = iscratch;
  c.sync = sync; c.cnt = 0; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
  const unsigned ncta = gridDim.x;
  for (int p = 0; p < nphases; ++p) {
    const Phase& ph = phases[p];
    switch (ph.type) {
      case PH_GEMV: gemv_phase(ph, c); break;
      case PH_EMBED: embed_phase(ph, c); break;
      case PH_ATTN: attn_phase(ph, c); break;
      default: sample_phase(ph, c); break;
    }
    if (p + 1 == nphases) break;
    // grid barrier: every CTA's writes of phase p are visible before anyone starts phase p+1
    csync<NCT, CBAR>();
    if (c.tid == 0) {
      __threadfence();
      red_release(&sync->counter, 1u);
      const unsigned target = ncta * (unsigned)(p + 1);
      for (unsigned spin = 0; ld_acquire(&sync->counter) < target; ++spin)
        if (spin > (1u << 24)) die(sync, 0x300);
      __threadfence();
    }
    csync<NCT, CBAR>();
  }
}

}  // namespace mega
