// Persistent decode megakernel: one audio frame (backbone step + codebook-0 sample + 31 depth-
// decoder steps, ~640 dependent phases) in ONE launch, for one stream.
//
// Why: at batch 1 a frame is ~800 tiny dependent GEMVs (0.3-10 us of HBM time each); launched one
// by one they are latency bound (v1: 17 % of the HBM roofline).  Here one CTA per SM stays
// resident for the whole frame:
//   * every warp streams ITS share of every weight matrix, in consumption order, through a
//     private shared-memory ring with bulk async copies (cp.async.bulk + mbarrier complete_tx).
//     Weight addresses are data independent, so the stream runs AHEAD of the dependency chain:
//     148 SMs x 192 KB = 28 MB of weights (~4 us of HBM time) stay in flight while grid barriers,
//     norms, attention and sampling resolve.  The schedule of what to fetch next comes from a
//     compact table passed in kernel-parameter (constant) space;
//   * the same warp computes its output-row groups out of shared memory (fp32 accumulate, warp
//     shuffle reduction) and runs the fused epilogue (RoPE + KV append / residual / SwiGLU /
//     logits) at the reference's bf16 rounding points, then refills the slot it just drained;
//   * phases are separated by a monotonic grid barrier (release/acquire counter in L2); the next
//     phase's descriptor is prefetched into shared memory while the current one computes.
// Every spin has a trip-count cap and traps instead of hanging the GPU.
#pragma once
#include "lm_kernels.cuh"

namespace mega {

#ifndef MEGA_NW
#define MEGA_NW 8
#endif
#ifndef MEGA_SLOTS
#define MEGA_SLOTS 2
#endif
#ifndef MEGA_CHUNK
#define MEGA_CHUNK 4096
#endif
#ifndef MEGA_L2_AHEAD
#define MEGA_L2_AHEAD 0  /* measured: an extra L2 prefetch stage slows the stream (5.15 vs 4.3 ms/frame) */
#endif
constexpr int NW = MEGA_NW;           // warps per CTA
constexpr int NCT = NW * 32;          // threads per CTA
constexpr int SLOTS = MEGA_SLOTS;     // ring slots per warp
constexpr int CHUNK_ELEMS = MEGA_CHUNK;  // bf16 per slot
constexpr int KC_MAX = CHUNK_ELEMS / 2;  // k-extent of a chunk (a chunk holds >= 2 rows)
constexpr int L2_AHEAD = MEGA_L2_AHEAD;  // chunks per warp that are pulled into L2 ahead of the smem ring
constexpr int MAXNB = 2;              // activation rows per phase (depth step 1 carries 2)
constexpr int XBUF_ELEMS = 2 * MAXNB * 8192;  // 64 KB: activation rows, or the fused attention's q / K / V staging
constexpr int CBAR = 0;               // all threads are consumers: plain CTA barrier
constexpr int MAX_GEMV = 640;         // rows of the prefetch table (kernel parameter space)
constexpr size_t SMEM_RING = (size_t)NW * SLOTS * CHUNK_ELEMS * 2;
constexpr size_t SMEM_X = (size_t)XBUF_ELEMS * 2;
constexpr size_t SMEM_MISC = 4096;
constexpr int MAX_SPLIT_TASKS = 128;  // (row group, k-chunk) tasks per CTA in a split-K phase
constexpr size_t SMEM_BYTES = SMEM_RING + SMEM_X + SMEM_MISC;

enum { PH_GEMV = 0, PH_EMBED = 1, PH_ATTN = 2, PH_SAMPLE = 3 };
enum { POS_FIXED = 0, POS_BACKBONE = 1 };

// Full description of a phase (global memory; staged into shared memory one phase ahead).
struct __align__(16) Phase {
  int type, epi, norm, nb;
  const bf16* W;
  int rows, K, R, KC, G, ldx;
  const bf16* x;  // [nb, ldx] activations in global memory
  const bf16* norm_scale;
  bf16* out;
  const bf16* resid;
  bf16* x_copy_out;  // CTA 0 publishes the staged (normalised) rows here
  float eps;
  int ldo;
  // RoPE / KV append / attention
  bf16 *q, *kc, *vc;
  const bf16* rope;
  int heads, kv_heads, hd, slots, pos_mode, pos0;
  int attn_prologue;  // GEMV: x = attention(q, cache) computed redundantly by every CTA (<= 32 keys)
  int cb;
  bf16* att_out;  // PH_ATTN: [heads*hd]
  // sample / embed
  int V, C, D, ldl;
  const bf16* logits;
  bf16* next_in;
  const bf16 *audio_emb, *text_emb;
  bf16* h_out;
  // PLAIN epilogue with two destinations: rows >= split_row go to out2[row - split_row] (stacked
  // [codebook0_head; projection] matrix: logits and the depth decoder's position-0 input in one phase)
  bf16* out2;
  int split_row;
  int next_ld;  // SAMPLE: row length of the gather table feeding next_in
  const bf16* next_table;  // SAMPLE: projection(embedding) table [(codebooks-1)*V][Dd], or null -> audio_emb rows
};
static_assert(sizeof(Phase) % 16 == 0, "Phase must be copyable in 16-byte units");

// What the weight prefetcher needs to know about a GEMV phase.
struct PfDesc {
  const bf16* W;
  int rows, K, G, R;  // R rows per chunk (small phases use half-size chunks so more warps share them)
};
struct PfTable {
  int n;
  int pad_[3];
  PfDesc d[MAX_GEMV];
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// Weights are streamed with an L2 evict-first policy: they are read once per use and must not push
// the small hot data (activations, KV cache, norm scales, RoPE tables) out of the 126 MB L2 --
// otherwise every dependent load of a phase queues behind ~28 MB of in-flight weight traffic.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// Second, deeper stage of the weight stream: pull a chunk into the 126 MB L2 well before its turn
// in the shared-memory ring.  HBM then keeps streaming while every CTA sits in a latency chain
// (barrier, activation load, epilogue), and the ring refills at L2 speed.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void arrive_release(unsigned* p) {
  // the fence also releases what the other threads of the CTA wrote before the CTA barrier
  asm volatile("fence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float ldcg_bf(const bf16* p) {
  return bf2f(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
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

// work distribution: the i-th row group owned by (cta, warp)
__device__ __forceinline__ int group_of(int cta, int ncta, int warp, int i) { return cta + ncta * (warp + NW * i); }

// ---- per-warp weight prefetcher -----------------------------------------------------------------------
// A phase's work is cut into tasks (row group j of this CTA, k-chunk kc), numbered u = j*nkc + kc and
// dealt round-robin to the warps (u = warp, warp + NW, ...).  For K <= KC_MAX a task is a whole row
// group; for the down projections (K = 8192) the k-chunks of one row pair go to different warps, so
// all warps stream in parallel and the partial sums meet in shared memory.
struct Prefetch {
  int gi;  // index into the GEMV table
  int u;   // task index inside the phase
  unsigned issued;
  bool done;
  uint64_t policy;  // L2 cache policy of the weight stream
};

__device__ __forceinline__ int nkc_of(int K) { return K <= KC_MAX ? 1 : K / KC_MAX; }

__device__ __forceinline__ void pf_seek(Prefetch& pf, const PfTable& tab, int cta, int ncta, int warp) {
  // position on the first (phase, task) at or after (gi, u) that this warp owns
  while (pf.gi < tab.n && cta + ncta * (pf.u / nkc_of(tab.d[pf.gi].K)) >= tab.d[pf.gi].G) {
    ++pf.gi;
    pf.u = warp;
  }
  pf.done = pf.gi >= tab.n;
}

// L2 prefetch of the chunk the cursor points at, then advance (no shared-memory slot involved)
__device__ __forceinline__ void pf_issue_l2(Prefetch& pf, const PfTable& tab, int cta, int ncta, int warp, int lane) {
  if (pf.done) return;
  const PfDesc& d = tab.d[pf.gi];
  const int K = d.K, KC = K < KC_MAX ? K : KC_MAX, R = d.R, nkc = K / KC;
  const int j = pf.u / nkc, kc = pf.u - j * nkc;
  const int r0 = (cta + ncta * j) * R;
  const int nr = min(R, d.rows - r0);
  if (nkc == 1) {
    if (lane == 0) bulk_prefetch_l2(d.W + (size_t)r0 * K, (uint32_t)nr * K * 2);
  } else if (lane < nr) {
    bulk_prefetch_l2(d.W + (size_t)(r0 + lane) * K + (size_t)kc * KC, (uint32_t)KC * 2);
  }
  pf.u += NW;
  pf_seek(pf, tab, cta, ncta, warp);
}

// issue the next chunk of this warp's stream into ring slot (issued % SLOTS); whole warp calls it
__device__ __forceinline__ void pf_issue(Prefetch& pf, const PfTable& tab, bf16* ring, uint64_t* full, int cta, int ncta,
                                         int warp, int lane) {
  if (pf.done) return;
  const PfDesc& d = tab.d[pf.gi];
  const int K = d.K, KC = K < KC_MAX ? K : KC_MAX, R = d.R, nkc = K / KC;
  const int j = pf.u / nkc, kc = pf.u - j * nkc;
  const int r0 = (cta + ncta * j) * R;
  const int nr = min(R, d.rows - r0);
  const int slot = pf.issued % SLOTS;
  if (lane == 0) {
    uint64_t* fb = &full[warp * SLOTS + slot];
    bf16* dst = ring + (size_t)(warp * SLOTS + slot) * CHUNK_ELEMS;
    if (nkc == 1) {  // whole rows are contiguous: one copy
      mbar_expect_tx(fb, (uint32_t)nr * K * 2);
      bulk_g2s(dst, d.W + (size_t)r0 * K, (uint32_t)nr * K * 2, fb, pf.policy);
    } else {
      mbar_expect_tx(fb, (uint32_t)nr * KC * 2);
      for (int r = 0; r < nr; ++r)
        bulk_g2s(dst + r * KC, d.W + (size_t)(r0 + r) * K + (size_t)kc * KC, (uint32_t)KC * 2, fb, pf.policy);
    }
  }
  ++pf.issued;
  pf.u += NW;
  pf_seek(pf, tab, cta, ncta, warp);
}

// ---- consumer pieces ---------------------------------------------------------------------------------
struct Ctx {
  const FrameParams* P;
  bf16* ring;
  bf16* xs;
  uint64_t* full;
  float* scratch;  // 33 floats
  float* psum;     // [MAX_SPLIT_TASKS][4] split-K partial sums
  int* iscratch;   // 40 ints
  Sync* sync;
  Prefetch pf;   // shared-memory ring cursor
  Prefetch pf2;  // L2 prefetch cursor, L2_AHEAD chunks further down the same stream
  unsigned cnt;  // chunks consumed by this warp so far
  unsigned long long* trp;  // fine-grained trace slots of the current phase (CTA 0, thread 0) or null
  int tid, warp, lane;
  int bb_pos, bb_slot;  // RoPE position / cache slot of the backbone row of this frame
  // operands of the CURRENT phase that were fetched while the previous grid barrier was spinning
  uint4 pre_scale;   // this thread's 16-byte unit of the RMSNorm scale
  float pre_a, pre_b;  // epilogue operands of this lane's pair in the warp's first row group
  bool pre_valid;
};

__device__ __forceinline__ void phase_pos(const Phase& ph, const Ctx& c, int n, int& pos, int& slot) {
  if (ph.pos_mode == POS_BACKBONE) {
    pos = c.bb_pos;
    slot = c.bb_slot;
  } else {
    pos = slot = ph.pos0 + n;
  }
}

// Operands the epilogue of a row pair needs from global memory (residual values / RoPE cos,sin).
// They are fetched when the warp STARTS a row group, so their L2 round trip -- ~1.5 us while the
// weight stream saturates the memory system -- overlaps the dot products instead of following them.
struct EpiPre {
  float a, b;
};
__device__ __forceinline__ EpiPre epilogue_prefetch(const Phase& ph, const Ctx& c, int r0, int n) {
  EpiPre e;
  e.a = e.b = 0.f;
  if (r0 >= ph.rows || n >= ph.nb) return e;
  if (ph.epi == EPI_RESID) {
    e.a = ldcg_bf(ph.resid + (size_t)n * ph.ldo + r0);
    if (r0 + 1 < ph.rows) e.b = ldcg_bf(ph.resid + (size_t)n * ph.ldo + r0 + 1);
  } else if (ph.epi == EPI_ROPE_KV) {
    const int hd = ph.hd;
    if (r0 < (ph.heads + ph.kv_heads) * hd) {
      int pos, slot;
      phase_pos(ph, c, n, pos, slot);
      const __nv_bfloat162 cs = *reinterpret_cast<const __nv_bfloat162*>(ph.rope + ((size_t)pos * (hd / 2) + ((r0 % hd) >> 1)) * 2);
      e.a = __low2float(cs);
      e.b = __high2float(cs);
    }
  }
  return e;
}

// fused epilogue of one output-row pair (row r0, r0+1) for activation row n
__device__ __forceinline__ void epilogue(const Phase& ph, const Ctx& c, int r0, int n, float a0, float a1,
                                         const EpiPre& pre) {
  const bool has1 = r0 + 1 < ph.rows;
  const float y0 = rbf(a0), y1 = rbf(a1);
  if (ph.epi == EPI_PLAIN) {
    if (ph.out2 && r0 >= ph.split_row) {
      ph.out2[r0 - ph.split_row] = f2bf(y0);
      if (has1) ph.out2[r0 + 1 - ph.split_row] = f2bf(y1);
    } else {
      ph.out[(size_t)n * ph.ldo + r0] = f2bf(y0);
      if (has1) ph.out[(size_t)n * ph.ldo + r0 + 1] = f2bf(y1);
    }
  } else if (ph.epi == EPI_RESID) {
    ph.out[(size_t)n * ph.ldo + r0] = f2bf(y0 + pre.a);
    if (has1) ph.out[(size_t)n * ph.ldo + r0 + 1] = f2bf(y1 + pre.b);
  } else if (ph.epi == EPI_SWIGLU) {
    ph.out[(size_t)n * ph.ldo + (r0 >> 1)] = f2bf(silu_bf(y0) * y1);
  } else {  // EPI_ROPE_KV
    const int hd = ph.hd, qrows = ph.heads * hd, krows = ph.kv_heads * hd;
    int pos, slot;
    phase_pos(ph, c, n, pos, slot);
    float o0 = y0, o1 = y1;
    if (r0 < qrows + krows) {
      o0 = rbf(__fsub_rn(__fmul_rn(y0, pre.a), __fmul_rn(y1, pre.b)));
      o1 = rbf(__fadd_rn(__fmul_rn(y1, pre.a), __fmul_rn(y0, pre.b)));
    }
    if (r0 < qrows) {
      *reinterpret_cast<__nv_bfloat162*>(ph.q + (size_t)n * qrows + r0) = __floats2bfloat162_rn(o0, o1);
    } else {
      const bool isk = r0 < qrows + krows;
      const int rr = r0 - (isk ? qrows : qrows + krows);
      const int kvh = rr / hd, d = rr % hd;
      bf16* dst = (isk ? ph.kc : ph.vc) + ((size_t)kvh * ph.slots + slot) * hd + d;  // stream 0
      *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
    }
  }
}

// dot products of one ring chunk ([R][KC] weights) with the staged activations
template <int R, int NB>
__device__ __forceinline__ void chunk_dot(const bf16* chunk, const bf16* xk, int K, int KC, int lane, float (&acc)[R][NB]) {
  // two independent accumulator sets (even / odd 256-element blocks): the FMA chains, not the
  // shared-memory bandwidth, bound this loop with only two warps per scheduler
  float acc2[R][NB];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc2[r][n] = 0.f;
  for (int k = lane * 8; k < KC; k += 512) {
    const bool two = k + 256 < KC;  // KC is a multiple of 256, not necessarily of 512
    const int k2 = two ? k + 256 : k;
    uint4 xa[NB], xb[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      xa[n] = *reinterpret_cast<const uint4*>(xk + (size_t)n * K + k);
      xb[n] = two ? *reinterpret_cast<const uint4*>(xk + (size_t)n * K + k2) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint4 wa = *reinterpret_cast<const uint4*>(chunk + r * KC + k);
      const uint4 wb = *reinterpret_cast<const uint4*>(chunk + r * KC + k2);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        acc[r][n] = dot8(wa, xa[n], acc[r][n]);
        acc2[r][n] = dot8(wb, xb[n], acc2[r][n]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[r][n] += acc2[r][n];
}

// The warp's tasks of this phase: chunks come from its ring; after the shuffle reduction lane
// (pair, n) runs that pair's epilogue (K <= KC_MAX), or the partial sums go to shared memory and a
// post pass adds the k-chunks in a fixed order and runs the epilogue (split-K phases).
template <int R, int NB>
__device__ __forceinline__ void gemv_groups(const Phase& ph, Ctx& c, const PfTable& tab) {
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int K = ph.K, KC = ph.KC, nkc = K / KC;
  const int my_pr = c.lane % (R / 2), my_n = c.lane / (R / 2);
  const bool epi_lane = c.lane < (R / 2) * NB;
  if (nkc == 1) {
    for (int i = 0;; ++i) {
      const int g = cta + ncta * (c.warp + NW * i);
      if (g >= ph.G) break;
      EpiPre pre;
      pre.a = pre.b = 0.f;
      if (i == 0 && c.pre_valid) {
        pre.a = c.pre_a;
        pre.b = c.pre_b;
      } else if (epi_lane) {
        pre = epilogue_prefetch(ph, c, g * R + 2 * my_pr, my_n);
      }
      float acc[R][NB];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[r][n] = 0.f;
      const int slot = c.cnt % SLOTS;
      mbar_wait(&c.full[c.warp * SLOTS + slot], (c.cnt / SLOTS) & 1, c.sync, 0x200 + c.warp);
      if (c.trp && i == 0 && !ph.attn_prologue) c.trp[1] = gtimer();
      chunk_dot<R, NB>(c.ring + (size_t)(c.warp * SLOTS + slot) * CHUNK_ELEMS, c.xs, K, KC, c.lane, acc);
      __syncwarp();
      ++c.cnt;
      if (c.trp && i == 0 && !ph.attn_prologue) c.trp[2] = gtimer();
      // the slot is drained: refill it with the chunk SLOTS ahead in this warp's stream
      pf_issue(c.pf, tab, c.ring, c.full, cta, ncta, c.warp, c.lane);
      if (L2_AHEAD > 0) pf_issue_l2(c.pf2, tab, cta, ncta, c.warp, c.lane);
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
      if (epi_lane) {
        const int r0 = g * R + 2 * my_pr;
        if (r0 < ph.rows && my_n < ph.nb) epilogue(ph, c, r0, my_n, y0, y1, pre);
      }
      if (c.trp && i == 0 && !ph.attn_prologue) c.trp[3] = gtimer();
    }
  } else {
    // split K: post-pass item t = (row group j of this CTA, pair, activation row)
    const int ngroups = ph.G > cta ? (ph.G - cta + ncta - 1) / ncta : 0;
    const int per_group = (R / 2) * ph.nb;
    const int items = ngroups * per_group;
    int pj = 0, pr0 = 0, pn = 0;
    EpiPre pre;
    pre.a = pre.b = 0.f;
    if (c.tid < items) {
      pj = c.tid / per_group;
      const int rem = c.tid - pj * per_group;
      pn = rem / (R / 2);
      pr0 = (cta + ncta * pj) * R + 2 * (rem % (R / 2));
      pre = epilogue_prefetch(ph, c, pr0, pn);
    }
    for (int u = c.warp;; u += NW) {
      const int j = u / nkc, kc = u - j * nkc;
      if (cta + ncta * j >= ph.G) break;
      float acc[R][NB];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[r][n] = 0.f;
      const int slot = c.cnt % SLOTS;
      mbar_wait(&c.full[c.warp * SLOTS + slot], (c.cnt / SLOTS) & 1, c.sync, 0x200 + c.warp);
      if (c.trp && u == c.warp) c.trp[1] = gtimer();
      chunk_dot<R, NB>(c.ring + (size_t)(c.warp * SLOTS + slot) * CHUNK_ELEMS, c.xs + (size_t)kc * KC, K, KC, c.lane, acc);
      __syncwarp();
      ++c.cnt;
      if (c.trp && u == c.warp) c.trp[2] = gtimer();
      pf_issue(c.pf, tab, c.ring, c.full, cta, ncta, c.warp, c.lane);
      if (L2_AHEAD > 0) pf_issue_l2(c.pf2, tab, cta, ncta, c.warp, c.lane);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          const float sr = warp_sum(acc[r][n]);
          if (c.lane == r * NB + n) c.psum[(u * R + r) * MAXNB + n] = sr;
        }
    }
    csync<NCT, CBAR>();
    if (c.trp) c.trp[3] = gtimer();
    if (c.tid < items) {
      const int pr = (pr0 - (cta + ncta * pj) * R);  // row inside the group (even)
      float y0 = 0.f, y1 = 0.f;
      for (int kc = 0; kc < nkc; ++kc) {  // fixed order -> deterministic rounding
        y0 += c.psum[((pj * nkc + kc) * R + pr) * MAXNB + pn];
        y1 += c.psum[((pj * nkc + kc) * R + pr + 1) * MAXNB + pn];
      }
      epilogue(ph, c, pr0, pn, y0, y1, pre);
    }
  }
}

// Attention over <= 32 cached keys for every (row, head), redundantly in every CTA, straight into
// the activation buffer of the output projection (depth decoder: 32 slots per stream, head_dim 128).
//   xs layout (bf16 elements): [0, nb*dim) output rows | [2048, +nb*dim) q | [4096, +kv*keys*136) K
//   rows padded to 136 (conflict-free 16-byte reads with lane == key) | [12800, +kv*32*128) V |
//   partial scores (fp32).
// Two parts.  attn_prefetch runs while the PREVIOUS phase's grid barrier is still spinning: the K/V
// rows of earlier positions are final, so they are copied into shared memory with cp.async there and
// their (loaded) L2 latency disappears behind the barrier.  attn_small_into_x then only fetches q and
// the current position's K/V rows, and computes out of shared memory: a warp owns (row, head, half of
// the head dims): partial q.k over its 64 dims with lane == key, halves summed through shared
// memory, softmax by shuffles, P.V for its 64 output dims with lane == 2 dims.
constexpr int A_HD = 128, A_KS = A_HD + 8, A_QOFF = 2048, A_KOFF = 4096, A_VOFF = A_KOFF + 2 * 32 * A_KS;
constexpr int A_PARTOFF = A_VOFF + 2 * 32 * A_HD;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void attn_prefetch(const Phase& ph, Ctx& c) {
  const int kvn = ph.kv_heads, nold = ph.pos0;  // positions [0, pos0) were written in earlier steps
  const int units = kvn * nold * (A_HD / 8);
  for (int u = c.tid; u < 2 * units; u += NCT) {
    const bool isv = u >= units;
    const int ku = isv ? u - units : u, row = ku >> 4, i = ku & 15;
    const int kvh = row / nold, j = row - kvh * nold;
    const bf16* src = (isv ? ph.vc : ph.kc) + (kvh * ph.slots + j) * A_HD + i * 8;
    bf16* dst = isv ? c.xs + A_VOFF + (kvh * 32 + j) * A_HD + i * 8 : c.xs + A_KOFF + (kvh * 32 + j) * A_KS + i * 8;
    cp_async16(dst, src);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void attn_small_into_x(const Phase& ph, Ctx& c) {
  const int heads = ph.heads, kvn = ph.kv_heads, grp = heads / kvn, nb = ph.nb;
  float* part = reinterpret_cast<float*>(c.xs + A_PARTOFF);  // [nb*heads*2][32]
  const int nitems = nb * heads * 2;
  {  // q rows and the K/V rows of the positions written by the phase that just finished
    const int qunits = nb * heads * (A_HD / 8), kunits = kvn * nb * (A_HD / 8), total = qunits + 2 * kunits;
    for (int u = c.tid; u < total; u += NCT) {
      if (u < qunits) {
        *reinterpret_cast<uint4*>(c.xs + A_QOFF + u * 8) = ldcg16(ph.q + u * 8);
      } else {
        const bool isv = u >= qunits + kunits;
        const int ku = u - qunits - (isv ? kunits : 0), row = ku >> 4, i = ku & 15;
        const int kvh = row / nb, j = ph.pos0 + row - kvh * nb;
        const uint4 v = ldcg16((isv ? ph.vc : ph.kc) + (kvh * ph.slots + j) * A_HD + i * 8);
        if (isv) *reinterpret_cast<uint4*>(c.xs + A_VOFF + (kvh * 32 + j) * A_HD + i * 8) = v;
        else *reinterpret_cast<uint4*>(c.xs + A_KOFF + (kvh * 32 + j) * A_KS + i * 8) = v;
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  csync<NCT, CBAR>();
  if (c.trp) c.trp[1] = gtimer();
  const float scale = 0.08838834764831845f;  // 1/sqrt(128)
  for (int it0 = 0; it0 < nitems; it0 += NW) {
    const int item = it0 + c.warp;
    if (item < nitems) {
      const int n = item / (heads * 2), h = (item >> 1) % heads, half = item & 1, kvh = h / grp;
      const int nkeys = ph.pos0 + n + 1;
      const bool own = c.lane < nkeys;
      const bf16* kr = c.xs + A_KOFF + (kvh * 32 + (own ? c.lane : 0)) * A_KS + half * 64;
      const bf16* qr = c.xs + A_QOFF + (n * heads + h) * A_HD + half * 64;
      float a0 = 0.f, a1 = 0.f;  // two chains: the dot product is latency, not throughput, bound
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        a0 = dot8(*reinterpret_cast<const uint4*>(kr + i * 8), *reinterpret_cast<const uint4*>(qr + i * 8), a0);
        a1 = dot8(*reinterpret_cast<const uint4*>(kr + i * 8 + 8), *reinterpret_cast<const uint4*>(qr + i * 8 + 8), a1);
      }
      part[item * 32 + c.lane] = a0 + a1;
    }
  }
  csync<NCT, CBAR>();
  if (c.trp) c.trp[2] = gtimer();
  for (int it0 = 0; it0 < nitems; it0 += NW) {
    const int item = it0 + c.warp;
    if (item < nitems) {
      const int n = item / (heads * 2), h = (item >> 1) % heads, half = item & 1, kvh = h / grp;
      const int nkeys = ph.pos0 + n + 1;
      const bool own = c.lane < nkeys;
      const float sc = own ? (part[(item & ~1) * 32 + c.lane] + part[(item | 1) * 32 + c.lane]) * scale : -INFINITY;
      const float mx = warp_max(sc);
      const float e = own ? expf(sc - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e);
      const bf16* vp = c.xs + A_VOFF + kvh * 32 * A_HD + half * 64 + c.lane * 2;
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent chains
      for (int j0 = 0; j0 < nkeys; j0 += 4) {  // only the visible keys, 4 independent chains
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = j0 + t;
          const float pj = __shfl_sync(0xffffffffu, e, j & 31);  // 0 for j >= nkeys
          const uint32_t v = j < nkeys ? *reinterpret_cast<const uint32_t*>(vp + j * A_HD) : 0u;  // rows past nkeys are stale
          o0[t] = fmaf(pj, bflo(v), o0[t]);
          o1[t] = fmaf(pj, bfhi(v), o1[t]);
        }
      }
      *reinterpret_cast<__nv_bfloat162*>(c.xs + n * ph.K + h * A_HD + half * 64 + c.lane * 2) =
          __floats2bfloat162_rn(((o0[0] + o0[1]) + (o0[2] + o0[3])) * inv, ((o1[0] + o1[1]) + (o1[2] + o1[3])) * inv);
    }
  }
  if (c.trp) c.trp[3] = gtimer();
  csync<NCT, CBAR>();
}

__device__ __forceinline__ float sumsq8(const uint4& v) {
  float s = 0.f, a;
  a = bflo(v.x); s = fmaf(a, a, s); a = bfhi(v.x); s = fmaf(a, a, s);
  a = bflo(v.y); s = fmaf(a, a, s); a = bfhi(v.y); s = fmaf(a, a, s);
  a = bflo(v.z); s = fmaf(a, a, s); a = bfhi(v.z); s = fmaf(a, a, s);
  a = bflo(v.w); s = fmaf(a, a, s); a = bfhi(v.w); s = fmaf(a, a, s);
  return s;
}
// torchtune RMSNorm on 8 elements: bf16( bf16(x * inv) * scale )
__device__ __forceinline__ uint4 norm8(const uint4& v, float inv, const uint4& sc) {
  __nv_bfloat162 o[4];
  o[0] = __floats2bfloat162_rn(rbf(bflo(v.x) * inv) * bflo(sc.x), rbf(bfhi(v.x) * inv) * bfhi(sc.x));
  o[1] = __floats2bfloat162_rn(rbf(bflo(v.y) * inv) * bflo(sc.y), rbf(bfhi(v.y) * inv) * bfhi(sc.y));
  o[2] = __floats2bfloat162_rn(rbf(bflo(v.z) * inv) * bflo(sc.z), rbf(bfhi(v.z) * inv) * bfhi(sc.z));
  o[3] = __floats2bfloat162_rn(rbf(bflo(v.w) * inv) * bflo(sc.w), rbf(bfhi(v.w) * inv) * bfhi(sc.w));
  return *reinterpret_cast<uint4*>(o);
}

// stage the phase's activation rows into shared memory (+ RMSNorm prologue)
__device__ __forceinline__ void stage_x(const Phase& ph, Ctx& c) {
  if (ph.attn_prologue) {
    attn_small_into_x(ph, c);
    return;
  }
  const int K = ph.K, nb = ph.nb;
  if (ph.norm) {
    // K <= 2048 for every normed phase: one 16-byte unit per thread; x and scale loads together
    const int k8 = c.tid;
    const bool on = k8 < K / 8;
    uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0, sc = x0;
    if (on) {
      x0 = ldcg16(ph.x + k8 * 8);
      if (nb == 2) x1 = ldcg16(ph.x + ph.ldx + k8 * 8);
      sc = c.pre_valid ? c.pre_scale : *reinterpret_cast<const uint4*>(ph.norm_scale + k8 * 8);
    }
    float s0 = sumsq8(x0), s1 = sumsq8(x1);
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (c.lane == 0) {
      c.scratch[c.warp] = s0;
      c.scratch[NW + c.warp] = s1;
    }
    csync<NCT, CBAR>();
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      t0 += c.scratch[w];
      t1 += c.scratch[NW + w];
    }
    if (on) {
      *reinterpret_cast<uint4*>(c.xs + k8 * 8) = norm8(x0, 1.0f / sqrtf(t0 / (float)K + ph.eps), sc);
      if (nb == 2) *reinterpret_cast<uint4*>(c.xs + K + k8 * 8) = norm8(x1, 1.0f / sqrtf(t1 / (float)K + ph.eps), sc);
    }
  } else {
    constexpr int MAXIT = 8192 / 8 / NCT;
    uint4 v0[MAXIT], v1[MAXIT];
#pragma unroll
    for (int t = 0; t < MAXIT; ++t) {
      const int k8 = c.tid + t * NCT;
      if (k8 < K / 8) {
        v0[t] = ldcg16(ph.x + k8 * 8);
        if (nb == 2) v1[t] = ldcg16(ph.x + ph.ldx + k8 * 8);
      }
    }
#pragma unroll
    for (int t = 0; t < MAXIT; ++t) {
      const int k8 = c.tid + t * NCT;
      if (k8 < K / 8) {
        *reinterpret_cast<uint4*>(c.xs + k8 * 8) = v0[t];
        if (nb == 2) *reinterpret_cast<uint4*>(c.xs + K + k8 * 8) = v1[t];
      }
    }
  }
  csync<NCT, CBAR>();
  if (ph.x_copy_out && blockIdx.x == 0) {
    for (int k8 = c.tid; k8 < nb * (K / 8); k8 += NCT)
      *reinterpret_cast<uint4*>(ph.x_copy_out + k8 * 8) = *reinterpret_cast<const uint4*>(c.xs + k8 * 8);
  }
}

__device__ __forceinline__ void gemv_phase(const Phase& ph, Ctx& c, const PfTable& tab) {
  stage_x(ph, c);
  if (c.trp) c.trp[0] = gtimer();
  if (ph.nb == 1) {
    if (ph.R == 2) gemv_groups<2, 1>(ph, c, tab);
    else if (ph.R == 4) gemv_groups<4, 1>(ph, c, tab);
    else if (ph.R == 8) gemv_groups<8, 1>(ph, c, tab);
    else gemv_groups<16, 1>(ph, c, tab);
  } else {
    if (ph.R == 2) gemv_groups<2, 2>(ph, c, tab);
    else if (ph.R == 4) gemv_groups<4, 2>(ph, c, tab);
    else if (ph.R == 8) gemv_groups<8, 2>(ph, c, tab);
    else gemv_groups<16, 2>(ph, c, tab);
  }
}

// backbone attention: CTA h < heads owns q-head h over keys [0 .. slot]
__device__ __forceinline__ void attn_phase(const Phase& ph, Ctx& c) {
  const int h = blockIdx.x;
  if (h >= ph.heads) return;
  const int hd = ph.hd;
  int pos, slot;
  phase_pos(ph, c, 0, pos, slot);
  const int nkeys = slot + 1;
  const int kvh = h / (ph.heads / ph.kv_heads);
  const bf16* kp = ph.kc + (size_t)kvh * ph.slots * hd;
  const bf16* vp = ph.vc + (size_t)kvh * ph.slots * hd;
  float* sc = reinterpret_cast<float*>(c.xs);  // [slots] scores (<= 8 KB)
  float* part = sc + ph.slots;                  // [NCT] partial outputs
  float* qs = part + NCT;                       // [hd]
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
  int j = g;
  for (; j + 7 * G < nkeys; j += 8 * G) {  // 8 independent loads in flight
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ldcg_bf(vp + (size_t)(j + u * G) * hd + d);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fmaf(sc[j + u * G], v[u], acc);
  }
  for (; j < nkeys; j += G) acc = fmaf(sc[j], ldcg_bf(vp + (size_t)j * hd + d), acc);
  part[c.tid] = acc;
  csync<NCT, CBAR>();
  if (g == 0) {
    for (int gg = 1; gg < G; ++gg) acc += part[gg * hd + d];
    ph.att_out[(size_t)h * hd + d] = f2bf(acc * (1.0f / sum));
  }
}

__device__ __forceinline__ void embed_phase(const Phase& ph, Ctx& c) {
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
  for (int c0 = 0; c0 <= ph.C; c0 += 11) {  // 11 independent gathers in flight, summed in column order
    uint4 v[11];
    bool on[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const int cb = c0 + j;
      on[j] = cb <= ph.C && msk[cb < ph.C + 1 ? cb : 0] != 0;
      v[j] = make_uint4(0, 0, 0, 0);
      if (on[j]) {
        const bf16* row = (cb < ph.C) ? ph.audio_emb + ((size_t)tok[cb] + (size_t)ph.V * cb) * ph.D
                                      : ph.text_emb + (size_t)tok[cb] * ph.D;
        v[j] = *reinterpret_cast<const uint4*>(row + u * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < 11; ++j)
      if (on[j]) {
        acc[0] += bflo(v[j].x); acc[1] += bfhi(v[j].x); acc[2] += bflo(v[j].y); acc[3] += bfhi(v[j].y);
        acc[4] += bflo(v[j].z); acc[5] += bfhi(v[j].z); acc[6] += bflo(v[j].w); acc[7] += bfhi(v[j].w);
      }
  }
  __nv_bfloat162 o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
  *reinterpret_cast<uint4*>(ph.h_out + u * 8) = *reinterpret_cast<uint4*>(o);
}

__device__ __forceinline__ void sample_phase(const Phase& ph, Ctx& c) {
  if (blockIdx.x != 0) return;
  const FrameParams* P = c.P;
  float* xs = reinterpret_cast<float*>(c.xs);                              // [4096]
  unsigned int* hist = reinterpret_cast<unsigned int*>(xs + SAMPLE_MAXV);  // [256]
  const int V = ph.V, C = ph.C, cb = ph.cb;
  if (P->logits_out)
    for (int i = c.tid; i < V; i += NCT) P->logits_out[(size_t)cb * V + i] = __ushort_as_bfloat16(__ldcg(
        reinterpret_cast<const unsigned short*>(ph.logits) + i));
  const bf16* nz = P->noise ? P->noise + (size_t)cb * V : nullptr;
  const unsigned long long ctr = ((P->offset * (unsigned long long)C + cb)) * 4096ull;
  if (c.trp) c.trp[0] = gtimer();
  int tok = sample_row<NCT, CBAR, true>(ph.logits, nz, V, P->temperature, P->topk, P->seed, ctr, xs, hist, c.scratch,
                                        c.iscratch, c.tid);
  if (c.trp) c.trp[1] = gtimer();
  if (c.tid == 0 && P->sampled_out) P->sampled_out[cb] = tok;
  if (P->forced) tok = P->forced[cb];
  if (c.tid == 0) P->out[cb] = tok;
  if (ph.next_in) {
    // next depth-decoder input: projection(embed_audio(cb, tok)) read from the table built at setup
    const int ld = ph.next_table ? ph.next_ld : ph.D;
    const bf16* row = (ph.next_table ? ph.next_table : ph.audio_emb) + ((size_t)tok + (size_t)cb * V) * ld;
    for (int d8 = c.tid; d8 < ld / 8; d8 += NCT)
      *reinterpret_cast<uint4*>(ph.next_in + d8 * 8) = *reinterpret_cast<const uint4*>(row + d8 * 8);
  }
}

// ---- the kernel ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NCT, 1)
k_frame_mega(const Phase* __restrict__ phases, int nphases, const FrameParams* __restrict__ P, Sync* sync,
             unsigned long long* __restrict__ trace /* optional [nphases][8] globaltimer ns of CTA 0 */,
             const __grid_constant__ PfTable tab) {
  extern __shared__ __align__(128) unsigned char smem[];
  bf16* ring = reinterpret_cast<bf16*>(smem);
  bf16* xs = reinterpret_cast<bf16*>(smem + SMEM_RING);
  unsigned char* misc = smem + SMEM_RING + SMEM_X;
  Phase* phbuf = reinterpret_cast<Phase*>(misc);                           // [2] double buffer
  uint64_t* full = reinterpret_cast<uint64_t*>(misc + 2 * sizeof(Phase));  // [NW*SLOTS]
  float* scratch = reinterpret_cast<float*>(full + NW * SLOTS);            // [34]
  int* iscratch = reinterpret_cast<int*>(scratch + 34);                    // [40]
  float* psum = reinterpret_cast<float*>(iscratch + 40);                   // [MAX_SPLIT_TASKS * 2 * MAXNB]
  static_assert(2 * sizeof(Phase) + NW * SLOTS * 8 + 34 * 4 + 40 * 4 + MAX_SPLIT_TASKS * 2 * MAXNB * 4 <= SMEM_MISC,
                "misc region too small");
  static_assert(NW <= 32, "block reductions assume <= 32 warps");

  Ctx c;
  c.P = P; c.trp = nullptr; c.ring = ring; c.xs = xs; c.full = full; c.scratch = scratch; c.iscratch = iscratch; c.psum = psum;
  c.sync = sync; c.cnt = 0; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
  c.pre_valid = false; c.pre_a = c.pre_b = 0.f; c.pre_scale = make_uint4(0, 0, 0, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * SLOTS; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  constexpr int PH16 = sizeof(Phase) / 16;
  if (c.tid < PH16) reinterpret_cast<uint4*>(&phbuf[0])[c.tid] = reinterpret_cast<const uint4*>(&phases[0])[c.tid];
  c.bb_pos = (int)P->pos[P->S - 1];  // batch 1: stream 0, last prompt row
  c.bb_slot = P->cache_len + P->S - 1;
  __syncthreads();

  // start this warp's weight stream: SLOTS chunks in flight from now on
  c.pf.gi = 0; c.pf.u = c.warp; c.pf.issued = 0; c.pf.done = false;
  c.pf.policy = policy_evict_first();
  pf_seek(c.pf, tab, blockIdx.x, gridDim.x, c.warp);
  for (int s = 0; s < SLOTS; ++s) pf_issue(c.pf, tab, ring, full, blockIdx.x, gridDim.x, c.warp, c.lane);
  c.pf2 = c.pf;  // continues where the ring cursor stands, then stays L2_AHEAD chunks in front
  for (int s = 0; s < L2_AHEAD; ++s) pf_issue_l2(c.pf2, tab, blockIdx.x, gridDim.x, c.warp, c.lane);

  const unsigned ncta = gridDim.x;
  const bool tr = trace != nullptr && blockIdx.x == 0 && c.tid == 0;
  for (int p = 0; p < nphases; ++p) {
    const Phase& ph = phbuf[p & 1];
    if (tr) {
      trace[p * 8 + 0] = gtimer();
      c.trp = trace + p * 8 + 4;
    }
    // stage the next descriptor while this phase runs (read after the barrier below)
    if (p + 1 < nphases && c.tid < PH16)
      reinterpret_cast<uint4*>(&phbuf[(p + 1) & 1])[c.tid] = reinterpret_cast<const uint4*>(&phases[p + 1])[c.tid];
    switch (ph.type) {
      case PH_GEMV: gemv_phase(ph, c, tab); break;
      case PH_EMBED: embed_phase(ph, c); break;
      case PH_ATTN: attn_phase(ph, c); break;
      default: sample_phase(ph, c); break;
    }
    if (tr) trace[p * 8 + 1] = gtimer();
    if (p + 1 == nphases) break;
    // grid barrier: every CTA's writes of phase p are visible before anyone starts phase p+1
    csync<NCT, CBAR>();
    if (tr) trace[p * 8 + 2] = gtimer();
    if (c.tid == 0) arrive_release(&sync->counter);
    {
      // While the barrier resolves: fetch what phase p+1 needs that is already final -- the norm
      // scale and the epilogue operands (residual values from two phases ago, RoPE cos/sin) of this
      // warp's first row group.  Their L2 round trip then overlaps the barrier instead of following it.
      const Phase& nx = phbuf[(p + 1) & 1];
      c.pre_valid = false;
      if (nx.type == PH_GEMV && nx.attn_prologue) attn_prefetch(nx, c);
      if (nx.type == PH_GEMV) {
        c.pre_valid = true;
        if (nx.norm && c.tid < nx.K / 8) c.pre_scale = *reinterpret_cast<const uint4*>(nx.norm_scale + c.tid * 8);
        const int g = group_of(blockIdx.x, gridDim.x, c.warp, 0);
        const int hp = nx.R / 2;
        c.pre_a = c.pre_b = 0.f;
        if (g < nx.G && c.lane < hp * nx.nb) {
          const EpiPre e = epilogue_prefetch(nx, c, g * nx.R + 2 * (c.lane % hp), c.lane / hp);
          c.pre_a = e.a;
          c.pre_b = e.b;
        }
      }
    }
    if (c.tid == 0) {
      const unsigned target = ncta * (unsigned)(p + 1);
      for (unsigned spin = 0; ld_acquire(&sync->counter) < target; ++spin)
        if (spin > (1u << 24)) die(sync, 0x300);
    }
    csync<NCT, CBAR>();
    if (tr) trace[p * 8 + 3] = gtimer();
  }
}

}  // namespace mega
