// Persistent decode megakernel: one audio frame (backbone step + codebook-0 sample + 31 depth-
// decoder steps, ~610 dependent phases) in ONE launch, for one stream.
//
// Why: at batch 1 a frame is ~610 tiny dependent GEMVs (0.3-10 us of HBM time each); launched one by
// one, or separated by grid barriers, they are latency bound.  Here one CTA per SM stays resident for
// the whole frame and three things keep the dependency chain short:
//   * WEIGHT STREAM.  Every consumer warp owns a shared-memory ring of 4 x 4 KB slots; a ninth,
//     producer warp walks the frame's schedule (a table in kernel-parameter space) and issues every
//     warp's next chunk with a bulk async copy (cp.async.bulk + mbarrier complete_tx) as soon as the
//     slot it goes to has been drained, pulling chunks further ahead into L2 while it waits.  Weight
//     addresses are data independent, so the stream runs AHEAD of the dependency chain: 148 SMs x
//     128 KB of weights are in flight or landed while activations resolve.
//   * FRAGMENT-MAJOR WEIGHTS + TENSOR-CORE DOT.  The matrices are re-packed once at setup into
//     [row group of R=8/16][32-wide k block][lane][16 B] order: a chunk is one contiguous bulk copy,
//     every lane reads a complete mma.m16n8k16 operand with one conflict-free LDS.128 (16-row groups:
//     the weights are the A operand, pre-packed (a0..a3) quads; 8-row groups: the B operand), and the
//     dot products with the (<= 2) activation rows are HMMA instructions whose accumulator already
//     holds finished sums -- no unpack/FMA chains, no register shuffles, no shuffle reductions.  All 8
//     warps of a CTA split the K extent of a row group, partial sums meet in shared memory in a fixed
//     order.
//   * TAGGED ACTIVATIONS INSTEAD OF GRID BARRIERS.  Every vector one phase hands to the next travels
//     as 32-bit words {tag16 | bf16}; the tag names the producing phase and frame.  A consumer
//     re-loads a word until its tag is the expected one: the data is its own ready flag, a phase
//     boundary costs one L2 round trip, and CTAs without work in a phase run ahead.
// Token-dependent linear maps are tables: projection(embedding) and the first decoder layer's RoPE'd
// [q;k;v] of it are gathered by the sampling phase instead of being computed by GEMV phases.
// Every spin has a trip-count cap; a wait that gives up records an error code and lets the launch drain
// (no __trap: the CUDA context must survive, see report_error).
#pragma once
#include "lm_kernels.cuh"

namespace mega {

constexpr int NW = 8;                 // warps per CTA
constexpr int NCT = NW * 32;          // consumer threads per CTA
#ifndef MEGA_KV_FENCE
#define MEGA_KV_FENCE 32  /* bit 5: release / acquire once per codebook step, gpu-scope half on the producer warp (kv_producer_poll); inline variants: bit 3: release / acquire once per codebook step (kv_step_sync), bit 4: as st.release / ld.acquire instead of fence + relaxed access; measurement variants: 0 = none, bit 0: fence in every QKV epilogue, bit 1: in every attn_prefetch, bit 2: in the sampling phase's gather */
#endif
constexpr bool KV_PROD = ((MEGA_KV_FENCE) & 32) != 0;  // bit 5: the release / acquire runs on the producer warp (kv_producer_poll)
constexpr int NTHREADS = NCT + 32;    // + one producer warp that feeds the weight rings
#ifndef MEGA_SLOTS
#define MEGA_SLOTS 4
#endif
#ifndef MEGA_SLOT_BYTES
#define MEGA_SLOT_BYTES 4096
#endif
constexpr int SLOTS = MEGA_SLOTS;            // ring slots per warp
constexpr int SLOT_BYTES = MEGA_SLOT_BYTES;  // bytes per slot (a chunk never exceeds it)
#ifndef MEGA_REP
#define MEGA_REP 1  /* measured: 1, 4 and 8 copies hand off equally fast; the mechanism stays */
#endif
constexpr int REP = MEGA_REP;         // copies of every broadcast activation vector (spreads the readers over L2 slices)
constexpr int MAXNB = 2;              // activation rows per phase (depth step 1 carries 2)
constexpr int XBUF_ELEMS = 25600;  // 50 KB: activation rows (<= 2 x 8192), or the fused attention's q / K / V staging (A_IOFF + 64)
constexpr int CBAR = 1;               // consumer warps synchronise on named barrier 1 (the producer warp never joins)
constexpr int MAX_GEMV = 640;         // rows of the prefetch table (kernel parameter space)
constexpr int MAX_LOCAL_GROUPS = 8;   // row groups one CTA owns in one phase
constexpr size_t SMEM_RING = (size_t)NW * SLOTS * SLOT_BYTES;
constexpr size_t SMEM_X = (size_t)XBUF_ELEMS * 2;
constexpr size_t SMEM_PSUM = (size_t)MAX_LOCAL_GROUPS * NW * 16 * MAXNB * 4;
constexpr size_t SMEM_MISC = 2048;
constexpr size_t SMEM_BYTES = SMEM_RING + SMEM_X + SMEM_PSUM + SMEM_MISC;

enum { PH_GEMV = 0, PH_EMBED = 1, PH_ATTN = 2, PH_SAMPLE = 3 };
enum { POS_FIXED = 0, POS_BACKBONE = 1 };

// Full description of a phase (global memory; staged into shared memory one phase ahead).
// t_* pointers address TAGGED words; *_src = index of the phase that produced the words consumed.
struct __align__(16) Phase {
  int type, epi, norm, nb;
  // GEMV: W is fragment-major packed, G groups of R rows, group g belongs to CTA (g + rot) % ncta
  const bf16* W;
  int rows, K, R, G;
  int rot, gq, gr;  // gq = G / ncta, gr = G % ncta
  int tb, chunk, nch, kchunk, nblk;  // bytes of a warp's slice of a group, bytes / count / k extent / blocks of its chunks
  int hd_shift;  // log2(hd)
  int grp_shift, heads_shift;  // log2(heads / kv_heads), log2(heads)
  float inv_K;   // 1/K when K is a power of two (exact), else 0
  int ldx;
  const uint32_t* t_x;  // input rows [nb][ldx]
  int x_src[2];
  const bf16* norm_scale;
  float eps;
  int ldo;
  uint32_t* t_out;   // output rows [nb][ldo]; EPI_RESID: the residual stream, updated in place
  uint32_t* t_out2;  // EPI_PLAIN with two destinations: rows >= split_row go to t_out2[row - split_row]
  int split_row;
  int resid_src[2];
  int attn_prologue;  // GEMV: x = attention(q, cache) computed redundantly by every CTA (<= 32 keys)
  // RoPE / KV append / attention
  uint32_t* t_q;   // [nb][heads*hd]
  uint32_t* t_kv;  // K and V rows of the positions written by the QKV phase: [nb][2][kv_heads*hd]
  int q_src;
  int heads, kv_heads, hd, slots, pos_mode, pos0;
  bf16 *kc, *vc;  // plain bf16 cache (later steps / frames)
  const bf16* rope;
  // sample / embed
  int V, C, D, cb;
  const uint32_t* t_logits;
  int logits_src;
  int next_ld;             // SAMPLE: row length of the gather table feeding t_next
  uint32_t* t_next;        // SAMPLE: next depth-decoder input row
  const bf16* next_table;  // SAMPLE: projection(embedding) table [(codebooks-1)*V][next_ld]
  // SAMPLE: [q;k;v] of the first decoder layer for the sampled token, RoPE applied at position cb + 1
  // (table [(codebooks-1)*V][(heads + 2 kv_heads) * hd]); the phase then also hands q / k / v to the next step
  // (t_q, t_kv, kc / vc at slot pos0) and that step has no QKV phase in its first layer.  Null: no table.
  const bf16* qkv_table;
  const bf16 *audio_emb, *text_emb;
  // distance in words between the REP copies of the tagged vectors (0: single copy)
  int x_rs, out_rs, out2_rs, q_rs, kv_rs, next_rs;
  int keep;  // this matrix is loaded with the L2 evict-last policy
  int TV;    // EMBED: text vocabulary (range check of the text column)
  int rope_len;  // EMBED: rows of the backbone RoPE table (range check of input_pos)
  // ordering of the depth decoder's plain KV rows, once per codebook step (MEGA_KV_FENCE bit 3, see kv_fence):
  // before this phase every CTA ... kv_sync & 255 = 1: releases its rows (done word t_done[cta] tagged with phase
  // done_src); 2: acquires (polls all CTAs' done words for the tag of phase done_src); 3: (producer-warp variant only)
  // waits until the acquire for done_src is acknowledged.  kv_sync >> 8 = done_src of the previous request (-1: none)
  int kv_sync, done_src;
  uint32_t* t_done;
};
static_assert(sizeof(Phase) % 16 == 0, "Phase must be copyable in 16-byte units");

// What the weight prefetcher needs to know about a GEMV phase.
struct PfDesc {
  const bf16* W;
  int G, rot;
  int group_bytes;  // R * K * 2
  int chunk_nch;    // chunk bytes | chunks per warp slice << 16 | (keep in L2: evict-last) << 30
};
static_assert(sizeof(PfDesc) == 24, "PfDesc layout");
struct PfTable {
  int n;
  int pad_;
  uint32_t* t_done;  // per-CTA done words of the KV ordering (kv_producer_poll)
  PfDesc d[MAX_GEMV];
};

using Sync = DevStatus;  // frame counter (tag salt, bumped by k_mega_prepare) + sticky error word

// Runs before the megakernel on the same stream: publishes the call's parameters, bumps the frame counter and
// checks the frame's inputs (token ids of the last prompt row against their tables, input_pos against the cache
// position and the RoPE table) -- here, not in the megakernel: a returning call in its prologue cost 2 % of the
// frame (round 2 bisect, profiles/r2_regression_bisect.txt).  The megakernel itself only clamps.
__global__ void k_mega_prepare(FrameParams* dst, FrameParams v, Sync* sync, int C, int V, int TV, int rope_len) {
  if (threadIdx.x == 0) {
    *dst = v;
    sync->seq += 1;
    sync->error = 0;  // per call; the host mirror stays sticky until csm_check_error reads it
  }
  __syncthreads();
  const size_t fr = (size_t)(v.S - 1) * (C + 1);  // batch 1: stream 0, last prompt row
  if ((int)threadIdx.x <= C && v.mask[fr + threadIdx.x]) checked_token(v.tokens + fr, threadIdx.x, C, V, TV, sync);
  if (threadIdx.x == 0) {
    const int64_t pos = v.pos[v.S - 1];
    const int slot = (v.lane_meta ? v.lane_meta[v.B] : v.cache_len) + v.S - 1;
    if (pos != slot || pos < 0 || pos >= rope_len) report_error(sync, 0x803);
  }
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
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking probe
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
// Weights are streamed with an L2 evict-first policy: they are read once per use and must not push
// the small hot data (tagged activations, KV cache, norm scales, RoPE tables) out of the 126 MB L2.
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
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float ldcg_bf(const bf16* p) {
  return bf2f(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// A wait that exceeds its trip cap must not hang the GPU -- and must not __trap() either: a trap poisons
// the CUDA context, and the reference's retry loops (tts_service.py:500-514) could never recover.  Instead the
// first wait to exceed its cap records a sticky error code (device word + mapped host word), and every wait
// looks at that word once per 1024 trips: from then on each wait of every thread returns within a millisecond,
// the launch drains to its end with garbage tokens (clamped before they index anything), the context
// survives, and the host raises from the code.  Everything is inline: a call that RETURNS inside or right
// after a poll loop makes the compiler keep the loop's live state across it (measured: +5 % and +14 % frame
// time for the two call-based variants tried in round 2; round 1's trap was free because it never returned).
constexpr unsigned SPIN_CAP = 1u << 22;  // ~ seconds of L2 round trips
__device__ __forceinline__ bool give_up(Sync* sync, unsigned spin, unsigned code, unsigned cap = SPIN_CAP) {
#ifdef MEGA_DIAG_TRAP  /* diagnosis only: round 1's behaviour, to price the soft abort */
  if (spin > cap) __trap();
  return false;
#endif
  if ((spin & 1023u) != 1023u) return false;
  if (*reinterpret_cast<volatile unsigned int*>(&sync->error) != 0u) return true;  // somebody gave up: drain
  if (spin < cap) return false;
  if (atomicCAS(&sync->error, 0u, code) == 0u && sync->host_error)
    *reinterpret_cast<volatile unsigned int*>(sync->host_error) = code;  // lands by the end of the launch at the latest
  return true;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, Sync* sync, unsigned code) {
  for (unsigned spin = 0; !mbar_try(bar, parity); ++spin)
    if (give_up(sync, spin, code)) break;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// D += A(16x16, row) * B(16x8, col), bf16 operands, fp32 accumulate (HMMA.16816.F32.BF16)
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- ordering of the plain KV-cache rows --------------------------------------------------------------
// The depth decoder's cache rows are written as plain bf16 (a QKV epilogue, or the sampling phase's table
// gather) and read by OTHER CTAs in later codebook steps of the SAME launch (attn_prefetch, cp.async.cg).  The
// tagged hand-off words do not order those plain stores.  The default build orders them once per codebook step: a
// release by every CTA after the last layer's gate/up phase, an acquire by every CTA before the sampling phase, the
// gpu-scope instructions issued by the producer warp (kv_producer_poll / kv_post below; +2.3 % of a frame).  Measured
// alternatives (profiles/r2_kv_fence_variants.txt), kept as compile-time variants: bits 3-4 run the same protocol
// inline in the consumer warps (kv_step_sync; +3.1 %: its code spills at the 168-register cap); bits 0-2 put fence.acq_rel.gpu into every QKV epilogue / attn_prefetch / sampling-phase
// gather (+0.105 / +0.156 / +0.039 ms per frame, together +7.4 %); -DMEGA_KV_FENCE=0 orders nothing (the hardware
// does deliver the rows -- a store reaches L2, the point of coherence, in < 1 us and the readers come >= 16 phases
// later -- but nothing in the PTX memory model says so); a tagged-word cache was built and measured too (+17 %: its
// reads cannot be asynchronous).  (The backbone cache is read by the NEXT launch.)
template <int WHO>
__device__ __forceinline__ void kv_fence() {
  if ((MEGA_KV_FENCE) & WHO) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ uint4 lda4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.acquire.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_acquire_cta(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release_cta(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- tagged activations ------------------------------------------------------------------------------
// Buffer reuse is safe without further synchronisation: seeing ANY word of phase p implies that its
// producer had staged the complete output of phase p-1, hence that every phase <= p-1 is complete;
// a buffer written in phase p is never rewritten before phase p+2.
__device__ __forceinline__ uint32_t tag_of(unsigned seq, int phase) {
  return ((uint32_t)(((seq & 31u) << 11) | (unsigned)(phase + 1))) << 16;
}
__device__ __forceinline__ uint32_t tword(uint32_t tag, float v) { return tag | (uint32_t)__bfloat16_as_ushort(f2bf(v)); }
__device__ __forceinline__ uint32_t tword_raw(uint32_t tag, uint32_t bf) { return tag | (bf & 0xffffu); }
__device__ __forceinline__ uint4 ldv4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ldv2(const uint32_t* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ldv1(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Tagged words are written with st.relaxed.gpu and polled with ld.volatile (= relaxed, system scope): both sides are
// strong operations in the PTX memory model, so the hand-off itself is race-free word by word (every word carries its
// own tag).  Same SASS as a .cg store (STG.E.STRONG.GPU): the qualifier costs nothing.
__device__ __forceinline__ void stv4(uint32_t* p, const uint4& v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stv2(uint32_t* p, const uint2& v) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ bool fresh4(const uint4& v, uint32_t tag) {
  return ((((v.x ^ tag) | (v.y ^ tag)) | ((v.z ^ tag) | (v.w ^ tag))) & 0xffff0000u) == 0;
}
__device__ __forceinline__ uint2 poll2(const uint32_t* p, uint32_t tag, Sync* sync) {
  uint2 v = ldv2(p);
  for (unsigned spin = 0; (((v.x ^ tag) | (v.y ^ tag)) & 0xffff0000u) != 0; ++spin) {
    if (give_up(sync, spin, 0x401)) break;
    v = ldv2(p);
  }
  return v;
}
__device__ __forceinline__ uint32_t poll1(const uint32_t* p, uint32_t tag, Sync* sync) {
  uint32_t v = ldv1(p);
  for (unsigned spin = 0; ((v ^ tag) & 0xffff0000u) != 0; ++spin) {
    if (give_up(sync, spin, 0x402)) break;
    v = ldv1(p);
  }
  return v;
}
// four tagged words -> four packed bf16 (the low halves)
__device__ __forceinline__ uint2 strip4(const uint4& v) {
  return make_uint2((v.x & 0xffffu) | (v.y << 16), (v.z & 0xffffu) | (v.w << 16));
}
__device__ __forceinline__ float tval(uint32_t w) { return __uint_as_float(w << 16); }
// store to every copy of a broadcast vector (rs = distance between copies in words, 0 = single copy)
__device__ __forceinline__ void rep_st1(uint32_t* p, int rs, uint32_t v) {
#pragma unroll
  for (int r = 0; r < REP; ++r)
    if (r == 0 || rs) st_relaxed_gpu(p + (size_t)r * rs, v);
}
__device__ __forceinline__ void rep_st2(uint32_t* p, int rs, const uint2& v) {
#pragma unroll
  for (int r = 0; r < REP; ++r)
    if (r == 0 || rs) stv2(p + (size_t)r * rs, v);
}
__device__ __forceinline__ void rep_st4(uint32_t* p, int rs, const uint4& v) {
#pragma unroll
  for (int r = 0; r < REP; ++r)
    if (r == 0 || rs) stv4(p + (size_t)r * rs, v);
}
// the copy this CTA reads
__device__ __forceinline__ const uint32_t* my_copy(const uint32_t* p, int rs) { return p + (size_t)(blockIdx.x % REP) * rs; }

// ---- weight stream: one producer warp per CTA -------------------------------------------------------
// A phase's row group g belongs to CTA (g + rot) % ncta; all NW consumer warps of that CTA split its
// K extent: warp w owns the contiguous slice [w/NW, (w+1)/NW) of the group's fragment-major bytes, cut
// into chunks of <= SLOT_BYTES.  Every consumer warp has a private ring of SLOTS slots with a full
// (complete_tx) and an empty (consumer lane 0 arrives) mbarrier per slot.  The producer warp walks
// the frame's schedule -- phases in order, the CTA's groups in order, chunks in order, exactly the
// order in which gemv_groups consumes -- and lane w issues warp w's chunk as soon as the slot it goes
// to has been drained.  Consumers never compute addresses or touch the schedule: their inner loop is
// wait(full) -> LDS + HMMA -> arrive(empty).
__device__ __forceinline__ int local_cta(int cta, int ncta, int rot) {
  const int cl = cta - rot;
  return cl < 0 ? cl + ncta : cl;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifndef MEGA_L2_AHEAD
#define MEGA_L2_AHEAD 6  /* measured: 3.16 (off) -> 3.09 ms/frame with 6 steps, 3.12 with 12 */
#endif
constexpr int L2_AHEAD = MEGA_L2_AHEAD;  // ring steps (NW chunks each) that the L2 prefetch cursor runs ahead of the ring cursor

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// position in the frame's schedule: (phase, group of this CTA, chunk)
struct Cursor {
  int gi, g, ch;
};
__device__ __forceinline__ bool cursor_valid(const Cursor& k, int ntab) { return k.gi < ntab; }
// first position at or after (gi, ., .) where this CTA owns a group
__device__ __forceinline__ void cursor_seek(Cursor& k, const PfDesc* tab, int ntab, int cta, int ncta) {
  while (k.gi < ntab) {
    if (k.g < 0) k.g = local_cta(cta, ncta, tab[k.gi].rot);
    if (k.g < tab[k.gi].G) return;
    ++k.gi;
    k.g = -1;
    k.ch = 0;
  }
}
__device__ __forceinline__ void cursor_next(Cursor& k, const PfDesc* tab, int ntab, int cta, int ncta) {
  if (++k.ch >= ((tab[k.gi].chunk_nch >> 16) & 0xff)) {
    k.ch = 0;
    k.g += ncta;
    cursor_seek(k, tab, ntab, cta, ncta);
  }
}
__device__ __forceinline__ const unsigned char* cursor_src(const Cursor& k, const PfDesc* tab, int lane, int& chunk, bool& keep) {
  const PfDesc d = tab[k.gi];
  chunk = d.chunk_nch & 0xffff;
  keep = (d.chunk_nch >> 30) & 1;
  const int tb = chunk * ((d.chunk_nch >> 16) & 0xff);
  return reinterpret_cast<const unsigned char*>(d.W) + (size_t)k.g * d.group_bytes + (size_t)lane * tb + (size_t)k.ch * chunk;
}

// KV ordering on the producer warp (MEGA_KV_FENCE bit 5): consumer thread NCT-1 posts a request word
// {tag of done_src | type} in shared memory (st.release.cta after the phase's end barrier); the producer warp looks
// at the word once per ring step and in its wait loop, and runs the gpu-scope half there -- type 1: st.release.gpu
// done[cta] = tag (cumulative over the consumers' cache stores through barrier -> release.cta -> acquire.cta);
// type 2: ld.acquire.gpu of every CTA's done word until it carries the tag -- then acknowledges the request in a second
// shared word (st.release.cta), which the consumers read (ld.acquire.cta) before the next step's attn_prefetch.
// A stall of this warp is absorbed by the weight rings (a warp's four slots hold its whole slice of a down phase).
#ifndef MEGA_KV_SERVE_NOINLINE
#define MEGA_KV_SERVE_NOINLINE 0  /* 1: one out-of-line copy of the request handler for the producer loop's three call sites */
#endif
#if MEGA_KV_SERVE_NOINLINE
#define KV_SERVE_ATTR __noinline__
#else
#define KV_SERVE_ATTR __forceinline__
#endif
__device__ KV_SERVE_ATTR void kv_producer_serve(uint32_t r, uint32_t* kv, uint32_t* done, Sync* sync, int lane);
__device__ __forceinline__ void kv_producer_poll(uint32_t& last, uint32_t* kv, uint32_t* done, Sync* sync, int lane) {
  const uint32_t r = lds_acquire_cta(&kv[0]);  // same word in every lane: warp-uniform
  if (r == last) return;
  last = r;
  if ((r & 3u) == 3u) return;  // the consumers have run their last phase: nothing to do, nothing to acknowledge
  kv_producer_serve(r, kv, done, sync, lane);
}
__device__ KV_SERVE_ATTR void kv_producer_serve(uint32_t r, uint32_t* kv, uint32_t* done, Sync* sync, int lane) {
  const uint32_t dtag = r & 0xffff0000u;
  if ((r & 3u) == 1u) {
    if (lane == 0) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(done + blockIdx.x), "r"(dtag) : "memory");
      if (blockIdx.x == gridDim.x - 1)  // pad to whole 16-byte units for the pollers
        for (unsigned q = gridDim.x; (q & 3u) != 0; ++q) st_relaxed_gpu(done + q, dtag);
    }
  } else {
    for (unsigned q = lane * 4; q < gridDim.x; q += 128) {
      uint4 v = lda4(done + q);
      for (unsigned spin = 0; !fresh4(v, dtag); ++spin) {
        if (give_up(sync, spin, 0x40b)) break;
        v = lda4(done + q);
      }
    }
  }
  __syncwarp();
  if (lane == 0) sts_release_cta(&kv[1], r);
}
__device__ __forceinline__ void producer_loop(const PfDesc* tab, int ntab, unsigned char* ring, uint64_t* full, uint64_t* empty,
                                              Sync* sync, int lane, uint32_t* kv, uint32_t* done) {
  const int cta = blockIdx.x, ncta = gridDim.x;
  // Weights that are used once per frame stream through L2 evict-first.  The depth decoder's 222 MB are
  // used 31 times per frame: the matrices marked "keep" are loaded evict-last, so that part of them
  // survives in the 126 MB L2 from one codebook step to the next and never touches HBM again.
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
  Cursor k, k2;
  k.gi = 0; k.g = -1; k.ch = 0;
  cursor_seek(k, tab, ntab, cta, ncta);
  k2 = k;
  unsigned ahead = 0;  // ring steps by which the L2 cursor k2 leads the ring cursor k
  uint32_t kv_last = 0;
  for (unsigned issued = 0; cursor_valid(k, ntab); ++issued) {
    if (KV_PROD) kv_producer_poll(kv_last, kv, done, sync, lane);
    const int slot = issued % SLOTS;
    uint64_t* eb = &empty[(lane < NW ? lane : 0) * SLOTS + slot];
    const uint32_t parity = ((issued / SLOTS) & 1) ^ 1;
    // Wait until all eight slots of this step are drained.  While waiting -- the CTA sits in a latency
    // chain and HBM would idle -- run the second, deeper stage of the weight stream: pull chunks up to
    // L2_AHEAD steps beyond the ring into the 126 MB L2, so that the rings later refill at L2 speed.
    for (unsigned spin = 0;; ++spin) {
      // (with the L2 stage on, probe without suspending so that the prefetches really go out while waiting)
      const bool ok = lane >= NW || (L2_AHEAD > 0 ? mbar_test(eb, parity) : mbar_try(eb, parity));
      if (__all_sync(0xffffffffu, ok)) break;
      if (__any_sync(0xffffffffu, give_up(sync, spin, 0x100, 1u << 26))) break;  // warp-uniform exit
      if (KV_PROD) kv_producer_poll(kv_last, kv, done, sync, lane);
      if (L2_AHEAD > 0 && ahead < SLOTS + L2_AHEAD && cursor_valid(k2, ntab)) {
        if (ahead >= SLOTS && lane < NW) {  // the first SLOTS steps ahead are in the ring (or on their way) already
          int chunk;
          bool keep;
          const unsigned char* src = cursor_src(k2, tab, lane, chunk, keep);
          bulk_prefetch_l2(src, (uint32_t)chunk, keep ? pol_last : pol_first);
        }
        cursor_next(k2, tab, ntab, cta, ncta);
        ++ahead;
      }
    }
    if (lane < NW) {
      int chunk;
      bool keep;
      const unsigned char* src = cursor_src(k, tab, lane, chunk, keep);
      uint64_t* fb = &full[lane * SLOTS + slot];
      mbar_expect_tx(fb, (uint32_t)chunk);
      bulk_g2s(ring + (size_t)(lane * SLOTS + slot) * SLOT_BYTES, src, (uint32_t)chunk, fb, keep ? pol_last : pol_first);
    }
    cursor_next(k, tab, ntab, cta, ncta);
    if (ahead > 0) --ahead;
    else k2 = k;
  }
  // a CTA whose last chunk goes out early (no groups in the frame's last phases) still has requests to serve
  if (KV_PROD)
    for (unsigned spin = 0; (kv_last & 3u) != 3u; ++spin) {
      kv_producer_poll(kv_last, kv, done, sync, lane);
      if (__any_sync(0xffffffffu, give_up(sync, spin, 0x40f, 1u << 26))) break;
    }
}

#define CK(i) do { if (c.trp) c.trp[i] = clock64(); } while (0)
// ---- consumer pieces ---------------------------------------------------------------------------------
struct Ctx {
  const FrameParams* P;
  unsigned char* ring;
  bf16* xs;
  uint64_t *full, *empty;
  float* scratch;  // 34 floats
  float* psum;     // [MAX_LOCAL_GROUPS][NW][16][MAXNB] split-K partial sums
  int* iscratch;   // 40 ints
  Sync* sync;
  unsigned cnt;  // chunks consumed by this warp so far
  unsigned long long* trp;  // trace slots of the current phase (CTA 0, thread 0) or null
  int tid, warp, lane;
  int bb_pos, bb_slot;  // RoPE position / cache slot of the backbone row of this frame
  int bb_lane;          // cache lane of the stream (continuous batching; 0 for the plain batch-1 use)
  unsigned seq;         // frame counter (tag salt)
  uint32_t tag;         // tag of the words the current phase produces
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
// They are requested at the START of the phase, so their L2 round trip overlaps the hand-off.
struct EpiPre {
  uint32_t a, b;  // raw words: converting here would wait for the load at the start of the phase
};
__device__ __forceinline__ EpiPre epilogue_prefetch(const Phase& ph, const Ctx& c, int r0, int n) {
  EpiPre e;
  e.a = e.b = 0u;
  if (ph.epi == EPI_RESID) {
    // written at least two phases ago: one load, verified (and re-polled in the unlikely stale case) in the epilogue
    const uint2 w = ldv2(my_copy(ph.t_out, ph.out_rs) + (size_t)n * ph.ldo + r0);
    e.a = w.x;
    e.b = w.y;
  } else if (ph.epi == EPI_ROPE_KV) {
    const int hd = ph.hd;
    if (r0 < (ph.heads + ph.kv_heads) * hd) {
      int pos, slot;
      phase_pos(ph, c, n, pos, slot);
      e.a = *reinterpret_cast<const uint32_t*>(ph.rope + ((size_t)pos * (hd / 2) + ((r0 & (hd - 1)) >> 1)) * 2);  // {cos, sin}
    }
  }
  return e;
}

// fused epilogue of one output-row pair (row r0, r0+1) for activation row n
__device__ __forceinline__ void epilogue(const Phase& ph, const Ctx& c, int r0, int n, float a0, float a1, const EpiPre& pre) {
  const float y0 = rbf(a0), y1 = rbf(a1);
  const uint32_t tag = c.tag;
  float pa, pb;
  if (ph.epi == EPI_RESID) {
    uint2 w = make_uint2(pre.a, pre.b);
    const uint32_t rtag = tag_of(c.seq, ph.resid_src[n]);
    if ((((w.x ^ rtag) | (w.y ^ rtag)) & 0xffff0000u) != 0)
      w = poll2(my_copy(ph.t_out, ph.out_rs) + (size_t)n * ph.ldo + r0, rtag, c.sync);
    pa = tval(w.x);
    pb = tval(w.y);
  } else {
    pa = bflo(pre.a);  // cos
    pb = bfhi(pre.a);  // sin
  }
  if (ph.epi == EPI_PLAIN) {
    const bool second = ph.t_out2 && r0 >= ph.split_row;
    uint32_t* dst = second ? ph.t_out2 + (r0 - ph.split_row) : ph.t_out + (size_t)n * ph.ldo + r0;
    rep_st2(dst, second ? ph.out2_rs : ph.out_rs, make_uint2(tword(tag, y0), tword(tag, y1)));
  } else if (ph.epi == EPI_RESID) {
    rep_st2(ph.t_out + (size_t)n * ph.ldo + r0, ph.out_rs, make_uint2(tword(tag, y0 + pa), tword(tag, y1 + pb)));
  } else if (ph.epi == EPI_SWIGLU) {
    rep_st1(ph.t_out + (size_t)n * ph.ldo + (r0 >> 1), ph.out_rs, tword(tag, silu_bf(y0) * y1));
  } else {  // EPI_ROPE_KV
    const int hd = ph.hd, qrows = ph.heads * hd, krows = ph.kv_heads * hd;
    int pos, slot;
    phase_pos(ph, c, n, pos, slot);
    float o0 = y0, o1 = y1;
    if (r0 < qrows + krows) {
      o0 = rbf(__fsub_rn(__fmul_rn(y0, pa), __fmul_rn(y1, pb)));
      o1 = rbf(__fadd_rn(__fmul_rn(y1, pa), __fmul_rn(y0, pb)));
    }
    if (r0 < qrows) {
      rep_st2(ph.t_q + (size_t)n * qrows + r0, ph.q_rs, make_uint2(tword(tag, o0), tword(tag, o1)));
    } else {
      const bool isk = r0 < qrows + krows;
      const int rr = r0 - (isk ? qrows : qrows + krows);
      const int kvh = rr >> ph.hd_shift, d = rr & (hd - 1);
      // this step's consumers read the tagged copy; the cache keeps plain bf16 for later steps / frames
      rep_st2(ph.t_kv + ((size_t)n * 2 + (isk ? 0 : 1)) * krows + rr, ph.kv_rs, make_uint2(tword(tag, o0), tword(tag, o1)));
      const int lane = ph.pos_mode == POS_BACKBONE ? c.bb_lane : 0;  // the depth decoder's cache is per-frame scratch
      bf16* dst = (isk ? ph.kc : ph.vc) + (((size_t)lane * ph.kv_heads + kvh) * ph.slots + slot) * hd + d;
      *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
      if (ph.pos_mode == POS_FIXED) kv_fence<1>();  // depth decoder: read by other CTAs later in this launch
    }
  }
}

// The CTA's row groups of this phase.  Warp w streams slice w of every group through its ring: a
// chunk is [nblk] fragment-major blocks of R rows x 32 k, each made of R/8 sub-blocks [lane][16 B] =
// W[8-row sub-block row lane/4][k = (lane%4)*8 .. +8].
// Tensor-core mapping (mma.m16n8k16, D[m][n] = sum_k A[m][k] B[k][n]): B = weights (n = the 8 rows of
// a sub-block), A = activations (m = activation row).  One 16-byte load per lane is a complete
// operand, no register shuffling:
//   * B: the lane's 8 consecutive k are the pairs P0..P3; HMMA "lo" takes (b0,b1) = (P0,P1), HMMA "hi"
//     takes (P2,P3);
//   * A: x is staged in shared memory with the pairs of every 8-group in the order (P0,P2,P1,P3), so
//     the loaded quad (a0,a1,a2,a3) = (xP0,xP2,xP1,xP3) serves BOTH: in "lo" the rows 0-7 operands
//     (a0,a2) = (xP0,xP1) meet (P0,P1) and accumulator rows 0-7 are right; in "hi" the rows 8-15
//     operands (a1,a3) = (xP2,xP3) meet (P2,P3) and accumulator rows 8-15 are right.  The other halves
//     are garbage and never read; lanes of unused activation rows load row 0 and only feed accumulator
//     columns nobody reads.  y[act row g][sub-block row 2q+e] = lo[e] + hi[2+e] in lane (g, q).
// The partial sums of the 8 warps meet in shared memory, and one thread per (group, row pair,
// activation row) adds them in a fixed order and runs the fused epilogue.
// ONE copy of this code serves every GEMV phase (R is a run-time value): the ~640 phases of a frame
// should stay inside the instruction cache, a miss is an L2 round trip in the dependency chain.
__device__ __forceinline__ void mma_x(float (&d)[4], const uint4& x, uint32_t b0, uint32_t b1) {
  mma16816(d, x.x, x.y, x.z, x.w, b0, b1);
}

__device__ __forceinline__ void gemv_groups(const Phase& ph, Ctx& c, int cl, int ngl, int j_item,
                                            int pair, int n_item, bool item_on, const EpiPre& pre) {
  const int ncta = gridDim.x;
  const int K = ph.K, R = ph.R;
  const bool r16 = R == 16;
  const int nch = ph.nch, kchunk = ph.kchunk, nblk = ph.nblk;
  const int g = c.lane >> 2, q = c.lane & 3;
  const bf16* xw = c.xs + c.warp * (K / NW) + (size_t)(g < ph.nb ? g : 0) * K + q * 8;
  const int ntask = ngl * nch;
  float acc[2][2][4];  // [block parity][lo / hi][fragment]
  int j = 0, ch = 0;
  // (measured: probing the NEXT chunk's barrier before the HMMAs of the current one, with try_wait or
  // test_wait, is slower: 3.06 -> 3.3 ms/frame)
  for (int t = 0; t < ntask; ++t) {
    if (ch == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][m][e] = 0.f;
    }
    const int slot = c.cnt % SLOTS;
    if (t == 0) CK(9);
    mbar_wait(&c.full[c.warp * SLOTS + slot], (c.cnt / SLOTS) & 1, c.sync, 0x200);
    if (t == 0) CK(10);
    {
      const unsigned char* wp = c.ring + (size_t)(c.warp * SLOTS + slot) * SLOT_BYTES + c.lane * 16;
      const bf16* xp = xw + ch * kchunk;
      const uint4 z = make_uint4(0, 0, 0, 0);
      if (r16) {
        // 16-row groups: weights are the A operand.  A block is [lo / hi half][lane][16 B]; the lane's
        // quad is (W[2g][Pa], W[2g+1][Pa], W[2g][Pb], W[2g+1][Pb]) with (Pa,Pb) = the pairs (P0,P1) of its
        // 8 consecutive k in the lo half and (P2,P3) in the hi half -- exactly (a0,a1,a2,a3).  x is the
        // B operand, staged in natural order: one 16-byte load = (b0,b1) of the lo and of the hi HMMA.
        // Four blocks per round: twelve loads go out first, then eight HMMA on four accumulators.
#pragma unroll 1
        for (int b = 0; b < nblk; b += 4) {
          uint4 wl[4], wh[4], xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool on = b + i < nblk;
            wl[i] = on ? *reinterpret_cast<const uint4*>(wp + (b + i) * 1024) : z;
            wh[i] = on ? *reinterpret_cast<const uint4*>(wp + (b + i) * 1024 + 512) : z;
            xv[i] = *reinterpret_cast<const uint4*>(xp + (on ? b + i : b) * 32);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mma16816(acc[i & 1][0], wl[i].x, wl[i].y, wl[i].z, wl[i].w, xv[i].x, xv[i].y);
            mma16816(acc[i & 1][1], wh[i].x, wh[i].y, wh[i].z, wh[i].w, xv[i].z, xv[i].w);
          }
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < nblk; b += 2) {
          const bool two = b + 1 < nblk;
          const uint4 x0 = *reinterpret_cast<const uint4*>(xp + b * 32);
          const uint4 wa = *reinterpret_cast<const uint4*>(wp + b * 512);
          const uint4 x1 = two ? *reinterpret_cast<const uint4*>(xp + b * 32 + 32) : x0;
          const uint4 wc = two ? *reinterpret_cast<const uint4*>(wp + b * 512 + 512) : z;
          mma_x(acc[0][0], x0, wa.x, wa.y);
          mma_x(acc[0][1], x0, wa.z, wa.w);
          mma_x(acc[1][0], x1, wc.x, wc.y);
          mma_x(acc[1][1], x1, wc.z, wc.w);
        }
      }
    }
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[c.warp * SLOTS + slot]);  // the slot is drained: the producer may refill it
    if (t == 0) CK(11);
    ++c.cnt;
    if (++ch == nch) {
      if (r16) {
        // accumulator of lane (g, 0): rows 2g (c0,c1) and 2g+1 (c2,c3) x activation rows 0,1
        if (q == 0) {
          float4 o;
          o.x = (acc[0][0][0] + acc[0][1][0]) + (acc[1][0][0] + acc[1][1][0]);
          o.y = (acc[0][0][1] + acc[0][1][1]) + (acc[1][0][1] + acc[1][1][1]);
          o.z = (acc[0][0][2] + acc[0][1][2]) + (acc[1][0][2] + acc[1][1][2]);
          o.w = (acc[0][0][3] + acc[0][1][3]) + (acc[1][0][3] + acc[1][1][3]);
          *reinterpret_cast<float4*>(c.psum + ((size_t)(j * NW + c.warp) * 16 + 2 * g) * MAXNB) = o;
        }
      } else if (g < MAXNB) {
        // lane (g, q): activation row g, rows 2q, 2q+1 of the 8-row group
        float* ps = c.psum + ((size_t)(j * NW + c.warp) * 16 + 2 * q) * MAXNB + g;
        {
          ps[0] = (acc[0][0][0] + acc[0][1][2]) + (acc[1][0][0] + acc[1][1][2]);
          ps[MAXNB] = (acc[0][0][1] + acc[0][1][3]) + (acc[1][0][1] + acc[1][1][3]);
        }
      }
      ch = 0;
      ++j;
    }
  }
  CK(12);
  csync<NCT, CBAR>();
  if (c.trp) c.trp[2] = gtimer();
  CK(13);
  if (item_on) {
    const float* ps = c.psum + ((size_t)(j_item * NW) * 16 + 2 * pair) * MAXNB + n_item;
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {  // fixed order -> deterministic rounding
      y0 += ps[(size_t)w * 16 * MAXNB];
      y1 += ps[(size_t)w * 16 * MAXNB + MAXNB];
    }
    epilogue(ph, c, (cl + ncta * j_item) * R + 2 * pair, n_item, y0, y1, pre);
  }
  CK(14);
}

// Attention over <= 32 cached keys for every (row, head), redundantly in every CTA, straight into
// the activation buffer of the output projection (depth decoder: 32 slots per stream, head_dim 128).
// Everything runs on the tensor cores out of shared memory; columns are (activation row, head within
// the kv group): nb * grp <= 8 = the N extent of mma.m16n8k16.
//   xs layout (bf16 elements):
//     [0, nb*dim)            output rows (the O projection's x)
//     A_QOFF  q    [kv][8 columns][160]      (rows padded to 160: conflict-free 16-byte fragment loads)
//     A_KOFF  K    [kv][32 keys][160]
//     A_VOFF  V    [kv][32 keys][136]        (rows padded to 136: conflict-free ldmatrix)
//     A_SOFF  S    [kv][8 columns][32] fp32 scaled scores
//     A_POFF  P    [kv][8 columns][40] bf16 exp(s - max), 0 beyond the visible keys
//     A_IOFF  1/sum [kv][8] fp32
// attn_prefetch runs right after the PREVIOUS phase ends: the K/V rows of earlier positions are final,
// so they are copied into shared memory with cp.async while this CTA waits for the q words.
// attn_small_into_x polls q and the current positions' K/V (tagged words of the QKV phase), then:
//   scores  warp (kv, 16-key tile): S = K q^T, 8 HMMA over the 128 dims
//   softmax warp per 2 (kv, column) rows, lane == key
//   P.V     warp per 4 (kv, 8-dim tile): V fragments by ldmatrix.trans, <= 2 HMMA each, scaled by 1/sum.
constexpr int A_HD = 128, A_QS = 160, A_KS = 160, A_VS = 136, A_PS = 40;
constexpr int A_QOFF = 2048, A_KOFF = A_QOFF + 2 * 8 * A_QS, A_VOFF = A_KOFF + 2 * 32 * A_KS;
constexpr int A_SOFF = A_VOFF + 2 * 32 * A_VS, A_POFF = A_SOFF + 2 * 8 * 32 * 2, A_IOFF = A_POFF + 2 * 8 * A_PS;
static_assert(A_IOFF + 64 <= XBUF_ELEMS, "attention staging must fit the activation buffer");

__device__ __forceinline__ void attn_prefetch(const Phase& ph, Ctx& c) {
  const int kvn = ph.kv_heads, nold = ph.pos0;  // positions [0, pos0) were written in earlier steps
  const int per = nold * (A_HD / 8);            // 16-byte units of one kv head's K (or V) rows
  if (nold > 0) kv_fence<2>();  // acquire side: this CTA has seen tagged words that follow the rows' release fences
  for (int u = c.tid; u < 2 * kvn * per; u += NCT) {
    int r = u;
    const bool isv = r >= kvn * per;
    if (isv) r -= kvn * per;
    const int kvh = r >= per ? 1 : 0;  // kvn <= 2
    r -= kvh * per;
    const int j = r >> 4, i = r & 15;
    const bf16* src = (isv ? ph.vc : ph.kc) + (kvh * ph.slots + j) * A_HD + i * 8;
    bf16* dst = isv ? c.xs + A_VOFF + (kvh * 32 + j) * A_VS + i * 8 : c.xs + A_KOFF + (kvh * 32 + j) * A_KS + i * 8;
    cp_async16(dst, src);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

#ifndef MEGA_PV_UNROLL
#define MEGA_PV_UNROLL 4  /* P.V tiles of a warp in flight together (final tree, ms per frame: 1 -> 2.953, 2 -> 2.943, 4 -> 2.930; tools/runs/gpu_check52.sh) */
#endif
constexpr int PV_UNROLL = MEGA_PV_UNROLL;
__device__ __forceinline__ void attn_small_into_x(const Phase& ph, Ctx& c) {
  const int heads = ph.heads, kvn = ph.kv_heads, nb = ph.nb;
  const int gsh = ph.grp_shift, grp = 1 << gsh;  // heads per kv head (a power of two)
  const int ncols = nb * grp;                    // columns per kv head: (activation row, head in group)
  const int nkmax = ph.pos0 + nb;                // visible keys of the last row
  const int nmt = nkmax > 16 ? 2 : 1;            // 16-key tiles in use
  float* S = reinterpret_cast<float*>(c.xs + A_SOFF);
  bf16* P = c.xs + A_POFF;
  float* inv = reinterpret_cast<float*>(c.xs + A_IOFF);
  {
    // V rows between the last visible key and the end of the last tile in use: P is 0 there, but stale
    // shared memory may hold Inf / NaN patterns (0 * Inf): clear them
    const int zrows = nmt * 16 - nkmax;
    for (int u = c.tid; u < kvn * zrows * 16; u += NCT) {
      const int kvh = u >= zrows * 16 ? 1 : 0, r = u - kvh * zrows * 16;
      *reinterpret_cast<uint4*>(c.xs + A_VOFF + (kvh * 32 + nkmax + (r >> 4)) * A_VS + (r & 15) * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  {  // q rows and the K/V rows of the positions written by the phase that just finished (tagged words)
    const uint32_t tag = tag_of(c.seq, ph.q_src);
    const int krows = kvn * A_HD, hsh = ph.heads_shift;
    const int qunits = nb * heads * (A_HD / 4), kunits = nb * 2 * krows / 4, total = qunits + kunits;
    const uint32_t* tq = my_copy(ph.t_q, ph.q_rs);
    const uint32_t* tkv = my_copy(ph.t_kv, ph.kv_rs);
#pragma unroll 1
    for (int u0 = 0; u0 < total; u0 += 2 * NCT) {
      uint4 v[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int u = u0 + t * NCT + c.tid;
        if (u < total) v[t] = ldv4(u < qunits ? tq + u * 4 : tkv + (u - qunits) * 4);
      }
      for (unsigned spin = 0;; ++spin) {  // both units together (see stage_x)
        bool ok = true;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int u = u0 + t * NCT + c.tid;
          if (u < total && !fresh4(v[t], tag)) {
            v[t] = ldv4(u < qunits ? tq + u * 4 : tkv + (u - qunits) * 4);
            ok = false;
          }
        }
        if (ok) break;
        if (give_up(c.sync, spin, 0x406)) break;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int u = u0 + t * NCT + c.tid;
        if (u < total) {
          bf16* dst;
          if (u < qunits) {
            const int d4 = u & 31, hn = u >> 5, h = hn & (heads - 1), n = hn >> hsh;  // [n][head][128]
            dst = c.xs + A_QOFF + ((h >> gsh) * 8 + n * grp + (h & (grp - 1))) * A_QS + d4 * 4;
          } else {
            const int e = (u - qunits) * 4;  // element in [nb][2][krows]
            const int n = e >= 2 * krows ? 1 : 0, r = e - n * 2 * krows;
            const bool isv = r >= krows;
            const int rr = isv ? r - krows : r, kvh = rr >> 7, d = rr & 127, j = ph.pos0 + n;
            dst = isv ? c.xs + A_VOFF + (kvh * 32 + j) * A_VS + d : c.xs + A_KOFF + (kvh * 32 + j) * A_KS + d;
          }
          *reinterpret_cast<uint2*>(dst) = strip4(v[t]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  csync<NCT, CBAR>();
  CK(6);
  const int g = c.lane >> 2, qd = c.lane & 3;
  if (c.warp < kvn * nmt) {
    // scores of one (kv head, 16-key tile): A = K rows (keys), B = q columns, k = the 128 dims
    const int kvh = c.warp / nmt, mt = c.warp - kvh * nmt;
    const bf16* kp = c.xs + A_KOFF + (kvh * 32 + mt * 16 + g) * A_KS + qd * 8;
    const bf16* qp = c.xs + A_QOFF + (kvh * 8 + (g < ncols ? g : 0)) * A_QS + qd * 8;
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    uint4 ka[4], kb[4], qv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ka[i] = *reinterpret_cast<const uint4*>(kp + i * 32);
      kb[i] = *reinterpret_cast<const uint4*>(kp + 8 * A_KS + i * 32);
      qv[i] = *reinterpret_cast<const uint4*>(qp + i * 32);
      if (g >= ncols) qv[i] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mma16816(s0, ka[i].x, kb[i].x, ka[i].y, kb[i].y, qv[i].x, qv[i].y);
      mma16816(s1, ka[i].z, kb[i].z, ka[i].w, kb[i].w, qv[i].z, qv[i].w);
    }
    const float scale = 0.08838834764831845f;  // 1/sqrt(128)
    float* sp = S + (kvh * 8 + 2 * qd) * 32 + mt * 16 + g;  // accumulator: keys g, g+8 x columns 2qd, 2qd+1
    sp[0] = (s0[0] + s1[0]) * scale;
    sp[32] = (s0[1] + s1[1]) * scale;
    sp[8] = (s0[2] + s1[2]) * scale;
    sp[40] = (s0[3] + s1[3]) * scale;
  }
  csync<NCT, CBAR>();
  CK(7);
  {
    // softmax rows: (kv head, column); lane == key.  The kvn * ncols live rows are dealt round-robin over the warps
    // (one row per warp for a one-row step: 2 kv heads x 4 heads of the group)
    const int csh = gsh + nb - 1;  // log2(ncols): nb is 1 or 2, grp a power of two
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int rv = c.warp + rr * NW;
      if (rv < (kvn << csh)) {
        const int kvh = rv >> csh, col = rv & (ncols - 1), row = kvh * 8 + col;
        const int nkeys = ph.pos0 + (col >> gsh) + 1;
        const bool own = c.lane < nkeys;
        const float sc = own ? S[row * 32 + c.lane] : -INFINITY;
        const float mx = warp_max(sc);
        const float e = own ? expf(sc - mx) : 0.f;
        const float sum = warp_sum(e);
        P[row * A_PS + c.lane] = f2bf(e);
        if (c.lane == 0) inv[row] = 1.0f / sum;
      }
    }
  }
  csync<NCT, CBAR>();
  CK(8);
  {
    // P.V: tiles of 8 output dims; A = P (columns as rows, 8 of 16 used), B = V^T by ldmatrix.trans
    const int tiles = kvn * 16, per = tiles / NW;  // <= 4 (kvn <= 2)
#pragma unroll PV_UNROLL
    for (int i = 0; i < 4; ++i) {
      if (i >= per) break;
      const int t = c.warp * per + i, kvh = t >> 4, n0 = (t & 15) * 8;
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, c.xs + A_VOFF + (kvh * 32 + c.lane) * A_VS + n0);
      const bf16* pp = P + (kvh * 8 + g) * A_PS + 2 * qd;
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(pp), a2 = *reinterpret_cast<const uint32_t*>(pp + 8);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(o, a0, 0u, a2, 0u, vb[0], vb[1]);
      if (nmt == 2) {
        const uint32_t a4 = *reinterpret_cast<const uint32_t*>(pp + 16), a6 = *reinterpret_cast<const uint32_t*>(pp + 24);
        mma16816(o, a4, 0u, a6, 0u, vb[2], vb[3]);
      }
      if (g < ncols) {
        const float is = inv[kvh * 8 + g];
        const int n = g >> gsh, h = kvh * grp + (g & (grp - 1));
        // an 8-row-group O projection reads x with the pairs of an 8-group as (P0,P2,P1,P3), a 16-row one in natural order
        const int ps = ph.R == 16 ? qd : (qd == 1 ? 2 : (qd == 2 ? 1 : qd));
        *reinterpret_cast<__nv_bfloat162*>(c.xs + n * ph.K + h * A_HD + n0 + 2 * ps) = __floats2bfloat162_rn(o[0] * is, o[1] * is);
      }
    }
  }
  csync<NCT, CBAR>();
}

__device__ __forceinline__ float sumsq4(const uint4& v) {
  float s = 0.f, a;
  a = tval(v.x); s = fmaf(a, a, s);
  a = tval(v.y); s = fmaf(a, a, s);
  a = tval(v.z); s = fmaf(a, a, s);
  a = tval(v.w); s = fmaf(a, a, s);
  return s;
}
// torchtune RMSNorm on 4 tagged elements: bf16( bf16(x * inv) * scale ), packed
__device__ __forceinline__ uint2 norm4(const uint4& v, float inv, uint32_t sc01, uint32_t sc23) {
  __nv_bfloat162 o[2];
  o[0] = __floats2bfloat162_rn(rbf(tval(v.x) * inv) * bflo(sc01), rbf(tval(v.y) * inv) * bfhi(sc01));
  o[1] = __floats2bfloat162_rn(rbf(tval(v.z) * inv) * bflo(sc23), rbf(tval(v.w) * inv) * bfhi(sc23));
  return *reinterpret_cast<uint2*>(o);
}

// stage the phase's activation rows into shared memory (+ RMSNorm prologue): poll the tagged words.
// Warp w stages exactly the K/8 slice of every row that ITS mma loop reads, so un-normed phases need
// no CTA barrier at all here (a __syncwarp orders the slice), and normed phases need one, for the
// sum of squares; normed phases (K <= 2048, <= 4 units per lane) keep the words in registers across it.
// A unit is 4 elements = two pairs; the pairs of every 8-group are stored in the order (P0,P2,P1,P3)
// the mma loop expects: unit 2i (P0,P1) -> pair slots 0 and 2, unit 2i+1 (P2,P3) -> slots 1 and 3.
__device__ __forceinline__ void store_unit(bf16* row, int k4, const uint2& pr, bool natural) {
  if (natural) {  // 16-row groups read x as the B operand, in natural order
    *reinterpret_cast<uint2*>(row + k4 * 4) = pr;
  } else {
    uint32_t* d = reinterpret_cast<uint32_t*>(row + (k4 >> 1) * 8) + (k4 & 1);
    d[0] = pr.x;
    d[2] = pr.y;
  }
}

__device__ __forceinline__ void stage_x(const Phase& ph, Ctx& c) {
  if (ph.attn_prologue) {
    attn_small_into_x(ph, c);
    return;
  }
  const int K = ph.K, nb = ph.nb;
  const uint32_t tag0 = tag_of(c.seq, ph.x_src[0]), tag1 = tag_of(c.seq, ph.x_src[1]);
  const bool norm = ph.norm != 0, natural = ph.R == 16;
  const int slice = K / NW, nu = slice >> 2, tw = nb * nu;  // units of 4 words per row / in all rows of the warp's slice
  const uint32_t* txw = my_copy(ph.t_x, ph.x_rs) + c.warp * slice;
  bf16* xsw = c.xs + c.warp * slice;
  uint4 v[4];  // (measured: eight units in flight per lane spill and are slower)
  uint2 sc[4];
  float ss0 = 0.f, ss1 = 0.f;
#pragma unroll 1
  for (int e0 = 0; e0 < tw; e0 += 128) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = e0 + t * 32 + c.lane;
      if (e < tw) {
        const int n = e >= nu ? 1 : 0, k4 = e - n * nu;
        v[t] = ldv4(txw + (size_t)n * ph.ldx + k4 * 4);
        if (norm) sc[t] = *reinterpret_cast<const uint2*>(ph.norm_scale + c.warp * slice + k4 * 4);
      }
    }
    // all units of the batch are polled TOGETHER: one round trip per round for every stale unit, not one
    // after the other (a lane that waits for data would otherwise pay a round trip per unit after it lands)
    for (unsigned spin = 0;; ++spin) {
      bool ok = true;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int e = e0 + t * 32 + c.lane;
        if (e < tw) {
          const int n = e >= nu ? 1 : 0, k4 = e - n * nu;
          if (!fresh4(v[t], n ? tag1 : tag0)) {
            v[t] = ldv4(txw + (size_t)n * ph.ldx + k4 * 4);
            ok = false;
          }
        }
      }
      if (ok) break;
      if (give_up(c.sync, spin, 0x405)) break;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = e0 + t * 32 + c.lane;
      if (e < tw) {
        const int n = e >= nu ? 1 : 0, k4 = e - n * nu;
        if (norm) {
          const float s = sumsq4(v[t]);
          if (n) ss1 += s;
          else ss0 += s;
        } else {
          store_unit(xsw + (size_t)n * K, k4, strip4(v[t]), natural);
        }
      }
    }
  }
  CK(6);
  if (norm) {
    ss0 = warp_sum(ss0);
    ss1 = warp_sum(ss1);
    if (c.lane == 0) {
      c.scratch[c.warp] = ss0;
      c.scratch[NW + c.warp] = ss1;
    }
    CK(7);
    csync<NCT, CBAR>();
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      t0 += c.scratch[w];
      t1 += c.scratch[NW + w];
    }
    const float ik = ph.inv_K;  // exact when K is a power of two
    // torch.rsqrt on a GPU is rsqrtf (<= 2 ulp), which is also ~50 instructions shorter than 1 / sqrtf in this chain
    const float i0 = rsqrtf((ik != 0.f ? t0 * ik : t0 / (float)K) + ph.eps);
    const float i1 = nb == 2 ? rsqrtf((ik != 0.f ? t1 * ik : t1 / (float)K) + ph.eps) : 0.f;
    CK(8);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = t * 32 + c.lane;
      if (e < tw) {
        const int n = e >= nu ? 1 : 0, k4 = e - n * nu;
        store_unit(xsw + (size_t)n * K, k4, norm4(v[t], n ? i1 : i0, sc[t].x, sc[t].y), natural);
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void gemv_phase(const Phase& ph, Ctx& c) {
  const int ncta = gridDim.x;
  const int cl = local_cta(blockIdx.x, ncta, ph.rot);
  const int ngl = ph.gq + (cl < ph.gr ? 1 : 0);
  if (ngl == 0) return;  // nothing to do here: run ahead
  // epilogue item of this thread: (local group, row pair, activation row)
  const int R = ph.R;
  const int ipg_shift = R == 16 ? 4 : 3;  // items per group = (R / 2) pairs x 2 activation rows
  const int j_item = c.tid >> ipg_shift, rem = c.tid & ((1 << ipg_shift) - 1);
  const int pair = rem >> 1, n_item = rem & 1;
  const int r0 = (cl + ncta * j_item) * R + 2 * pair;
  const bool item_on = j_item < ngl && n_item < ph.nb && r0 < ph.rows;
  EpiPre pre;
  pre.a = pre.b = 0u;
  if (item_on) pre = epilogue_prefetch(ph, c, r0, n_item);
  CK(5);
  stage_x(ph, c);
  if (c.trp) c.trp[1] = gtimer();
  gemv_groups(ph, c, cl, ngl, j_item, pair, n_item, item_on, pre);
}

// Long contexts (a voice prompt: 1568 frames): with one CTA per q-head 32 of the 148 CTAs walk the whole cache -- 29 us
// per layer at 1568 keys, +0.47 ms per frame.  From ATTN_SPLIT_MIN keys on, the cache of a head is cut into ATTN_SPLITS
// key ranges, one CTA each (128 CTAs): every CTA computes the softmax numerator / denominator of its range against
// its own maximum (m, l, o[hd]: the flash-decoding partial), publishes them as tagged words -- an fp32 value is two
// words, 16 payload bits each -- and the CTAs of range 0 combine the ranges of their head in range order
// (deterministic) into the phase's output.  One more hand-off inside the phase, taken only on this path; a separate
// function so that the short-context path keeps its registers and code.
constexpr int ATTN_SPLITS = 4;
constexpr int ATTN_SPLIT_MIN = 512;
constexpr int ATTN_PART_WORDS = 2 * 64 + 4;  // o[64] + m + l, two words per value (head_dim 64)
__device__ __forceinline__ void st_f32_tagged(uint32_t* p, uint32_t tag, float v) {
  const uint32_t b = __float_as_uint(v);
  stv2(p, make_uint2(tag | (b >> 16), tag | (b & 0xffffu)));
}
__device__ __forceinline__ float ld_f32_tagged(const uint32_t* p, uint32_t tag, Sync* sync) {
  const uint32_t hi = poll1(p, tag, sync), lo = poll1(p + 1, tag, sync);
  return __uint_as_float((hi << 16) | (lo & 0xffffu));
}
// (everything it needs from the kernel's context comes by value: a reference would force that struct into local memory)
__device__ __noinline__ void attn_split(const Phase& ph, int slot, int tid, int bb_lane, unsigned seq, uint32_t otag, bf16* xs,
                                        float* scratch, Sync* sync) {
  const int cta = blockIdx.x;
  if (cta >= ATTN_SPLITS * ph.heads) return;
  const int h = cta % ph.heads, sp = cta / ph.heads;
  constexpr int hd = 64;
  const int nold = slot;  // keys in the cache; the current position's K / V are tagged words of the QKV phase
  const int chunk = (nold + ATTN_SPLITS - 1) / ATTN_SPLITS;
  const int k0 = sp * chunk < nold ? sp * chunk : nold;
  const int k1 = k0 + chunk < nold ? k0 + chunk : nold;
  const bool last = sp == ATTN_SPLITS - 1;  // this range also takes the current position
  const int kvh = h / (ph.heads / ph.kv_heads);
  const int krows = ph.kv_heads * hd;
  const bf16* kp = ph.kc + ((size_t)bb_lane * ph.kv_heads + kvh) * ph.slots * hd;
  const bf16* vp = ph.vc + ((size_t)bb_lane * ph.kv_heads + kvh) * ph.slots * hd;
  float* sc = reinterpret_cast<float*>(xs);   // scores of this range: sc[j - k0], the current position at [k1 - k0]
  float* part = sc + ((chunk + 1 + 3) & ~3);     // [NCT / 8][hd] partial outputs
  float* qs = part + (NCT / 8) * hd;             // [hd], then kcur [hd], vcur [hd]
  float* kcur = qs + hd;
  float* vcur = kcur + hd;
  const float scale = 1.0f / sqrtf((float)hd);
  const uint32_t tag = tag_of(seq, ph.q_src);
  for (int d = tid; d < (last ? 3 * hd : hd); d += NCT) {
    const int which = d / hd, dd = d - which * hd;
    const uint32_t* src = which == 0 ? my_copy(ph.t_q, ph.q_rs) + (size_t)h * hd + dd
                                     : my_copy(ph.t_kv, ph.kv_rs) + (size_t)(which - 1) * krows + (size_t)kvh * hd + dd;
    qs[d] = tval(poll1(src, tag, sync));
  }
  csync<NCT, CBAR>();
  const int n = k1 - k0 + (last ? 1 : 0);  // scores of this range
  float mx = -INFINITY;
  for (int j = tid; j < n; j += NCT) {
    float s = 0.f;
    if (j == k1 - k0) {
      for (int i = 0; i < hd; ++i) s = fmaf(qs[i], kcur[i], s);
    } else {
      const bf16* kr = kp + (size_t)(k0 + j) * hd;
#pragma unroll
      for (int i = 0; i < hd / 8; ++i) {
        const uint4 v = ldcg16(kr + i * 8);
        const float* q8 = qs + i * 8;
        s = fmaf(q8[0], bflo(v.x), s); s = fmaf(q8[1], bfhi(v.x), s);
        s = fmaf(q8[2], bflo(v.y), s); s = fmaf(q8[3], bfhi(v.y), s);
        s = fmaf(q8[4], bflo(v.z), s); s = fmaf(q8[5], bfhi(v.z), s);
        s = fmaf(q8[6], bflo(v.w), s); s = fmaf(q8[7], bfhi(v.w), s);
      }
    }
    s *= scale;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_max<NCT, CBAR>(mx, scratch, tid);
  float sum = 0.f;
  for (int j = tid; j < n; j += NCT) {
    const float e = expf(sc[j] - mx);  // (an empty range: no iterations, mx = -inf, sum = 0)
    sc[j] = e;
    sum += e;
  }
  sum = block_sum<NCT, CBAR>(sum, scratch, tid);
  // P.V over the cached keys of the range: 8 dims per thread, 8 threads per key row, NCT / 8 key groups
  constexpr int tpr = hd >> 3, ngroups = NCT / tpr;
  const int kg = tid / tpr, d8 = (tid - kg * tpr) * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  {
    const int nr = k1 - k0;
    const bf16* vr = vp + (size_t)k0 * hd + d8;
    int j = kg;
    for (; j + 3 * ngroups < nr; j += 4 * ngroups) {  // four V rows in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldcg16(vr + (size_t)(j + u * ngroups) * hd);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float p = sc[j + u * ngroups];
        acc[0] = fmaf(p, bflo(v[u].x), acc[0]); acc[1] = fmaf(p, bfhi(v[u].x), acc[1]);
        acc[2] = fmaf(p, bflo(v[u].y), acc[2]); acc[3] = fmaf(p, bfhi(v[u].y), acc[3]);
        acc[4] = fmaf(p, bflo(v[u].z), acc[4]); acc[5] = fmaf(p, bfhi(v[u].z), acc[5]);
        acc[6] = fmaf(p, bflo(v[u].w), acc[6]); acc[7] = fmaf(p, bfhi(v[u].w), acc[7]);
      }
    }
    for (; j < nr; j += ngroups) {
      const uint4 v = ldcg16(vr + (size_t)j * hd);
      const float p = sc[j];
      acc[0] = fmaf(p, bflo(v.x), acc[0]); acc[1] = fmaf(p, bfhi(v.x), acc[1]);
      acc[2] = fmaf(p, bflo(v.y), acc[2]); acc[3] = fmaf(p, bfhi(v.y), acc[3]);
      acc[4] = fmaf(p, bflo(v.z), acc[4]); acc[5] = fmaf(p, bfhi(v.z), acc[5]);
      acc[6] = fmaf(p, bflo(v.w), acc[6]); acc[7] = fmaf(p, bfhi(v.w), acc[7]);
    }
  }
  if (last && kg == 0) {
    const float p = sc[k1 - k0];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vcur[d8 + e], acc[e]);
  }
  *reinterpret_cast<float4*>(part + kg * hd + d8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(part + kg * hd + d8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  csync<NCT, CBAR>();
  uint32_t* base = reinterpret_cast<uint32_t*>(sync->att_part);
  uint32_t* mine = base + (size_t)(sp * ph.heads + h) * ATTN_PART_WORDS;
  if (tid < hd) {
    float o = 0.f;
    for (int gg = 0; gg < ngroups; ++gg) o += part[gg * hd + tid];  // fixed order
    st_f32_tagged(mine + 2 * tid, otag, o);
  } else if (tid == hd) {
    st_f32_tagged(mine + 2 * hd, otag, mx);
    st_f32_tagged(mine + 2 * hd + 2, otag, sum);
  }
  if (sp != 0) return;
  // range 0 combines the ranges of its head, in range order
  if (tid < hd) {
    float m[ATTN_SPLITS], l[ATTN_SPLITS], o[ATTN_SPLITS];
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < ATTN_SPLITS; ++r) {
      const uint32_t* pr = base + (size_t)(r * ph.heads + h) * ATTN_PART_WORDS;
      o[r] = ld_f32_tagged(pr + 2 * tid, otag, sync);
      m[r] = ld_f32_tagged(pr + 2 * hd, otag, sync);
      l[r] = ld_f32_tagged(pr + 2 * hd + 2, otag, sync);
      M = fmaxf(M, m[r]);
    }
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int r = 0; r < ATTN_SPLITS; ++r) {
      const float w = expf(m[r] - M);  // (an empty range: exp(-inf) = 0)
      L = fmaf(l[r], w, L);
      O = fmaf(o[r], w, O);
    }
    rep_st1(ph.t_out + (size_t)h * hd + tid, ph.out_rs, tword(otag, O * (1.0f / L)));
  }
}

// backbone attention: CTA h < heads owns q-head h over keys [0 .. slot]; the current position's K/V
// come from the QKV phase's tagged words, earlier positions from the cache (previous launches)
__device__ __forceinline__ void attn_phase(const Phase& ph, Ctx& c) {
  const int h = blockIdx.x;
  const int hd = ph.hd;
  int pos, slot;
  phase_pos(ph, c, 0, pos, slot);
  if (slot + 1 >= ATTN_SPLIT_MIN && hd == 64 && ATTN_SPLITS * ph.heads <= (int)gridDim.x) {
    attn_split(ph, slot, c.tid, c.bb_lane, c.seq, c.tag, c.xs, c.scratch, c.sync);
    return;
  }
  if (h >= ph.heads) return;
  const int nkeys = slot + 1;
  const int kvh = h / (ph.heads / ph.kv_heads);
  const int krows = ph.kv_heads * hd;
  const bf16* kp = ph.kc + ((size_t)c.bb_lane * ph.kv_heads + kvh) * ph.slots * hd;
  const bf16* vp = ph.vc + ((size_t)c.bb_lane * ph.kv_heads + kvh) * ph.slots * hd;
  float* sc = reinterpret_cast<float*>(c.xs);  // [slots] scores (<= 8 KB)
  float* part = sc + ((ph.slots + 3) & ~3);      // [NCT / (hd/8)][hd] = 8 NCT partial outputs (16-byte aligned)
  float* qs = part + 8 * NCT;                   // [hd]
  float* kcur = qs + hd;                        // [hd]
  float* vcur = kcur + hd;                      // [hd]
  const float scale = 1.0f / sqrtf((float)hd);
  const uint32_t tag = tag_of(c.seq, ph.q_src);
  for (int d = c.tid; d < 3 * hd; d += NCT) {
    const int which = d / hd, dd = d - which * hd;
    const uint32_t* src = which == 0 ? my_copy(ph.t_q, ph.q_rs) + (size_t)h * hd + dd
                                     : my_copy(ph.t_kv, ph.kv_rs) + (size_t)(which - 1) * krows + (size_t)kvh * hd + dd;
    qs[d] = tval(poll1(src, tag, c.sync));  // qs, kcur, vcur are contiguous
  }
  csync<NCT, CBAR>();
  float mx = -INFINITY;
  for (int j = c.tid; j < nkeys; j += NCT) {
    float s = 0.f;
    if (j == slot) {
      for (int i = 0; i < hd; ++i) s = fmaf(qs[i], kcur[i], s);
    } else {
      const bf16* kr = kp + (size_t)j * hd;
      for (int i = 0; i < hd / 8; ++i) {
        const uint4 v = ldcg16(kr + i * 8);
        const float* q8 = qs + i * 8;
        s = fmaf(q8[0], bflo(v.x), s); s = fmaf(q8[1], bfhi(v.x), s);
        s = fmaf(q8[2], bflo(v.y), s); s = fmaf(q8[3], bfhi(v.y), s);
        s = fmaf(q8[4], bflo(v.z), s); s = fmaf(q8[5], bfhi(v.z), s);
        s = fmaf(q8[6], bflo(v.w), s); s = fmaf(q8[7], bfhi(v.w), s);
      }
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
  // P.V: 8 dims (one 16-byte load) per thread, hd/8 threads per key row, NCT / (hd/8) key groups:
  // a context of a few hundred keys is one batch of loads in flight per thread
  const int tpr = hd >> 3, ngroups = NCT / tpr;
  const int kg = c.tid / tpr, d8 = (c.tid - kg * tpr) * 8;
  const int nold = nkeys - 1;  // keys in the cache
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  int j = kg;
  for (; j + 3 * ngroups < nold; j += 4 * ngroups) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldcg16(vp + (size_t)(j + u * ngroups) * hd + d8);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float p = sc[j + u * ngroups];
      acc[0] = fmaf(p, bflo(v[u].x), acc[0]); acc[1] = fmaf(p, bfhi(v[u].x), acc[1]);
      acc[2] = fmaf(p, bflo(v[u].y), acc[2]); acc[3] = fmaf(p, bfhi(v[u].y), acc[3]);
      acc[4] = fmaf(p, bflo(v[u].z), acc[4]); acc[5] = fmaf(p, bfhi(v[u].z), acc[5]);
      acc[6] = fmaf(p, bflo(v[u].w), acc[6]); acc[7] = fmaf(p, bfhi(v[u].w), acc[7]);
    }
  }
  for (; j < nold; j += ngroups) {
    const uint4 v = ldcg16(vp + (size_t)j * hd + d8);
    const float p = sc[j];
    acc[0] = fmaf(p, bflo(v.x), acc[0]); acc[1] = fmaf(p, bfhi(v.x), acc[1]);
    acc[2] = fmaf(p, bflo(v.y), acc[2]); acc[3] = fmaf(p, bfhi(v.y), acc[3]);
    acc[4] = fmaf(p, bflo(v.z), acc[4]); acc[5] = fmaf(p, bfhi(v.z), acc[5]);
    acc[6] = fmaf(p, bflo(v.w), acc[6]); acc[7] = fmaf(p, bfhi(v.w), acc[7]);
  }
  if (kg == 0) {
    const float p = sc[slot];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vcur[d8 + e], acc[e]);
  }
  *reinterpret_cast<float4*>(part + kg * hd + d8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(part + kg * hd + d8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  csync<NCT, CBAR>();
  if (c.tid < hd) {
    float o = 0.f;
    for (int gg = 0; gg < ngroups; ++gg) o += part[gg * hd + c.tid];  // fixed order
    rep_st1(ph.t_out + (size_t)h * hd + c.tid, ph.out_rs, tword(c.tag, o * (1.0f / sum)));
  }
}

__device__ __forceinline__ void embed_phase(const Phase& ph, Ctx& c) {
  // one 8-element unit of h per thread: unit u = cta + ncta * tid
  const FrameParams* P = c.P;
  const int u = blockIdx.x + gridDim.x * c.tid;
  if (u >= ph.D / 8) return;
  const size_t fr = (size_t)(P->S - 1);  // last prompt row of stream 0
  const int64_t* tok = P->tokens + fr * (ph.C + 1);
  const uint8_t* msk = P->mask + fr * (ph.C + 1);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int c0 = 0; c0 <= ph.C; c0 += 11) {  // 11 independent gathers in flight, summed in column order (17 spill: slower overall)
    uint4 v[11];
    bool on[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const int cb = c0 + j;
      on[j] = cb <= ph.C && msk[cb < ph.C + 1 ? cb : 0] != 0;
      v[j] = make_uint4(0, 0, 0, 0);
      if (on[j]) {
        // (ids the reference would raise on are reported once, by the kernel prologue, and read as row 0 here)
        const unsigned long long tt = (unsigned long long)tok[cb];
        const size_t t = tt < (unsigned long long)(cb < ph.C ? ph.V : ph.TV) ? (size_t)tt : 0;
        const bf16* row = (cb < ph.C) ? ph.audio_emb + (t + (size_t)ph.V * cb) * ph.D : ph.text_emb + t * ph.D;
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
  const uint32_t tag = c.tag;
  rep_st4(ph.t_out + u * 8, ph.out_rs, make_uint4(tword(tag, acc[0]), tword(tag, acc[1]), tword(tag, acc[2]), tword(tag, acc[3])));
  rep_st4(ph.t_out + u * 8 + 4, ph.out_rs, make_uint4(tword(tag, acc[4]), tword(tag, acc[5]), tword(tag, acc[6]), tword(tag, acc[7])));
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void sample_phase(const Phase& ph, Ctx& c) {
  if (blockIdx.x != 0) return;
  const FrameParams* P = c.P;
  float* xs = reinterpret_cast<float*>(c.xs);                              // [4096]
  unsigned int* hist = reinterpret_cast<unsigned int*>(xs + SAMPLE_MAXV);  // [256]
  bf16* lg = reinterpret_cast<bf16*>(hist + 256);                          // [V] logits of this step
  const int V = ph.V, C = ph.C, cb = ph.cb;
  const bool fast = P->topk == 1 && !P->forced;  // greedy: arg-max straight from the polled registers
  const float inv_t = 1.0f / P->temperature;
  float best = -INFINITY;
  int besti = 0x7fffffff, bestn = 0;
  {
    // the head phase wrote every word below the 16-row padded vocabulary: poll whole 4-word units
    const uint32_t tag = tag_of(c.seq, ph.logits_src);
    constexpr int MAXU = SAMPLE_MAXV / 4 / NCT;
    const int nunits = (V + 3) >> 2;
    uint4 w[MAXU];
#pragma unroll
    for (int t = 0; t < MAXU; ++t) {
      const int u = c.tid + t * NCT;
      if (u < nunits) w[t] = ldv4(ph.t_logits + u * 4);
    }
    for (unsigned spin = 0;; ++spin) {  // all units together (see stage_x)
      bool ok = true;
#pragma unroll
      for (int t = 0; t < MAXU; ++t) {
        const int u = c.tid + t * NCT;
        if (u < nunits && !fresh4(w[t], tag)) {
          w[t] = ldv4(ph.t_logits + u * 4);
          ok = false;
        }
      }
      if (ok) break;
      if (give_up(c.sync, spin, 0x407)) break;
    }
#pragma unroll
    for (int t = 0; t < MAXU; ++t) {
      const int u = c.tid + t * NCT;
      if (u < nunits) {
        const uint32_t ww[4] = {w[t].x, w[t].y, w[t].z, w[t].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = u * 4 + e;
          if (i < V) {
            lg[i] = __ushort_as_bfloat16((unsigned short)(ww[e] & 0xffffu));
            const float xv = rbf(tval(ww[e]) * inv_t);  // the value sample_row compares (x * (1/T) rounded to bf16)
            if (xv > best) { best = xv; besti = i; bestn = 1; }
            else if (xv == best) ++bestn;
          }
        }
      }
    }
  }
  if (c.trp) c.trp[1] = gtimer();
  int tok = -1;
  if (fast) {
    // warp arg-max (first index on ties, tie count), then the 8 warp results through shared memory
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      const int on = __shfl_xor_sync(0xffffffffu, bestn, o);
      if (ob > best) { best = ob; besti = oi; bestn = on; }
      else if (ob == best) { bestn += on; besti = min(besti, oi); }
    }
    if (c.lane == 0) {
      c.scratch[c.warp] = best;
      c.iscratch[c.warp] = besti;
      c.iscratch[NW + c.warp] = bestn;
    }
    // speculation: start pulling this warp's candidate rows of the gather tables towards L2, one 128-byte
    // line per lane (every lane holds the warp's arg-max after the butterfly)
    if (ph.t_next) {
      const int hb = ph.next_ld * 2;  // bytes of a projection(embedding) row
      const char* row = reinterpret_cast<const char*>(ph.next_table) + ((size_t)besti + (size_t)cb * V) * hb;
      for (int o = c.lane * 128; o < hb; o += 32 * 128) prefetch_l2(row + o);
      if (ph.qkv_table) {
        const int qb = (ph.heads + 2 * ph.kv_heads) * ph.hd * 2;  // bytes of a [q;k;v] row
        const char* qr = reinterpret_cast<const char*>(ph.qkv_table) + ((size_t)besti + (size_t)cb * V) * qb;
        for (int o = c.lane * 128; o < qb; o += 32 * 128) prefetch_l2(qr + o);
      }
    }
    csync<NCT, CBAR>();
    float b = c.scratch[0];
    int bi = c.iscratch[0], bn = c.iscratch[NW];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      const float ob = c.scratch[w];
      const int oi = c.iscratch[w], on = c.iscratch[NW + w];
      if (ob > b) { b = ob; bi = oi; bn = on; }
      else if (ob == b) { bn += on; bi = min(bi, oi); }
    }
    // a UNIQUE maximum is the token (probability exactly 1, see sample_row); ties take the general path
    if (bn == 1) tok = bi;
  }
  csync<NCT, CBAR>();  // lg[] complete; scratch reads done
  if (P->logits_out)
    for (int i = c.tid; i < V; i += NCT) P->logits_out[(size_t)cb * V + i] = lg[i];
  if (tok < 0) {
    const bf16* nz = P->noise ? P->noise + (size_t)cb * V : nullptr;
    const unsigned long long ctr = ((P->offset * (unsigned long long)C + cb)) * 4096ull;
    tok = sample_row<NCT, CBAR, false>(lg, nz, V, P->temperature, P->topk, P->seed, ctr, xs, hist, c.scratch, c.iscratch,
                                       c.tid);
  }
  if (c.trp) c.trp[2] = gtimer();
  if (c.tid == 0 && P->sampled_out) P->sampled_out[cb] = tok;
  if (P->forced) tok = P->forced[cb];
  if ((unsigned)tok >= (unsigned)V) {  // a teacher-forced id out of range, or garbage logits of a draining launch
    if (P->forced && c.tid == 0) report_error(c.sync, 0x802);
    tok = 0;
  }
  if (c.tid == 0) P->out[cb] = tok;
  if (ph.t_next) {
    // next depth-decoder input: projection(embed_audio(cb, tok)) read from the table built at setup, and -- when
    // the [q;k;v] table exists -- the first layer's q / k / v of the next step, a pure function of (codebook,
    // token) too: one more row gather instead of a GEMV phase.  All loads of both rows go out before any store
    // (the rows are cold in HBM: one latency, not two).
    const int ld = ph.next_ld;
    const bf16* row = ph.next_table + ((size_t)tok + (size_t)cb * V) * ld;
    const uint32_t tag = c.tag;
    const int hd = ph.hd, qrows = ph.heads * hd, krows = ph.kv_heads * hd;
    const bf16* qrow = ph.qkv_table ? ph.qkv_table + ((size_t)tok + (size_t)cb * V) * (qrows + 2 * krows) : nullptr;
    const int nh = ld / 4, nq = qrow ? (qrows + 2 * krows) / 4 : 0;  // 4-element units
    constexpr int MAXH = 2, MAXQ = 3;  // rows up to 2048 / 3072 elements
    uint2 hv[MAXH], qv[MAXQ];
#pragma unroll
    for (int t = 0; t < MAXH; ++t)
      if (c.tid + t * NCT < nh) hv[t] = *reinterpret_cast<const uint2*>(row + (c.tid + t * NCT) * 4);
#pragma unroll
    for (int t = 0; t < MAXQ; ++t)
      if (c.tid + t * NCT < nq) qv[t] = *reinterpret_cast<const uint2*>(qrow + (c.tid + t * NCT) * 4);
#pragma unroll
    for (int t = 0; t < MAXH; ++t) {
      const int d4 = c.tid + t * NCT;
      if (d4 < nh)
        rep_st4(ph.t_next + d4 * 4, ph.next_rs, make_uint4(tword_raw(tag, hv[t].x), tword_raw(tag, hv[t].x >> 16),
                                                            tword_raw(tag, hv[t].y), tword_raw(tag, hv[t].y >> 16)));
    }
    // q and the new k / v go out as tagged words for the attention of the next phase, k / v also as plain rows
    // into the cache (slot pos0) for the later steps
#pragma unroll
    for (int t = 0; t < MAXQ; ++t) {
      const int u = c.tid + t * NCT;
      if (u < nq) {
        const int e = u * 4;
        const uint2 v = qv[t];
        const uint4 w = make_uint4(tword_raw(tag, v.x), tword_raw(tag, v.x >> 16), tword_raw(tag, v.y), tword_raw(tag, v.y >> 16));
        if (e < qrows) {
          rep_st4(ph.t_q + e, ph.q_rs, w);
        } else {
          const bool isk = e < qrows + krows;
          const int rr = e - (isk ? qrows : qrows + krows);
          rep_st4(ph.t_kv + (isk ? 0 : krows) + rr, ph.kv_rs, w);
          bf16* dst = (isk ? ph.kc : ph.vc) + ((size_t)(rr >> ph.hd_shift) * ph.slots + ph.pos0) * hd + (rr & (hd - 1));
          *reinterpret_cast<uint2*>(dst) = v;
          kv_fence<4>();
        }
      }
    }
  }
}

// ---- the kernel ------------------------------------------------------------------------------------
// Ordering of the depth decoder's plain KV rows, once per codebook step (MEGA_KV_FENCE bit 3).  Called by every CTA --
// with or without work in the phase -- after the CTA barrier that ends the phase BEFORE one flagged kv_sync (the flag
// sits on the next descriptor because that one stays put until this warp passes the next barrier):
//  1 (before the last layer's down phase of step i, i.e. after its gate/up phase: every KV row of steps <= i that
//    this CTA wrote -- QKV epilogues, the sampling phase's table gather -- precedes the barrier): ONE thread runs the
//    release  st.release.gpu done[cta] = tag(done_src)  (bit 4 clear: fence.acq_rel.gpu ; st.relaxed.gpu -- same
//    cost).  The release is cumulative over the other threads' stores through the barrier.  Measured: 0.75 us in that
//    down phase (the thread's warp joins the phase late).
//  2 (before the sampling phase of step i, two phases later): the last warp runs the acquire  ld.acquire.gpu
//    done[all CTAs] until they carry that tag  (bit 4 clear: ld.volatile ... ; fence.acq_rel.gpu).  The words were
//    stored ~10 us earlier, so this is one L2 round trip, and only ONE CTA works in the sampling phase (+0.45 us there).  Every attn_prefetch of step i + 1 comes after
//    the barrier that ends the sampling phase, i.e. after this fence in causality order, and reads rows of
//    steps <= i only (the row of step i + 1 travels as tagged words).
// Release -> observed done word -> acquire is a direct synchronizes-with edge between EVERY writer CTA and EVERY
// reader CTA; no chain of relaxed hand-offs is relied on.
__device__ __forceinline__ void kv_step_sync(const Phase& ph, const Ctx& c) {
  constexpr bool FUSED = ((MEGA_KV_FENCE) & 16) != 0;  // st.release / ld.acquire instead of fence + relaxed access
  const uint32_t dtag = tag_of(c.seq, ph.done_src);
  if ((ph.kv_sync & 255) == 1) {
    if (c.tid == NCT - 1) {
      if (FUSED) {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ph.t_done + blockIdx.x), "r"(dtag) : "memory");
      } else {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        st_relaxed_gpu(ph.t_done + blockIdx.x, dtag);
      }
      if (blockIdx.x == gridDim.x - 1)  // pad to whole 16-byte units for the pollers
        for (unsigned q = gridDim.x; (q & 3u) != 0; ++q) st_relaxed_gpu(ph.t_done + q, dtag);
    }
  } else if ((ph.kv_sync & 255) == 2 && c.warp == NW - 1) {
    for (unsigned q = c.lane * 4; q < gridDim.x; q += 128) {
      uint4 v = FUSED ? lda4(ph.t_done + q) : ldv4(ph.t_done + q);
      for (unsigned spin = 0; !fresh4(v, dtag); ++spin) {
        if (give_up(c.sync, spin, 0x40b)) break;
        v = FUSED ? lda4(ph.t_done + q) : ldv4(ph.t_done + q);
      }
    }
    if (!FUSED) asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
}

// Consumer side of the producer-warp variant (kv_producer_poll).  kv_sync = type | previous request's done_src << 8:
// type 1 / 2: thread NCT-1 posts the request (after checking that the previous one was acknowledged: long done);
// type 3 (the next step's first phase): every thread waits for the acknowledgement of the step's acquire.
__device__ __forceinline__ void kv_post(const Phase& nx, const Ctx& c, uint32_t* kvw) {
  const int type = nx.kv_sync & 255, prev_src = nx.kv_sync >> 8;
  const uint32_t dtag = tag_of(c.seq, nx.done_src);
  if (type == 3) {
    for (unsigned spin = 0; lds_acquire_cta(&kvw[1]) != (dtag | 2u); ++spin)
      if (give_up(c.sync, spin, 0x40d)) break;
  } else if (c.tid == NCT - 1) {
    const uint32_t prev = type == 2 ? (dtag | 1u) : (prev_src >= 0 ? (tag_of(c.seq, prev_src) | 2u) : 0u);
    for (unsigned spin = 0; lds_acquire_cta(&kvw[1]) != prev; ++spin)
      if (give_up(c.sync, spin, 0x40e)) break;
    sts_release_cta(&kvw[0], dtag | (uint32_t)type);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_frame_mega(const Phase* __restrict__ phases, int nphases, const FrameParams* __restrict__ P, Sync* sync,
             unsigned long long* __restrict__ trace /* optional [ncta][nphases][16]: 4 globaltimer ns + 12 clock64 marks, thread 0 of every CTA */,
             const __grid_constant__ PfTable tab) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  bf16* xs = reinterpret_cast<bf16*>(smem + SMEM_RING);
  float* psum = reinterpret_cast<float*>(smem + SMEM_RING + SMEM_X);
  unsigned char* misc = smem + SMEM_RING + SMEM_X + SMEM_PSUM;
  Phase* phbuf = reinterpret_cast<Phase*>(misc);                           // [2] double buffer
  uint64_t* full = reinterpret_cast<uint64_t*>(misc + 2 * sizeof(Phase));  // [NW*SLOTS]
  uint64_t* empty = full + NW * SLOTS;                                     // [NW*SLOTS]
  float* scratch = reinterpret_cast<float*>(empty + NW * SLOTS);           // [34]
  int* iscratch = reinterpret_cast<int*>(scratch + 34);                    // [40]
  uint32_t* kvw = reinterpret_cast<uint32_t*>(iscratch + 40);              // [2] KV ordering on the producer warp: request, acknowledgement
  static_assert(2 * sizeof(Phase) + 2 * NW * SLOTS * 8 + 34 * 4 + 40 * 4 + 2 * 4 <= SMEM_MISC, "misc region too small");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NW * SLOTS; ++i) mbar_init(&full[i], 1);  // full[] and empty[] are contiguous
    kvw[0] = kvw[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  constexpr int PH16 = sizeof(Phase) / 16;
  if (threadIdx.x < PH16)
    reinterpret_cast<uint4*>(&phbuf[0])[threadIdx.x] = reinterpret_cast<const uint4*>(&phases[0])[threadIdx.x];
  __syncthreads();

  if (threadIdx.x >= NCT) {  // producer warp: feeds the eight weight rings for the whole frame, then retires
    producer_loop(tab.d, tab.n, ring, full, empty, sync, threadIdx.x - NCT, kvw, tab.t_done);  // the schedule is read from parameter space: only this warp walks it
    return;
  }

  Ctx c;
  c.P = P; c.trp = nullptr; c.ring = ring; c.xs = xs; c.full = full; c.empty = empty; c.scratch = scratch; c.iscratch = iscratch;
  c.psum = psum; c.sync = sync; c.cnt = 0; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
  c.bb_slot = row_len(P, 0) + P->S - 1;
  c.bb_lane = row_lane(P, 0);
  {
    // batch 1: stream 0, last prompt row.  Keys are masked by cache slot; the reference masks by input_pos:
    // identical when input_pos == cache position (its only use).  k_mega_prepare reports anything else; here
    // the position is only clamped into the RoPE table.
    const int64_t pos = P->pos[P->S - 1];
    const int lim = phbuf[0].rope_len - 1;
    c.bb_pos = pos < 0 ? 0 : (pos > lim ? lim : (int)pos);
  }
  c.seq = sync->seq;

  const bool tr = trace != nullptr && c.tid == 0;
  for (int p = 0; p < nphases; ++p) {
    const Phase& ph = phbuf[p & 1];
    c.tag = tag_of(c.seq, p);
    if (tr) {
      c.trp = trace + ((size_t)blockIdx.x * nphases + p) * 16;
      c.trp[0] = gtimer();
      c.trp[1] = c.trp[2] = 0;
      c.trp[4] = clock64();
    }
    // stage the next descriptor while this phase runs (asynchronously: nobody waits for the load here)
    if (p + 1 < nphases && c.tid < PH16)
      cp_async16(reinterpret_cast<uint4*>(&phbuf[(p + 1) & 1]) + c.tid, reinterpret_cast<const uint4*>(&phases[p + 1]) + c.tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    switch (ph.type) {
      case PH_GEMV: gemv_phase(ph, c); break;
      case PH_EMBED: embed_phase(ph, c); break;
      case PH_ATTN: attn_phase(ph, c); break;
      default: sample_phase(ph, c); break;
    }
    if (tr) {
      c.trp[3] = gtimer();
      c.trp[15] = clock64();
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // the next descriptor has landed
    // end of phase inside the CTA: shared activations / partial sums may be overwritten from here on
    csync<NCT, CBAR>();
    if (p + 1 < nphases) {
      const Phase& nx = phbuf[(p + 1) & 1];  // stable until this warp passes the NEXT phase's barrier (ph is not: the others re-stage it)
      if (((MEGA_KV_FENCE) & 8) && nx.kv_sync) kv_step_sync(nx, c);
      if (KV_PROD && nx.kv_sync) kv_post(nx, c, kvw);
      if (nx.type == PH_GEMV && nx.attn_prologue) {
        const int cl = local_cta(blockIdx.x, gridDim.x, nx.rot);
        if (nx.gq + (cl < nx.gr ? 1 : 0) > 0) attn_prefetch(nx, c);
      }
    }
  }
  if (KV_PROD && c.tid == NCT - 1) sts_release_cta(&kvw[0], 3u);  // lets the producer warp retire
}

}  // namespace mega
