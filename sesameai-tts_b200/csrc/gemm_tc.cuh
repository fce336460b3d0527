// Dense bf16 GEMM on the 5th-generation tensor cores: Y[rows, out] = X[rows, K] . W[out, K]^T
// (both operands K-major, fp32 accumulation in TMEM, bf16 result), for the places where the hot
// path is a real contraction: prompt prefill and batched decode.
//
//   warp 0      TMA producer  : cp.async.bulk.tensor.2d (128B swizzle) of a 128 x 64 X tile and a
//                               128 x 64 W tile per stage into a 6-deep shared-memory ring
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                               (M128 x N128 x K16, 4 per stage) into a 128-column TMEM accumulator;
//                               tcgen05.commit releases the stage / signals the epilogue
//   warps 2..5  epilogue      : tcgen05.ld (32 lanes x 32 columns per warp per step) -> registers ->
//                               fused epilogue (bf16 round, + residual, SwiGLU on interleaved
//                               gate/up columns) -> global
// Three kernels share these pieces:
//   k_gemm_tc        one output tile per CTA (the layout above).  Serves the decode-sized GEMMs (<= 512 rows), whose
//                    tiles all fit on the SMs at once, with CLUSTER SPLIT-K: the z extent of the grid is one
//                    thread-block cluster of 2 / 4 / 8 CTAs, CTA z accumulates k slice z, the fp32 partial tiles meet
//                    through distributed shared memory and are added in slice order (deterministic).  An older
//                    split-K through global memory (partial tiles in a workspace, a counter per tile, the last CTA of
//                    a tile adds the slices in order) remains behind CSM_TC_SPLITK.
//   k_gemm_tc_p<128> persistent: one CTA per SM walks the tiles, two TMEM accumulators, eight epilogue warps, coalesced
//                    epilogue through a shared-memory tile -- every GEMM with more tiles than SMs at decode row counts.
//   k_gemm_tc_p<256> the same with 128 x 256 tiles (three 48 KB stages, all 512 TMEM columns): prompt-sized GEMMs.
// Every wait is trip-capped and traps instead of hanging.  All three take part in programmatic dependent launch
// (common.cuh): W boxes may be requested before griddepcontrol.wait, X boxes and the epilogue come after it.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tc {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 6;
constexpr int UMMA_K = 16;
constexpr int THREADS = 192;
constexpr uint32_t STAGE_BYTES = (BM + BN) * BK * 2;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { EPI_STORE = 0, EPI_ADD_RESID = 1, EPI_SWIGLU_PAIRS = 2, EPI_ROPE_KV = 3 };

// EPI_ROPE_KV (cluster split-K epilogue of the fused [q;k;v] projection only): the arithmetic of k_rope_kv_rows --
// rotate the (even, odd) pairs of q and k with the row's position, q -> q_out [rows, heads * hd], k / v -> the cache
// at the row's slot -- without the round trip of the projection through global memory and without the launch.
struct RopeKV {
  const bf16* rope;
  const int *row_stream, *row_pos, *row_slot;  // null: stream n % imp_B, position = slot = imp_pos + n / imp_B
  int imp_B, imp_pos, heads, kv_heads, hd, slots;
  bf16 *q_out, *k_cache, *v_cache;
};

struct Args {
  bf16* out;
  long long ldo;
  const bf16* resid;  // EPI_ADD_RESID: [rows, ldo] (may alias out)
  int rows, n_out, K;
  int epi;
  // split-K (gridDim.z > 1)
  float* part;             // [splits][m_tiles * BM][ldp] fp32 partial tiles
  unsigned int* counters;  // [m_tiles * n_tiles], zero between launches (the last CTA of a tile resets its counter)
  long long ldp;           // n_tiles * BN
  int cluster;             // gridDim.z > 1 with the z extent launched as ONE thread-block cluster: slices meet through DSMEM
  RopeKV rk;               // EPI_ROPE_KV
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  for (unsigned spin = 0; !mbar_try(b, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(const void* p) {
  uint64_t d = 0;
  d |= (uint64_t)((s32(p) & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, N and M of the tile
__device__ __forceinline__ uint32_t instr_desc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of the accumulator -> registers (one row per lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// thread-block cluster helpers (split-K through distributed shared memory)
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();  // (.aligned: the warp must be converged; the TMA / MMA warps come here from single-lane branches)
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(s32(local)), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float4 dsmem_ld4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
constexpr int CK_LD = BN + 4;  // padded fp32 row of a slice's partial tile in shared memory

// EPI_ROPE_KV for four consecutive columns n .. n+3 (two (even, odd) pairs) of output row rr; y = the bf16-rounded
// projection values.  Same arithmetic as k_rope_kv_rows / the skinny kernel's epilogue.
__device__ __forceinline__ void rope_kv_store4(const RopeKV& k, int rr, int n, int n_out, const float (&y)[4]) {
  const int hd = k.hd, qrows = k.heads * hd, krows = k.kv_heads * hd;
  const int pos = k.row_stream ? k.row_pos[rr] : k.imp_pos + rr / k.imp_B;
  const int slot = k.row_stream ? k.row_slot[rr] : k.imp_pos + rr / k.imp_B;
  const int stream = k.row_stream ? k.row_stream[rr] : rr % k.imp_B;
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    const int r0 = n + 2 * pr;
    if (r0 >= n_out) break;  // (n_out is even: whole pairs)
    const float y0 = y[2 * pr], y1 = y[2 * pr + 1];
    float o0 = y0, o1 = y1;
    if (r0 < qrows + krows) {
      const __nv_bfloat162 cs = *reinterpret_cast<const __nv_bfloat162*>(k.rope + ((size_t)pos * (hd / 2) + ((r0 % hd) >> 1)) * 2);
      const float c = __low2float(cs), sn = __high2float(cs);
      o0 = rbf(__fsub_rn(__fmul_rn(y0, c), __fmul_rn(y1, sn)));
      o1 = rbf(__fadd_rn(__fmul_rn(y1, c), __fmul_rn(y0, sn)));
    }
    const __nv_bfloat162 ov = __floats2bfloat162_rn(o0, o1);
    if (r0 < qrows) {
      *reinterpret_cast<__nv_bfloat162*>(k.q_out + (size_t)rr * qrows + r0) = ov;
    } else {
      const bool isk = r0 < qrows + krows;
      const int rl = r0 - (isk ? qrows : qrows + krows);
      const int kvh = rl / hd, d = rl % hd;
      bf16* dst = (isk ? k.k_cache : k.v_cache) + (((size_t)stream * k.kv_heads + kvh) * k.slots + slot) * hd + d;
      *reinterpret_cast<__nv_bfloat162*>(dst) = ov;
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int splits = gridDim.z;
  const int num_kb = a.K / BK / splits;        // k blocks of this CTA's slice (the launcher keeps the division exact)
  const int kb0 = blockIdx.z * num_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {  // one warp allocates the accumulator columns (and frees them at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      pdl_trigger();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + BM * BK * 2;
        mbar_expect(&full[s], STAGE_BYTES);
        tma_load_2d(sa, &map_x, &full[s], (kb0 + kb) * BK, m0);
        tma_load_2d(sb, &map_w, &full[s], (kb0 + kb) * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        const unsigned char* sb = sa + BM * BK * 2;
        const uint64_t ad = smem_desc(sa), bd = smem_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)  // +32 bytes along K inside the 128-byte swizzle atom
          umma(tmem_base, ad + (uint64_t)(k * UMMA_K * 2 >> 4), bd + (uint64_t)(k * UMMA_K * 2 >> 4), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
      }
      umma_commit(acc_full);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32) -> output rows m0 + that range
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    pdl_wait();
    pdl_trigger();
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    bool finish = true;  // this CTA runs the epilogue (always, without split-K)
    if (a.cluster) {
      // Cluster split-K: the z extent of the grid is one thread-block cluster, CTA z accumulated k slice z.  Every
      // CTA parks its fp32 partial tile in its own shared memory (the operand ring is drained: all MMAs have
      // completed), the cluster meets, and CTA z adds rows [z * 128 / splits, ...) of all slices in slice order
      // through distributed shared memory -- no global partials, no counters, no last-CTA pass.
      float* ptile = reinterpret_cast<float*>(smem);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        float* trow = ptile + (size_t)(q * 32 + lane) * CK_LD + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(trow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      finish = false;
    } else if (splits > 1) {
      // 1. park the fp32 partial tile (one 128-byte line per lane and step)
      float* prow = a.part + ((size_t)blockIdx.z * gridDim.y * BM + (m0 - 0) + q * 32 + lane) * a.ldp + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(prow + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      // 2. count this slice in; the CTA that completes the tile adds the slices up
      __threadfence();
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
      __shared__ unsigned int s_last;
      if (warp == 2 && lane == 0) {
        unsigned int* ctr = a.counters + blockIdx.y * gridDim.x + blockIdx.x;
        const unsigned int seen = atomicAdd(ctr, 1u);
        s_last = seen == (unsigned)splits - 1;
        if (s_last) *ctr = 0;  // ready for the next launch on this stream
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      finish = s_last != 0;
      if (finish) __threadfence();
    }
    if (finish && splits > 1) {
      // The tile's last CTA adds the slices in order 0 .. splits-1 (deterministic whoever finishes) and runs the
      // epilogue in a COALESCED layout: a warp takes whole rows, lane l the four columns 4l .. 4l+3, four rows in
      // flight -- the loads of all slices of a row batch go out together (the row-per-lane layout of the TMEM
      // path would make this pass a chain of dependent L2 round trips).
      const int et = threadIdx.x - 64;  // 0 .. 127 among the epilogue warps
      const int ew = et >> 5;
      const size_t zs = (size_t)gridDim.y * BM * a.ldp;
      const int n = n0 + lane * 4;
#pragma unroll 1
      for (int r0 = ew * 32; r0 < ew * 32 + 32; r0 += 4) {
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = __ldcg(reinterpret_cast<const float4*>(a.part + ((size_t)m0 + r0 + i) * a.ldp + n));
        for (int z = 1; z < splits; ++z) {
          float4 t[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) t[i] = __ldcg(reinterpret_cast<const float4*>(a.part + z * zs + ((size_t)m0 + r0 + i) * a.ldp + n));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i].x += t[i].x; acc[i].y += t[i].y; acc[i].z += t[i].z; acc[i].w += t[i].w;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = m0 + r0 + i;
          if (rr >= a.rows) continue;
          const float y[4] = {rbf(acc[i].x), rbf(acc[i].y), rbf(acc[i].z), rbf(acc[i].w)};
          if (a.epi == EPI_SWIGLU_PAIRS) {
            bf16* o = a.out + (long long)rr * a.ldo + (n >> 1);
            if (n + 1 < a.n_out) o[0] = f2bf(silu_bf(y[0]) * y[1]);
            if (n + 3 < a.n_out) o[1] = f2bf(silu_bf(y[2]) * y[3]);
          } else {
            bf16* o = a.out + (long long)rr * a.ldo + n;
            const bf16* r = a.resid + (long long)rr * a.ldo + n;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < a.n_out) o[j] = f2bf(a.epi == EPI_ADD_RESID ? y[j] + bf2f(r[j]) : y[j]);
          }
        }
      }
    } else if (finish) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      }
      if (row < a.rows) {
        const int n = n0 + c0;
        if (a.epi == EPI_SWIGLU_PAIRS) {
          // columns (2i, 2i+1) = (gate_i, up_i): bf16( bf16(silu(bf16 gate)) * bf16 up )
          bf16* o = a.out + (long long)row * a.ldo + (n >> 1);
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            if (n + j + 1 < a.n_out) o[j >> 1] = f2bf(silu_bf(rbf(__uint_as_float(v[j]))) * rbf(__uint_as_float(v[j + 1])));
        } else {
          bf16* o = a.out + (long long)row * a.ldo + n;
          const bf16* r = a.resid + (long long)row * a.ldo + n;
          if (n + 32 <= a.n_out && (a.ldo & 7) == 0) {  // 4 x 16-byte stores
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8) {
              uint4 rv = make_uint4(0, 0, 0, 0);
              if (a.epi == EPI_ADD_RESID) rv = *reinterpret_cast<const uint4*>(r + j8);
              const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
              __nv_bfloat162 pk[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float y0 = rbf(__uint_as_float(v[j8 + 2 * t])), y1 = rbf(__uint_as_float(v[j8 + 2 * t + 1]));
                if (a.epi == EPI_ADD_RESID) {
                  y0 += bflo(rr[t]);
                  y1 += bfhi(rr[t]);
                }
                pk[t] = __floats2bfloat162_rn(y0, y1);
              }
              *reinterpret_cast<uint4*>(o + j8) = *reinterpret_cast<uint4*>(pk);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (n + j >= a.n_out) break;
              float y = rbf(__uint_as_float(v[j]));
              if (a.epi == EPI_ADD_RESID) y += bf2f(r[j]);
              o[j] = f2bf(y);
            }
          }
        }
      }
    }
    }
  }
  if (a.cluster) {
    cluster_sync_all();  // every slice's partial tile is in its CTA's shared memory
    if (warp >= 2) {
      const float* ptile = reinterpret_cast<const float*>(smem);
      const int et = threadIdx.x - 64;  // 0 .. 127
      const int rows_per = BM / splits;
      const int n = n0 + (et & 31) * 4;
#pragma unroll 1
      for (int r = (int)blockIdx.z * rows_per + (et >> 5); r < ((int)blockIdx.z + 1) * rows_per; r += 4) {
        const float* lp = ptile + (size_t)r * CK_LD + (et & 31) * 4;  // the same offset in every CTA of the cluster
        float4 t[8];  // all slices' loads go out together (a DSMEM round trip each), then they are added in slice order
#pragma unroll
        for (int z = 0; z < 8; ++z)
          if (z < splits) t[z] = dsmem_ld4(dsmem_addr(lp, (uint32_t)z));
        float4 acc = t[0];
#pragma unroll
        for (int z = 1; z < 8; ++z)
          if (z < splits) { acc.x += t[z].x; acc.y += t[z].y; acc.z += t[z].z; acc.w += t[z].w; }
        const int rr = m0 + r;
        if (rr >= a.rows) continue;
        const float y[4] = {rbf(acc.x), rbf(acc.y), rbf(acc.z), rbf(acc.w)};
        if (a.epi == EPI_ROPE_KV) {
          rope_kv_store4(a.rk, rr, n, a.n_out, y);
        } else if (a.epi == EPI_SWIGLU_PAIRS) {
          bf16* o = a.out + (long long)rr * a.ldo + (n >> 1);
          if (n + 1 < a.n_out) o[0] = f2bf(silu_bf(y[0]) * y[1]);
          if (n + 3 < a.n_out) o[1] = f2bf(silu_bf(y[2]) * y[3]);
        } else {
          bf16* o = a.out + (long long)rr * a.ldo + n;
          const bf16* rs = a.resid + (long long)rr * a.ldo + n;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < a.n_out) o[j] = f2bf(a.epi == EPI_ADD_RESID ? y[j] + bf2f(rs[j]) : y[j]);
        }
      }
    }
    cluster_sync_all();  // nobody leaves (or frees its shared memory) while a peer still reads it
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// ---- persistent variant (round 2): prompt prefill and every GEMM that is not split-K --------------------------
// One CTA per SM walks the output tiles (n tile fastest: the CTAs at work share their X rows in L2).  TMA warp, MMA
// lane and EIGHT epilogue warps run concurrently across tiles: the accumulator is double buffered in TMEM
// (2 x 128 columns; acc_full / acc_empty barriers), the shared-memory ring keeps counting k blocks across tiles,
// so the loads and MMAs of tile i+1 run under the epilogue of tile i.  The epilogue drains the accumulator row-per-
// lane into a shared-memory tile and then walks it in row-major order, four columns per lane: residual loads and
// bf16 stores are coalesced (256 bytes per warp and row) instead of 32 lines per store instruction.
// BNT = 256 (prompt-sized GEMMs): a 128 x 256 tile needs 48 KB of operands per 4.2 MFLOP instead of 2 x 32 KB -- the
// 128 x 128 tile is bound by the L2 -> shared-memory path at 44-47 % tensor activity.  Three 48 KB stages, both
// accumulators fill the 512 TMEM columns, the epilogue drains a tile in two 128-column halves through the same
// staging tile.
constexpr int P_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int P_TLD = 128 + 4;
constexpr size_t P_TILE_BYTES = (size_t)BM * P_TLD * 4;
template <int BNT> struct PCfg {
  static constexpr int STAGES = BNT == 128 ? 4 : 3;
  static constexpr uint32_t STAGE = (BM + BNT) * BK * 2;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE + P_TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};
constexpr size_t P_SMEM_BYTES = PCfg<128>::SMEM;

template <int BNT>
__global__ void __launch_bounds__(P_THREADS, 1)
k_gemm_tc_p(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, Args a) {
  constexpr int P_STAGES = PCfg<BNT>::STAGES;
  constexpr uint32_t STAGE_BYTES = PCfg<BNT>::STAGE;  // (shadows the 128-wide constant of the one-tile kernel)
  constexpr int BN = BNT;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* tile = reinterpret_cast<float*>(smem + (size_t)P_STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)P_STAGES * STAGE_BYTES + P_TILE_BYTES);
  uint64_t* empty = full + P_STAGES;
  uint64_t* acc_full = empty + P_STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = a.K / BK;
  const int n_tiles = (a.n_out + BN - 1) / BN;
  const int m_tiles = (a.rows + BM - 1) / BM;
  const long long tiles = (long long)m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // Programmatic dependent launch: the W boxes of the first ring pass do not depend on the previous kernel --
      // they are requested before pdl_wait(), the X boxes of the same stages after it (both count on full[s]).
      // (Tile coordinates advance incrementally: a division per k block in this one thread starves the ring.)
      const long long mine = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const unsigned total = (unsigned)(mine * num_kb);
      const unsigned pre = total < (unsigned)P_STAGES ? total : (unsigned)P_STAGES;
      auto owed_x = [&]() {  // once: wait for the previous kernel, then the X boxes of the stages filled so far
        pdl_wait();
        pdl_trigger();
        for (unsigned j = 0; j < pre; ++j) {
          const long long t = blockIdx.x + (long long)(j / num_kb) * gridDim.x;
          tma_load_2d(smem + (size_t)j * STAGE_BYTES, &map_x, &full[j], (int)(j % num_kb) * BK, (int)(t / n_tiles) * BM);
        }
      };
      unsigned kc = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int n0 = (int)(t % n_tiles) * BN, m0 = (int)(t / n_tiles) * BM;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          if (kc == pre) owed_x();
          const int s = kc % P_STAGES;
          mbar_wait(&empty[s], ((kc / P_STAGES) & 1) ^ 1);
          unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
          unsigned char* sb = sa + BM * BK * 2;
          mbar_expect(&full[s], STAGE_BYTES);
          if (kc >= pre) tma_load_2d(sa, &map_x, &full[s], kb * BK, m0);
          tma_load_2d(sb, &map_w, &full[s], kb * BK, n0);
        }
      }
      if (total <= pre) owed_x();  // fewer k blocks than stages
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(BM, BN);
      unsigned kc = 0, it = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const unsigned buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % P_STAGES;
          mbar_wait(&full[s], (kc / P_STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
          const unsigned char* sb = sa + BM * BK * 2;
          const uint64_t ad = smem_desc(sa), bd = smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma(acc, ad + (uint64_t)(k * UMMA_K * 2 >> 4), bd + (uint64_t)(k * UMMA_K * 2 >> 4), idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int q = warp & 3;            // the TMEM lane quarter is tied to the warp's index in the CTA
    const int half = (warp - 2) >> 2;  // warps 2..5 drain columns [0, 64), warps 6..9 columns [64, 128)
    const int et = threadIdx.x - 64;   // 0 .. 255
    const int c = (et & 31) * 4;       // this thread's four columns of the tile
    unsigned it = 0;
    pdl_wait();
    pdl_trigger();
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int n0 = (int)(t % n_tiles) * BN, m0 = (int)(t / n_tiles) * BM;
      const unsigned buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int mrows = a.rows - m0 < BM ? a.rows - m0 : BM;
#pragma unroll 1
      for (int hcol = 0; hcol < BN; hcol += 128) {  // 128 accumulator columns at a time through the staging tile
      if (hcol) asm volatile("bar.sync 2, 256;" ::: "memory");  // the previous half has left the staging tile
#pragma unroll 1
      for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + buf * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(hcol + c0), v);
        float* trow = tile + (size_t)(q * 32 + lane) * P_TLD + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(trow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (hcol + 128 >= BN) {  // the accumulator is drained: hand it back to the MMA lane
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&acc_empty[buf])) : "memory");
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the half tile is complete in shared memory
      const int n = n0 + hcol + c;
      const bool full4 = n + 4 <= a.n_out;
      const bool vec = full4 && (a.ldo & 3) == 0;  // 8-byte stores / residual loads
      if (n < a.n_out) {
#pragma unroll 1
        for (int r0 = et >> 5; r0 < mrows; r0 += 4 * 8) {
          uint2 rr[4];
          if (a.epi == EPI_ADD_RESID && vec) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * 8;
              if (r < mrows) rr[u] = *reinterpret_cast<const uint2*>(a.resid + (long long)(m0 + r) * a.ldo + n);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * 8;
            if (r >= mrows) break;
            const long long row = m0 + r;
            const float4 tv = *reinterpret_cast<const float4*>(tile + (size_t)r * P_TLD + c);
            const float y[4] = {rbf(tv.x), rbf(tv.y), rbf(tv.z), rbf(tv.w)};
            if (a.epi == EPI_ROPE_KV) {
              rope_kv_store4(a.rk, (int)row, n, a.n_out, y);
            } else if (a.epi == EPI_SWIGLU_PAIRS) {
              // columns (2i, 2i+1) = (gate_i, up_i): bf16( bf16(silu(bf16 gate)) * bf16 up )
              bf16* o = a.out + row * a.ldo + (n >> 1);
              if (full4) {
                *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(silu_bf(y[0]) * y[1], silu_bf(y[2]) * y[3]);
              } else if (n + 1 < a.n_out) {
                o[0] = f2bf(silu_bf(y[0]) * y[1]);
              }
            } else if (vec) {
              float z[4] = {y[0], y[1], y[2], y[3]};
              if (a.epi == EPI_ADD_RESID) {
                z[0] += bflo(rr[u].x); z[1] += bfhi(rr[u].x); z[2] += bflo(rr[u].y); z[3] += bfhi(rr[u].y);
              }
              __nv_bfloat162 pk[2] = {__floats2bfloat162_rn(z[0], z[1]), __floats2bfloat162_rn(z[2], z[3])};
              *reinterpret_cast<uint2*>(a.out + row * a.ldo + n) = *reinterpret_cast<uint2*>(pk);
            } else {
              for (int j = 0; j < 4 && n + j < a.n_out; ++j) {
                float z = y[j];
                if (a.epi == EPI_ADD_RESID) z += bf2f(a.resid[row * a.ldo + n + j]);
                a.out[row * a.ldo + n + j] = f2bf(z);
              }
            }
          }
        }
      }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the staging tile is free for the next tile
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

}  // namespace tc
