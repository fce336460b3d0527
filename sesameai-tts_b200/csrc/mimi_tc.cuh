// Mimi decode GEMMs on the 5th-generation tensor cores: C[M, N] = A[M, K] . B[N, K]^T in TF32 (fp32 operands
// read by tcgen05.mma.kind::tf32, fp32 accumulation in TMEM), for the causal convolutions / transposed
// convolutions of the SEANet decoder and the linear layers of the Mimi transformer (moshi MimiModel.decode,
// reference sesameai/generator.py:116,299).
//
//   warp 0      TMA producer : cp.async.bulk.tensor.3d of a 128 x 32 fp32 A tile and .2d of a 128 x 32 B tile per
//                              stage (128-byte swizzle) into a 4-deep shared-memory ring.  The A map is THREE
//                              dimensional {channel, tap, row} with the tap stride equal to the row stride: row m
//                              of the implicit im2col matrix is the contiguous slice x[m .. m + taps - 1] of the
//                              time-major activation, so a causal Conv1d(k) / ConvTranspose1d(2s, s) is a plain GEMM
//                              over OVERLAPPING rows and no im2col buffer exists (rows beyond M read as zero)
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x N128 x K8, four
//                              per stage); tcgen05.commit releases the stage / signals the epilogue
//   warps 2..9  epilogue     : tcgen05.ld -> bias / GELU / LayerScale / residual -> fp32 stores; optionally the
//                              ELU of the result (rounded to TF32) as a second output, because the consumer of a
//                              SEANet activation applies ELU to its input and a TMA-fed operand cannot be touched
//                              on its way into the tensor core
// Persistent: one CTA per SM walks the tiles; two accumulators in TMEM (2 x 128 columns) let the epilogue of one
// tile run under the loads and MMAs of the next.
// Operands are rounded to TF32 (round-to-nearest) where they are PRODUCED (weights at pack time, activations in
// the producing epilogue), so the tensor core's truncation of the low mantissa bits is exact.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mtc {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 4;
constexpr int UMMA_K = 8;
constexpr int THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr uint32_t STAGE_BYTES = (BM + BN) * BK * 4;
constexpr int TLD = BN + 4;                                    // padded row of the epilogue staging tile
constexpr size_t TILE_BYTES = (size_t)BM * TLD * 4;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { F_GELU = 2, F_RESID = 4, F_LAYERSCALE = 8, F_OUT_ELU = 16, F_ROUND = 32 };

struct Args {
  float* C;  // [M, ldc]: the result (F_OUT_ELU: its ELU, TF32-rounded; F_ROUND: TF32-rounded), or null
  long long ldc;
  float* C2;  // [M, ldc2] or null: TF32-rounded ELU of the result in addition to C
  long long ldc2;
  int M, N, K;
  const float* bias;  // bias[n % bias_period] or null
  int bias_period;
  const float* R;  // residual [M, ldr]
  long long ldr;
  const float* scale;  // LayerScale [N]
  int flags;
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(const void* p) {
  uint64_t d = 0;
  d |= (uint64_t)((s32(p) & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same for an operand that starts ``rows`` rows into a 1024-byte-aligned swizzled block: only the start address
// moves (rows * 128 bytes).  The 128-byte swizzle is a function of the shared-memory ADDRESS (chunk index xor
// address bits 7-9), the same function TMA applied when it wrote the block, so a row-shifted view reads the right
// chunks.  Measured on B200: with the descriptor's base-offset field set to the row phase the product is wrong
// (SNR 3 dB), with the field left 0 it is exact (profiles/r2_mimi.txt).
__device__ __forceinline__ uint64_t smem_desc_rows(const void* block, int rows) {
  return smem_desc(reinterpret_cast<const unsigned char*>(block) + rows * 128);
}
// instruction descriptor: D fp32, A / B TF32 (format 2), both K-major
__device__ __forceinline__ uint32_t instr_desc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// (the result is rounded to TF32's 10 mantissa bits right away: exp(x) - 1 from the fast exponential, absolute
// error ~1e-7, is far inside that)
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : __expf(x) - 1.f; }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// cin: channels per tap of the A operand (K = taps * cin, cin % 32 == 0).
// PERSISTENT: one CTA per SM walks the output tiles t = blockIdx.x, + gridDim.x, ... (n tile fastest, so the CTAs
// working at the same time share their A rows in L2).  The three roles run concurrently across tiles: the TMA warp
// is already loading tile i+1 while the MMA lane works on tile i and the epilogue warps drain tile i-1 -- the
// accumulator is double buffered in TMEM (2 x 128 columns, acc_full / acc_empty barriers), the shared-memory ring
// simply keeps counting k blocks across tiles.  The SEANet tail is 11 250 tiles of 6 (or 1) k blocks each: with
// one tile per CTA the fixed cost of a CTA (barrier init, TMEM allocation, pipeline fill, drain) was the whole run time.
__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Args a, int cin) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* tile = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);  // [BM][TLD] epilogue staging
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES + TILE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = a.K / BK;
  const int n_tiles = (a.N + BN - 1) / BN;
  const long long m_tiles = ((long long)a.M + BM - 1) / BM;
  const long long tiles = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
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
      unsigned kc = 0;  // k blocks issued so far, over all tiles
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int n0 = (int)(t % n_tiles) * BN;
        const int m0 = (int)(t / n_tiles) * BM;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % STAGES;
          mbar_wait(&empty[s], ((kc / STAGES) & 1) ^ 1);
          unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
          unsigned char* sb = sa + BM * BK * 4;
          mbar_expect(&full[s], STAGE_BYTES);
          const int k = kb * BK;
          tma_load_3d(sa, &map_a, &full[s], k % cin, k / cin, m0);
          tma_load_2d(sb, &map_b, &full[s], k, n0);
        }
      }
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
          const int s = kc % STAGES;
          mbar_wait(&full[s], (kc / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
          const unsigned char* sb = sa + BM * BK * 4;
          const uint64_t ad = smem_desc(sa), bd = smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)  // +32 bytes along K inside the 128-byte swizzle atom
            umma_tf32(acc, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // epilogue (8 warps), two steps per tile.  (1) warp w may touch TMEM lanes [32*(w%4), +32): the two warps of a
    // lane quarter split the columns, each lane drains ITS row into the shared-memory tile and the accumulator goes
    // back to the MMA lane.  (2) the warps walk the tile in row-major order, four consecutive columns per lane:
    // bias / GELU / LayerScale / residual and the stores are COALESCED -- with one row per lane a 32- or 64-column
    // output (the SEANet tail) made every store instruction touch 32 different lines.  The valid width of a tile is
    // 32, 64 or 128 (launcher), so a thread keeps its column for the whole tile: no division in the loop, bias and
    // LayerScale are loaded once per tile.  (With four warps and a division per element this step, not HBM, was the
    // run time of the persistent kernel: one warp per scheduler hides no latency.)
    const int q = warp & 3;           // the TMEM lane quarter is tied to the warp's index in the CTA
    const int half = (warp - 2) >> 2; // warps 2..5 take the first half of the columns, 6..9 the second
    const int et = threadIdx.x - 64;  // 0 .. 255
    unsigned it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int n0 = (int)(t % n_tiles) * BN;
      const long long m0 = (t / n_tiles) * BM;
      const unsigned buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nvalid = a.N - n0 < BN ? a.N - n0 : BN;  // 32, 64 or 128
      const int cper = nvalid > 32 ? nvalid >> 1 : 32;   // columns each of the quarter's two warps drains
#pragma unroll 1
      for (int c0 = half * cper; c0 < (half + 1) * cper && c0 < nvalid; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* trow = tile + (size_t)(q * 32 + lane) * TLD + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(trow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      // this warp is done with the accumulator: hand it back (the MMA lane may start tile it + 2 in it)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&acc_empty[buf])) : "memory");
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the tile is complete in shared memory
      const int csh = nvalid == 128 ? 5 : (nvalid == 64 ? 4 : 3);  // log2(float4 columns per row)
      const int c = (et & ((1 << csh) - 1)) * 4;                   // this thread's column for the whole tile
      const int rstep = 256 >> csh;                                // rows covered by the 256 threads per round
      const int mrows = a.M - m0 < BM ? (int)(a.M - m0) : BM;
      const int n = n0 + c;
      float bv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {1.f, 1.f, 1.f, 1.f};
      if (a.bias) {
        const int bi = n % a.bias_period;  // the period is a multiple of 4: the four columns stay inside it
        bv[0] = a.bias[bi]; bv[1] = a.bias[bi + 1]; bv[2] = a.bias[bi + 2]; bv[3] = a.bias[bi + 3];
      }
      if (a.flags & F_LAYERSCALE) {
        sv[0] = a.scale[n]; sv[1] = a.scale[n + 1]; sv[2] = a.scale[n + 2]; sv[3] = a.scale[n + 3];
      }
#pragma unroll 1
      for (int r0 = et >> csh; r0 < mrows; r0 += 4 * rstep) {
        // four rows per thread per round: the residual loads of the round go out together
        float4 rr[4];
        if (a.flags & F_RESID) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * rstep;
            if (r < mrows) rr[u] = *reinterpret_cast<const float4*>(a.R + (m0 + r) * a.ldr + n);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * rstep;
          if (r >= mrows) break;
          const long long row = m0 + r;
          const float4 tv = *reinterpret_cast<const float4*>(tile + (size_t)r * TLD + c);
          float y[4] = {tv.x + bv[0], tv.y + bv[1], tv.z + bv[2], tv.w + bv[3]}, e[4];
          if (a.flags & F_GELU) {
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = gelu_f(y[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) y[k] *= sv[k];
          if (a.flags & F_RESID) {
            y[0] += rr[u].x; y[1] += rr[u].y; y[2] += rr[u].z; y[3] += rr[u].w;
          }
          if (a.C2 || (a.flags & F_OUT_ELU)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] = round_tf32(elu_f(y[k]));
          }
          if (a.C2) *reinterpret_cast<float4*>(a.C2 + row * a.ldc2 + n) = make_float4(e[0], e[1], e[2], e[3]);
          if (a.C) {
            float4 o;
            if (a.flags & F_OUT_ELU) o = make_float4(e[0], e[1], e[2], e[3]);
            else if (a.flags & F_ROUND) o = make_float4(round_tf32(y[0]), round_tf32(y[1]), round_tf32(y[2]), round_tf32(y[3]));
            else o = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(a.C + row * a.ldc + n) = o;
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

// ---- weight-resident variant: the residual-block convolutions of the SEANet tail ---------------------------------
// Those GEMMs are one n tile wide (N = 32 / 64 / 128) over a small weight matrix (K x N fp32 <= 96 KB) and 10^4
// m tiles.  In k_gemm_tf32 every tile re-loads the weights as a full 128-row B box (zero filled above N) next to
// its A boxes: half of the ring's bytes in flight -- the resource that bounds these launches -- carry weights, and
// the MMA is 128 columns wide whatever N.  Here
//   * the weights are loaded ONCE per CTA into their own shared-memory region, the ring (up to 8 stages) carries
//     only A boxes, and the MMA is exactly N wide;
//   * the residual operand of a tile (128 x N fp32, 32 KB at N = 64) comes through TMA as well, into a ring of its
//     own that the producer fills tiles ahead: loaded by the epilogue threads it was 16 KB in flight per SM behind
//     a dependent HBM round trip per tile, and that -- not the tensor core, not the A stream -- was the run time;
//   * tap-shift mode (Conv1d k = 3; ra.taps > 1): the taps of a conv read the SAME activation rows, shifted by one
//     row each.  Instead of one A box per (tap, 32 channels) -- every row crossing L2 -> shared memory three times --
//     a stage holds the 128 + taps - 1 rows of a 32-channel block once (map_a is then a plain 2-d map of the
//     activation) and the MMAs of tap t read it through a descriptor that starts t rows in (smem_desc_rows);
//   * fused final convolution (stage-3 residual conv, N = 64; fw != null): the decoder ends in Conv1d(64 -> 1,
//     k = 3) over ELU(y).  While a thread holds four ELU'd columns of a row it forms their products with the three
//     taps, the 16 lanes of the row add them up (xor shuffles, fixed order), and sample t of the waveform is
//     (P0[t-2] + P1[t-1]) + (P2[t] + bias) of those per-row sums: the 368 MB activation of a 60 s utterance is
//     neither written nor read back.  The first two samples of a tile need rows of the previous tile: both tiles
//     atomicAdd their part onto the zero-initialised output -- two addends, so the sum does not depend on their
//     order.  Rows -2 / -1 of the first tile come from ``halo`` (the previous chunk of a streamed decode, zeros
//     otherwise) through the same arithmetic, the last two ELU'd rows go to ``tail`` for the next chunk; tiles
//     start at multiples of 128 samples and a frame is 1920 = 15 * 128 samples, so streamed and one-shot decodes
//     compute every sample by the same expression.
constexpr int R_MAX_STAGES = 8, R_MAX_RES = 3;
constexpr int R_EPI_WARPS = 16, R_EPI_THREADS = R_EPI_WARPS * 32, R_THREADS = 64 + R_EPI_THREADS;  // TMA warp, MMA warp, epilogue
constexpr uint32_t A_STAGE_BYTES = BM * BK * 4;

struct RArgs {
  Args e;             // N = the whole output width; C2 / GELU / LayerScale are not supported here
  int stages;         // depth of the A ring
  int nres;           // depth of the residual ring (0: no residual)
  int taps;           // > 1: tap-shift mode (see k_gemm_tf32_r) with this many taps; else 0
  const float* fw;    // fused final conv: weights [3][64] tap-major, or null (needs N == 64, M % 128 == 0)
  float fb;           // its bias
  float* wav;         // [M], zero-initialised
  const float* halo;  // [2][64] ELU'd rows -2, -1
  float* tail;        // [2][64] receives ELU'd rows M-2, M-1 (or null)
};

// tap-shift stage: 128 + taps - 1 (<= 136) rows of 32 channels
constexpr uint32_t A_SHIFT_STAGE_BYTES = 136 * BK * 4;
__host__ __device__ constexpr size_t r_smem_bytes(int K, int N, int stages, int nres, bool shift = false) {
  return (size_t)K * N * 4 + (size_t)stages * (shift ? A_SHIFT_STAGE_BYTES : A_STAGE_BYTES) + (size_t)BM * (N + 4) * 4 + (size_t)nres * BM * N * 4 +
         3 * BM * 4 /*row sums of the fused conv*/ + 1024 /*align*/ + 256 /*barriers*/;
}

// products of four ELU'd columns of a row with the three taps of the final conv, summed over the row's 16 lanes
__device__ __forceinline__ void final_row_sums(const float4& o, const float (&fw)[3][4], float (&p)[3]) {
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    float v = o.x * fw[t][0];
    v = fmaf(o.y, fw[t][1], v);
    v = fmaf(o.z, fw[t][2], v);
    v = fmaf(o.w, fw[t][3], v);
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    p[t] = v;
  }
}

// NB: output width; RESID: residual ring; FUSED: final conv in the epilogue (compile-time: the epilogue of the
// run-time-flag version spent ~45 % of its instructions on integer / predicate bookkeeping)
template <int NB, bool RESID, bool FUSED>
__global__ void __launch_bounds__(R_THREADS, 1)
k_gemm_tf32_r(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_r, RArgs ra, int cin) {
  const Args& a = ra.e;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int tld = NB + 4;
  const int stages = ra.stages, nres = RESID ? ra.nres : 0;
  const int taps = ra.taps > 1 ? ra.taps : 0;
  const uint32_t a_stage = taps ? A_SHIFT_STAGE_BYTES : A_STAGE_BYTES;
  const int num_ab = taps ? cin / BK : a.K / BK;  // A stages per tile
  const int num_kb = a.K / BK;
  constexpr uint32_t b_kb_bytes = (uint32_t)NB * BK * 4;  // one k block of the weights: NB rows of 128 bytes
  constexpr uint32_t r_bytes = (uint32_t)BM * NB * 4;     // one residual tile, dense rows
  unsigned char* bres = smem;
  unsigned char* ring = bres + (size_t)num_kb * b_kb_bytes;
  float* tile = reinterpret_cast<float*>(ring + (size_t)stages * a_stage);
  unsigned char* rring = reinterpret_cast<unsigned char*>(tile) + (size_t)BM * tld * 4;
  float* psum = reinterpret_cast<float*>(rring + (size_t)nres * r_bytes);  // [3][BM]
  uint64_t* full = reinterpret_cast<uint64_t*>(psum + 3 * BM);
  uint64_t* empty = full + R_MAX_STAGES;
  uint64_t* acc_full = empty + R_MAX_STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* b_full = acc_empty + 2;
  uint64_t* r_full = b_full + 1;              // [R_MAX_RES]
  uint64_t* r_empty = r_full + R_MAX_RES;     // [R_MAX_RES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + R_MAX_RES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles = ((long long)a.M + BM - 1) / BM;
  constexpr uint32_t tmem_cols = NB == 128 ? 256u : (NB == 64 ? 128u : 64u);  // two accumulators, a power of two

  if (threadIdx.x == 0) {
    for (int s = 0; s < R_MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], R_EPI_WARPS);
    }
    for (int i = 0; i < R_MAX_RES; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], R_EPI_WARPS);
    }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (RESID) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect(b_full, (uint32_t)num_kb * b_kb_bytes);
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(bres + (size_t)kb * b_kb_bytes, &map_b, b_full, kb * BK, 0);
      int s = 0, rs = 0;
      uint32_t ph = 1, rph = 1;  // parities that pass on fresh "empty" barriers
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (int)(t * BM);
        if (RESID) {  // the tile's residual rows, ahead of its A boxes
          mbar_wait(&r_empty[rs], rph);
          mbar_expect(&r_full[rs], r_bytes);
          tma_load_2d(rring + (size_t)rs * r_bytes, &map_r, &r_full[rs], 0, m0);
          if (++rs == nres) { rs = 0; rph ^= 1; }
        }
        for (int ab = 0; ab < num_ab; ++ab) {
          mbar_wait(&empty[s], ph);
          mbar_expect(&full[s], a_stage);
          const int k = ab * BK;
          if (taps) tma_load_2d(ring + (size_t)s * a_stage, &map_a, &full[s], k, m0);  // rows m0 .. m0 + 135 of channel block ab
          else tma_load_3d(ring + (size_t)s * a_stage, &map_a, &full[s], k % cin, k / cin, m0);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(BM, NB);
      mbar_wait(b_full, 0);
      int s = 0;
      uint32_t ph = 0;
      unsigned it = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const unsigned buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * NB;
        for (int ab = 0; ab < num_ab; ++ab) {
          mbar_wait(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned char* sa = ring + (size_t)s * a_stage;
          if (taps) {
            for (int t = 0; t < taps; ++t) {  // K index of the weights = tap * cin + channel
              const uint64_t ad = smem_desc_rows(sa, t), bd = smem_desc(bres + (size_t)(t * num_ab + ab) * b_kb_bytes);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_tf32(acc, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (ab | t | k) != 0);
            }
          } else {
            const uint64_t ad = smem_desc(sa), bd = smem_desc(bres + (size_t)ab * b_kb_bytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_tf32(acc, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (ab | k) != 0);
          }
          umma_commit(&empty[s]);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // SIXTEEN epilogue warps: the per-tile work is a chain of short dependent steps (TMEM load, staging, ELU,
    // shuffles) and with two warps per scheduler its latency, not its instruction count, was the tile time
    const int q = warp & 3;            // TMEM lane quarter of this warp
    const int part = (warp - 2) >> 2;  // which of the quarter's four warps: columns [part * NB / 4, + NB / 4)
    const int et = threadIdx.x - 64;   // 0 .. 511
    constexpr int cw = NB >> 2;        // 8, 16 or 32 columns per warp
    constexpr int csh = NB == 128 ? 5 : (NB == 64 ? 4 : 3);
    const int c = (et & ((1 << csh) - 1)) * 4;  // this thread's four columns, for every tile
    constexpr int rstep = R_EPI_THREADS >> csh, rround = 4 * rstep;
    const int rfirst = et >> csh;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias) {
      const int bi = c % a.bias_period;
      bv[0] = a.bias[bi]; bv[1] = a.bias[bi + 1]; bv[2] = a.bias[bi + 2]; bv[3] = a.bias[bi + 3];
    }
    float fw[3][4];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) fw[t][k] = FUSED ? ra.fw[t * 64 + c + k] : 0.f;
    unsigned it = 0;
    int rs = 0;
    uint32_t rph = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const long long m0 = t * BM;
      const unsigned buf = it & 1;
      const int mrows = a.M - m0 < BM ? (int)(a.M - m0) : BM;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const int c0 = part * cw;
        const uint32_t taddr = tmem_base + buf * NB + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        float* trow = tile + (size_t)(q * 32 + lane) * tld + c0;
        if constexpr (cw == 32) {
          uint32_t v[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(trow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else if constexpr (cw == 16) {
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(trow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          uint32_t v[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; j += 4) *reinterpret_cast<uint4*>(trow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&acc_empty[buf])) : "memory");
      asm volatile("bar.sync 2, 512;" ::: "memory");  // the tile is complete in shared memory
      const float* rbuf = reinterpret_cast<const float*>(rring + (size_t)rs * r_bytes);
      if (RESID) mbar_wait(&r_full[rs], rph);
      const bool last_tile = m0 + BM >= a.M;
#pragma unroll 1
      for (int r0 = rfirst; r0 < BM; r0 += rround) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * rstep;
          if (rstep * 4 > BM && r >= BM) break;           // (32-column tiles: 64 rows per step, two steps)
          if (!FUSED && mrows < BM && r >= mrows) break;  // partial last tile (the fused conv requires whole tiles)
          const long long row = m0 + r;
          const float4 tv = *reinterpret_cast<const float4*>(tile + (size_t)r * tld + c);
          float y[4] = {tv.x + bv[0], tv.y + bv[1], tv.z + bv[2], tv.w + bv[3]};
          if (RESID) {
            const float4 rv = *reinterpret_cast<const float4*>(rbuf + (size_t)r * NB + c);
            y[0] += rv.x; y[1] += rv.y; y[2] += rv.z; y[3] += rv.w;
          }
          float4 o;
          if (a.flags & F_OUT_ELU) o = make_float4(round_tf32(elu_f(y[0])), round_tf32(elu_f(y[1])), round_tf32(elu_f(y[2])), round_tf32(elu_f(y[3])));
          else if (a.flags & F_ROUND) o = make_float4(round_tf32(y[0]), round_tf32(y[1]), round_tf32(y[2]), round_tf32(y[3]));
          else o = make_float4(y[0], y[1], y[2], y[3]);
          if (!FUSED || a.C) *reinterpret_cast<float4*>(a.C + row * a.ldc + c) = o;
          if (FUSED) {
            float p[3];
            final_row_sums(o, fw, p);
            if ((et & 15) == 0) {
              psum[r] = p[0]; psum[BM + r] = p[1]; psum[2 * BM + r] = p[2];
            }
            if (last_tile && ra.tail && r >= BM - 2) *reinterpret_cast<float4*>(ra.tail + (r - (BM - 2)) * 64 + c) = o;
          }
        }
      }
      if (RESID) {  // this warp is done with the residual tile
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&r_empty[rs])) : "memory");
        if (++rs == nres) { rs = 0; rph ^= 1; }
      }
      if (FUSED) {
        asm volatile("bar.sync 2, 512;" ::: "memory");  // the row sums of the tile are complete
        const float* P0 = psum;
        const float* P1 = psum + BM;
        const float* P2 = psum + 2 * BM;
        if (et < BM) {
          const int j = et;
          if (j >= 2) ra.wav[m0 + j] = (P0[j - 2] + P1[j - 1]) + (P2[j] + ra.fb);
          else if (j == 1) atomicAdd(ra.wav + m0 + 1, P1[0] + (P2[1] + ra.fb));
          else atomicAdd(ra.wav + m0, P2[0] + ra.fb);
        } else if (et == BM) {  // this tile's last rows in the next tile's first two samples
          if (m0 + BM < a.M) atomicAdd(ra.wav + m0 + BM, P0[BM - 2] + P1[BM - 1]);
        } else if (et == BM + 1) {
          if (m0 + BM + 1 < a.M) atomicAdd(ra.wav + m0 + BM + 1, P0[BM - 1]);
        } else if (et >= 384 && et < 416 && m0 == 0) {
          // rows -2, -1 (the previous chunk's tail) through the same arithmetic: lanes 0..15 row -2, 16..31 row -1
          const float4 o = *reinterpret_cast<const float4*>(ra.halo + (lane >> 4) * 64 + c);
          float p[3];
          final_row_sums(o, fw, p);
          const float p1_m1 = __shfl_sync(0xffffffffu, p[1], 16);
          if (lane == 0) atomicAdd(ra.wav, p[0] + p1_m1);
          if (lane == 16) atomicAdd(ra.wav + 1, p[0]);
        }
      }
      asm volatile("bar.sync 2, 512;" ::: "memory");  // staging tile and row sums are free for the next tile
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// element-wise helpers of the tensor-core decode path
__global__ void k_round_tf32(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = round_tf32(x[i]);
}

}  // namespace mtc
