// Mimi decode GEMMs on the 5th-generation tensor cores: C[M, N] = A[M, K] . B[N, K]^T in TF32 (fp32 operands
// read by tcgen05.mma.kind::tf32, fp32 accumulation in TMEM), for the causal convolutions / transposed
// convolutions of the SEANet decoder and the linear layers of the Mimi transformer (moshi MimiModel.decode,
// reference sesameai/generator.py:116,299).
//
//   warp 0      TMA producer : cp.async.bulk.tensor.3d of a 128 x 32 fp32 A tile and .2d of a 128 x 32 B tile per
//                              stage (128-byte swizzle) into a 3-deep shared-memory ring.  The A map is THREE
//                              dimensional {channel, tap, row} with the tap stride equal to the row stride: row m
//                              of the implicit im2col matrix is the contiguous slice x[m .. m + taps - 1] of the
//                              time-major activation, so a causal Conv1d(k) / ConvTranspose1d(2s, s) is a plain GEMM
//                              over OVERLAPPING rows and no im2col buffer exists (rows beyond M read as zero)
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x N128 x K8, four
//                              per stage); tcgen05.commit releases the stage / signals the epilogue
//   warps 2..5  epilogue     : tcgen05.ld -> bias / GELU / LayerScale / residual -> fp32 stores; optionally the
//                              ELU of the result (rounded to TF32) as a second output, because the consumer of a
//                              SEANet activation applies ELU to its input and a TMA-fed operand cannot be touched
//                              on its way into the tensor core
// 96 KB of shared memory and 128 TMEM columns per CTA: two CTAs per SM, one's epilogue under the other's MMAs.
// Operands are rounded to TF32 (round-to-nearest) where they are PRODUCED (weights at pack time, activations in
// the producing epilogue), so the tensor core's truncation of the low mantissa bits is exact.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mtc {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int UMMA_K = 8;
constexpr int THREADS = 192;
constexpr uint32_t STAGE_BYTES = (BM + BN) * BK * 4;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

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
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// cin: channels per tap of the A operand (K = taps * cin, cin % 32 == 0)
__global__ void __launch_bounds__(THREADS, 2)
k_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Args a, int cin) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const long long m0 = (long long)blockIdx.y * BM;
  const int num_kb = a.K / BK;

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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + BM * BK * 4;
        mbar_expect(&full[s], STAGE_BYTES);
        const int k = kb * BK;
        tma_load_3d(sa, &map_a, &full[s], k % cin, k / cin, (int)m0);
        tma_load_2d(sb, &map_b, &full[s], k, n0);
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
        const unsigned char* sb = sa + BM * BK * 4;
        const uint64_t ad = smem_desc(sa), bd = smem_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)  // +32 bytes along K inside the 128-byte swizzle atom
          umma_tf32(tmem_base, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // epilogue, two steps.  (1) warp w may touch TMEM lanes [32*(w%4), +32): each lane drains ITS row of the
    // accumulator into a shared-memory tile (the pipeline stages are free by now).  (2) the four warps walk the
    // tile in row-major order, four consecutive columns per lane: bias / GELU / LayerScale / residual and the
    // stores are COALESCED -- with one row per lane a 32- or 64-column output (the SEANet tail) made every store
    // instruction touch 32 different lines and the epilogue, not HBM, bounded those GEMMs.
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int nvalid = a.N - n0 < BN ? a.N - n0 : BN;  // a multiple of 4 (launcher)
    float* tile = reinterpret_cast<float*>(smem);      // [BM][TLD]
    constexpr int TLD = BN + 4;
#pragma unroll 1
    for (int c0 = 0; c0 < nvalid; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
    asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
    const int et = threadIdx.x - 64;                 // 0 .. 127
    const int c4n = nvalid >> 2;                     // float4 columns per row
    const int mrows = a.M - m0 < BM ? (int)(a.M - m0) : BM;
#pragma unroll 1
    for (int i = et; i < mrows * c4n; i += 128) {
      const int r = i / c4n, c = (i - r * c4n) * 4;
      const long long row = m0 + r;
      const int n = n0 + c;
      const float4 t = *reinterpret_cast<const float4*>(tile + (size_t)r * TLD + c);
      float y[4] = {t.x, t.y, t.z, t.w}, e[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float x = y[k];
        if (a.bias) x += a.bias[(n + k) % a.bias_period];
        if (a.flags & F_GELU) x = gelu_f(x);
        if (a.flags & F_LAYERSCALE) x *= a.scale[n + k];
        y[k] = x;
      }
      if (a.flags & F_RESID) {
        const float4 rr = *reinterpret_cast<const float4*>(a.R + row * a.ldr + n);
        y[0] += rr.x; y[1] += rr.y; y[2] += rr.z; y[3] += rr.w;
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// element-wise helpers of the tensor-core decode path
__global__ void k_round_tf32(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = round_tf32(x[i]);
}

}  // namespace mtc
