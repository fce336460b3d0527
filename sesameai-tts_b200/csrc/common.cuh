// Shared device helpers for libcsm_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define CSM_WARP 32

__device__ __forceinline__ float bf2f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 f2bf(float x) { return __float2bfloat16_rn(x); }
// value after a round trip through bf16 (one of the reference's rounding points)
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// two packed bf16 -> two fp32 (exact)
__device__ __forceinline__ float bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bfhi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// streaming 16-byte load: weights are read once per use, keep them out of L1
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Programmatic dependent launch (the row-batched decode step is ~1100 dependent launches of a few microseconds
// each): a kernel launched with the programmatic-serialisation attribute starts while its predecessor
// drains.  pdl_wait() returns once the previous kernel of the stream has completed and its writes are
// visible; pdl_trigger() lets the NEXT kernel's CTAs become resident once every CTA of this one has
// issued it.  Rules in this library: before pdl_wait() a kernel touches only immutable data (weights,
// tables: L2 prefetch, barrier / TMEM set-up, weight TMA boxes); EVERY thread that reads or writes
// activations calls pdl_wait() first; and the trigger comes AFTER the wait, so at most two consecutive
// kernels overlap and no ordering is inherited through a chain of waits.  Without the launch attribute
// both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Barrier among the first NT threads of the CTA on hardware barrier BAR (BAR 0 with NT ==
// blockDim.x is __syncthreads()).  Warp-specialised kernels sync their consumer warps on BAR 1.
template <int NT, int BAR>
__device__ __forceinline__ void csync() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

// reductions over NT threads (tid in [0, NT)) through a caller-provided 33-float smem scratch;
// every participating thread gets the result
template <int NT, int BAR>
__device__ __forceinline__ float block_sum(float v, float* scratch, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int nw = NT / 32;
  v = warp_sum(v);
  csync<NT, BAR>();
  if (lane == 0) scratch[warp] = v;
  csync<NT, BAR>();
  float t = (lane < nw) ? scratch[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
template <int NT, int BAR>
__device__ __forceinline__ float block_max(float v, float* scratch, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int nw = NT / 32;
  v = warp_max(v);
  csync<NT, BAR>();
  if (lane == 0) scratch[warp] = v;
  csync<NT, BAR>();
  float t = (lane < nw) ? scratch[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// dot of 8 packed bf16 weights with 8 packed bf16 activations, fp32 accumulate
__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc) {
  acc = fmaf(bflo(w.x), bflo(x.x), acc);
  acc = fmaf(bfhi(w.x), bfhi(x.x), acc);
  acc = fmaf(bflo(w.y), bflo(x.y), acc);
  acc = fmaf(bfhi(w.y), bfhi(x.y), acc);
  acc = fmaf(bflo(w.z), bflo(x.z), acc);
  acc = fmaf(bfhi(w.z), bfhi(x.z), acc);
  acc = fmaf(bflo(w.w), bflo(x.w), acc);
  acc = fmaf(bfhi(w.w), bfhi(x.w), acc);
  return acc;
}

// torch F.silu on a bf16 tensor: fp32 x/(1+exp(-x)), rounded to bf16
__device__ __forceinline__ float silu_bf(float g) { return rbf(g / (1.0f + expf(-g))); }
