// HMMA (mma.sync m16n8k16 bf16) issue rate and latency on B200, per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void k(int iters, long long* out, float* sink) {
  float acc[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  uint32_t a = threadIdx.x * 0x3c003c00u, b = 0x3c003c00u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma16816(acc[i], a, a + i, a, a, b, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) for (int e = 0; e < 4; ++e) s += acc[i][e];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
// the megakernel's chunk loop: 4 KB chunk from shared memory, R = 16
__global__ void kchunk(int iters, long long* out, float* sink) {
  __shared__ __align__(16) unsigned char ring[8 * 4096];
  __shared__ __align__(16) unsigned char xs[4096];
  for (int i = threadIdx.x; i < 8 * 4096 / 4; i += blockDim.x) ((uint32_t*)ring)[i] = 0x3c003c00u;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) ((uint32_t*)xs)[i] = 0x3c003c00u;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  float acc[4][4];
  for (int i = 0; i < 4; ++i) for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const unsigned char* wp = ring + warp * 4096 + lane * 16;
    const unsigned char* xp = xs + q * 16;
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 1
    for (int b = 0; b < 4; b += 2) {
      const uint4 w0 = *(const uint4*)(wp + b * 1024), w2 = *(const uint4*)(wp + (b + 1) * 1024);
      const uint4 w1 = *(const uint4*)(wp + b * 1024 + 512), w3 = *(const uint4*)(wp + (b + 1) * 1024 + 512);
      const uint4 x0 = g < 1 ? *(const uint4*)(xp + b * 64) : z, x1 = g < 1 ? *(const uint4*)(xp + (b + 1) * 64) : z;
      mma16816(acc[0], w0.x, w1.x, w0.y, w1.y, x0.x, x0.y);
      mma16816(acc[1], w0.z, w1.z, w0.w, w1.w, x0.z, x0.w);
      mma16816(acc[2], w2.x, w3.x, w2.y, w3.y, x1.x, x1.y);
      mma16816(acc[3], w2.z, w3.z, w2.w, w3.w, x1.z, x1.w);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 4; ++i) for (int e = 0; e < 4; ++e) s += acc[i][e];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* out; float* sink;
  cudaMallocManaged(&out, 64); cudaMalloc(&sink, 1 << 20);
  const int iters = 4096;
  for (int warps : {1, 4, 8, 16}) {
    k<1><<<148, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize();
    printf("warps/SM %2d  dependent HMMA chain : %.1f clk per HMMA\n", warps, (double)out[0] / iters);
    k<4><<<148, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize();
    printf("warps/SM %2d  4 independent chains : %.1f clk per HMMA per warp\n", warps, (double)out[0] / iters / 4);
    k<8><<<148, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize();
    printf("warps/SM %2d  8 independent chains : %.1f clk per HMMA per warp\n", warps, (double)out[0] / iters / 8);
  }
  kchunk<<<148, 256>>>(iters, out, sink); cudaDeviceSynchronize();
  printf("megakernel chunk loop (8 warps, 4 KB chunk = 8 HMMA + 12 LDS.128): %.0f clk per chunk\n", (double)out[0] / iters);
  kchunk<<<148, 32>>>(iters, out, sink); cudaDeviceSynchronize();
  printf("megakernel chunk loop (1 warp): %.0f clk per chunk\n", (double)out[0] / iters);
  return 0;
}
