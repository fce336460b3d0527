// Hand-off latency between CTAs through L2 on B200: the floor for the megakernel's phase boundary.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_handoff microbench_handoff.cu
// Tests: (1) ping-pong between two CTAs with different load/store flavours; (2) one producer CTA,
// all other CTAs poll a 4 KB tagged vector with 256 threads (or one lane per warp), time until the
// last consumer has seen all of it; (3) the same while a weight stream (bulk loads) saturates HBM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
template <int MODE>
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  if (MODE == 0) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else if (MODE == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else if (MODE == 2) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <int MODE>
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  if (MODE == 0) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  else if (MODE == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  else if (MODE == 2) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  else asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// (1) ping-pong: CTA a and CTA b (one thread each)
template <int MODE>
__global__ void k_pingpong(uint32_t* f, int a, int b, int iters, unsigned long long* out) {
  if (threadIdx.x != 0) return;
  uint32_t* fa = f;
  uint32_t* fb = f + 64;
  if (blockIdx.x == a) {
    const unsigned long long t0 = gtimer();
    const long long c0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      st_flag<MODE>(fa, i);
      for (unsigned sp = 0; ld_flag<MODE>(fb) != (uint32_t)i; ++sp) if (sp > (1u << 20)) { out[2] = i; return; }
    }
    out[0] = gtimer() - t0;
    out[1] = clock64() - c0;
  } else if (blockIdx.x == b) {
    for (int i = 1; i <= iters; ++i) {
      for (unsigned sp = 0; ld_flag<MODE>(fa) != (uint32_t)i; ++sp) if (sp > (1u << 20)) { out[3] = i; return; }
      st_flag<MODE>(fb, i);
    }
  }
}

// single-thread dependent load latency (pointer chase on one address = RTT)
template <int MODE>
__global__ void k_rtt(uint32_t* f, int iters, unsigned long long* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long c0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) acc += ld_flag<MODE>(f + (acc & 1));
  out[0] = clock64() - c0;
  out[1] = acc;
}

// (2) broadcast: CTA 0 writes a vector of n tagged words in round r; every other CTA polls it
// (pollers = 256 threads, each its own 16-byte unit) and records when it has seen everything.
// ts[r][cta] = globaltimer at completion; ts[r][0] = producer's timestamp right after its stores.
__device__ __forceinline__ uint4 ldv4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ldr4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
template <int LD>
__global__ void k_bcast(uint32_t* vec, int n, int rounds, unsigned long long* ts, int sentinel, const uint4* stream,
                        size_t stream_n16, int stream_ctas) {
  __shared__ uint4 sink[256];
  const int ncta = gridDim.x;
  if ((int)blockIdx.x >= ncta - stream_ctas) {
    // background HBM stream: plain 16-byte loads over a large buffer until round counter ends
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t per = stream_n16 / stream_ctas;
    const uint4* base = stream + (size_t)(blockIdx.x - (ncta - stream_ctas)) * per;
    for (int rep = 0; rep < rounds; ++rep)
      for (size_t i = threadIdx.x; i < per; i += 256 * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + u * 256 < per) ? __ldcs(base + i + u * 256) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc.x ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
      }
    sink[threadIdx.x] = acc;
    return;
  }
  const int nc = ncta - stream_ctas;
  for (int r = 1; r <= rounds; ++r) {
    const uint32_t tag = (uint32_t)r << 16;
    if (blockIdx.x == 0) {
      // wait until everybody reported the previous round (so rounds do not overlap)
      if (threadIdx.x == 0 && r > 1)
        for (int c = 1; c < nc; ++c)
          for (unsigned sp = 0; *((volatile unsigned long long*)&ts[(size_t)(r - 1) * ncta + c]) == 0; ++sp) if (sp > (1u << 20)) break;
      __syncthreads();
      for (int i = threadIdx.x * 4; i < n; i += 1024) {
        uint4 v = make_uint4(tag | 1, tag | 2, tag | 3, tag | 4);
        __stcg(reinterpret_cast<uint4*>(vec + i), v);
      }
      if (threadIdx.x == 0) ts[(size_t)r * ncta] = gtimer();
      __syncthreads();
    } else {
      if (sentinel) {
        if ((threadIdx.x & 31) == 0) {
          const uint32_t* p = vec + ((threadIdx.x * 4) % n);
          uint4 v;
          unsigned sp = 0;
          do { v = LD ? ldr4(p) : ldv4(p); } while (((v.x ^ tag) >> 16) && ++sp < (1u << 20));
        }
        __syncwarp();
      }
      for (int i = threadIdx.x * 4; i < n; i += 1024) {
        uint4 v;
        unsigned sp = 0;
        do {
          v = LD ? ldr4(vec + i) : ldv4(vec + i);
        } while (((((v.x ^ tag) | (v.y ^ tag)) | ((v.z ^ tag) | (v.w ^ tag))) >> 16) && ++sp < (1u << 20));
      }
      __syncthreads();
      if (threadIdx.x == 0) ts[(size_t)r * ncta + blockIdx.x] = gtimer();
    }
  }
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  uint32_t* f;
  unsigned long long* out;
  cudaMalloc(&f, 1 << 20);
  cudaMemset(f, 0, 1 << 20);
  cudaMallocManaged(&out, 4096);
  const int iters = 2000;
  const char* names[4] = {"volatile", "relaxed.gpu", "acq/rel.gpu", "cg"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int b : {1, 2, 37, 74, 147}) {
      cudaMemset(f, 0, 1024);
      out[0] = out[1] = out[2] = out[3] = 0;
      switch (mode) {
        case 0: k_pingpong<0><<<sms, 32>>>(f, 0, b, iters, out); break;
        case 1: k_pingpong<1><<<sms, 32>>>(f, 0, b, iters, out); break;
        case 2: k_pingpong<2><<<sms, 32>>>(f, 0, b, iters, out); break;
        default: k_pingpong<3><<<sms, 32>>>(f, 0, b, iters, out); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      printf("pingpong %-12s cta0<->cta%-3d one-way %.0f ns (%.0f clk)  %s\n", names[mode], b, out[0] / (2.0 * iters),
             out[1] / (2.0 * iters), e == cudaSuccess ? (out[2] | out[3] ? "SPIN CAP HIT" : "") : cudaGetErrorString(e));
    }
    switch (mode) {
      case 0: k_rtt<0><<<1, 32>>>(f, iters, out); break;
      case 1: k_rtt<1><<<1, 32>>>(f, iters, out); break;
      case 2: k_rtt<2><<<1, 32>>>(f, iters, out); break;
      default: k_rtt<3><<<1, 32>>>(f, iters, out); break;
    }
    cudaDeviceSynchronize();
    printf("rtt      %-12s %.0f clk per dependent load\n", names[mode], out[0] / (double)iters);
  }
  // broadcast
  const int rounds = 200;
  unsigned long long* dts;
  const size_t ts_bytes = (size_t)(rounds + 1) * sms * 8;
  cudaMalloc(&dts, ts_bytes);
  std::vector<unsigned long long> ts((size_t)(rounds + 1) * sms);
  uint4* stream;
  const size_t stream_bytes = (size_t)4 << 30;
  cudaMalloc(&stream, stream_bytes);
  cudaMemset(stream, 1, stream_bytes);
  for (int ld = 0; ld < 2; ++ld)
    for (int n : {1024, 8192})
      for (int sentinel = 0; sentinel < 2; ++sentinel)
        for (int sc : {0, 100}) {
          cudaMemset(f, 0, 1 << 20);
          cudaMemset(dts, 0, ts_bytes);
          const int r_eff = sc ? 20 : rounds;
          if (ld) k_bcast<1><<<sms, 256>>>(f, n, r_eff, dts, sentinel, stream, stream_bytes / 16 / 8, sc);
          else k_bcast<0><<<sms, 256>>>(f, n, r_eff, dts, sentinel, stream, stream_bytes / 16 / 8, sc);
          cudaError_t e = cudaDeviceSynchronize();
          cudaMemcpy(ts.data(), dts, ts_bytes, cudaMemcpyDeviceToHost);
          const int nc = sms - sc;
          double first = 0, last = 0, med = 0;
          int cnt = 0;
          for (int r = 2; r <= r_eff; ++r) {
            std::vector<double> d;
            for (int c = 1; c < nc; ++c) d.push_back((double)ts[(size_t)r * sms + c] - (double)ts[(size_t)r * sms]);
            std::sort(d.begin(), d.end());
            first += d.front(); last += d.back(); med += d[d.size() / 2];
            ++cnt;
          }
          printf("bcast ld=%s n=%5d words sentinel=%d stream_ctas=%3d: first %.0f ns  median %.0f ns  last %.0f ns  %s\n",
                 ld ? "relaxed.gpu" : "volatile", n, sentinel, sc, first / cnt, med / cnt, last / cnt,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
