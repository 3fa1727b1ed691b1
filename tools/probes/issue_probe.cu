// Issue-rate probe for the Jacobi sweep's instruction mix on sm_100a.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o issue_probe issue_probe.cu && ./issue_probe
//
// Each kernel runs one CTA per SM with W warps; every thread carries ILP
// independent dependency chains of depth 4 (the shape of one pixel's Jacobi
// update: fma(R,q, fma(L,q, fma(D,q, fma(U,q, h))))).  Reports warp-level
// instructions per clock per SM sub-partition and fp32 FMA lanes per clock
// per SM, for scalar FFMA, packed FFMA2 (fma.rn.f32x2), and both mixed with
// the shuffles / shared-memory loads of the real sweep body.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                                  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// MODE 0: scalar FFMA chains; 1: FFMA2 chains; 2: scalar + 2 SHFL per 16 FFMA; 3: FFMA2 + 2 SHFL per 8 FFMA2;
// 4: FFMA2 + 2 SHFL + 1 LDS.128 per 8 FFMA2; 5: scalar + 2 SHFL + 1 LDS.128 per 16 FFMA
template <int MODE, int ILP>
__global__ void probe(float *out, long long *cycles, int iters) {
  __shared__ float4 sm[1024];
  const int tid = threadIdx.x;
  sm[tid & 1023] = make_float4(tid, 1, 2, 3);
  __syncthreads();
  float acc[ILP * 2];
#pragma unroll
  for (int i = 0; i < ILP * 2; ++i) acc[i] = tid * 0.001f + i;
  float extra = 0.f;
  const float q = 0.25f;
  u64 q2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(q2) : "f"(q));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // MODE >= 4: the addend of every chain's first FMA comes from shared memory (the quarter-gradient of the
    // "h in shared memory" design): one LDS.128 per 4 chains
    float hv[ILP * 2];
    if (MODE >= 4) {
#pragma unroll
      for (int g = 0; g < (ILP * 2) / 4; ++g) {
        const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(tid + g * 32 + it) & 1023]);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(hv[4 * g]), "=f"(hv[4 * g + 1]), "=f"(hv[4 * g + 2]), "=f"(hv[4 * g + 3])
                     : "r"(addr));
      }
    }
    float nw[ILP * 2];
    if (MODE == 0 || MODE == 2 || MODE == 5) {
#pragma unroll
      for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int i = 0; i < ILP * 2; ++i) {
          const float add = (d == 0) ? ((MODE >= 4) ? hv[i] : acc[i]) : nw[i];
          nw[i] = ffma(acc[(i + d + 1) % (ILP * 2)], q, add);
        }
    } else {
#pragma unroll
      for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
          u64 a, b;
          if (d == 0 && MODE >= 4)
            asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(hv[2 * i]), "f"(hv[2 * i + 1]));
          else if (d == 0)
            asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
          else
            asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(nw[2 * i]), "f"(nw[2 * i + 1]));
          const int j = (i + d + 1) % ILP;
          asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(acc[2 * j]), "f"(acc[2 * j + 1]));
          const u64 r = ffma2(b, q2, a);
          asm("mov.b64 {%0, %1}, %2;" : "=f"(nw[2 * i]), "=f"(nw[2 * i + 1]) : "l"(r));
        }
    }
#pragma unroll
    for (int i = 0; i < ILP * 2; ++i) acc[i] = nw[i];
    if (MODE >= 2) {
      // per 16 scalar FMAs (= 8 packed): 2 shuffles feeding the next iteration's chains
#pragma unroll
      for (int g = 0; g < (ILP * 2) / 4; ++g) {
        extra += __shfl_up_sync(0xffffffffu, acc[4 * g], 1);
        extra += __shfl_down_sync(0xffffffffu, acc[4 * g + 1], 1);
      }
    }
  }
  const long long t1 = clock64();
  float s = extra;
#pragma unroll
  for (int i = 0; i < ILP * 2; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int ILP>
void run(const char *name, int warps, int sms, float *out, long long *cyc) {
  const int iters = 2000;
  probe<MODE, ILP><<<sms, warps * 32>>>(out, cyc, iters);
  CK(cudaDeviceSynchronize());
  probe<MODE, ILP><<<sms, warps * 32>>>(out, cyc, iters);
  CK(cudaDeviceSynchronize());
  long long h[256];
  CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += (double)h[i];
  mean /= sms;
  const bool packed = (MODE == 1 || MODE == 3 || MODE == 4);
  const double fma_inst = (double)iters * 4 * (packed ? ILP : 2 * ILP);  // per warp
  const double fmas = (double)iters * 4 * 2 * ILP;                        // scalar FMAs per thread
  const double other = (MODE >= 2) ? (double)iters * ((ILP * 2) / 4) * ((MODE >= 4) ? 5.0 : 4.0) : 0.0;  // shfl + fadd (+ lds + fadd)
  const double wps = warps / 4.0;
  printf("%-28s warps/SMSP %.0f ILP %2d: %8.0f clk  fma-inst/clk/SMSP %.3f  FMA lanes/clk/SM %6.1f  (+%.2f other inst/clk/SMSP)\n", name, wps,
         ILP * 2, mean, fma_inst * wps / mean, fmas * warps * 32 / mean, other * wps / mean);
}

int main() {
  int dev = 0;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  float *out;
  long long *cyc;
  CK(cudaMalloc(&out, sizeof(float) * sms * 1024));
  CK(cudaMalloc(&cyc, sizeof(long long) * sms));
  for (int warps : {4, 8, 16}) {
    run<0, 4>("FFMA", warps, sms, out, cyc);
    run<0, 8>("FFMA", warps, sms, out, cyc);
    run<0, 16>("FFMA", warps, sms, out, cyc);
    run<1, 4>("FFMA2", warps, sms, out, cyc);
    run<1, 8>("FFMA2", warps, sms, out, cyc);
    run<1, 16>("FFMA2", warps, sms, out, cyc);
    run<2, 16>("FFMA + 2 SHFL/16", warps, sms, out, cyc);
    run<3, 16>("FFMA2 + 2 SHFL/8", warps, sms, out, cyc);
    run<5, 16>("FFMA + 2 SHFL + LDS.128", warps, sms, out, cyc);
    run<4, 16>("FFMA2 + 2 SHFL + LDS.128", warps, sms, out, cyc);
  }
  return 0;
}
