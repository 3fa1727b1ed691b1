// Shared helpers for the fpie_b200 CUDA sources (sm_100a only).
#pragma once
#include <algorithm>

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <stdexcept>
#include <string>

namespace fpie {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    throw Error(buf);
  }
}
#define CUDA_CHECK(expr) ::fpie::cuda_check((expr), #expr, __FILE__, __LINE__)
#define FPIE_REQUIRE(cond, msg)          \
  do {                                   \
    if (!(cond)) throw ::fpie::Error(msg); \
  } while (0)

// RAII device buffer (cudaMalloc / cudaFree on a fixed device).  resize() keeps an allocation that is
// large enough and not wastefully so: repeated reset() calls (the GUI, batch services) then do no
// cudaMalloc / cudaFree at all -- both synchronise the device and cost milliseconds.
template <typename T>
struct DeviceBuffer {
  T *ptr = nullptr;
  size_t count = 0;     // logical element count
  size_t capacity = 0;  // allocated element count
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = capacity = 0;
  }
  void resize(size_t n) {
    const size_t keep_bytes = std::max<size_t>(2 * n * sizeof(T), (size_t)32 << 20);
    if (ptr && n && n <= capacity && capacity * sizeof(T) <= keep_bytes) {
      count = n;
      return;
    }
    release();
    if (n) {
      CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(T)));
      count = capacity = n;
    }
  }
  size_t bytes() const { return count * sizeof(T); }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != dev) CUDA_CHECK(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Geometry of the padded planar layout every grid kernel works on.
//   pixel (r, c) of plane p  ->  base[p * plane + (r + padr) * pitch + (c + padc)]
//   mask bit of (r, c)       ->  bits[(r + padr) * wpitch + ((c + padc) >> 5)] >> ((c + padc) & 31)
// padc and pitch are multiples of 32, so 128-bit accesses at 4-pixel groups
// are aligned and a group's 4 mask bits never straddle a word.
struct PlaneGeom {
  int n, m;         // logical rows / cols of the grid
  int padr, padc;   // padding before logical (0, 0)
  int rows, pitch;  // allocated rows, floats per row
  int wpitch;       // mask words per row (= pitch / 32)
  int groups;       // 4-pixel groups per logical row (= ceil(m / 4))
  long long plane;  // floats per plane (= rows * pitch)
};

#ifdef __CUDACC__
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// predicated (branch-free) 128-bit global store
__device__ __forceinline__ void st4_if(void *p, float4 v, uint32_t on) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.u32 p, %5, 0;\n"
      "@p st.global.v4.f32 [%0], {%1, %2, %3, %4};\n"
      "}\n" ::"l"(p),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(on)
      : "memory");
}

// One Jacobi update in the reference's add order, evaluated on quarter-scaled
// operands:  ((((g + U) + D) + L) + R) / 4  ==  fma(R,q, fma(L,q, fma(D,q, fma(U,q, g/4))))
// with q = 0.25.  Scaling by a power of two commutes with round-to-nearest, so
// each fma returns exactly a quarter of the reference's partial sum and the
// result is bit-identical (fpie/np_solver.py:83-88); hq = g/4 is precomputed.
__device__ __forceinline__ float jacobi_q(float hq, float up, float dn, float lf, float rt) {
  float t = __fmaf_rn(up, 0.25f, hq);
  t = __fmaf_rn(dn, 0.25f, t);
  t = __fmaf_rn(lf, 0.25f, t);
  return __fmaf_rn(rt, 0.25f, t);
}
#endif

}  // namespace fpie
