// EquSolver kernels + host driver (sm_100a).
//
// Replaces fpie/core/cuda/equ.cu (reference).  Differences in design:
//   * true Jacobi with ping-pong buffers (fpie/np_solver.py:33-41), not the
//     reference CUDA backend's racy in-place update (equ.cu:193-197);
//   * X and B are stored as three channel planes so that the four gathers of a
//     warp are 128-byte coalesced runs whenever ids are row-major;
//   * partition is a device prefix scan (equ.cu:36-54 is a serial host loop);
//   * residual = warp-shuffle + block reduction into fp64 (equ.cu:162-179 is a
//     single block with a serial tail).

#include <algorithm>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

#include <cstdlib>

#include "equ_solver.cuh"
#include "prep.cuh"

namespace fpie {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

// ---------------------------------------------------------------------------
// partition: inclusive scan of (mask > 0), reduce-then-scan in three launches
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;  // inclusive over warps
  }
  __syncthreads();
  const uint32_t warp_off = (w > 0) ? warp_sums[w - 1] : 0u;
  if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return warp_off + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_count_kernel(const int32_t *__restrict__ mask, long long count, uint32_t *__restrict__ block_sums) {
  const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < count) c += mask[base + i] > 0;
  uint32_t total;
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// exclusive scan of the block sums, one CTA, carries across chunks
__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums_kernel(uint32_t *__restrict__ block_sums, int nblocks) {
  uint32_t carry = 0;
  for (int base = 0; base < nblocks; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < nblocks) ? block_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, &total);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_write_kernel(const int32_t *__restrict__ mask, long long count, const uint32_t *__restrict__ block_offsets,
                  int32_t *__restrict__ ids) {
  const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_ITEMS;
  uint32_t f[SCAN_ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    f[i] = (base + i < count) ? (mask[base + i] > 0) : 0u;
    c += f[i];
  }
  uint32_t run = block_offsets[blockIdx.x] + block_exclusive_scan(c, nullptr);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    run += f[i];
    if (base + i < count) ids[base + i] = (int32_t)run;
  }
}

// ---------------------------------------------------------------------------
// upload helpers
// ---------------------------------------------------------------------------
__global__ void rows3_to_planes_kernel(long long N, long long pitch, const float *__restrict__ aos,
                                       float *__restrict__ dst0, float *__restrict__ dst1) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float v = aos[i * 3 + ch];
    dst0[ch * pitch + i] = v;
    if (dst1) dst1[ch * pitch + i] = v;
  }
}

__global__ void planes_to_rows3_kernel(long long N, long long pitch, const float *__restrict__ src,
                                       float *__restrict__ aos) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) aos[i * 3 + ch] = src[ch * pitch + i];
}

__global__ void check_index_kernel(long long N, const int4 *__restrict__ A, int *__restrict__ flag) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int4 a = A[i];
  const bool bad = a.x < 0 || a.y < 0 || a.z < 0 || a.w < 0 || a.x >= N || a.y >= N || a.z >= N || a.w >= N;
  if (bad) atomicExch(flag, 1);
}

// Compact neighbour table for row-major ids: when left / right are always "absent" or i-1 / i+1
// (true for every system built from a row-major partition) only up / down need to be stored;
// bit 31 of each carries "left present" / "right present".  16 -> 8 bytes per unknown per sweep.
// A second, 4-byte form holds the up / down neighbours as 15-bit DISTANCES (up = i - du, down = i + dd,
// 0 = absent; bits 15 / 31 = left / right present): row-major ids put them less than one image row of unknowns
// away.  flags bit 0: not structured; bit 1: some distance does not fit (the 8-byte table is used then).
__global__ void compact_index_kernel(long long N, const int4 *__restrict__ A, int2 *__restrict__ UD,
                                     uint32_t *__restrict__ D16, int *__restrict__ flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int4 a = A[i];
  const bool ok = (a.z == 0 || a.z == i - 1) && (a.w == 0 || a.w == i + 1);
  if (!ok) atomicOr(flags, 1);
  UD[i] = make_int2(a.x | (a.z != 0 ? (int)0x80000000 : 0), a.y | (a.w != 0 ? (int)0x80000000 : 0));
  const long long du = a.x ? i - a.x : 0, dd = a.y ? a.y - i : 0;
  if ((a.x && (du <= 0 || du > 0x7fff)) || (a.y && (dd <= 0 || dd > 0x7fff))) atomicOr(flags, 2);
  D16[i] = (uint32_t)(du & 0x7fff) | (a.z != 0 ? 0x8000u : 0u) | ((uint32_t)(dd & 0x7fff) << 16) | (a.w != 0 ? 0x80000000u : 0u);
}

// fp16 copy of B for the compact gather kernel; *inexact is raised when a value does not survive the round trip
// (B = gradient + boundary targets of uint8 images: halves below 1024 always do)
__global__ void equ_b_to_half_kernel(long long N, long long pitch, const float *__restrict__ src,
                                     __half *__restrict__ dst, int *__restrict__ inexact) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= pitch * 3 || (i % pitch) >= N) return;  // (the padding behind row N - 1 is never read)
  const float v = src[i];
  const __half h = __float2half_rn(v);
  dst[i] = h;
  if (!(__half2float(h) == v)) atomicOr(inexact, 4);
}

// Does A describe exactly the 4-neighbour structure of the mask that partition() labelled?
// (row-major ids, absent neighbours = 0, no masked pixel on the frame)
__global__ void check_grid_structure_kernel(int n, int m, const int32_t *__restrict__ mask,
                                            const int32_t *__restrict__ ids, const int4 *__restrict__ A,
                                            int *__restrict__ bad) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= (long long)n * m) return;
  if (!(mask[p] > 0)) return;
  const int r = (int)(p / m), c = (int)(p % m);
  if (r == 0 || c == 0 || r == n - 1 || c == m - 1) {
    atomicExch(bad, 1);
    return;
  }
  const int4 a = A[ids[p]];
  const long long q[4] = {p - m, p + m, p - 1, p + 1};
  const int got[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int want = mask[q[k]] > 0 ? ids[q[k]] : 0;
    if (got[k] != want) atomicExch(bad, 1);
  }
}

// ---------------------------------------------------------------------------
// sweep / residual / output
// ---------------------------------------------------------------------------
// Same update as equ_sweep_kernel, reading the compact table: left / right are the adjacent
// unknowns i-1 / i+1 (coalesced) or the constant row 0.
__global__ void __launch_bounds__(1024)
equ_sweep_lr_kernel(long long N, long long pitch, const int2 *__restrict__ UD, const float *__restrict__ B,
                    const float *__restrict__ xin, float *__restrict__ xout) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // programmatic dependent launch: the next sweep may be scheduled now; the table and B are not written by
  // sweeps and are fetched before this grid waits for the previous sweep's X
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (i >= N) return;
  const int2 ud = UD[i];
  const int up = ud.x & 0x7fffffff, dn = ud.y & 0x7fffffff;
  const long long lf = (ud.x < 0) ? i - 1 : 0, rt = (ud.y < 0) ? i + 1 : 0;
  float b[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) b[ch] = B[ch * pitch + i];
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float *x = xin + ch * pitch;
    float s = __fadd_rn(b[ch], x[up]);
    s = __fadd_rn(s, x[dn]);
    s = __fadd_rn(s, x[lf]);
    s = __fadd_rn(s, x[rt]);
    xout[ch * pitch + i] = __fmul_rn(s, 0.25f);
  }
}

// The same update from the 4-byte distance table and (BH) the fp16 copy of B: 4 + 6 + 12 + 12 = 34 bytes per
// unknown and sweep instead of the 52 of the reference layout (A 16 + B 12 + X 12 + X' 12) -- same operands, same
// add order, same bits.  A thread owns FOUR consecutive unknowns: table, B and the centre values arrive as one
// 128-bit (64-bit) load each, left / right neighbours of the inner three come from registers, and only the up /
// down gathers stay scalar -- 9 instead of 17 memory requests per unknown (the scalar form of this kernel was
// request-bound: 4.6 TB/s of DRAM traffic against 5.4 TB/s for the 8-byte table).
template <bool BH>
__global__ void __launch_bounds__(256)
equ_sweep_d16_kernel(long long N, long long pitch, const uint32_t *__restrict__ D16, const float *__restrict__ B,
                     const __half *__restrict__ B16, const float *__restrict__ xin, float *__restrict__ xout) {
  const long long i0 = 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (see equ_sweep_lr_kernel)
  if (i0 >= N) return;
  const uint4 tv = *reinterpret_cast<const uint4 *>(D16 + i0);  // (the table is zero-padded to the pitch)
  const uint32_t t[4] = {tv.x, tv.y, tv.z, tv.w};
  float b[3][4];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (BH) {
      const uint2 raw = *reinterpret_cast<const uint2 *>(B16 + ch * pitch + i0);
      const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
      b[ch][0] = lo.x, b[ch][1] = lo.y, b[ch][2] = hi.x, b[ch][3] = hi.y;
    } else {
      const float4 v = ld4(B + ch * pitch + i0);
      b[ch][0] = v.x, b[ch][1] = v.y, b[ch][2] = v.z, b[ch][3] = v.w;
    }
  }
  long long up[4], dn[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t du = t[j] & 0x7fffu, dd = (t[j] >> 16) & 0x7fffu;
    up[j] = du ? i0 + j - du : 0;
    dn[j] = dd ? i0 + j + dd : 0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float *x = xin + ch * pitch;
    const float4 cv = ld4(x + i0);
    const float c[4] = {cv.x, cv.y, cv.z, cv.w};
    const float zero_row = x[0];  // what an absent neighbour reads (row 0, the constant)
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lf = (t[j] & 0x8000u) ? (j > 0 ? c[j > 0 ? j - 1 : 0] : x[i0 - 1]) : zero_row;
      const float rt = (t[j] >> 31) ? (j < 3 ? c[j < 3 ? j + 1 : 3] : x[i0 + 4]) : zero_row;
      float s = __fadd_rn(b[ch][j], x[up[j]]);
      s = __fadd_rn(s, x[dn[j]]);
      s = __fadd_rn(s, lf);
      s = __fadd_rn(s, rt);
      o[j] = __fmul_rn(s, 0.25f);
    }
    if (i0 + 3 < N) {
      st4(xout + ch * pitch + i0, make_float4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j < N) xout[ch * pitch + i0 + j] = o[j];
    }
  }
}

// The 4-byte-table sweep as a PERSISTENT, software-pipelined kernel.  A sweep of equ_sweep_d16_kernel costs the
// same ~260 us on config 3 as the 8-byte and the 16-byte table kernels although it moves a third less: ncu shows no
// saturated unit (DRAM 72 %, L2 39 %, L1 wavefronts 74 % -- and halving those with 128-bit gathers made it slower),
// only warps waiting on memory: a thread's gathers cannot start before its table entry has arrived, two dependent
// trips to DRAM per unknown.  Here a CTA strides over the system and a thread fetches the table entry of its
// NEXT four unknowns before it gathers for the current four, so the two trips of consecutive chunks overlap:
// 262 -> 235 us per sweep on config 3 (135 -> 150 Gupd/s, 5.0 TB/s of DRAM traffic).  Measured around it: fetching
// two chunks ahead helps at equal occupancy (124 -> 142 Gupd/s at three CTAs per SM) but needs 80 registers, and
// four CTAs per SM with one chunk ahead is faster (150); prefetching the centre vectors as well: 114; half the
// occupancy: 117 -- the kernel is bound by memory latency x resident warps, under the board's power cap
// (SM clock 1.6 GHz in these runs).
// D = how many chunks ahead the TABLE entry is fetched (FPIE_B200_D16_DEPTH; 1 is the default: 155 Gupd/s on config 3,
// two ahead 151).  Only the table entry gates the gathers; B is needed when the sums are formed and is loaded with
// the chunk's own centre vectors and gathers (prefetching it as well measured the same and costs 6 registers).
// read-once streams (table, B) and the written state with streaming cache hints (-DFPIE_D16_STREAM=1: experiment)
#if defined(FPIE_D16_STREAM) && FPIE_D16_STREAM
#define D16_LD4(p) __ldcs(p)
#define D16_LD2(p) __ldcs(p)
#define D16_ST4(p, v) __stcs(p, v)
#else
#define D16_LD4(p) (*(p))
#define D16_LD2(p) (*(p))
#define D16_ST4(p, v) (*(p) = (v))
#endif
#ifndef FPIE_D16_BLOCK
#define FPIE_D16_BLOCK 256  // threads per CTA of the persistent gather kernel (1024 threads per SM with the fp16 B stream)
#endif
template <bool BH, int D>
__global__ void __launch_bounds__(FPIE_D16_BLOCK, (BH ? 1024 : 768) / FPIE_D16_BLOCK)
equ_sweep_d16p_kernel(long long N, long long pitch, const uint32_t *__restrict__ D16, const float *__restrict__ B,
                      const __half *__restrict__ B16, const float *__restrict__ xin, float *__restrict__ xout) {
  const long long stride = 4ll * gridDim.x * blockDim.x;
  long long i0 = 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (see equ_sweep_lr_kernel)
  if (i0 >= N) return;
  uint4 tq[D + 1];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    tq[d] = make_uint4(0u, 0u, 0u, 0u);
    if (i0 + d * stride < N) tq[d] = D16_LD4(reinterpret_cast<const uint4 *>(D16 + i0 + d * stride));  // (zero-padded to the pitch)
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (; i0 < N; i0 += stride) {
    const long long inext = i0 + D * stride;
    tq[D] = make_uint4(0u, 0u, 0u, 0u);
    if (inext < N) tq[D] = D16_LD4(reinterpret_cast<const uint4 *>(D16 + inext));
    uint2 bh[3];
    float4 bf[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (BH)
        bh[ch] = D16_LD2(reinterpret_cast<const uint2 *>(B16 + ch * pitch + i0));
      else
        bf[ch] = ld4(B + ch * pitch + i0);
    }
    const uint32_t t[4] = {tq[0].x, tq[0].y, tq[0].z, tq[0].w};
    int up[4], dn[4];  // (ids are int32: EquSolver::reset requires N < 2^31)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t du = t[j] & 0x7fffu, dd = (t[j] >> 16) & 0x7fffu;
      up[j] = du ? (int)i0 + j - (int)du : 0;
      dn[j] = dd ? (int)i0 + j + (int)dd : 0;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float *x = xin + ch * pitch;
      const float4 cv = ld4(x + i0);
      const float c[4] = {cv.x, cv.y, cv.z, cv.w};
      const float zero_row = x[0];  // what an absent neighbour reads (row 0, the constant)
      float b[4];
      if (BH) {
        const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&bh[ch].x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&bh[ch].y));
        b[0] = lo.x, b[1] = lo.y, b[2] = hi.x, b[3] = hi.y;
      } else {
        b[0] = bf[ch].x, b[1] = bf[ch].y, b[2] = bf[ch].z, b[3] = bf[ch].w;
      }
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lf = (t[j] & 0x8000u) ? (j > 0 ? c[j > 0 ? j - 1 : 0] : x[i0 - 1]) : zero_row;
        const float rt = (t[j] >> 31) ? (j < 3 ? c[j < 3 ? j + 1 : 3] : x[i0 + 4]) : zero_row;
        float sum = __fadd_rn(b[j], x[up[j]]);
        sum = __fadd_rn(sum, x[dn[j]]);
        sum = __fadd_rn(sum, lf);
        sum = __fadd_rn(sum, rt);
        o[j] = __fmul_rn(sum, 0.25f);
      }
      if (i0 + 3 < N) {
        D16_ST4(reinterpret_cast<float4 *>(xout + ch * pitch + i0), make_float4(o[0], o[1], o[2], o[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i0 + j < N) xout[ch * pitch + i0 + j] = o[j];
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) tq[d] = tq[d + 1];
  }
}

// X'[i] = ((((B[i] + X[up]) + X[down]) + X[left]) + X[right]) / 4   (np_solver.py:33-41)
__global__ void __launch_bounds__(1024)
equ_sweep_kernel(long long N, long long pitch, const int4 *__restrict__ A, const float *__restrict__ B,
                 const float *__restrict__ xin, float *__restrict__ xout) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (see equ_sweep_lr_kernel)
  if (i >= N) return;
  const int4 a = A[i];
  float b[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) b[ch] = B[ch * pitch + i];
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float *x = xin + ch * pitch;
    float s = __fadd_rn(b[ch], x[a.x]);
    s = __fadd_rn(s, x[a.y]);
    s = __fadd_rn(s, x[a.z]);
    s = __fadd_rn(s, x[a.w]);
    xout[ch * pitch + i] = __fmul_rn(s, 0.25f);
  }
}

// Red-black Gauss-Seidel half-sweep, in place over ids [lo, hi) (fpie/core/openmp/equ.cc:67-80, 109-118).
// With ids from the red-black partition every neighbour of an updated unknown has the other
// colour, so the in-place update is race-free and thread-count independent.
__global__ void __launch_bounds__(1024)
equ_rb_half_kernel(long long lo, long long hi, long long pitch, const int4 *__restrict__ A,
                   const float *__restrict__ B, float *__restrict__ x) {
  const long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int4 a = A[i];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float *xc = x + ch * pitch;
    float s = __fadd_rn(B[ch * pitch + i], xc[a.x]);
    s = __fadd_rn(s, xc[a.y]);
    s = __fadd_rn(s, xc[a.z]);
    s = __fadd_rn(s, xc[a.w]);
    xc[i] = __fmul_rn(s, 0.25f);
  }
}

// parity-filtered flags for the red-black partition: out = mask > 0 && ((row + col) & 1) == parity
__global__ void rb_flags_kernel(const int32_t *__restrict__ mask, long long count, int m, int parity,
                                int32_t *__restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const int r = (int)(idx / m), c = (int)(idx % m);
  out[idx] = (mask[idx] > 0 && ((r + c) & 1) == parity) ? 1 : 0;
}

// ids = odd pixel ? odd_rank : n_odd + even_rank   (0 on unmasked pixels, openmp/equ.cc:30-34)
__global__ void rb_combine_kernel(const int32_t *__restrict__ mask, long long count, int m,
                                  const int32_t *__restrict__ odd_ids, const int32_t *__restrict__ even_ids,
                                  int32_t *__restrict__ ids) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const int r = (int)(idx / m), c = (int)(idx % m);
  const int n_odd = odd_ids[count - 1];
  int v = 0;
  if (mask[idx] > 0) v = ((r + c) & 1) ? odd_ids[idx] : n_odd + even_ids[idx];
  ids[idx] = v;
}

// err[c] = sum_i |((((B + X[up]) + X[down]) + X[left]) + X[right]) - 4 X[i]|   (np_solver.py:42-50)
__global__ void __launch_bounds__(256)
equ_residual_kernel(long long lo, long long N, long long pitch, const int4 *__restrict__ A,
                    const float *__restrict__ B, const float *__restrict__ x, double *__restrict__ err) {
  // rows [lo, N): the whole system, or the rows a rank owns of an id-range shard (set_window)
  const long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double v[3] = {0.0, 0.0, 0.0};
  if (i < N) {
    const int4 a = A[i];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float *xc = x + ch * pitch;
      float s = __fadd_rn(B[ch * pitch + i], xc[a.x]);
      s = __fadd_rn(s, xc[a.y]);
      s = __fadd_rn(s, xc[a.z]);
      s = __fadd_rn(s, xc[a.w]);
      s = __fsub_rn(s, __fmul_rn(4.0f, xc[i]));
      v[ch] = (double)fabsf(s);
    }
  }
  __shared__ double partial[3][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[ch] += __shfl_down_sync(0xffffffffu, v[ch], o);
    if (lane == 0) partial[ch][w] = v[ch];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      double t = (lane < (blockDim.x >> 5)) ? partial[ch][lane] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (lane == 0 && t != 0.0) atomicAdd(&err[ch], t);
    }
  }
}

__device__ __forceinline__ uint8_t clip_byte(float v) { return (uint8_t)__float2uint_rz(fminf(fmaxf(v, 0.f), 255.f)); }

// img[N, 3] u8 (equ.cu:181-187; row 0 is written too, as zeros)
__global__ void __launch_bounds__(256)
equ_to_u8_kernel(long long N, long long pitch, const float *__restrict__ x, uint8_t *__restrict__ img) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) img[i * 3 + ch] = clip_byte(x[ch * pitch + i]);
}

// Rows of X by index, packed [n, 3]: the data plane of the id-range sharded solver (fpie_b200/shard.py) -- what a
// rank sends are the rows its peers hold as ghosts, what it receives lands in its own ghost rows.  The reference's
// MPI EquSolver moves ALL of X through rank 0 instead (mpi/equ.cc:136-146).
__global__ void __launch_bounds__(256)
equ_gather_rows_kernel(long long n, long long N, long long pitch, const int32_t *__restrict__ idx,
                       const float *__restrict__ x, float *__restrict__ out, int *__restrict__ bad) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long i = idx[j];
  if (i < 0 || i >= N) {
    atomicExch(bad, 1);
    return;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) out[j * 3 + ch] = x[ch * pitch + i];
}

__global__ void __launch_bounds__(256)
equ_scatter_rows_kernel(long long n, long long N, long long pitch, const int32_t *__restrict__ idx,
                        const float *__restrict__ in, float *__restrict__ x, int *__restrict__ bad) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long i = idx[j];
  if (i <= 0 || i >= N) {  // (row 0 is the constant: never a ghost)
    atomicExch(bad, 1);
    return;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) x[ch * pitch + i] = in[j * 3 + ch];
}

// ---------------------------------------------------------------------------
// fused Processor-level reset: A / X / B straight from uint8 images
// (fpie/process.py:227-266)
// ---------------------------------------------------------------------------
__global__ void crop_flags_kernel(BlendImages b, int32_t *__restrict__ flags, uint8_t *__restrict__ canvas) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)b.n * b.m) return;
  const int i = (int)(idx / b.m), j = (int)(idx % b.m);
  flags[idx] = canonical_mask_at(b, b.x0 + i, b.y0 + j) ? 1 : 0;
  const long long tp = ((long long)(b.h1 + b.x0 + i) * b.tw + (b.w1 + b.y0 + j)) * 3;
  canvas[idx * 3 + 0] = b.tgt[tp + 0];
  canvas[idx * 3 + 1] = b.tgt[tp + 1];
  canvas[idx * 3 + 2] = b.tgt[tp + 2];
}

__global__ void __launch_bounds__(256)
equ_build_kernel(BlendImages b, long long pitch, const int32_t *__restrict__ flags, const int32_t *__restrict__ ids,
                 int4 *__restrict__ A, float *__restrict__ X0, float *__restrict__ X1, float *__restrict__ B,
                 int32_t *__restrict__ pix) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)b.n * b.m) return;
  if (!flags[idx]) return;
  const int i = (int)(idx / b.m), j = (int)(idx % b.m);
  const int id = ids[idx];
  // masked pixels are never on the crop frame, so the four neighbours exist
  const long long nb[4] = {idx - b.m, idx + b.m, idx - 1, idx + 1};
  const int di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
  int a[4];
  bool out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    out[k] = flags[nb[k]] == 0;
    a[k] = out[k] ? 0 : ids[nb[k]];
  }
  A[id] = make_int4(a[0], a[1], a[2], a[3]);
  pix[id - 1] = (int32_t)idx;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float bv = pixel_gradient(b, i, j, ch);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (out[k]) bv += target_at(b, i + di[k], j + dj[k], ch);  // Dirichlet values folded into B
    const float xv = target_at(b, i, j, ch);
    B[ch * pitch + id] = bv;
    X0[ch * pitch + id] = xv;
    X1[ch * pitch + id] = xv;
  }
}

// canvas[pix[i-1]] = u8(X[i])   (process.py:278)
__global__ void __launch_bounds__(256)
equ_paste_kernel(long long K, long long pitch, const float *__restrict__ x, const int32_t *__restrict__ pix,
                 uint8_t *__restrict__ canvas) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= K) return;
  const long long p = pix[i];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) canvas[p * 3 + ch] = clip_byte(x[ch * pitch + i + 1]);
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
static int blocks_for(long long work, int threads) { return (int)ceil_div(work, threads); }

EquSolver::EquSolver(int device, cudaStream_t stream, int block_size) : device_(device), stream_(stream) {
  const char *no_graph = getenv("FPIE_B200_NO_GRAPH");
  graph_off_ = no_graph && no_graph[0] && no_graph[0] != '0';
  const char *no_d16 = getenv("FPIE_B200_NO_DELTA16");  // A/B: keep the 8-byte (up, down) table and fp32 B
  no_delta16_ = no_d16 && no_d16[0] && no_d16[0] != '0';
  // The 4-byte table pays off when the sweep streams from HBM (config 3: +10 %); an L2-resident system is faster
  // with one unknown per thread (config 1: 5.6 vs 7.1 us per sweep), so small systems keep the 8-byte table.
  const char *d16_pipe = getenv("FPIE_B200_D16_PIPE");  // A/B: 0 = one launch-wide pass, N = persistent with N CTAs per SM
  if (d16_pipe && d16_pipe[0]) {
    d16_pipe_ = d16_pipe[0] != '0';
    if (d16_pipe_) d16_ctas_per_sm_ = std::max(1, atoi(d16_pipe));
  }
  const char *d16_min = getenv("FPIE_B200_DELTA16_MIN");
  delta16_min_ = d16_min && d16_min[0] ? atoll(d16_min) : (1ll << 21);
  int count = 0;
  CUDA_CHECK(cudaGetDeviceCount(&count));
  FPIE_REQUIRE(device >= 0 && device < count, "fpie_b200: no such CUDA device");
  DeviceGuard guard(device_);
  cudaDeviceProp prop{};
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device_));
  FPIE_REQUIRE(prop.major >= 10, "fpie_b200 is built for sm_100a (Blackwell) only");
  sm_count_ = prop.multiProcessorCount;
  // the reference's -z flag (fpie/args.py block-size, default 1024); we accept
  // any multiple of 32 up to 1024 and default to 256
  if (block_size >= 100000) {  // 100000 + z: block size z, always the generic int4 table (cross-checks)
    force_generic_ = true;
    block_size -= 100000;
  }
  if (block_size <= 0) block_size = 256;
  block_ = std::min(1024, std::max(32, block_size / 32 * 32));
  err_.resize(3);
  flag_.resize(1);
  CUDA_CHECK(cudaMallocHost(&host_err_, 3 * sizeof(double)));
  CUDA_CHECK(cudaMallocHost(&host_flag_, sizeof(int)));
}

EquSolver::~EquSolver() {
  int prev = -1;  // (see GridSolver::~GridSolver)
  cudaGetDevice(&prev);
  cudaSetDevice(device_);
  drop_graphs();
  upload_.destroy_stream();
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  if (host_err_) cudaFreeHost(host_err_);
  if (host_flag_) cudaFreeHost(host_flag_);
  if (prev >= 0 && prev != device_) cudaSetDevice(prev);
}

void EquSolver::require_ready() const { FPIE_REQUIRE(ready_, "EquSolver: step/state called before reset"); }

void EquSolver::scan_ids(const int32_t *dev_mask, int64_t count, int32_t *dev_ids) {
  const int nblocks = (int)ceil_div(count, SCAN_CHUNK);
  block_sums_.resize(std::max(nblocks, 1));
  scan_count_kernel<<<nblocks, SCAN_THREADS, 0, stream_>>>(dev_mask, count, block_sums_.ptr);
  scan_block_sums_kernel<<<1, SCAN_THREADS, 0, stream_>>>(block_sums_.ptr, nblocks);
  scan_write_kernel<<<nblocks, SCAN_THREADS, 0, stream_>>>(dev_mask, count, block_sums_.ptr, dev_ids);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 3;
}

// ids of a contiguous device mask [n, m]: row-major running count (Jacobi mode) or the
// reference OpenMP backend's odd-then-even labelling (red-black mode).  Sets n_mid_.
void EquSolver::label(const int32_t *dev_mask, int n, int m, int32_t *dev_ids) {
  const long long count = (long long)n * m;
  if (mode_ == 0) {
    scan_ids(dev_mask, count, dev_ids);
    n_mid_ = 0;
    return;
  }
  rb_tmp_.resize((size_t)count * 3);
  int32_t *flags = rb_tmp_.ptr, *odd = rb_tmp_.ptr + count, *even = rb_tmp_.ptr + 2 * count;
  rb_flags_kernel<<<blocks_for(count, 256), 256, 0, stream_>>>(dev_mask, count, m, 1, flags);
  scan_ids(flags, count, odd);
  rb_flags_kernel<<<blocks_for(count, 256), 256, 0, stream_>>>(dev_mask, count, m, 0, flags);
  scan_ids(flags, count, even);
  rb_combine_kernel<<<blocks_for(count, 256), 256, 0, stream_>>>(dev_mask, count, m, odd, even, dev_ids);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 3;
  int32_t n_odd = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_odd, odd + (count - 1), 4, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  n_mid_ = (int64_t)n_odd + 1;
}

void EquSolver::set_mode(int mode) {
  FPIE_REQUIRE(mode >= 0 && mode <= 2, "EquSolver mode must be 0 (Jacobi), 1 (red-black Gauss-Seidel) or 2 (Jacobi, gather only)");
  mode_ = (mode == 1) ? 1 : 0;
  no_promote_ = (mode == 2);
  part_n_ = part_m_ = 0;
  promoted_ = false;
  ready_ = false;
}

void EquSolver::partition(int n, int m, const int32_t *mask, int64_t mask_rs, int64_t mask_cs, int32_t *out_ids) {
  FPIE_REQUIRE(n >= 1 && m >= 1 && mask && out_ids, "partition: bad arguments");
  FPIE_REQUIRE(mask_cs == 1 && mask_rs >= m, "partition: mask rows must be contiguous (column stride 1)");
  FPIE_REQUIRE((int64_t)n * m < (int64_t)1 << 31, "partition: more than 2^31 pixels");
  DeviceGuard guard(device_);
  if (promoted_) {  // the promoted state indexes through ids_: bring it home before they are overwritten
    pull_tiled_state();
    promoted_ = false;
  }
  const size_t count = (size_t)n * m;
  istage_.resize(count);
  ids_.resize(count);
  CUDA_CHECK(cudaMemcpy2DAsync(istage_.ptr, (size_t)m * 4, mask, (size_t)mask_rs * 4, (size_t)m * 4, n,
                               cudaMemcpyHostToDevice, stream_));
  label(istage_.ptr, n, m, ids_.ptr);
  part_n_ = n;
  part_m_ = m;
  CUDA_CHECK(cudaMemcpyAsync(out_ids, ids_.ptr, count * 4, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void EquSolver::allocate(int64_t N) {
  drop_graphs();  // (they hold the old buffers and sizes)
  N_ = N;
  pitch_ = round_up(N, 32);
  A_.resize((size_t)N);
  for (auto &b : X_) b.resize((size_t)pitch_ * 3);
  B_.resize((size_t)pitch_ * 3);
  img_.resize((size_t)N * 3);
  cur_ = 0;
  win_lo_ = 0;
  win_hi_ = N;
  rows_checked_ = false;
}

void EquSolver::reset(int64_t N, const int32_t *A, const float *X, const float *B) {
  FPIE_REQUIRE(N >= 1 && A && X && B, "EquSolver.reset: bad arguments");
  FPIE_REQUIRE(N < ((int64_t)1 << 31), "EquSolver.reset: ids are int32");
  DeviceGuard guard(device_);
  ready_ = false;
  fused_ = false;
  promoted_ = false;
  allocate(N);
  stage_.resize((size_t)N * 3);
  CUDA_CHECK(cudaMemcpyAsync(A_.ptr, A, (size_t)N * 16, cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaMemsetAsync(flag_.ptr, 0, sizeof(int), stream_));
  check_index_kernel<<<blocks_for(N, 256), 256, 0, stream_>>>(N, A_.ptr, flag_.ptr);
  CUDA_CHECK(cudaMemcpyAsync(host_flag_, flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaMemcpyAsync(stage_.ptr, X, (size_t)N * 12, cudaMemcpyHostToDevice, stream_));
  rows3_to_planes_kernel<<<blocks_for(N, 256), 256, 0, stream_>>>(N, pitch_, stage_.ptr, X_[0].ptr, X_[1].ptr);
  CUDA_CHECK(cudaMemcpyAsync(stage_.ptr, B, (size_t)N * 12, cudaMemcpyHostToDevice, stream_));
  rows3_to_planes_kernel<<<blocks_for(N, 256), 256, 0, stream_>>>(N, pitch_, stage_.ptr, B_.ptr, nullptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 3;
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  FPIE_REQUIRE(*host_flag_ == 0, "EquSolver.reset: A holds an index outside [0, N)");
  stats_.unknowns = N - 1;
  compact_tables();
  promoted_ = false;
  bool zero_row = true;  // the grid embedding needs X[0] = B[0] = 0 and A[0] = 0 (the Processor's row 0)
  for (int k = 0; k < 3; ++k) zero_row = zero_row && X[k] == 0.f && B[k] == 0.f;
  for (int k = 0; k < 4; ++k) zero_row = zero_row && A[k] == 0;
  if (structured_ && zero_row && part_n_ > 0) {
    // does A match the mask this solver labelled in partition()?  (the reference flow: process.py:187, 270)
    const long long px = (long long)part_n_ * part_m_;
    int32_t last = 0;
    CUDA_CHECK(cudaMemcpyAsync(&last, ids_.ptr + (px - 1), 4, cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaMemsetAsync(flag_.ptr, 0, sizeof(int), stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    if ((int64_t)last == N - 1) {
      check_grid_structure_kernel<<<blocks_for(px, 256), 256, 0, stream_>>>(part_n_, part_m_, istage_.ptr, ids_.ptr,
                                                                            A_.ptr, flag_.ptr);
      CUDA_CHECK(cudaGetLastError());
      stats_.launches += 1;
      CUDA_CHECK(cudaMemcpyAsync(host_flag_, flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream_));
      CUDA_CHECK(cudaStreamSynchronize(stream_));
      if (*host_flag_ == 0) try_promote(part_n_, part_m_);
    }
  }
  ready_ = true;
}

// Promotion: an Equ system whose ids are the row-major labels of a known crop IS a grid problem --
// state X on the masked pixels and 0 elsewhere, gradient B -- and
//   ((((B + X[up]) + X[down]) + X[left]) + X[right]) / 4
// is exactly what the grid kernel evaluates there (absent neighbours read the constant 0 in both
// formulations), so the temporally blocked kernel produces the same bits several times faster.
void EquSolver::try_promote(int n, int m) {
  if (mode_ != 0 || no_promote_) return;
  if (!tiled_) tiled_.reset(new GridSolver(device_, stream_, 0, 0));
  EquEmbed e{n, m, istage_.ptr, ids_.ptr, X_[cur_].ptr, B_.ptr, pitch_};
  tiled_->reset_from_equ(e);
  promoted_ = true;
  tiled_dirty_ = false;
}

void EquSolver::pull_tiled_state() {
  if (!promoted_ || !tiled_dirty_) return;
  EquEmbed e{part_n_, part_m_, istage_.ptr, ids_.ptr, X_[cur_].ptr, B_.ptr, pitch_};
  tiled_->export_to_equ(e);
  tiled_dirty_ = false;
}

void EquSolver::compact_tables() {
  structured_ = false;
  if (mode_ != 0) return;
  ud_.resize((size_t)N_);
  d16_.resize((size_t)pitch_);
  CUDA_CHECK(cudaMemsetAsync(d16_.ptr, 0, d16_.bytes(), stream_));
  b16_.resize((size_t)pitch_ * 3);
  CUDA_CHECK(cudaMemsetAsync(flag_.ptr, 0, sizeof(int), stream_));
  compact_index_kernel<<<blocks_for(N_, 256), 256, 0, stream_>>>(N_, A_.ptr, ud_.ptr, d16_.ptr, flag_.ptr);
  equ_b_to_half_kernel<<<blocks_for(pitch_ * 3, 256), 256, 0, stream_>>>(N_, pitch_, B_.ptr, b16_.ptr, flag_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 2;
  CUDA_CHECK(cudaMemcpyAsync(host_flag_, flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  structured_ = ((*host_flag_ & 1) == 0) && !force_generic_;
  delta16_ = structured_ && ((*host_flag_ & 2) == 0) && !no_delta16_ && N_ >= delta16_min_;
  b16_ok_ = delta16_ && ((*host_flag_ & 4) == 0);
}

void EquSolver::reset_from_images(const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh, int mw, int mc,
                                  const uint8_t *tgt, int th, int tw, int h0, int w0, int h1, int w1, int grad_mode,
                                  int64_t *out_n, int32_t *out_box4) {
  DeviceGuard guard(device_);
  ready_ = false;
  BlendUpload &up = upload_;  // device copies of the images are kept between resets (no malloc / free per call)
  up.upload(stream_, src, sh, sw, mask, mh, mw, mc, tgt, th, tw, h0, w0, h1, w1, grad_mode);
  const BlendImages &b = up.images();
  const long long count = (long long)b.n * b.m;
  if (count >= (1ll << 31)) up.wait_copies();  // (the caller's buffers are its own again when this call returns)
  FPIE_REQUIRE(count < (1ll << 31), "reset: crop has more than 2^31 pixels");
  istage_.resize((size_t)count);
  ids_.resize((size_t)count);
  canvas_.resize((size_t)count * 3);
  crop_flags_kernel<<<blocks_for(count, 256), 256, 0, stream_>>>(b, istage_.ptr, canvas_.ptr);
  CUDA_CHECK(cudaGetLastError());
  label(istage_.ptr, b.n, b.m, ids_.ptr);
  // K = number of masked pixels: the running count of the flags (the crop frame is never masked,
  // so the last pixel carries the total in Jacobi mode; red-black ids are not monotone -> count again)
  int32_t last = 0;
  if (mode_ == 0) {
    CUDA_CHECK(cudaMemcpyAsync(&last, ids_.ptr + (count - 1), 4, cudaMemcpyDeviceToHost, stream_));
  } else {
    scan_ids(istage_.ptr, count, rb_tmp_.ptr);
    CUDA_CHECK(cudaMemcpyAsync(&last, rb_tmp_.ptr + (count - 1), 4, cudaMemcpyDeviceToHost, stream_));
  }
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  const int64_t K = last;
  allocate(K + 1);
  pix_.resize((size_t)std::max<int64_t>(K, 1));
  // row 0 of A / X / B is the zero constant (process.py:248-250)
  CUDA_CHECK(cudaMemsetAsync(A_.ptr, 0, 16, stream_));
  for (int ch = 0; ch < 3; ++ch) {
    CUDA_CHECK(cudaMemsetAsync(X_[0].ptr + ch * pitch_, 0, 4, stream_));
    CUDA_CHECK(cudaMemsetAsync(X_[1].ptr + ch * pitch_, 0, 4, stream_));
    CUDA_CHECK(cudaMemsetAsync(B_.ptr + ch * pitch_, 0, 4, stream_));
  }
  equ_build_kernel<<<blocks_for(count, 256), 256, 0, stream_>>>(b, pitch_, istage_.ptr, ids_.ptr, A_.ptr, X_[0].ptr,
                                                               X_[1].ptr, B_.ptr, pix_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 2;
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  crop_n_ = b.n;
  crop_m_ = b.m;
  fused_ = true;
  stats_.unknowns = K;
  compact_tables();
  part_n_ = b.n;
  part_m_ = b.m;
  promoted_ = false;
  try_promote(b.n, b.m);
  ready_ = true;
  if (out_n) *out_n = K + 1;
  if (out_box4) {
    out_box4[0] = b.h1 + b.x0;
    out_box4[1] = b.h1 + b.x0 + b.n;
    out_box4[2] = b.w1 + b.y0;
    out_box4[3] = b.w1 + b.y0 + b.m;
  }
}

void EquSolver::sweeps_async(int iters) {
  require_ready();
  FPIE_REQUIRE(iters >= 0, "step: negative iteration count");
  DeviceGuard guard(device_);
  if (mode_ == 1) {
    // red-black Gauss-Seidel: odd ids [1, n_mid) then even ids [n_mid, N), in place (openmp/equ.cc:107-118)
    FPIE_REQUIRE(n_mid_ >= 1 && n_mid_ <= N_, "red-black mode needs ids from this solver's partition()");
    float *x = X_[cur_].ptr;
    for (int i = 0; i < iters; ++i) {
      if (n_mid_ > 1)
        equ_rb_half_kernel<<<blocks_for(n_mid_ - 1, block_), block_, 0, stream_>>>(1, n_mid_, pitch_, A_.ptr, B_.ptr, x);
      if (N_ > n_mid_)
        equ_rb_half_kernel<<<blocks_for(N_ - n_mid_, block_), block_, 0, stream_>>>(n_mid_, N_, pitch_, A_.ptr, B_.ptr, x);
    }
    stats_.launches += 2 * (int64_t)iters;
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (promoted_) {
    tiled_->sweeps_async(iters);
    tiled_dirty_ = tiled_dirty_ || iters > 0;
    return;
  }
  // One launch per sweep, chained by programmatic dependent launch (the next sweep's CTAs are resident and have
  // their table rows and B in registers when the previous sweep drains); long runs replay a captured graph of
  // kGraphSweeps sweeps, so that an L2-resident system (config 1: a sweep is a few microseconds) is not bound by
  // the host's launch rate.
  auto one_sweep = [&](cudaStream_t st, int &cur) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks_for(N_, block_));
    cfg.blockDim = dim3(block_);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const float *xin = X_[cur].ptr;
    float *xout = X_[cur ^ 1].ptr;
    const float *b = B_.ptr;
    if (delta16_) {
      const uint32_t *d16 = d16_.ptr;
      const __half *b16 = b16_.ptr;
      cfg.blockDim = dim3(256);
      cfg.gridDim = dim3((unsigned)blocks_for((N_ + 3) / 4, 256));
      static const int depth = getenv("FPIE_B200_D16_DEPTH") ? atoi(getenv("FPIE_B200_D16_DEPTH")) : 1;
      auto kern = d16_pipe_ ? (b16_ok_ ? (depth == 2 ? equ_sweep_d16p_kernel<true, 2> : equ_sweep_d16p_kernel<true, 1>)
                                       : (depth == 2 ? equ_sweep_d16p_kernel<false, 2> : equ_sweep_d16p_kernel<false, 1>))
                            : (b16_ok_ ? equ_sweep_d16_kernel<true> : equ_sweep_d16_kernel<false>);
      if (d16_pipe_) {  // persistent: d16_ctas_per_sm_ x 256 threads per SM stride over the system
        const long long per_sm = (long long)std::min(d16_ctas_per_sm_, b16_ok_ ? 4 : 3) * 256 / FPIE_D16_BLOCK;
        cfg.blockDim = dim3(FPIE_D16_BLOCK);
        cfg.gridDim = dim3((unsigned)std::min<long long>(blocks_for((N_ + 3) / 4, FPIE_D16_BLOCK), sm_count_ * std::max(1ll, per_sm)));
      }
      CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, (long long)N_, (long long)pitch_, d16, b, b16, xin, xout));
    } else if (structured_) {
      const int2 *ud = ud_.ptr;
      CUDA_CHECK(cudaLaunchKernelEx(&cfg, equ_sweep_lr_kernel, (long long)N_, (long long)pitch_, ud, b, xin, xout));
    } else {
      const int4 *a = A_.ptr;
      CUDA_CHECK(cudaLaunchKernelEx(&cfg, equ_sweep_kernel, (long long)N_, (long long)pitch_, a, b, xin, xout));
    }
    cur ^= 1;
  };
  int left = iters;
  if (!graph_off_ && left >= 2 * kGraphSweeps) {
    if (!graph_warm_) {  // (first launches outside a capture: lazy kernel loading)
      for (int i = 0; i < 2; ++i) one_sweep(stream_, cur_);
      left -= 2;
      graph_warm_ = true;
    }
    if (!graph_[cur_]) {
      if (!cap_stream_) CUDA_CHECK(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
      int c = cur_;
      CUDA_CHECK(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal));
      cudaGraph_t captured = nullptr;
      try {
        for (int i = 0; i < kGraphSweeps; ++i) one_sweep(cap_stream_, c);
      } catch (...) {
        cudaStreamEndCapture(cap_stream_, &captured);
        if (captured) cudaGraphDestroy(captured);
        throw;
      }
      CUDA_CHECK(cudaStreamEndCapture(cap_stream_, &captured));
      const cudaError_t rc = cudaGraphInstantiate(&graph_[cur_], captured, 0);
      cudaGraphDestroy(captured);
      CUDA_CHECK(rc);
    }
    while (left >= kGraphSweeps) {  // (an even number of sweeps: the graph starts and ends on the same buffer)
      CUDA_CHECK(cudaGraphLaunch(graph_[cur_], stream_));
      left -= kGraphSweeps;
    }
  }
  for (; left > 0; --left) one_sweep(stream_, cur_);
  stats_.launches += iters;
  CUDA_CHECK(cudaGetLastError());
}

void EquSolver::drop_graphs() {
  for (auto &gx : graph_) {
    if (gx) cudaGraphExecDestroy(gx);
    gx = nullptr;
  }
  graph_warm_ = false;
}

void EquSolver::finish_async() {
  require_ready();
  DeviceGuard guard(device_);
  pull_tiled_state();
  CUDA_CHECK(cudaMemsetAsync(err_.ptr, 0, 3 * sizeof(double), stream_));
  equ_residual_kernel<<<blocks_for(std::max<int64_t>(win_hi_ - win_lo_, 1), 256), 256, 0, stream_>>>(
      win_lo_, win_hi_, pitch_, A_.ptr, B_.ptr, X_[cur_].ptr, err_.ptr);
  equ_to_u8_kernel<<<blocks_for(N_, 256), 256, 0, stream_>>>(N_, pitch_, X_[cur_].ptr, img_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 2;
  CUDA_CHECK(cudaMemcpyAsync(host_err_, err_.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
}

void EquSolver::sync() {
  DeviceGuard guard(device_);
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void EquSolver::fetch(uint8_t *out_img, float *out_err3) {
  require_ready();
  DeviceGuard guard(device_);
  if (out_img) CUDA_CHECK(cudaMemcpyAsync(out_img, img_.ptr, (size_t)N_ * 3, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (out_err3)
    for (int c = 0; c < 3; ++c) out_err3[c] = (float)host_err_[c];
}

void EquSolver::set_window(int64_t lo, int64_t hi) {
  require_ready();
  FPIE_REQUIRE(lo >= 0 && lo <= hi && hi <= N_, "set_window: rows outside [0, N]");
  win_lo_ = lo;
  win_hi_ = hi;
}

void EquSolver::fetch_rows(int64_t lo, int64_t hi, uint8_t *out_img, float *out_err3) {
  require_ready();
  FPIE_REQUIRE(lo >= 0 && lo <= hi && hi <= N_, "fetch_rows: rows outside [0, N]");
  DeviceGuard guard(device_);
  if (out_img && hi > lo)
    CUDA_CHECK(cudaMemcpyAsync(out_img, img_.ptr + lo * 3, (size_t)(hi - lo) * 3, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (out_err3)
    for (int c = 0; c < 3; ++c) out_err3[c] = (float)host_err_[c];
}

void EquSolver::gather_rows(const int32_t *dev_idx, int64_t n, float *dev_out) {
  require_ready();
  FPIE_REQUIRE(n >= 0 && (n == 0 || (dev_idx && dev_out)), "gather_rows: bad arguments");
  if (n == 0) return;
  DeviceGuard guard(device_);
  pull_tiled_state();
  CUDA_CHECK(cudaMemsetAsync(flag_.ptr, 0, sizeof(int), stream_));
  equ_gather_rows_kernel<<<blocks_for(n, 256), 256, 0, stream_>>>(n, N_, pitch_, dev_idx, X_[cur_].ptr, dev_out, flag_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
  check_rows_flag("gather_rows");
}

void EquSolver::scatter_rows(const int32_t *dev_idx, int64_t n, const float *dev_in) {
  require_ready();
  FPIE_REQUIRE(n >= 0 && (n == 0 || (dev_idx && dev_in)), "scatter_rows: bad arguments");
  FPIE_REQUIRE(!promoted_ && mode_ == 0, "scatter_rows: only the index-mapped Jacobi path takes rows from outside");
  if (n == 0) return;
  DeviceGuard guard(device_);
  CUDA_CHECK(cudaMemsetAsync(flag_.ptr, 0, sizeof(int), stream_));
  equ_scatter_rows_kernel<<<blocks_for(n, 256), 256, 0, stream_>>>(n, N_, pitch_, dev_idx, dev_in, X_[cur_].ptr, flag_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
  check_rows_flag("scatter_rows");
}

// the index check of gather / scatter is read back lazily: at the next call, and by sync() / fetch()
void EquSolver::check_rows_flag(const char *who) {
  if (!rows_checked_) {  // the first call with a new index list is checked synchronously
    CUDA_CHECK(cudaMemcpyAsync(host_flag_, flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    FPIE_REQUIRE(*host_flag_ == 0, std::string(who) + ": an index is outside the system");
  }
}

// Sweep until every channel's residual is <= tol (checked every `check_every` sweeps) or `max_iters`
// sweeps have run (red-black mode: one "sweep" = both half sweeps, as in step()).
int EquSolver::solve(int max_iters, int check_every, float tol, float *out_err3) {
  require_ready();
  FPIE_REQUIRE(max_iters >= 0 && check_every >= 1, "solve: max_iters must be >= 0 and check_every >= 1");
  int done = 0;
  while (true) {
    finish_async();
    sync();
    const double worst = std::max(host_err_[0], std::max(host_err_[1], host_err_[2]));
    if (worst <= (double)tol || done >= max_iters) break;
    const int s = std::min(check_every, max_iters - done);
    sweeps_async(s);
    done += s;
  }
  if (out_err3)
    for (int c = 0; c < 3; ++c) out_err3[c] = (float)host_err_[c];
  return done;
}

void EquSolver::step(int iters, uint8_t *out_img, float *out_err3) {
  sweeps_async(iters);
  finish_async();
  fetch(out_img, out_err3);
}

void EquSolver::step_paste(int iters, uint8_t *out_crop, float *out_err3, int64_t row_stride) {
  require_ready();
  FPIE_REQUIRE(fused_, "step_paste needs a solver reset with reset_from_images");
  DeviceGuard guard(device_);
  sweeps_async(iters);
  pull_tiled_state();
  CUDA_CHECK(cudaMemsetAsync(err_.ptr, 0, 3 * sizeof(double), stream_));
  equ_residual_kernel<<<blocks_for(std::max<int64_t>(win_hi_ - win_lo_, 1), 256), 256, 0, stream_>>>(
      win_lo_, win_hi_, pitch_, A_.ptr, B_.ptr, X_[cur_].ptr, err_.ptr);
  const int64_t K = N_ - 1;
  if (K > 0)
    equ_paste_kernel<<<blocks_for(K, 256), 256, 0, stream_>>>(K, pitch_, X_[cur_].ptr, pix_.ptr, canvas_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 2;
  CUDA_CHECK(cudaMemcpyAsync(host_err_, err_.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  const size_t row_bytes = (size_t)crop_m_ * 3;
  if (row_stride <= 0) row_stride = (int64_t)row_bytes;
  FPIE_REQUIRE((size_t)row_stride >= row_bytes, "step_paste: destination row stride is smaller than a row");
  if (out_crop)
    CUDA_CHECK(cudaMemcpy2DAsync(out_crop, (size_t)row_stride, canvas_.ptr, row_bytes, row_bytes, crop_n_,
                                 cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (out_err3)
    for (int c = 0; c < 3; ++c) out_err3[c] = (float)host_err_[c];
}

void EquSolver::state(float *out) {
  require_ready();
  FPIE_REQUIRE(out, "state: null output");
  DeviceGuard guard(device_);
  pull_tiled_state();
  stage_.resize((size_t)N_ * 3);
  planes_to_rows3_kernel<<<blocks_for(N_, 256), 256, 0, stream_>>>(N_, pitch_, X_[cur_].ptr, stage_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
  CUDA_CHECK(cudaMemcpyAsync(out, stage_.ptr, (size_t)N_ * 12, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void EquSolver::system(int32_t *out_A, float *out_X, float *out_B) {
  require_ready();
  DeviceGuard guard(device_);
  pull_tiled_state();
  stage_.resize((size_t)N_ * 3);
  if (out_A) CUDA_CHECK(cudaMemcpyAsync(out_A, A_.ptr, (size_t)N_ * 16, cudaMemcpyDeviceToHost, stream_));
  if (out_X) {
    planes_to_rows3_kernel<<<blocks_for(N_, 256), 256, 0, stream_>>>(N_, pitch_, X_[cur_].ptr, stage_.ptr);
    CUDA_CHECK(cudaMemcpyAsync(out_X, stage_.ptr, (size_t)N_ * 12, cudaMemcpyDeviceToHost, stream_));
  }
  if (out_B) {
    planes_to_rows3_kernel<<<blocks_for(N_, 256), 256, 0, stream_>>>(N_, pitch_, B_.ptr, stage_.ptr);
    CUDA_CHECK(cudaMemcpyAsync(out_B, stage_.ptr, (size_t)N_ * 12, cudaMemcpyDeviceToHost, stream_));
  }
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

}  // namespace fpie
