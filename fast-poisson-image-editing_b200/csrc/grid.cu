// GridSolver kernels + host driver (sm_100a).
//
// Replaces fpie/core/cuda/grid.cu (reference) with a different design:
//   * planar, padded fp32 state (3 channel planes), 1-bit mask, quarter-scaled
//     gradient planes -- the interleaved [n,m,3] layout exists only at the API;
//   * true Jacobi (ping-pong buffers), bit-identical to fpie/np_solver.py:81-88;
//   * temporally blocked sweeps: a CTA keeps a 128-wide register tile of one
//     plane, runs k sweeps on it with a (k)-deep redundant halo and writes the
//     interior back, so HBM is touched once per k sweeps;
//   * tiles without masked pixels are never scheduled, fully masked tiles run
//     a select-free instruction stream.

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "grid_solver.cuh"
#include "prep.cuh"
#include "tma.cuh"

#ifndef FPIE_SWEEP_UNROLL
#define FPIE_SWEEP_UNROLL 4
#endif

namespace fpie {

// copies of the sweep body in the tile loop: the register rotation at the loop edge is paid once per four sweeps
constexpr int kSweepUnroll = FPIE_SWEEP_UNROLL;
#ifndef FPIE_PULL_AHEAD
#define FPIE_PULL_AHEAD 4
#endif
constexpr int kPullAhead = FPIE_PULL_AHEAD;  // rows before the strip's end at which the neighbours' edge rows are pulled
#ifndef FPIE_PULL_TALL
#define FPIE_PULL_TALL 8  // ... for strips of 18 rows and more
#endif

// ---------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------

// int32 mask [n, m] (device, contiguous) -> 1 bit per pixel in the padded
// geometry; the outer frame is forced to 0 (fpie/process.py:342-351 guarantees
// it; base_solver.h:101-140 does not check).  One warp per output word.
__global__ void pack_mask_kernel(PlaneGeom g, const int32_t *__restrict__ mask, uint32_t *__restrict__ bits,
                                 unsigned long long *__restrict__ count) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long words = (long long)g.rows * g.wpitch;
  if (warp >= words) return;
  const int prow = (int)(warp / g.wpitch);
  const int pcol = (int)(warp % g.wpitch) * 32 + lane;
  const int r = prow - g.padr, c = pcol - g.padc;
  bool on = false;
  if (r > 0 && r < g.n - 1 && c > 0 && c < g.m - 1) on = mask[(long long)r * g.m + c] != 0;
  const uint32_t word = __ballot_sync(0xffffffffu, on);
  if (lane == 0) {
    bits[warp] = word;
    if (word) atomicAdd(count, (unsigned long long)__popc(word));
  }
}

// interleaved fp32 [n, m, 3] -> three planes, scaled (1 for the state, 0.25
// for the gradient); optionally mirrored into a second destination.
__global__ void aos_to_planes_kernel(PlaneGeom g, const float *__restrict__ aos, float scale, float *__restrict__ dst0,
                                     float *__restrict__ dst1) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)g.n * g.m;
  if (idx >= total) return;
  const int r = (int)(idx / g.m), c = (int)(idx % g.m);
  const long long off = (long long)(r + g.padr) * g.pitch + (c + g.padc);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float v = aos[idx * 3 + ch] * scale;
    dst0[ch * g.plane + off] = v;
    if (dst1) dst1[ch * g.plane + off] = v;
  }
}

// fp32 quarter-gradient planes -> fp16 copy for the temporally blocked kernel.
// *inexact is raised if any value does not survive the round trip (then the
// fp32 planes are streamed instead).  Gradients built from uint8 images are
// multiples of 1/8 below 256 and always survive.
__global__ void planes_to_half_kernel(long long count, const float *__restrict__ src, __half *__restrict__ dst,
                                      int *__restrict__ inexact) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float v = src[i];
  const __half h = __float2half_rn(v);
  dst[i] = h;
  if (!(__half2float(h) == v)) atomicOr(inexact, 1);
}

__global__ void planes_to_aos_kernel(PlaneGeom g, const float *__restrict__ src, float *__restrict__ aos) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)g.n * g.m;
  if (idx >= total) return;
  const int r = (int)(idx / g.m), c = (int)(idx % g.m);
  const long long off = (long long)(r + g.padr) * g.pitch + (c + g.padc);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) aos[idx * 3 + ch] = src[ch * g.plane + off];
}

// ---------------------------------------------------------------------------
// one Jacobi sweep per launch (fallback / cross-check path, variant 1)
// thread = one 4-pixel group of one plane; groups without masked pixels exit
// before touching the state (both ping-pong buffers hold the same constants).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_sweep1_kernel(PlaneGeom g, int nplanes, const uint32_t *__restrict__ bits, const float *__restrict__ xin,
                   float *__restrict__ xout, const float *__restrict__ hq) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long per_plane = (long long)g.n * g.groups;
  if (idx >= per_plane * nplanes) return;
  const int p = (int)(idx / per_plane);
  const long long q = idx % per_plane;
  const int r = (int)(q / g.groups), grp = (int)(q % g.groups);
  const int pc = g.padc + 4 * grp;
  const uint32_t nib = (bits[(long long)(r + g.padr) * g.wpitch + (pc >> 5)] >> (pc & 31)) & 0xFu;
  if (!nib) return;
  const long long off = (long long)p * g.plane + (long long)(r + g.padr) * g.pitch + pc;
  const float4 c = ld4(xin + off), u = ld4(xin + off - g.pitch), d = ld4(xin + off + g.pitch), h = ld4(hq + off);
  const float lf = xin[off - 1], rt = xin[off + 4];
  float4 o;
  o.x = (nib & 1u) ? jacobi_q(h.x, u.x, d.x, lf, c.y) : c.x;
  o.y = (nib & 2u) ? jacobi_q(h.y, u.y, d.y, c.x, c.z) : c.y;
  o.z = (nib & 4u) ? jacobi_q(h.z, u.z, d.z, c.y, c.w) : c.z;
  o.w = (nib & 8u) ? jacobi_q(h.w, u.w, d.w, c.z, rt) : c.w;
  st4(xout + off, o);
}

// ---------------------------------------------------------------------------
// temporally blocked sweeps: nsweeps (<= K) Jacobi sweeps per pass over HBM.
//
// Tile = (R * NW) rows x 128 cols of one plane, held in registers: thread
// (warp w, lane l) owns rows [w*R, (w+1)*R) x cols [4l, 4l+4).  Per sweep a
// thread needs the row above / below its strip (other warps: exchanged through
// a 2-deep shared-memory mailbox, one __syncthreads per sweep) and the
// columns left / right of it (other lanes: warp shuffles).  Values within s
// pixels of the tile edge are wrong after s sweeps; the tile origin is chosen
// so that only the inner (TH-2K) x (128-2HK) region is stored.
// ---------------------------------------------------------------------------
template <int R, bool MIXED>
__device__ __forceinline__ void tile_sweep(float4 (&x)[R], const float4 (&h)[R], const uint32_t (&mb)[(R + 7) / 8],
                                           float4 up, float4 dn) {
  float4 prev = up;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float4 cur = x[i];
    const float4 nxt = (i + 1 < R) ? x[i + 1] : dn;
    const float lf = __shfl_up_sync(0xffffffffu, cur.w, 1);
    const float rt = __shfl_down_sync(0xffffffffu, cur.x, 1);
    float4 o;
    o.x = jacobi_q(h[i].x, prev.x, nxt.x, lf, cur.y);
    o.y = jacobi_q(h[i].y, prev.y, nxt.y, cur.x, cur.z);
    o.z = jacobi_q(h[i].z, prev.z, nxt.z, cur.y, cur.w);
    o.w = jacobi_q(h[i].w, prev.w, nxt.w, cur.z, rt);
    if (MIXED) {
      const uint32_t nib = mb[i / 8] >> ((i % 8) * 4);
      o.x = (nib & 1u) ? o.x : cur.x;
      o.y = (nib & 2u) ? o.y : cur.y;
      o.z = (nib & 4u) ? o.z : cur.z;
      o.w = (nib & 8u) ? o.w : cur.w;
    }
    x[i] = o;
    prev = cur;
  }
}

// sweeps + store of one register tile (shared by the direct-load and the
// TMA-pipelined kernels).  `parity` carries the mailbox double-buffer phase
// across tiles: there is exactly one __syncthreads per sweep.
template <int R, int NW>
__device__ __forceinline__ void tile_run(float4 (&x)[R], const float4 (&h)[R], uint32_t (&mb)[(R + 7) / 8], bool full,
                                         float4 (*mailbox)[2][NW][32], int &parity, int nsweeps, int halo_y,
                                         int halo_x, float *__restrict__ out, int pitch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int TH = R * NW;
  for (int s = 0; s < nsweeps; ++s) {
    mailbox[parity][0][w][lane] = x[0];
    mailbox[parity][1][w][lane] = x[R - 1];
    __syncthreads();
    const float4 up = (w > 0) ? mailbox[parity][1][w - 1][lane] : x[0];
    const float4 dn = (w + 1 < NW) ? mailbox[parity][0][w + 1][lane] : x[R - 1];
    parity ^= 1;
    if (full)
      tile_sweep<R, false>(x, h, mb, up, dn);
    else
      tile_sweep<R, true>(x, h, mb, up, dn);
  }
  // store the inner region; groups without masked pixels keep their constants
  const bool lane_in = (4 * lane >= halo_x) && (4 * lane < TILE_W - halo_x);
  if (lane_in) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int tr = w * R + i;
      const uint32_t nib = (mb[i / 8] >> ((i % 8) * 4)) & 0xFu;
      if (tr >= halo_y && tr < TH - halo_y && nib) st4(out + (long long)i * pitch, x[i]);
    }
  }
}

template <int R>
__device__ __forceinline__ void load_mask_bits(uint32_t (&mb)[(R + 7) / 8], bool full,
                                               const uint32_t *__restrict__ bits, int wpitch, int prow0, int pcol) {
#pragma unroll
  for (int i = 0; i < (R + 7) / 8; ++i) mb[i] = full ? 0xffffffffu : 0u;
  if (!full) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const uint32_t word = bits[(long long)(prow0 + i) * wpitch + (pcol >> 5)];
      mb[i / 8] |= ((word >> (pcol & 31)) & 0xFu) << ((i % 8) * 4);
    }
  }
}

// Tile descriptors are packed into 8 bytes so that the two descriptors a CTA
// keeps in flight cost four registers:  x = plane row | flags << 28,
// y = plane col | plane << 20.
__host__ __device__ __forceinline__ int2 pack_tile(int prow, int pcol, int plane, int flags) {
  return make_int2(prow | (flags << 28), pcol | (plane << 20));
}
struct TileRef {
  int prow, pcol, plane;
  bool full;
};
__device__ __forceinline__ TileRef unpack_tile(int2 d) {
  TileRef t;
  t.prow = d.x & 0x0fffffff;
  t.full = ((d.x >> 28) & 1) != 0;
  t.pcol = d.y & 0xfffff;
  t.plane = (d.y >> 20) & 0xfff;
  return t;
}

// direct-load variant: registers are filled straight from global memory
template <int R, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
grid_sweepk_kernel(PlaneGeom g, const uint32_t *__restrict__ bits, const float *__restrict__ xin,
                   float *__restrict__ xout, const float *__restrict__ hq, const int2 *__restrict__ tiles, int ntiles,
                   int nsweeps, int halo_y, int halo_x) {
  __shared__ float4 mailbox[2][2][NW][32];  // [sweep parity][0 = top row, 1 = bottom row][warp][lane]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int parity = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const TileRef td = unpack_tile(tiles[t]);
    const int prow0 = td.prow + w * R;
    const int pcol = td.pcol + 4 * lane;
    const long long base = (long long)td.plane * g.plane + (long long)prow0 * g.pitch + pcol;
    const bool full = td.full;
    float4 x[R], h[R];
    uint32_t mb[(R + 7) / 8];
#pragma unroll
    for (int i = 0; i < R; ++i) x[i] = ld4(xin + base + (long long)i * g.pitch);
#pragma unroll
    for (int i = 0; i < R; ++i) h[i] = ld4(hq + base + (long long)i * g.pitch);
    load_mask_bits<R>(mb, full, bits, g.wpitch, prow0, pcol);
    tile_run<R, NW>(x, h, mb, full, mailbox, parity, nsweeps, halo_y, halo_x, xout + base, g.pitch);
  }
}

// One row of the register tile: 4 pixels, neighbours left/right by shuffle.
template <bool MIXED>
__device__ __forceinline__ void row_update(float4 &xi, const float4 &hi, const float4 &prev, const float4 &nxt,
                                           uint32_t nib) {
  const float4 cur = xi;
  const float lf = __shfl_up_sync(0xffffffffu, cur.w, 1);
  const float rt = __shfl_down_sync(0xffffffffu, cur.x, 1);
  float4 o;
  o.x = jacobi_q(hi.x, prev.x, nxt.x, lf, cur.y);
  o.y = jacobi_q(hi.y, prev.y, nxt.y, cur.x, cur.z);
  o.z = jacobi_q(hi.z, prev.z, nxt.z, cur.y, cur.w);
  o.w = jacobi_q(hi.w, prev.w, nxt.w, cur.z, rt);
  if (MIXED) {
    o.x = (nib & 1u) ? o.x : cur.x;
    o.y = (nib & 2u) ? o.y : cur.y;
    o.z = (nib & 4u) ? o.z : cur.z;
    o.w = (nib & 8u) ? o.w : cur.w;
  }
  xi = o;
}

// One sweep of a register tile with a split-phase edge exchange: publish the
// strip's first / last row, arrive on an mbarrier, update the R-2 interior rows
// (they need no other warp), and only then wait for the neighbours' rows to
// finish rows 0 and R-1.  The barrier latency hides behind (R-2)/R of the work.
template <int R, int NW, bool MIXED>
__device__ __forceinline__ void tile_sweep_split(float4 (&x)[R], const float4 (&h)[R],
                                                 const uint32_t (&mb)[(R + 7) / 8], float4 (*mailbox)[2][NW][32],
                                                 uint64_t *mail_bar, int parity, uint32_t &mphase) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  mailbox[parity][0][w][lane] = x[0];
  mailbox[parity][1][w][lane] = x[R - 1];
  __syncwarp();
  if (lane == 0) mbar_arrive(&mail_bar[parity]);
  const float4 first_old = x[1];
  float4 prev = x[0];
  float4 up, dn;
  // the neighbours' rows are pulled a few interior rows before they are needed, so that the
  // barrier check and the shared-memory latency hide behind the remaining interior rows
  constexpr int PULL_ROW = (R >= 18) ? R - FPIE_PULL_TALL : (R >= 8) ? R - kPullAhead : R - 2;
#pragma unroll
  for (int i = 1; i < R - 1; ++i) {
    if (i == PULL_ROW) {
      mbar_wait(&mail_bar[parity], (mphase >> parity) & 1u);
      mphase ^= 1u << parity;
      up = (w > 0) ? mailbox[parity][1][w - 1][lane] : x[0];
      dn = (w + 1 < NW) ? mailbox[parity][0][w + 1][lane] : x[R - 1];
    }
    const float4 cur = x[i];
    row_update<MIXED>(x[i], h[i], prev, x[i + 1], mb[i / 8] >> ((i % 8) * 4));
    prev = cur;
  }
  row_update<MIXED>(x[0], h[0], up, first_old, mb[0]);
  row_update<MIXED>(x[R - 1], h[R - 1], prev, dn, mb[(R - 1) / 8] >> (((R - 1) % 8) * 4));
}

// Shared-memory layout of the pipelined kernel (all sections 128-byte aligned).
template <int R, int NW, bool H16>
struct PipeSmem {
  static constexpr uint32_t align128(uint32_t v) { return (v + 127u) & ~127u; }
  static constexpr int TH = R * NW;
  static constexpr int H16_W = TILE_W + 8;  // fp16 box: start rounded down to 8 columns, 8 columns wider
  static constexpr uint32_t X_BYTES = TH * TILE_W * 4;
  static constexpr uint32_t H_BYTES = H16 ? TH * H16_W * 2 : X_BYTES;
  static constexpr uint32_t M_BYTES = TH * MASK_BOX_WORDS * 4;
  static constexpr uint32_t X_OFF = 0;
  static constexpr uint32_t H_OFF = align128(X_OFF + X_BYTES);
  static constexpr uint32_t M_OFF = align128(H_OFF + H_BYTES);
  static constexpr uint32_t MAIL_OFF = align128(M_OFF + M_BYTES);
  static constexpr uint32_t TOTAL = MAIL_OFF + 2 * 2 * NW * 32 * 16;
};

// TMA-pipelined variant: while the CTA sweeps the tile it holds in registers,
// the TMA engine streams the next tile (state, quarter-gradient and -- for
// tiles that are not fully masked -- the mask words) into shared memory; one
// elected thread arms an mbarrier with the byte count and issues the
// cp.async.bulk.tensor loads.  smem -> registers is a conflict-free 128-bit
// copy (each warp reads one 512-byte row).  Nothing at the top of the tile loop
// depends on a global load: descriptors are fetched two tiles ahead.
template <int R, int NW, int OCC, bool H16>
__global__ void __launch_bounds__(NW * 32, OCC)
grid_sweepk_pipe_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h,
                        const __grid_constant__ CUtensorMap tm_m, PlaneGeom g, float *__restrict__ xout,
                        const int2 *__restrict__ tiles, int ntiles, int nsweeps, int halo_y, int halo_x, int reverse) {
  constexpr int TH = R * NW;
  using L = PipeSmem<R, NW, H16>;
  constexpr int H16_W = L::H16_W;
  constexpr uint32_t TILE_BYTES = L::X_BYTES, H_BYTES = L::H_BYTES, MASK_BYTES = L::M_BYTES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *sx = reinterpret_cast<float *>(smem_raw + L::X_OFF);
  unsigned char *sh = smem_raw + L::H_OFF;
  uint32_t *sm = reinterpret_cast<uint32_t *>(smem_raw + L::M_OFF);
  float4(*mailbox)[2][NW][32] = reinterpret_cast<float4(*)[2][NW][32]>(smem_raw + L::MAIL_OFF);
  __shared__ uint64_t bars[3];  // [0] TMA landing, [1..2] edge exchange per sweep parity
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], NW);
    mbar_init(&bars[2], NW);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int2 d) {  // one thread: arm the barrier and start the tile's bulk loads
    const TileRef r = unpack_tile(d);
    mbar_expect_tx(&bars[0], TILE_BYTES + H_BYTES + (r.full ? 0u : MASK_BYTES));
    tma_load_3d(sx, &tm_x, r.pcol, r.prow, r.plane, &bars[0]);
    tma_load_3d(sh, &tm_h, H16 ? (r.pcol & ~7) : r.pcol, r.prow, r.plane, &bars[0]);
    // the bulk copy must start on a 16-byte boundary: round the word column down to a multiple of 4
    if (!r.full) tma_load_2d(sm, &tm_m, (r.pcol >> 5) & ~3, r.prow, &bars[0]);
  };

  int t = blockIdx.x;
  if (t >= ntiles) return;
  const int stride = gridDim.x;
  // every other pass walks the tile list backwards: it starts on the tiles the previous pass wrote last,
  // which are still in L2
  if (reverse) tiles += ntiles - 1;
  const int dir = reverse ? -1 : 1;
  int2 cur = tiles[dir * t];  // (the tile list is not written by the sweep kernels: safe before the dependency wait)
  int2 nxt = (t + stride < ntiles) ? tiles[dir * (t + stride)] : cur;
  // programmatic dependent launch: let the next pass get scheduled, then wait for the previous pass --
  // it wrote the state this pass reads and read the buffer this pass overwrites
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) issue(cur);
  int parity = 0;
  uint32_t phase = 0, mphase = 0;
  for (; t < ntiles; t += stride) {
    // descriptor of the tile after next: consumed one full iteration from now
    const int2 nxt2 = (t + 2 * stride < ntiles) ? tiles[dir * (t + 2 * stride)] : nxt;
    const TileRef td = unpack_tile(cur);
    const int pcol = td.pcol + 4 * lane;
    const long long base = (long long)td.plane * g.plane + (long long)(td.prow + w * R) * g.pitch + pcol;
    float4 x[R], h[R];
    uint32_t mb[(R + 7) / 8];

    mbar_wait(&bars[0], phase);
    phase ^= 1;
    const int soff = (w * R) * TILE_W + 4 * lane;
#pragma unroll
    for (int i = 0; i < R; ++i) x[i] = ld4(sx + soff + i * TILE_W);
    if (H16) {
      const __half *hrow = reinterpret_cast<const __half *>(sh) + (w * R) * H16_W + (td.pcol & 7) + 4 * lane;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const uint2 raw = *reinterpret_cast<const uint2 *>(hrow + i * H16_W);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
        h[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
    } else {
      const float *hrow = reinterpret_cast<const float *>(sh) + soff;
#pragma unroll
      for (int i = 0; i < R; ++i) h[i] = ld4(hrow + i * TILE_W);
    }
#pragma unroll
    for (int i = 0; i < (R + 7) / 8; ++i) mb[i] = td.full ? 0xffffffffu : 0u;
    if (!td.full) {
      const uint32_t *mrow = sm + (w * R) * MASK_BOX_WORDS + ((pcol >> 5) - ((td.pcol >> 5) & ~3));
#pragma unroll
      for (int i = 0; i < R; ++i)
        mb[i / 8] |= ((mrow[i * MASK_BOX_WORDS] >> (pcol & 31)) & 0xFu) << ((i % 8) * 4);
    }
    __syncthreads();  // every thread has drained the staging buffers
    if (threadIdx.x == 0 && t + stride < ntiles) issue(nxt);

    // (tall register tiles: the body is long enough that unrolling past 2 only costs instruction cache --
    // measured 853 -> 877 Gupd/s on cfg2 for R = 21; 2 is the minimum that keeps `parity` static)
    constexpr int UNROLL = (R >= 18) ? 2 : kSweepUnroll;
#pragma unroll UNROLL
    for (int s = 0; s < nsweeps; ++s) {
      if (td.full)
        tile_sweep_split<R, NW, false>(x, h, mb, mailbox, &bars[1], parity, mphase);
      else
        tile_sweep_split<R, NW, true>(x, h, mb, mailbox, &bars[1], parity, mphase);
      parity ^= 1;
    }
    // store the inner region; groups without masked pixels keep their constants.
    // Branch-free: one row-validity bitmask per thread, predicated 128-bit stores.
    {
      const int lo = max(halo_y - w * R, 0), hi = min(TH - halo_y - w * R, R);
      uint32_t rows_ok = (hi > lo) ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
      if ((4 * lane < halo_x) || (4 * lane >= TILE_W - halo_x)) rows_ok = 0u;
      float *out = xout + base;
      const uint32_t pitch_bytes = (uint32_t)g.pitch * 4u;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const uint32_t nib = td.full ? 1u : (mb[i / 8] >> ((i % 8) * 4)) & 0xFu;
        const uint32_t on = ((rows_ok >> i) & 1u) && nib;
        st4_if(reinterpret_cast<char *>(out) + (size_t)i * pitch_bytes, x[i], on);
      }
    }
    cur = nxt;
    nxt = nxt2;
  }
}

}  // namespace fpie
#include "grid_pair.cuh"
#include "patch.cuh"
#include "grid_cluster.cuh"
namespace fpie {

// Classify the tile grid: flag bit0 = some masked pixel in the stored (inner)
// region, bit1 = every pixel of the whole tile masked.  One CTA per tile.
__global__ void classify_tiles_kernel(PlaneGeom g, const uint32_t *__restrict__ bits, int tiles_x, int tile_h,
                                      int step_y, int step_x, int halo_y, int halo_x, uint32_t *__restrict__ flags) {
  const int tile = blockIdx.x;
  const int ty = tile / tiles_x, tx = tile % tiles_x;
  const int prow0 = g.padr + ty * step_y - halo_y;
  const int pcol0 = g.padc + tx * step_x - halo_x;
  int any_inner = 0, all_full = 1;
  for (int q = threadIdx.x; q < tile_h * 32; q += blockDim.x) {
    const int tr = q >> 5, l = q & 31;
    const int pc = pcol0 + 4 * l;
    const uint32_t nib = (bits[(long long)(prow0 + tr) * g.wpitch + (pc >> 5)] >> (pc & 31)) & 0xFu;
    if (nib != 0xFu) all_full = 0;
    if (nib && tr >= halo_y && tr < tile_h - halo_y && 4 * l >= halo_x && 4 * l < TILE_W - halo_x) any_inner = 1;
  }
  any_inner = __syncthreads_or(any_inner);
  all_full = __syncthreads_and(all_full);
  if (threadIdx.x == 0) flags[tile] = (any_inner ? 1u : 0u) | (all_full ? 2u : 0u);
}

// ---------------------------------------------------------------------------
// epilogue: residual (fpie/np_solver.py:90-96) and uint8 image (grid.cu:115-131)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float resid_term(float t, float hq, float up, float dn, float lf, float rt) {
  float v = __fsub_rn(__fmul_rn(4.0f, t), __fmul_rn(4.0f, hq));  // 4t - g
  v = __fsub_rn(v, up);
  v = __fsub_rn(v, dn);
  v = __fsub_rn(v, lf);
  v = __fsub_rn(v, rt);
  return fabsf(v);
}

// the EquSolver's expression of the same residual, |((((B + U) + D) + L) + R) - 4 X| (np_solver.py:42-50,
// equ.cu equ_residual_kernel), on the grid embedding of an Equ system (B = 4 hq exactly, absent neighbours
// hold the constant 0): rounds differently from the grid expression once the residual is rounding noise
__device__ __forceinline__ float resid_term_equ(float t, float hq, float up, float dn, float lf, float rt) {
  float s = __fadd_rn(__fmul_rn(4.0f, hq), up);
  s = __fadd_rn(s, dn);
  s = __fadd_rn(s, lf);
  s = __fadd_rn(s, rt);
  s = __fsub_rn(s, __fmul_rn(4.0f, t));
  return fabsf(s);
}

// grid: (blocks over rows x groups, plane).  err[plane % 3] accumulates in double.
template <bool EQU>
__global__ void __launch_bounds__(256)
grid_residual_kernel(PlaneGeom g, BatchMap bm, int row_lo, int row_hi, const uint32_t *__restrict__ bits,
                     const float *__restrict__ x, const float *__restrict__ hq, double *__restrict__ err) {
  const int p = blockIdx.y;
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long per_plane = (long long)(row_hi - row_lo) * g.groups;
  float acc = 0.f;
  if (q < per_plane) {
    const int r = row_lo + (int)(q / g.groups), grp = (int)(q % g.groups);
    const int pc = g.padc + 4 * grp;
    const uint32_t nib = (bits[(long long)(r + g.padr) * g.wpitch + (pc >> 5)] >> (pc & 31)) & 0xFu;
    if (nib) {
      const long long off = (long long)p * g.plane + (long long)(r + g.padr) * g.pitch + pc;
      const float4 c = ld4(x + off), u = ld4(x + off - g.pitch), d = ld4(x + off + g.pitch), h = ld4(hq + off);
      const float lf = x[off - 1], rt = x[off + 4];
      if (EQU) {
        if (nib & 1u) acc += resid_term_equ(c.x, h.x, u.x, d.x, lf, c.y);
        if (nib & 2u) acc += resid_term_equ(c.y, h.y, u.y, d.y, c.x, c.z);
        if (nib & 4u) acc += resid_term_equ(c.z, h.z, u.z, d.z, c.y, c.w);
        if (nib & 8u) acc += resid_term_equ(c.w, h.w, u.w, d.w, c.z, rt);
      } else {
        if (nib & 1u) acc += resid_term(c.x, h.x, u.x, d.x, lf, c.y);
        if (nib & 2u) acc += resid_term(c.y, h.y, u.y, d.y, c.x, c.z);
        if (nib & 4u) acc += resid_term(c.z, h.z, u.z, d.z, c.y, c.w);
        if (nib & 8u) acc += resid_term(c.w, h.w, u.w, d.w, c.z, rt);
      }
    }
  }
  if (bm.batch > 0) {
    // one residual triple per patch: groups never straddle patches (pw % 4 == 0), lanes of a warp may.  A warp
    // that lies inside one patch (the common case) reduces by shuffles and issues ONE atomic -- per-thread
    // atomics on three addresses per patch cost 8.9 ms per step on config 5 (profiles/r02_cfg5_launches.csv).
    int id = -1;
    if (q < per_plane) {
      const int r = row_lo + (int)(q / g.groups), grp = (int)(q % g.groups);
      id = (r / bm.ph) * bm.bcols + (4 * grp) / bm.pw;
    }
    const int id0 = __shfl_sync(0xffffffffu, id, 0);
    if (__all_sync(0xffffffffu, id == id0)) {
      double v = (double)acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && id0 >= 0 && v != 0.0) atomicAdd(&err[(long long)id0 * 3 + p], v);
    } else if (acc != 0.f) {
      atomicAdd(&err[(long long)id * 3 + p], (double)acc);
    }
    return;
  }
  // warp shuffle reduction, then one partial per warp through shared memory
  double v = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __shared__ double partial[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) partial[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < (blockDim.x >> 5)) ? partial[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0 && v != 0.0) atomicAdd(&err[p % 3], v);
  }
}

__device__ __forceinline__ uint32_t clip_u8(float v) { return __float2uint_rz(fminf(fmaxf(v, 0.f), 255.f)); }

// thread = 4 pixels x 3 channels -> 12 interleaved bytes of img[n, m, 3]
__global__ void __launch_bounds__(256)
grid_to_u8_kernel(PlaneGeom g, BatchMap bm, const float *__restrict__ x, uint8_t *__restrict__ img) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= (long long)g.n * g.groups) return;
  const int r = (int)(q / g.groups), grp = (int)(q % g.groups);
  const long long off = (long long)(r + g.padr) * g.pitch + g.padc + 4 * grp;
  const float4 a = ld4(x + off), b = ld4(x + g.plane + off), c = ld4(x + 2 * g.plane + off);
  const uint32_t a0 = clip_u8(a.x), a1 = clip_u8(a.y), a2 = clip_u8(a.z), a3 = clip_u8(a.w);
  const uint32_t b0 = clip_u8(b.x), b1 = clip_u8(b.y), b2 = clip_u8(b.z), b3 = clip_u8(b.w);
  const uint32_t c0 = clip_u8(c.x), c1 = clip_u8(c.y), c2 = clip_u8(c.z), c3 = clip_u8(c.w);
  long long o = ((long long)r * g.m + 4 * grp) * 3;
  if (bm.batch > 0) {  // mosaic -> [batch, ph, pw, 3]
    const int by = r / bm.ph, bx = (4 * grp) / bm.pw;
    const int id = by * bm.bcols + bx;
    if (id >= bm.batch) return;
    o = (((long long)id * bm.ph + (r - by * bm.ph)) * bm.pw + (4 * grp - bx * bm.pw)) * 3;
  }
  const int valid = min(4, g.m - 4 * grp);
  if (valid == 4 && (o & 3) == 0) {
    uint32_t *dst = reinterpret_cast<uint32_t *>(img + o);
    dst[0] = a0 | (b0 << 8) | (c0 << 16) | (a1 << 24);
    dst[1] = b1 | (c1 << 8) | (a2 << 16) | (b2 << 24);
    dst[2] = c2 | (a3 << 8) | (b3 << 16) | (c3 << 24);
  } else {
    const uint32_t px[4][3] = {{a0, b0, c0}, {a1, b1, c1}, {a2, b2, c2}, {a3, b3, c3}};
    for (int k = 0; k < valid; ++k)
      for (int ch = 0; ch < 3; ++ch) img[o + k * 3 + ch] = (uint8_t)px[k][ch];
  }
}

// Fused Processor-level reset (fpie/process.py:338-378): one warp builds 32
// consecutive plane columns of one crop row -- mask word by ballot, state from
// the target crop, quarter-scaled mixed gradient on masked pixels.
__global__ void __launch_bounds__(256)
grid_build_kernel(PlaneGeom g, BlendImages b, int equ_form, int row_lo, int row_hi, uint32_t *__restrict__ bits,
                  float *__restrict__ x0, float *__restrict__ x1, float *__restrict__ hq,
                  unsigned long long *__restrict__ count) {
  // crop rows [row_lo, row_hi): the whole grid, or the rows whose source / target rows have arrived (the upload
  // comes in row chunks and the build of a chunk runs beside the copy of the next)
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)(row_hi - row_lo) * g.wpitch) return;
  const int r = row_lo + (int)(warp / g.wpitch);
  const int pcol = (int)(warp % g.wpitch) * 32 + lane;
  const int c = pcol - g.padc;
  BlendImages pb;
  int pr = r, pc2 = c;
  const bool inside = c >= 0 && c < g.m && patch_view(b, r, c, pb, pr, pc2);
  const bool on = inside && canonical_mask_at(pb, pb.x0 + pr, pb.y0 + pc2);
  const uint32_t word = __ballot_sync(0xffffffffu, on);
  const long long prow = r + g.padr;
  if (lane == 0) {
    bits[prow * g.wpitch + (pcol >> 5)] = word;
    if (word) atomicAdd(count, (unsigned long long)__popc(word));
  }
  if (!inside) return;
  const long long off = prow * g.pitch + pcol;
  if (equ_form) {
    // EquSolver arithmetic on the grid (the promoted form, see grid_from_equ_kernel): an unknown carries
    // X = target and B = grad + sum of the targets of its neighbours OUTSIDE the mask (process.py:255-263;
    // small integers and halves: exact in any order); every other pixel is the constant 0 (A's row 0).
    // In slab mode the halo frame rows hold the neighbour band's unknowns: live state, cleared mask bit.
    const int mr = pb.x0 + pr, mc2 = pb.y0 + pc2;
    const bool live = pb.slab ? raw_mask_at(pb, mr, mc2) : on;
    const int dr[4] = {-1, 1, 0, 0}, dc[4] = {0, 0, -1, 1};
    bool outside[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      outside[k] = on && !(pb.slab ? raw_mask_at(pb, mr + dr[k], mc2 + dc[k]) : canonical_mask_at(pb, mr + dr[k], mc2 + dc[k]));
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float t = live ? target_at(pb, pr, pc2, ch) : 0.f;
      float bsum = 0.f;
      if (on) {
        bsum = pixel_gradient(pb, pr, pc2, ch);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (outside[k]) bsum += target_at(pb, pr + dr[k], pc2 + dc[k], ch);
      }
      x0[ch * g.plane + off] = t;
      x1[ch * g.plane + off] = t;
      hq[ch * g.plane + off] = 0.25f * bsum;
    }
    return;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float t = target_at(pb, pr, pc2, ch);
    x0[ch * g.plane + off] = t;
    x1[ch * g.plane + off] = t;
    hq[ch * g.plane + off] = on ? 0.25f * pixel_gradient(pb, pr, pc2, ch) : 0.f;
  }
}

// EquSolver promotion (see equ.cu): embed a row-major Equ system into the grid.  Unknown i sits on
// its pixel with state X[i] and quarter-gradient B[i]/4; every other pixel is 0, because the Equ
// formulation folds the boundary values into B and gathers the constant row 0 for absent neighbours.
__global__ void __launch_bounds__(256)
grid_from_equ_kernel(PlaneGeom g, EquEmbed e, uint32_t *__restrict__ bits, float *__restrict__ x0,
                     float *__restrict__ x1, float *__restrict__ hq, unsigned long long *__restrict__ count) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)g.n * g.wpitch) return;
  const int r = (int)(warp / g.wpitch);
  const int pcol = (int)(warp % g.wpitch) * 32 + lane;
  const int c = pcol - g.padc;
  const bool inside = c >= 0 && c < g.m;
  const long long p = (long long)r * g.m + c;
  const bool on = inside && e.mask[p] > 0;
  const uint32_t word = __ballot_sync(0xffffffffu, on);
  const long long prow = r + g.padr;
  if (lane == 0) {
    bits[prow * g.wpitch + (pcol >> 5)] = word;
    if (word) atomicAdd(count, (unsigned long long)__popc(word));
  }
  if (!on) return;  // planes were zero-filled
  const long long off = prow * g.pitch + pcol;
  const long long id = e.ids[p];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float t = e.X[ch * e.pitch + id];
    x0[ch * g.plane + off] = t;
    x1[ch * g.plane + off] = t;
    hq[ch * g.plane + off] = 0.25f * e.B[ch * e.pitch + id];
  }
}

__global__ void __launch_bounds__(256)
grid_to_equ_kernel(PlaneGeom g, EquEmbed e, const float *__restrict__ x) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= (long long)g.n * g.m) return;
  if (!(e.mask[p] > 0)) return;
  const int r = (int)(p / g.m), c = (int)(p % g.m);
  const long long off = (long long)(r + g.padr) * g.pitch + c + g.padc;
  const long long id = e.ids[p];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) e.X[ch * e.pitch + id] = x[ch * g.plane + off];
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
static int blocks_for(long long work, int threads) { return (int)ceil_div(work, threads); }

GridSolver::GridSolver(int device, cudaStream_t stream, int block_k, int variant)
    : device_(device), stream_(stream), variant_(variant) {
  int count = 0;
  CUDA_CHECK(cudaGetDeviceCount(&count));
  FPIE_REQUIRE(device >= 0 && device < count, "fpie_b200: no such CUDA device");
  DeviceGuard guard(device_);
  cudaDeviceProp prop{};
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device_));
  FPIE_REQUIRE(prop.major >= 10, "fpie_b200 is built for sm_100a (Blackwell) only");
  sm_count_ = prop.multiProcessorCount;
  if (variant_ >= 100) {  // 100 + v: variant v, always streaming the fp32 gradient planes
    force_h32_ = true;
    variant_ -= 100;
  }
  // variant 0 = automatic tile shape, block_k 0 = automatic blocking depth: both follow the grid size at
  // reset (variant 39 is the fixed 14 x 12 shape that variant 0 used to name)
  auto_tune_ = (variant_ == 0);
  auto_k_ = (block_k <= 0);
  if (block_k <= 0) block_k = 8;
  FPIE_REQUIRE(block_k <= MAX_BLOCK_K, "block_k must be in 1..16");
  configure(variant_ == 0 ? 24 : variant_, block_k);
  err_.resize(4);
  CUDA_CHECK(cudaMallocHost(&host_err_, 4 * sizeof(double)));
  const char *no_serp = getenv("FPIE_B200_NO_SERPENTINE");
  serpentine_ = !(no_serp && no_serp[0] && no_serp[0] != '0');
  const char *no_graph = getenv("FPIE_B200_NO_GRAPH");
  graph_off_ = no_graph && no_graph[0] && no_graph[0] != '0';
  // the persistent small-image kernel: FPIE_B200_PATCH=0 switches it off (A/B against the mosaic path),
  // FPIE_B200_PATCH_ROWS=4|8 overrides the rows per thread
  const char *patch = getenv("FPIE_B200_PATCH");
  patch_off_ = patch && patch[0] == '0';
  const char *prow = getenv("FPIE_B200_PATCH_ROWS");
  patch_rows_ = prow ? atoi(prow) : 0;
  if (patch_rows_ != 4 && patch_rows_ != 8) patch_rows_ = 0;
}

GridSolver::~GridSolver() {
  // (runs from Python's garbage collector on whatever thread and current device it happens to be:
  // restore the caller's device; no exceptions out of a destructor, hence no DeviceGuard)
  halo_disconnect();
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device_);
  drop_graphs();
  upload_.destroy_stream();
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  if (host_err_) cudaFreeHost(host_err_);
  if (prev >= 0 && prev != device_) cudaSetDevice(prev);
}

void GridSolver::configure(int variant, int block_k) {
  variant_ = variant;
  block_k_ = block_k;
  halo_x_ = (int)round_up(block_k_, 4);
  shape_ = shape_for(variant_);
  FPIE_REQUIRE(shape_.tile_h() > 2 * block_k_, "block_k too deep for the tile height");
}

// Candidate (tile shape, depth) pairs of the automatic configuration and the cost model that ranks them
// on grids of up to 12 Mpx, where the number of waves a pass needs on 148 SMs decides (a pass over 543
// tiles on 296 CTA slots costs two full waves).  Per-wave time of one pass = a + b * k microseconds, fitted
// to tools/tune_grid.py runs at 512^2 .. 3072^2 (profiles/r01_tune_small_grids.json; the picks lose 0-6 % against
// the best measured configuration of each size and gain up to 18 % over a size-only rule); `launch` = per-pass launch cost.
struct TileChoice {
  int variant, k, occ;
  double a, b;
};
static const TileChoice kTileChoices[] = {
    {12, 6, 2, 1.68, 0.427},  {12, 8, 2, 1.68, 0.427},  {12, 10, 2, 1.68, 0.427}, {12, 12, 2, 1.68, 0.427},
    {12, 16, 2, 1.68, 0.427}, {36, 6, 2, 2.22, 0.531},  {36, 8, 2, 2.22, 0.531},  {36, 10, 2, 2.22, 0.531},
    {36, 12, 2, 2.22, 0.531}, {36, 16, 2, 2.22, 0.531}, {24, 8, 1, 2.70, 0.551},  {24, 10, 1, 2.70, 0.551},
    {24, 12, 1, 2.70, 0.551}, {24, 16, 1, 2.70, 0.551},
};
constexpr double kLaunchUs = 4.0;
// the last, partial wave of a two-CTAs-per-SM shape is cheaper when it leaves one CTA per SM
constexpr double kLoneCtaWave = 0.85;
constexpr long long kModelMaxPixels = 12000000;

// Automatic configuration (measured on B200, tools/tune_grid.py, profiles/r01_tune_variants_k.json): small
// grids cannot fill 148 SMs with 168-row tiles and pay the per-launch latency once per pass, so they get
// 64-row tiles on two CTAs per SM and deeper temporal blocking; mid-size grids 84-row tiles (21 rows per
// thread, four warps, two CTAs per SM); large grids the 168-row tile as 8 warps x 21 rows with k = 8.
void GridSolver::auto_configure(int n, int m) {
  const long long px = (long long)n * m;
  int v, k;
  if (px <= 450000)
    v = 12, k = 16;
  else if (px <= 1600000)
    v = 12, k = 8;
  else if (px <= 6000000)
    v = 36, k = 8;
  else
    v = 24, k = 8;
  // very wide, flat grids: a tile much taller than the grid sweeps mostly padding
  if (v != 12 && n <= 36) v = 12;
  else if (v == 24 && n <= 72) v = 36;
  if (!auto_k_) k = block_k_;  // the caller's depth wins; take the tallest tile if the small one cannot hold it
  if (shape_for(v).tile_h() <= 2 * k) v = 24;
  configure(v, k);
}

void GridSolver::require_ready() const { FPIE_REQUIRE(ready_, "GridSolver: step/state called before reset"); }

void GridSolver::layout(int n, int m) {
  FPIE_REQUIRE(n >= 1 && m >= 1, "GridSolver.reset: empty grid");
  if (auto_tune_) auto_configure(n, m);
  const int step_x = TILE_W - 2 * halo_x_, step_y = shape_.tile_h() - 2 * block_k_;
  const int tiles_x = (int)ceil_div(m, step_x), tiles_y = (int)ceil_div(n, step_y);
  PlaneGeom g{};
  g.n = n;
  g.m = m;
  g.padr = PAD_ROWS;
  g.padc = PAD_COLS;
  // multiple of 128 floats: mask rows (pitch / 32 words) stay 16-byte aligned for TMA
  long long need_cols = (long long)tiles_x * step_x + halo_x_ + 4;
  int need_rows = tiles_y * step_y + block_k_ + 1;
  if (auto_tune_ && auto_k_) {
    // shape and depth may still change once the mask is on the device (choose_by_model, the k = 12 rule
    // of after_state_loaded): make the padded planes large enough for every tiling that can be picked
    for (const TileChoice &c : kTileChoices) {
      const int hx = (int)round_up(c.k, 4), sx = TILE_W - 2 * hx, sy = shape_for(c.variant).tile_h() - 2 * c.k;
      need_cols = std::max(need_cols, (long long)ceil_div(m, sx) * sx + hx + 4);
      need_rows = std::max(need_rows, (int)ceil_div(n, sy) * sy + c.k + 1);
    }
  }
  g.pitch = (int)round_up(g.padc + need_cols, 128);
  g.rows = g.padr + need_rows;
  g.wpitch = g.pitch / 32;
  g.groups = (int)ceil_div(m, 4);
  g.plane = (long long)g.rows * g.pitch;
  // tile descriptors pack the plane row into 28 bits and the plane column into 20 (pack_tile)
  FPIE_REQUIRE(g.pitch < (1 << 20) && g.rows < (1 << 28), "GridSolver.reset: grid too wide or too tall for the tile descriptors (columns < 2^20, rows < 2^28)");
  geom_ = g;
  win_lo_ = 0;
  win_hi_ = n;
}

void GridSolver::reset(int n, int m, const int32_t *mask, int64_t mask_rs, int64_t mask_cs, const float *tgt,
                       const float *grad) {
  FPIE_REQUIRE(mask && tgt && grad, "GridSolver.reset: null input");
  FPIE_REQUIRE(mask_cs == 1 && mask_rs >= m, "GridSolver.reset: mask rows must be contiguous (column stride 1)");
  DeviceGuard guard(device_);
  ready_ = false;
  resid_equ_ = false;
  zeroed_ = PlaneGeom{};
  batch_ = BatchMap{0, 0, 0, 0};
  layout(n, m);
  const PlaneGeom &g = geom_;
  const size_t pixels = (size_t)n * m;
  for (auto &b : x_) b.resize((size_t)g.plane * 3);
  hq_.resize((size_t)g.plane * 3);
  bits_.resize((size_t)g.rows * g.wpitch);
  stage_.resize(pixels * 3);
  mask_stage_.resize(pixels);
  img_.resize(pixels * 3);

  CUDA_CHECK(cudaMemsetAsync(x_[0].ptr, 0, x_[0].bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(x_[1].ptr, 0, x_[1].bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(hq_.ptr, 0, hq_.bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(err_.ptr, 0, err_.bytes(), stream_));

  CUDA_CHECK(cudaMemcpy2DAsync(mask_stage_.ptr, (size_t)m * 4, mask, (size_t)mask_rs * 4, (size_t)m * 4, n,
                               cudaMemcpyHostToDevice, stream_));
  {
    const long long warps = (long long)g.rows * g.wpitch;
    pack_mask_kernel<<<blocks_for(warps * 32, 256), 256, 0, stream_>>>(
        g, mask_stage_.ptr, bits_.ptr, reinterpret_cast<unsigned long long *>(err_.ptr + 3));
    CUDA_CHECK(cudaGetLastError());
  }
  CUDA_CHECK(cudaMemcpyAsync(stage_.ptr, tgt, pixels * 12, cudaMemcpyHostToDevice, stream_));
  aos_to_planes_kernel<<<blocks_for(pixels, 256), 256, 0, stream_>>>(g, stage_.ptr, 1.0f, x_[0].ptr, x_[1].ptr);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(stage_.ptr, grad, pixels * 12, cudaMemcpyHostToDevice, stream_));
  aos_to_planes_kernel<<<blocks_for(pixels, 256), 256, 0, stream_>>>(g, stage_.ptr, 0.25f, hq_.ptr, nullptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 3;
  after_state_loaded();
}

void GridSolver::reset_from_images(const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh, int mw, int mc,
                                   const uint8_t *tgt, int th, int tw, int h0, int w0, int h1, int w1, int grad_mode,
                                   int64_t *out_n, int32_t *out_box4, bool crop) {
  DeviceGuard guard(device_);
  ready_ = false;
  BlendUpload &up = upload_;  // device copies of the images are kept between resets (no malloc / free per call)
  up.upload(stream_, src, sh, sw, mask, mh, mw, mc, tgt, th, tw, h0, w0, h1, w1, grad_mode, crop, &chunks_);
  batch_ = BatchMap{0, 0, 0, 0};
  try {
    build_from_upload();
  } catch (...) {
    up.wait_copies();
    throw;
  }
  const BlendImages &b = up.images();
  if (out_n) *out_n = (int64_t)geom_.n * geom_.m;
  if (out_box4) {
    out_box4[0] = b.h1 + b.x0;
    out_box4[1] = b.h1 + b.x0 + b.n;
    out_box4[2] = b.w1 + b.y0;
    out_box4[3] = b.w1 + b.y0 + b.m;
  }
}

// `batch` independent patches of ph x pw pixels, solved together as one mosaic grid
// (bcols patches per mosaic row).  Every patch keeps its own unmasked frame, so
// the patches do not interact; the mosaic just feeds the tiled kernel full-width rows.
void GridSolver::reset_batch(const uint8_t *src, const uint8_t *mask, const uint8_t *tgt, int batch, int ph, int pw,
                             int mc, int grad_mode) {
  DeviceGuard guard(device_);
  ready_ = false;
  FPIE_REQUIRE(pw % 4 == 0, "reset_batch: patch width must be a multiple of 4");
  // aim for a roughly square mosaic that is at least 4096 pixels wide when the batch allows it
  int bcols = 1;
  while (bcols * pw < 4096 && bcols < batch) bcols *= 2;
  upload_.upload_batch(stream_, src, mask, tgt, batch, ph, pw, mc, grad_mode, bcols);
  chunks_.count = 0;
  batch_ = BatchMap{batch, ph, pw, bcols};
  batch_err_.resize((size_t)batch * 3);
  build_from_upload();
}

void GridSolver::reset_from_equ(const EquEmbed &e) {
  DeviceGuard guard(device_);
  ready_ = false;
  resid_equ_ = false;
  zeroed_ = PlaneGeom{};
  batch_ = BatchMap{0, 0, 0, 0};
  layout(e.n, e.m);
  const PlaneGeom &g = geom_;
  for (auto &buf : x_) buf.resize((size_t)g.plane * 3);
  hq_.resize((size_t)g.plane * 3);
  bits_.resize((size_t)g.rows * g.wpitch);
  img_.resize((size_t)g.n * g.m * 3);
  CUDA_CHECK(cudaMemsetAsync(x_[0].ptr, 0, x_[0].bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(x_[1].ptr, 0, x_[1].bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(hq_.ptr, 0, hq_.bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(bits_.ptr, 0, bits_.bytes(), stream_));
  CUDA_CHECK(cudaMemsetAsync(err_.ptr, 0, err_.bytes(), stream_));
  const long long warps = (long long)g.n * g.wpitch;
  grid_from_equ_kernel<<<blocks_for(warps * 32, 256), 256, 0, stream_>>>(
      g, e, bits_.ptr, x_[0].ptr, x_[1].ptr, hq_.ptr, reinterpret_cast<unsigned long long *>(err_.ptr + 3));
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
  after_state_loaded();
}

void GridSolver::export_to_equ(const EquEmbed &e) {
  require_ready();
  DeviceGuard guard(device_);
  const long long px = (long long)geom_.n * geom_.m;
  grid_to_equ_kernel<<<blocks_for(px, 256), 256, 0, stream_>>>(geom_, e, x_[cur_].ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
}

void GridSolver::build_from_upload() {
  const BlendImages &b = upload_.images();
  FPIE_REQUIRE(!(equ_form_ && b.batch > 0), "the EquSolver formulation is not available for batched patches");
  resid_equ_ = equ_form_;
  layout(b.n, b.m);
  const PlaneGeom &g = geom_;
  for (auto &buf : x_) buf.resize((size_t)g.plane * 3);
  hq_.resize((size_t)g.plane * 3);
  bits_.resize((size_t)g.rows * g.wpitch);
  img_.resize((size_t)g.n * g.m * 3);
  // The build kernel writes every pixel and mask word of the n x m grid; only the padding around it has to be
  // zeroed, and it stays zero (no kernel ever writes padding).  A reset on the geometry of the previous one
  // (the GUI, repeated blends of one size) therefore skips 0.6 GB of memsets at 4096^2.
  const bool clean = zeroed_.n == g.n && zeroed_.m == g.m && zeroed_.rows == g.rows && zeroed_.pitch == g.pitch &&
                     zeroed_ptr_[0] == x_[0].ptr && zeroed_ptr_[1] == x_[1].ptr && zeroed_ptr_[2] == hq_.ptr &&
                     zeroed_ptr_[3] == (float *)bits_.ptr && batch_.batch == 0 && !zeroed_batch_;
  if (!clean) {
    CUDA_CHECK(cudaMemsetAsync(x_[0].ptr, 0, x_[0].bytes(), stream_));
    CUDA_CHECK(cudaMemsetAsync(x_[1].ptr, 0, x_[1].bytes(), stream_));
    CUDA_CHECK(cudaMemsetAsync(hq_.ptr, 0, hq_.bytes(), stream_));
    CUDA_CHECK(cudaMemsetAsync(bits_.ptr, 0, bits_.bytes(), stream_));
  }
  zeroed_ = g;
  zeroed_ptr_[0] = x_[0].ptr;
  zeroed_ptr_[1] = x_[1].ptr;
  zeroed_ptr_[2] = hq_.ptr;
  zeroed_ptr_[3] = (float *)bits_.ptr;
  zeroed_batch_ = batch_.batch > 0;
  CUDA_CHECK(cudaMemsetAsync(err_.ptr, 0, err_.bytes(), stream_));
  auto build_rows = [&](int lo, int hi) {
    if (hi <= lo) return;
    const long long warps = (long long)(hi - lo) * g.wpitch;
    grid_build_kernel<<<blocks_for(warps * 32, 256), 256, 0, stream_>>>(
        g, b, equ_form_ ? 1 : 0, lo, hi, bits_.ptr, x_[0].ptr, x_[1].ptr, hq_.ptr,
        reinterpret_cast<unsigned long long *>(err_.ptr + 3));
    CUDA_CHECK(cudaGetLastError());
    stats_.launches += 1;
  };
  if (chunks_.count > 0) {
    // a crop row reads the source / target rows above and below it: the build lags the copies by one row
    int lo = 0;
    for (int k = 0; k < chunks_.count; ++k) {
      const int hi = (k + 1 == chunks_.count) ? g.n : std::max(chunks_.row_hi[k] - 1, lo);
      CUDA_CHECK(cudaStreamWaitEvent(stream_, chunks_.ev[k], 0));
      build_rows(lo, hi);
      lo = hi;
    }
    chunks_.count = 0;
  } else {
    build_rows(0, g.n);
  }
  stats_.launches += 1;
  after_state_loaded();
}

// Everything reset still has to decide once the state is on the device -- fp16 or fp32 gradient stream,
// blocking depth, tile shape, the list of active tiles -- from ONE round trip: the candidate tilings are
// classified speculatively, their flags come back together with the unknown count and the fp16 flag,
// and the host picks.  (Round 1 synchronised five times here.)
void GridSolver::after_state_loaded() {
  const PlaneGeom &g = geom_;
  // fp16 copy of the quarter-gradient (halves that stream when every value is exactly representable)
  const long long count = g.plane * 3;
  hq16_.resize((size_t)count);
  flag_.resize(1);
  CUDA_CHECK(cudaMemsetAsync(flag_.ptr, 0, sizeof(int), stream_));
  planes_to_half_kernel<<<blocks_for(count, 256), 256, 0, stream_>>>(count, hq_.ptr, hq16_.ptr, flag_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;

  // candidate (variant, depth) pairs
  struct Cand {
    int variant, k, occ;
    double a, b;
  };
  std::vector<Cand> cands;
  const long long px = (long long)g.n * g.m;
  const bool free_choice = auto_k_ && auto_tune_;
  if (free_choice && px > kModelMaxPixels) {
    cands.push_back({variant_, block_k_, 1, 0, 0});
    cands.push_back({variant_, 12, 1, 0, 0});  // (picked when the grid is >= 90 % masked, see below)
  } else if (free_choice && batch_.batch == 0) {
    for (const TileChoice &c : kTileChoices) cands.push_back({c.variant, c.k, c.occ, c.a, c.b});
  } else {
    cands.push_back({variant_, block_k_, 1, 0, 0});
  }
  const int NC = (int)cands.size();
  std::vector<size_t> offset(NC + 1, 0);
  std::vector<int> tiles_x(NC), tile_h(NC), hx(NC);
  for (int i = 0; i < NC; ++i) {
    tile_h[i] = shape_for(cands[i].variant).tile_h();
    hx[i] = (int)round_up(cands[i].k, 4);
    tiles_x[i] = (int)ceil_div(g.m, TILE_W - 2 * hx[i]);
    offset[i + 1] = offset[i] + (size_t)tiles_x[i] * ceil_div(g.n, tile_h[i] - 2 * cands[i].k);
  }
  tile_flags_.resize(offset[NC]);
  for (int i = 0; i < NC; ++i) {
    const int n_i = (int)(offset[i + 1] - offset[i]);
    classify_tiles_kernel<<<n_i, 256, 0, stream_>>>(g, bits_.ptr, tiles_x[i], tile_h[i], tile_h[i] - 2 * cands[i].k,
                                                    TILE_W - 2 * hx[i], cands[i].k, hx[i], tile_flags_.ptr + offset[i]);
  }
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += NC;
  std::vector<uint32_t> flags(offset[NC]);
  int inexact = 0;
  CUDA_CHECK(cudaMemcpyAsync(flags.data(), tile_flags_.ptr, flags.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             stream_));
  CUDA_CHECK(cudaMemcpyAsync(&inexact, flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  // unknown count (stored as a 64-bit integer in err_[3])
  CUDA_CHECK(cudaMemcpyAsync(host_err_ + 3, err_.ptr + 3, sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  h16_ok_ = (inexact == 0) && !force_h32_;
  unsigned long long cnt;
  memcpy(&cnt, host_err_ + 3, sizeof(cnt));
  stats_.unknowns = (int64_t)cnt;

  int best = 0;
  if (free_choice && px > kModelMaxPixels) {
    // large and (almost) fully masked grids: every tile is a select-free full tile, and the per-tile cost
    // amortises better over 12 sweeps than over 8 (measured 894 -> 934 Gupd/s at 4096^2, 2011 -> 2114 on two
    // 16384 x 32768 bands); masks with a long boundary prefer 8 (fewer partially filled boundary tiles)
    best = ((double)stats_.unknowns >= 0.9 * (double)g.n * (double)g.m) ? 1 : 0;
  } else if (free_choice && batch_.batch == 0 && stats_.unknowns > 0) {
    // small and mid-size grids: the lowest modelled time per sweep, (waves x per-wave time + launch) / k,
    // from the number of active tiles of every candidate tiling
    double best_cost = 0.0;
    best = -1;
    for (int i = 0; i < NC; ++i) {
      long long active = 0;
      for (size_t t = offset[i]; t < offset[i + 1]; ++t) active += flags[t] & 1u;
      const long long slots = (long long)sm_count_ * cands[i].occ, items = 3 * active;
      const long long full = items / slots, rest = items - full * slots;
      double waves = (double)full;
      if (rest > 0) waves += (cands[i].occ == 2 && rest <= sm_count_) ? kLoneCtaWave : 1.0;
      const double cost = (waves * (cands[i].a + cands[i].b * cands[i].k) + kLaunchUs) / cands[i].k;
      if (best < 0 || cost < best_cost) best = i, best_cost = cost;
    }
  } else if (free_choice && batch_.batch == 0) {
    best = -1;  // no unknowns: keep the size-based configuration
  }
  if (best >= 0 && (cands[best].variant != variant_ || cands[best].k != block_k_)) configure(cands[best].variant, cands[best].k);
  if (best < 0) {  // (the flags of the current configuration are not among the candidates: classify it)
    build_tiles(nullptr);
  } else {
    build_tiles(flags.data() + offset[best]);
  }
  make_tensor_maps();
  cur_ = 0;
  ready_ = true;
}

// The active-tile list of the current configuration, from its classification flags (classified here
// when the caller has none).
void GridSolver::build_tiles(const uint32_t *flags_in) {
  const PlaneGeom &g = geom_;
  const int step_x = TILE_W - 2 * halo_x_, step_y = shape_.tile_h() - 2 * block_k_;
  const int tiles_x = (int)ceil_div(g.m, step_x), tiles_y = (int)ceil_div(g.n, step_y);
  const int ntiles = tiles_x * tiles_y;
  std::vector<uint32_t> own;
  if (!flags_in) {
    tile_flags_.resize(ntiles);
    classify_tiles_kernel<<<ntiles, 256, 0, stream_>>>(g, bits_.ptr, tiles_x, shape_.tile_h(), step_y, step_x, block_k_,
                                                       halo_x_, tile_flags_.ptr);
    CUDA_CHECK(cudaGetLastError());
    stats_.launches += 1;
    own.resize(ntiles);
    CUDA_CHECK(cudaMemcpyAsync(own.data(), tile_flags_.ptr, ntiles * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    flags_in = own.data();
  }
  std::vector<int2> &list = host_tiles_;
  list.clear();
  host_tile_row_.clear();
  list.reserve((size_t)ntiles * 3);
  int64_t active = 0;
  for (int t = 0; t < ntiles; ++t) {
    if (!(flags_in[t] & 1u)) continue;
    ++active;
    const int ty = t / tiles_x, tx = t % tiles_x;
    for (int ch = 0; ch < 3; ++ch) {
      list.push_back(pack_tile(g.padr + ty * step_y - block_k_, g.padc + tx * step_x - halo_x_, ch,
                               (flags_in[t] & 2u) ? 1 : 0));
      host_tile_row_.push_back(ty);
    }
  }
  edge_rows_ = 0;
  n_part_[0] = n_part_[1] = 0;
  drop_graphs();  // they hold the old tile list, tensor maps and tile shape
  stats_.active_tiles = active;
  stats_.total_tiles = ntiles;
  n_tile_entries_ = (int)list.size();
  tiles_.resize(std::max<size_t>(list.size(), 1));
  // (pageable source: the driver stages the bytes before the call returns, and host_tiles_ outlives it anyway)
  if (!list.empty())
    CUDA_CHECK(cudaMemcpyAsync(tiles_.ptr, list.data(), list.size() * sizeof(int2), cudaMemcpyHostToDevice, stream_));
}

namespace {

struct SweepArgs {
  int grid;
  cudaStream_t stream;
  PlaneGeom g;
  const uint32_t *bits;
  const float *xin;
  float *xout;
  const float *hq;
  const CUtensorMap *tm_x;
  const CUtensorMap *tm_h;
  const int2 *tiles;
  const CUtensorMap *tm_m;
  bool h16;
  int reverse;  // walk the tile list backwards (alternate passes: L2 reuse of the previous pass's last tiles)
  int ntiles, nsweeps, halo_y, halo_x;
  bool load_only;  // resolve (load) the kernel and set its attributes, launch nothing
};

template <int R, int NW>
void launch_direct(const SweepArgs &a) {
  if (a.load_only) {
    cudaFuncAttributes attr;
    CUDA_CHECK(cudaFuncGetAttributes(&attr, grid_sweepk_kernel<R, NW>));
    return;
  }
  grid_sweepk_kernel<R, NW><<<std::min(a.ntiles, a.grid), NW * 32, 0, a.stream>>>(a.g, a.bits, a.xin, a.xout, a.hq, a.tiles, a.ntiles,
                                                              a.nsweeps, a.halo_y, a.halo_x);
}

template <int R, int NW, int OCC, bool H16>
void launch_pipe_h(const SweepArgs &a) {
  auto kernel = grid_sweepk_pipe_kernel<R, NW, OCC, H16>;
  constexpr size_t smem = PipeSmem<R, NW, H16>::TOTAL;
  static int configured_device = -1;  // the attribute is per function and per device
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  if (configured_device != dev) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = dev;
  }
  if (a.load_only) return;
  const int grid = std::min(a.ntiles, a.grid * OCC);
  // programmatic dependent launch: the next pass may be scheduled while this one drains; its CTAs run
  // their prologue (barrier init, descriptor fetch) and block in griddepcontrol.wait until this grid
  // has completed, so the per-launch latency overlaps the tail instead of following it
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // (measured: a win once there is at least one tile per SM, a loss for grids of a few dozen tiles)
  cfg.numAttrs = (a.ntiles >= a.grid) ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, *a.tm_x, *a.tm_h, *a.tm_m, a.g, a.xout, a.tiles, a.ntiles, a.nsweeps,
                                a.halo_y, a.halo_x, a.reverse));
}

template <int R, int NW, int OCC>
void launch_pipe(const SweepArgs &a) {
  if (a.h16)
    launch_pipe_h<R, NW, OCC, true>(a);
  else
    launch_pipe_h<R, NW, OCC, false>(a);
}

template <int R, int NW, int OCC, bool H16>
void launch_pair_h(const SweepArgs &a) {
  auto kernel = grid_sweepk_pair_kernel<R, NW, OCC, H16>;
  constexpr size_t smem = PairSmem<R, NW, H16>::TOTAL;
  static int configured_device = -1;
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  if (configured_device != dev) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = dev;
  }
  if (a.load_only) return;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(std::min(a.ntiles, a.grid * OCC));
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // (see launch_pipe_h)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (a.ntiles >= a.grid) ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, *a.tm_x, *a.tm_h, *a.tm_m, a.g, a.xout, a.tiles, a.ntiles, a.nsweeps,
                                a.halo_y, a.halo_x, a.reverse));
}

template <int R, int NW, int OCC>
void launch_pair(const SweepArgs &a) {
  if (a.h16)
    launch_pair_h<R, NW, OCC, true>(a);
  else
    launch_pair_h<R, NW, OCC, false>(a);
}

template <int R, int NW, int OCC, bool H16>
void launch_cluster_h(const SweepArgs &a, int cluster) {
  auto kernel = grid_sweepk_cluster_kernel<R, NW, OCC, H16>;
  constexpr size_t smem = ClusterSmem<R, NW, H16>::TOTAL;
  static int configured_device = -1;
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  if (configured_device != dev) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = dev;
  }
  if (a.load_only) return;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // clusters the device holds at once (a cluster lives inside one GPC): asked once per cluster size
  static int max_clusters[9] = {0};
  if (!max_clusters[cluster]) {
    cfg.gridDim = dim3(cluster * a.grid * OCC);
    CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters[cluster], kernel, &cfg));
    FPIE_REQUIRE(max_clusters[cluster] > 0, "the cluster kernel does not fit this device");
  }
  const int clusters = std::min(a.ntiles, max_clusters[cluster]);
  cfg.gridDim = dim3(clusters * cluster);
  cfg.numAttrs = (a.ntiles >= a.grid / cluster) ? 2 : 1;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, *a.tm_x, *a.tm_h, *a.tm_m, a.g, a.xout, a.tiles, a.ntiles, a.nsweeps,
                                a.halo_y, a.halo_x, a.reverse));
}

template <int R, int NW, int OCC>
void launch_cluster(const SweepArgs &a, int cluster) {
  if (a.h16)
    launch_cluster_h<R, NW, OCC, true>(a, cluster);
  else
    launch_cluster_h<R, NW, OCC, false>(a, cluster);
}

struct VariantInfo {
  int rows, warps, occ;
  bool pipe;
  int cluster = 1;
};

// kernel variants (fpie_b200_grid_create `variant`): register-tile shape
// rows/thread x warps, CTAs per SM; "pipe" = TMA-staged + split-phase exchange.
// The default build carries the three shapes the solver picks by itself (12, 24, 36) plus two cross-check kernels
// (1, 4); -DFPIE_ALL_VARIANTS adds every measured-and-dominated kernel of DESIGN.md section 5 -- round 1's tile
// shapes, the packed-pair (FFMA2) kernels, the cluster kernels -- for test / tuning builds.
VariantInfo variant_info(int v) {
  switch (v) {
    case 0:  // (automatic: replaced by a concrete shape at reset)
    // two warps per scheduler: each sub-partition's 16384 registers then allow up to 255 per thread
    case 24: return {21, 8, 1, true};
    case 1: return {16, 12, 1, false};  // (tile shape unused: one sweep per launch)
    case 4: return {16, 12, 1, false};  // direct global -> register loads, no TMA staging (cross-check)
    case 12: return {8, 8, 2, true};
    // four warps per CTA, several CTAs per SM: small tiles for small grids without giving up rows per thread
    case 36: return {21, 4, 2, true};
#ifdef FPIE_ALL_VARIANTS
    case 39: return {14, 12, 1, true};  // round 1's first default
    // packed-pair (FFMA2) kernels: rows = 2 x pair-rows per thread (grid_pair.cuh)
    case 40: return {22, 8, 1, true};
    case 41: return {20, 8, 1, true};
    case 42: return {20, 4, 2, true};
    case 43:
    case 20: return {14, 12, 1, true};
    case 44: return {10, 16, 1, true};
    case 21: return {16, 12, 1, true};
    case 22: return {12, 12, 1, true};
    // cluster kernels: `cluster` CTAs of 4 warps stacked on one tile (grid_cluster.cuh)
    case 50: return {21, 4, 2, true, 2};
    case 51: return {21, 4, 2, true, 4};
    case 52: return {21, 4, 2, true, 3};
    case 53: return {14, 4, 3, true, 4};
    case 2: return {16, 8, 1, false};
    case 3: return {8, 16, 1, false};
    case 5: return {16, 12, 1, true};
    case 6: return {16, 6, 2, true};
    case 7: return {14, 6, 2, true};
    case 8: return {12, 8, 2, true};
    case 9: return {12, 14, 1, true};
    case 10: return {10, 16, 1, true};
    case 11: return {8, 16, 1, true};
    case 17: return {15, 12, 1, true};
    case 18: return {20, 8, 1, true};
    case 25: return {19, 8, 1, true};
    case 29: return {16, 4, 2, true};
    case 37: return {12, 4, 3, true};
#endif
    default: throw Error("fpie_b200: unknown grid kernel variant (measured-and-dominated shapes need a -DFPIE_ALL_VARIANTS build)");
  }
}

// one pass of the tiled kernel in the register-tile shape `variant` names
static void launch_variant(int variant, const SweepArgs &a) {
  switch (variant) {
    case 24: launch_pipe<21, 8, 1>(a); break;
    case 4: launch_direct<16, 12>(a); break;
    case 12: launch_pipe<8, 8, 2>(a); break;
    case 36: launch_pipe<21, 4, 2>(a); break;
#ifdef FPIE_ALL_VARIANTS
    case 39: launch_pipe<14, 12, 1>(a); break;
    case 40: launch_pair<11, 8, 1>(a); break;
    case 41: launch_pair<10, 8, 1>(a); break;
    case 42: launch_pair<10, 4, 2>(a); break;
    case 43:
    case 20: launch_pair<7, 12, 1>(a); break;
    case 44: launch_pair<5, 16, 1>(a); break;
    case 21: launch_pair<8, 12, 1>(a); break;
    case 22: launch_pair<6, 12, 1>(a); break;
    case 50: launch_cluster<21, 4, 2>(a, 2); break;
    case 51: launch_cluster<21, 4, 2>(a, 4); break;
    case 52: launch_cluster<21, 4, 2>(a, 3); break;
    case 53: launch_cluster<14, 4, 3>(a, 4); break;
    case 2: launch_direct<16, 8>(a); break;
    case 3: launch_direct<8, 16>(a); break;
    case 5: launch_pipe<16, 12, 1>(a); break;
    case 6: launch_pipe<16, 6, 2>(a); break;
    case 7: launch_pipe<14, 6, 2>(a); break;
    case 8: launch_pipe<12, 8, 2>(a); break;
    case 9: launch_pipe<12, 14, 1>(a); break;
    case 10: launch_pipe<10, 16, 1>(a); break;
    case 11: launch_pipe<8, 16, 1>(a); break;
    case 17: launch_pipe<15, 12, 1>(a); break;
    case 18: launch_pipe<20, 8, 1>(a); break;
    case 25: launch_pipe<19, 8, 1>(a); break;
    case 29: launch_pipe<16, 4, 2>(a); break;
    case 37: launch_pipe<12, 4, 3>(a); break;
#endif
    default: throw Error("fpie_b200: unknown grid kernel variant (measured-and-dominated shapes need a -DFPIE_ALL_VARIANTS build)");
  }
}

}  // namespace

void GridSolver::config(int *variant, int *rows, int *warps, int *occ) const {
  const VariantInfo v = variant_info(variant_);
  if (variant) *variant = variant_;
  if (rows) *rows = v.rows;
  if (warps) *warps = v.warps;
  if (occ) *occ = v.occ;
}

TileShape GridSolver::shape_for(int variant) {
  const VariantInfo v = variant_info(variant);
  return {v.rows, v.warps, v.cluster};
}

void GridSolver::make_tensor_maps() {
  const PlaneGeom &g = geom_;
  for (int i = 0; i < 2; ++i)
    tm_x_[i] = make_plane_tensor_map(x_[i].ptr, g.pitch, g.rows, 3, g.plane, TILE_W, shape_.box_h());
  tm_h_ = make_plane_tensor_map(hq_.ptr, g.pitch, g.rows, 3, g.plane, TILE_W, shape_.box_h());
  tm_h16_ = make_plane_tensor_map(hq16_.ptr, g.pitch, g.rows, 3, g.plane, TILE_W + 8, shape_.box_h(), true);
  tm_m_ = make_mask_tensor_map(bits_.ptr, g.wpitch, g.rows, MASK_BOX_WORDS, shape_.box_h());
}

namespace {

struct PatchArgs {
  PlaneGeom g;
  BatchMap bm;
  float *x;
  const float *hq;
  const uint32_t *bits;
  int nsweeps, nitems, cluster, sm_count;
  cudaStream_t stream;
};

// one persistent launch: every cluster of `cluster` CTAs walks over (patch, plane) items
template <int R, int NW, int CPT, bool FRAME>
int launch_patch_t(const PatchArgs &a) {
  auto kernel = grid_patch_kernel<R, NW, CPT, FRAME>;
  constexpr size_t smem = sizeof(PatchSmem<NW, CPT>);
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // how many clusters the device can hold at once (a cluster lives inside one GPC)
  cfg.gridDim = dim3(a.cluster);
  int max_clusters = 0;
  cfg.gridDim = dim3(a.cluster * a.sm_count);  // (the query wants a grid to reason about)
  CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
  FPIE_REQUIRE(max_clusters > 0, "the persistent patch kernel does not fit this device");
  const int clusters = std::min(a.nitems, max_clusters);
  cfg.gridDim = dim3(clusters * a.cluster);
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, a.g, a.bm, a.x, a.hq, a.bits, a.nsweeps, a.nitems));
  return clusters;
}

template <int R, int CPT>
int launch_patch_f(const PatchArgs &a, bool frame) {
  return frame ? launch_patch_t<R, 8, CPT, true>(a) : launch_patch_t<R, 8, CPT, false>(a);
}

}  // namespace

// Small patches / small images: the persistent cluster kernel of patch.cuh.  A (patch, plane) needs
// ceil(rows / (8 warps x R rows)) CTAs of one cluster (<= 8, the portable limit) and at most 256 columns.
bool GridSolver::patch_shape(int *rows_per_thread, int *cols_per_thread, int *cluster) const {
  // (an explicit tile shape or blocking depth is a request for the tiled kernel)
  if (patch_off_ || !auto_tune_ || !auto_k_) return false;
  const bool single = batch_.batch == 0;
  const int ph = single ? geom_.n : batch_.ph, pw = single ? geom_.m : batch_.pw;
  if (pw > 256 || ph > 512) return false;
  // Measured on B200 (tools/patch_bench.py, profiles/r02_patch_bench_v2.json).  A sweep of the persistent kernel costs
  // 0.53 us per cluster with 8 rows per thread (a 256-row plane on 4 CTAs) and 0.42 us with 4 (8 CTAs), whatever the
  // number of items as long as all clusters are resident; the tiled kernel needs 0.52 us per sweep for ONE 256^2
  // image and 0.79 us for three.  So: 4 rows per thread while every item's cluster fits the device at once (two
  // such CTAs per SM), 8 rows per thread -- half the hand-overs per pixel -- for larger batches (512 patches of
  // 256^2: 1336 vs 1123 Gupd/s; the tiled kernel: 809); one full-square 256^2 image: 155 vs 125 Gupd/s, 128^2: 61 vs
  // 32; with an arbitrary mask (per-pixel selects: tools/single_image_bench.py) a single 160^2 ... 256^2 image still
  // runs 6 % faster here (0.49 vs 0.52 us per sweep).
  const int items = (single ? 1 : batch_.batch) * 3;
  const int cpt = pw <= 128 ? 4 : 8;
  int r = patch_rows_;
  if (r == 0) {
    const long long ctas4 = (long long)items * ceil_div(ph, 32);
    r = (ph <= 128 || (ceil_div(ph, 32) <= 8 && ctas4 <= 2ll * sm_count_)) ? 4 : 8;
  }
  if (ceil_div(ph, 8 * r) > 8) r = 8;
  const int cl = (int)ceil_div(ph, 8 * r);
  if (cl > 8) return false;
  if (rows_per_thread) *rows_per_thread = r;
  if (cols_per_thread) *cols_per_thread = cpt;
  if (cluster) *cluster = cl;
  return true;
}

void GridSolver::patch_sweeps(int iters) {
  int r = 0, cpt = 0, cl = 0;
  patch_shape(&r, &cpt, &cl);
  const bool single = batch_.batch == 0;
  PatchArgs a{};
  a.g = geom_;
  a.bm = single ? BatchMap{1, geom_.n, geom_.m, 1} : batch_;
  a.x = x_[cur_].ptr;
  a.hq = hq_.ptr;
  a.bits = bits_.ptr;
  a.nsweeps = iters;
  a.nitems = a.bm.batch * 3;
  a.cluster = cl;
  a.sm_count = sm_count_;
  a.stream = stream_;
  // every pixel but the 1-pixel frame of every patch is an unknown, and the patch fills the cluster's rows and
  // the warp's columns exactly: the select-free instruction stream
  const bool all_unknown = stats_.unknowns == (int64_t)a.bm.batch * (a.bm.ph - 2) * (a.bm.pw - 2);
  const bool frame = all_unknown && a.bm.ph == cl * 8 * r && a.bm.pw == 32 * cpt;
  if (r == 4 && cpt == 8)
    patch_clusters_ = launch_patch_f<4, 8>(a, frame);
  else if (r == 8 && cpt == 8)
    patch_clusters_ = launch_patch_f<8, 8>(a, frame);
  else if (r == 4 && cpt == 4)
    patch_clusters_ = launch_patch_f<4, 4>(a, frame);
  else if (r == 8 && cpt == 4)
    patch_clusters_ = launch_patch_f<8, 4>(a, frame);
  else
    throw Error("fpie_b200: no persistent patch kernel for this shape");
  stats_.launches += 1;
  patch_launches_ += 1;
  CUDA_CHECK(cudaGetLastError());
}

void GridSolver::sweeps_async(int iters) {
  require_ready();
  FPIE_REQUIRE(iters >= 0, "step: negative iteration count");
  DeviceGuard guard(device_);
  const PlaneGeom &g = geom_;
  if (stats_.unknowns == 0 || iters == 0) return;
  // batches of small patches and single small images: one persistent launch, the state never leaves the SMs
  if (iters >= patch_min_iters_ && patch_shape(nullptr, nullptr, nullptr)) {
    patch_sweeps(iters);
    return;
  }
  if (variant_ == 1) {
    const long long work = (long long)g.n * g.groups * 3;
    for (int i = 0; i < iters; ++i) {
      grid_sweep1_kernel<<<blocks_for(work, 256), 256, 0, stream_>>>(g, 3, bits_.ptr, x_[cur_].ptr, x_[cur_ ^ 1].ptr,
                                                                     hq_.ptr);
      cur_ ^= 1;
    }
    stats_.launches += iters;
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  SweepArgs a{};
  a.grid = sm_count_;
  a.stream = stream_;
  a.g = g;
  a.bits = bits_.ptr;
  a.hq = hq_.ptr;
  a.h16 = h16_ok_;
  a.tm_h = h16_ok_ ? &tm_h16_ : &tm_h_;
  a.tm_m = &tm_m_;
  a.tiles = tiles_.ptr;
  a.ntiles = n_tile_entries_;
  a.halo_y = block_k_;
  a.halo_x = halo_x_;
  int left = iters;
  auto one_pass = [&](SweepArgs &args, int &cur, int nsweeps) {
    args.nsweeps = nsweeps;
    args.xin = x_[cur].ptr;
    args.xout = x_[cur ^ 1].ptr;
    args.tm_x = &tm_x_[cur];
    args.reverse = (serpentine_ && cur) ? 1 : 0;
    launch_variant(variant_, args);
    cur ^= 1;
  };
  // long runs: replay a captured graph of kGraphPasses full passes (an even number, so the graph starts
  // and ends on the same state buffer) -- the launch-bound inner loop of a step costs one graph launch
  // per kGraphPasses kernels
  if (!graph_off_ && left >= 2 * kGraphPasses * block_k_) {
    if (!graph_warm_) {  // per-function attributes are set lazily at the first launch: not inside a capture
      for (int i = 0; i < 2; ++i) one_pass(a, cur_, block_k_);
      left -= 2 * block_k_;
      stats_.launches += 2;
      graph_warm_ = true;
    }
    if (!graph_[cur_]) {
      if (!cap_stream_) CUDA_CHECK(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
      SweepArgs b = a;
      b.stream = cap_stream_;  // (the solver's own stream may be the legacy default stream, which cannot capture)
      int c = cur_;
      CUDA_CHECK(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal));
      cudaGraph_t captured = nullptr;
      try {
        for (int i = 0; i < kGraphPasses; ++i) one_pass(b, c, block_k_);
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
    while (left >= kGraphPasses * block_k_) {
      CUDA_CHECK(cudaGraphLaunch(graph_[cur_], stream_));
      left -= kGraphPasses * block_k_;
      stats_.launches += kGraphPasses;
    }
  }
  while (left > 0) {
    const int ns = std::min(left, block_k_);
    one_pass(a, cur_, ns);
    left -= ns;
    stats_.launches += 1;
  }
  CUDA_CHECK(cudaGetLastError());
}

void GridSolver::drop_graphs() {
  for (auto &gx : graph_) {
    if (gx) cudaGraphExecDestroy(gx);
    gx = nullptr;
  }
  graph_warm_ = false;
}

// ---- split passes (row-band sharding: overlap the halo exchange with the interior of a pass) ----------
// Tiles whose stored rows intersect the first / last `rows` grid rows form the "edge" part of the tile
// list, the rest the "interior" part.  pass_async() runs ONE pass (<= block_k sweeps) over one part without
// flipping the state buffers; the caller runs both parts in the order it needs, then flip().
void GridSolver::set_edge_rows(int rows) {
  require_ready();
  FPIE_REQUIRE(rows >= 0, "set_edge_rows: negative row count");
  DeviceGuard guard(device_);
  const int step_y = shape_.tile_h() - 2 * block_k_;
  std::vector<int2> part[2];
  for (size_t i = 0; i < host_tiles_.size(); ++i) {
    const int lo = host_tile_row_[i] * step_y, hi = lo + step_y;  // stored rows of the tile
    const bool edge = lo < rows || hi > geom_.n - rows;
    part[edge ? 0 : 1].push_back(host_tiles_[i]);
  }
  for (int p = 0; p < 2; ++p) {
    tiles_part_[p].resize(std::max<size_t>(part[p].size(), 1));
    n_part_[p] = (int)part[p].size();
    if (!part[p].empty())
      CUDA_CHECK(cudaMemcpyAsync(tiles_part_[p].ptr, part[p].data(), part[p].size() * sizeof(int2),
                                 cudaMemcpyHostToDevice, stream_));
  }
  CUDA_CHECK(cudaStreamSynchronize(stream_));  // (the host vectors go out of scope)
  edge_rows_ = rows;
}


void GridSolver::pass_async(int nsweeps, int part) {
  require_ready();
  FPIE_REQUIRE(part == 0 || part == 1, "pass_async: part must be 0 (edge tiles) or 1 (interior tiles)");
  FPIE_REQUIRE(nsweeps >= 1 && nsweeps <= block_k_, "pass_async: a pass runs 1..block_k sweeps");
  FPIE_REQUIRE(edge_rows_ > 0, "pass_async needs set_edge_rows");
  FPIE_REQUIRE(variant_ != 1, "pass_async: not available for the one-sweep-per-launch kernels");
  DeviceGuard guard(device_);
  run_pass(nsweeps, tiles_part_[part].ptr, n_part_[part]);
  CUDA_CHECK(cudaGetLastError());
}

// CUDA loads kernels lazily, at their first launch, and loading may synchronise the whole context.  A band
// whose stream already waits for a neighbour's rows must never trigger that (several bands of one process
// would deadlock: the load waits for the stream, the stream for a band that cannot launch), so everything a
// step launches is resolved up front.
void GridSolver::preload_kernels() {
  DeviceGuard guard(device_);
  cudaFuncAttributes attr;
  CUDA_CHECK(cudaFuncGetAttributes(&attr, grid_residual_kernel<false>));
  CUDA_CHECK(cudaFuncGetAttributes(&attr, grid_residual_kernel<true>));
  CUDA_CHECK(cudaFuncGetAttributes(&attr, grid_to_u8_kernel));
  CUDA_CHECK(cudaFuncGetAttributes(&attr, planes_to_aos_kernel));
  CUDA_CHECK(cudaFuncGetAttributes(&attr, grid_sweep1_kernel));
  if (variant_ != 1) {
    SweepArgs a{};
    a.load_only = true;
    a.h16 = h16_ok_;
    launch_variant(variant_, a);
  }
}

// one pass (<= block_k sweeps) over the tiles of `tiles`, current buffer -> other buffer, no flip
void GridSolver::run_pass(int nsweeps, const int2 *tiles, int ntiles) {
  if (stats_.unknowns == 0 || ntiles == 0) return;
  SweepArgs a{};
  a.grid = sm_count_;
  a.stream = stream_;
  a.g = geom_;
  a.bits = bits_.ptr;
  a.hq = hq_.ptr;
  a.h16 = h16_ok_;
  a.tm_h = h16_ok_ ? &tm_h16_ : &tm_h_;
  a.tm_m = &tm_m_;
  a.tiles = tiles;
  a.ntiles = ntiles;
  a.halo_y = block_k_;
  a.halo_x = halo_x_;
  a.nsweeps = nsweeps;
  a.xin = x_[cur_].ptr;
  a.xout = x_[cur_ ^ 1].ptr;
  a.tm_x = &tm_x_[cur_];
  a.reverse = (serpentine_ && cur_) ? 1 : 0;  // (alternate passes walk their tile list backwards, as sweeps_async)
  launch_variant(variant_, a);
  stats_.launches += 1;
}

void GridSolver::flip() {
  require_ready();
  if (stats_.unknowns > 0) cur_ ^= 1;
}

void GridSolver::finish_async() {
  require_ready();
  DeviceGuard guard(device_);
  const PlaneGeom &g = geom_;
  CUDA_CHECK(cudaMemsetAsync(err_.ptr, 0, 3 * sizeof(double), stream_));
  if (batch_.batch > 0) CUDA_CHECK(cudaMemsetAsync(batch_err_.ptr, 0, batch_err_.bytes(), stream_));
  const long long work = (long long)(win_hi_ - win_lo_) * g.groups;
  if (work > 0 && stats_.unknowns > 0) {
    dim3 grid(blocks_for(work, 256), 3);
    if (resid_equ_)  // (an image-level reset in the EquSolver's formulation: the EquSolver's residual expression too)
      grid_residual_kernel<true><<<grid, 256, 0, stream_>>>(g, batch_, win_lo_, win_hi_, bits_.ptr, x_[cur_].ptr, hq_.ptr,
                                                            batch_.batch > 0 ? batch_err_.ptr : err_.ptr);
    else
      grid_residual_kernel<false><<<grid, 256, 0, stream_>>>(g, batch_, win_lo_, win_hi_, bits_.ptr, x_[cur_].ptr, hq_.ptr,
                                                             batch_.batch > 0 ? batch_err_.ptr : err_.ptr);
    CUDA_CHECK(cudaGetLastError());
    stats_.launches += 1;
  }
  grid_to_u8_kernel<<<blocks_for((long long)g.n * g.groups, 256), 256, 0, stream_>>>(g, batch_, x_[cur_].ptr, img_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
  CUDA_CHECK(cudaMemcpyAsync(host_err_, err_.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
}

void GridSolver::sync() {
  DeviceGuard guard(device_);
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void GridSolver::fetch(uint8_t *out_img, float *out_err3, int64_t row_stride, int row_lo, int row_hi) {
  require_ready();
  DeviceGuard guard(device_);
  if (row_hi < 0) row_hi = geom_.n;
  FPIE_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= geom_.n, "fetch: row range outside the grid");
  FPIE_REQUIRE(batch_.batch == 0 || (row_lo == 0 && row_hi == geom_.n), "fetch: row ranges are not available for batched patches");
  const size_t row_bytes = (size_t)geom_.m * 3;
  if (row_stride <= 0) row_stride = (int64_t)row_bytes;
  FPIE_REQUIRE((size_t)row_stride >= row_bytes, "fetch: destination row stride is smaller than a row");
  if (out_img && batch_.batch > 0)  // [batch, ph, pw, 3], packed
    CUDA_CHECK(cudaMemcpyAsync(out_img, img_.ptr, (size_t)batch_.batch * batch_.ph * batch_.pw * 3,
                               cudaMemcpyDeviceToHost, stream_));
  else if (out_img && row_hi > row_lo)
    CUDA_CHECK(cudaMemcpy2DAsync(out_img, (size_t)row_stride, img_.ptr + (size_t)row_lo * row_bytes, row_bytes, row_bytes,
                                 row_hi - row_lo, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (out_err3) {
    if (batch_.batch > 0) {  // [batch, 3]
      std::vector<double> e((size_t)batch_.batch * 3);
      CUDA_CHECK(cudaMemcpy(e.data(), batch_err_.ptr, e.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < e.size(); ++i) out_err3[i] = (float)e[i];
    } else {
      for (int c = 0; c < 3; ++c) out_err3[c] = (float)host_err_[c];
    }
  }
}

// Sweep until every channel's residual is <= tol (checked every `check_every` sweeps) or `max_iters`
// sweeps have run.  The state stays on the device; step(0) afterwards returns image and err.
int GridSolver::solve(int max_iters, int check_every, float tol, float *out_err3) {
  require_ready();
  FPIE_REQUIRE(max_iters >= 0 && check_every >= 1, "solve: max_iters must be >= 0 and check_every >= 1");
  FPIE_REQUIRE(batch_.batch == 0, "solve: not available for batched patches (per-patch residuals)");
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

void GridSolver::step(int iters, uint8_t *out_img, float *out_err3, int64_t row_stride) {
  sweeps_async(iters);
  finish_async();
  fetch(out_img, out_err3, row_stride);
}

void GridSolver::state(float *out) {
  require_ready();
  FPIE_REQUIRE(out, "state: null output");
  DeviceGuard guard(device_);
  const size_t pixels = (size_t)geom_.n * geom_.m;
  stage_.resize(pixels * 3);
  planes_to_aos_kernel<<<blocks_for(pixels, 256), 256, 0, stream_>>>(geom_, x_[cur_].ptr, stage_.ptr);
  CUDA_CHECK(cudaGetLastError());
  stats_.launches += 1;
  CUDA_CHECK(cudaMemcpyAsync(out, stage_.ptr, pixels * 12, cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void GridSolver::set_row_window(int lo, int hi) {
  require_ready();
  FPIE_REQUIRE(0 <= lo && lo <= hi && hi <= geom_.n, "row window outside the grid");
  win_lo_ = lo;
  win_hi_ = hi;
}

void GridSolver::band_view(int which, float **base, int64_t *plane_stride, int64_t *row_pitch, int *pad_rows,
                           int *pad_cols) {
  require_ready();
  FPIE_REQUIRE(which == 0 || which == 1, "band_view: buffer index must be 0 or 1");
  if (base) *base = x_[which].ptr;
  if (plane_stride) *plane_stride = geom_.plane;
  if (row_pitch) *row_pitch = geom_.pitch;
  if (pad_rows) *pad_rows = geom_.padr;
  if (pad_cols) *pad_cols = geom_.padc;
}

}  // namespace fpie
