// Packed-pair variant of the temporally blocked GridSolver kernel (sm_100a FFMA2).
//
// Blackwell adds `fma.rn.f32x2` (SASS FFMA2): one instruction, two independent fp32 FMAs on a
// 64-bit register pair.  The FMA pipe rate is unchanged (tools/probes/issue_probe.cu: 125 of 128
// lanes per clock and SM at half the issue slots of scalar FFMA), so the instruction count of the
// sweep -- what bounds the scalar kernel -- drops by ~40 % and the sweep becomes FMA-pipe bound.
// The two lanes of a pair must be lattice sites with identical neighbour structure, so a thread
// pairs row r of the tile's UPPER half with row r + TH/2 of its LOWER half (same columns): up /
// down / left / right neighbours of a pair are again aligned pairs, except across the seam
// between the halves, which the first / last warp patch by reading the other end's mailbox row
// with swapped components.  The tile, its halo logic, the TMA staging, the programmatic dependent
// launch and the HBM layout are exactly those of the scalar kernel (grid.cu).
//
// With two warps per scheduler nothing hides a shuffle's latency but the thread's own independent
// work, so the left / right neighbour shuffles of a row are issued kShuffleAhead rows before the
// row is updated (they read old values: Jacobi).
#pragma once

#include "tma.cuh"

namespace fpie {

typedef unsigned long long pair64;

__device__ __forceinline__ pair64 pk2(float lo, float hi) {
  pair64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void up2(pair64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float lo2(pair64 v) {
  float lo, hi;
  up2(v, lo, hi);
  return lo;
}
__device__ __forceinline__ float hi2(pair64 v) {
  float lo, hi;
  up2(v, lo, hi);
  return hi;
}
__device__ __forceinline__ pair64 ffma2(pair64 a, pair64 b, pair64 c) {
  pair64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ pair64 shfl_up_pair(pair64 v) {
  float lo, hi;
  up2(v, lo, hi);
  return pk2(__shfl_up_sync(0xffffffffu, lo, 1), __shfl_up_sync(0xffffffffu, hi, 1));
}
__device__ __forceinline__ pair64 shfl_down_pair(pair64 v) {
  float lo, hi;
  up2(v, lo, hi);
  return pk2(__shfl_down_sync(0xffffffffu, lo, 1), __shfl_down_sync(0xffffffffu, hi, 1));
}

struct alignas(16) PairRow {
  pair64 p[4];  // 4 pixels; .lo = row in the upper half, .hi = same row of the lower half
};

// mask nibbles of a thread's R pair-rows, 8 rows per word, upper / lower half
template <int R>
struct PairMask {
  static constexpr int W = (R + 7) / 8;
  uint32_t lo[W], hi[W];
  __device__ __forceinline__ uint32_t nib_lo(int i) const { return lo[i / 8] >> ((i % 8) * 4); }
  __device__ __forceinline__ uint32_t nib_hi(int i) const { return hi[i / 8] >> ((i % 8) * 4); }
};

// both components: ((((g + U) + D) + L) + R) / 4 on quarter-scaled operands (see jacobi_q)
__device__ __forceinline__ pair64 jacobi_q2(pair64 hq, pair64 up, pair64 dn, pair64 lf, pair64 rt, pair64 q) {
  pair64 t = ffma2(up, q, hq);
  t = ffma2(dn, q, t);
  t = ffma2(lf, q, t);
  return ffma2(rt, q, t);
}

// one pair-row; lf / rt = the left neighbour of pixel 0 / right neighbour of pixel 3 (already shuffled)
template <bool MIXED>
__device__ __forceinline__ void pair_row_update(PairRow &xi, const PairRow &hi, const PairRow &prev, const PairRow &nxt,
                                                pair64 lf, pair64 rt, uint32_t nib_lo, uint32_t nib_hi, pair64 q) {
  const PairRow cur = xi;
  PairRow o;
  o.p[0] = jacobi_q2(hi.p[0], prev.p[0], nxt.p[0], lf, cur.p[1], q);
  o.p[1] = jacobi_q2(hi.p[1], prev.p[1], nxt.p[1], cur.p[0], cur.p[2], q);
  o.p[2] = jacobi_q2(hi.p[2], prev.p[2], nxt.p[2], cur.p[1], cur.p[3], q);
  o.p[3] = jacobi_q2(hi.p[3], prev.p[3], nxt.p[3], cur.p[2], rt, q);
  if (MIXED) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float nl, nh, cl, ch;
      up2(o.p[j], nl, nh);
      up2(cur.p[j], cl, ch);
      nl = ((nib_lo >> j) & 1u) ? nl : cl;
      nh = ((nib_hi >> j) & 1u) ? nh : ch;
      o.p[j] = pk2(nl, nh);
    }
  }
  xi = o;
}

#ifndef FPIE_SHUFFLE_AHEAD
#define FPIE_SHUFFLE_AHEAD 2
#endif
constexpr int kShuffleAhead = FPIE_SHUFFLE_AHEAD;

// One sweep over the R pair-rows of a thread (= 2R tile rows), split-phase like tile_sweep_split:
// publish the strip's edge pair-rows, arrive, update the interior, then wait for the neighbours.
template <int R, int NW, bool MIXED>
__device__ __forceinline__ void pair_sweep(PairRow (&x)[R], const PairRow (&h)[R], const PairMask<R> &mb,
                                           PairRow (*mailbox)[2][NW][32], uint64_t *mail_bar, int parity,
                                           uint32_t &mphase, pair64 q) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  mailbox[parity][0][w][lane] = x[0];
  mailbox[parity][1][w][lane] = x[R - 1];
  __syncwarp();
  if (lane == 0) mbar_arrive(&mail_bar[parity]);
  const PairRow first_old = x[1];
  PairRow prev = x[0];
  PairRow up, dn;
  pair64 lf[R], rt[R];  // (fully unrolled: only the rows in flight are live)
  constexpr int PULL_ROW = (R >= 9) ? R - 4 : (R >= 6) ? R - 3 : R - 2;
  constexpr int LOOK = kShuffleAhead;
#pragma unroll
  for (int i = 1; i < 1 + LOOK && i < R - 1; ++i) {
    lf[i] = shfl_up_pair(x[i].p[3]);
    rt[i] = shfl_down_pair(x[i].p[0]);
  }
#pragma unroll
  for (int i = 1; i < R - 1; ++i) {
    if (i + LOOK < R - 1) {  // rows below i still hold the previous sweep's values
      lf[i + LOOK] = shfl_up_pair(x[i + LOOK].p[3]);
      rt[i + LOOK] = shfl_down_pair(x[i + LOOK].p[0]);
    }
    if (i == PULL_ROW) {
      lf[0] = shfl_up_pair(x[0].p[3]);
      rt[0] = shfl_down_pair(x[0].p[0]);
      lf[R - 1] = shfl_up_pair(x[R - 1].p[3]);
      rt[R - 1] = shfl_down_pair(x[R - 1].p[0]);
      mbar_wait(&mail_bar[parity], (mphase >> parity) & 1u);
      mphase ^= 1u << parity;
      // row above the strip: the previous warp's last pair-row; for warp 0 the upper half has no row
      // above it inside the tile (rim, any value) and the lower half continues the upper half's last
      // row = the LAST warp's bottom pair-row, upper component
      up = mailbox[parity][1][(w > 0) ? w - 1 : NW - 1][lane];
      dn = mailbox[parity][0][(w + 1 < NW) ? w + 1 : 0][lane];
      if (w == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) up.p[j] = pk2(lo2(up.p[j]), lo2(up.p[j]));
      }
      if (w == NW - 1) {  // below the upper half's last row lies the lower half's first row (warp 0, lower component)
#pragma unroll
        for (int j = 0; j < 4; ++j) dn.p[j] = pk2(hi2(dn.p[j]), hi2(dn.p[j]));
      }
    }
    const PairRow cur = x[i];
    pair_row_update<MIXED>(x[i], h[i], prev, x[i + 1], lf[i], rt[i], mb.nib_lo(i), mb.nib_hi(i), q);
    prev = cur;
  }
  pair_row_update<MIXED>(x[0], h[0], up, first_old, lf[0], rt[0], mb.nib_lo(0), mb.nib_hi(0), q);
  pair_row_update<MIXED>(x[R - 1], h[R - 1], prev, dn, lf[R - 1], rt[R - 1], mb.nib_lo(R - 1), mb.nib_hi(R - 1), q);
}

// Shared-memory layout: staging as in PipeSmem (tile of 2*R*NW rows), mailbox of pair-rows.
template <int R, int NW, bool H16>
struct PairSmem {
  static constexpr uint32_t align128(uint32_t v) { return (v + 127u) & ~127u; }
  static constexpr int TH = 2 * R * NW;
  static constexpr int H16_W = TILE_W + 8;
  static constexpr uint32_t X_BYTES = TH * TILE_W * 4;
  static constexpr uint32_t H_BYTES = H16 ? TH * H16_W * 2 : X_BYTES;
  static constexpr uint32_t M_BYTES = TH * MASK_BOX_WORDS * 4;
  static constexpr uint32_t X_OFF = 0;
  static constexpr uint32_t H_OFF = align128(X_OFF + X_BYTES);
  static constexpr uint32_t M_OFF = align128(H_OFF + H_BYTES);
  static constexpr uint32_t MAIL_OFF = align128(M_OFF + M_BYTES);
  static constexpr uint32_t TOTAL = MAIL_OFF + 2 * 2 * NW * 32 * sizeof(PairRow);
};

template <int R, int NW, int OCC, bool H16>
__global__ void __launch_bounds__(NW * 32, OCC)
grid_sweepk_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h,
                        const __grid_constant__ CUtensorMap tm_m, PlaneGeom g, float *__restrict__ xout,
                        const int2 *__restrict__ tiles, int ntiles, int nsweeps, int halo_y, int halo_x, int reverse) {
  constexpr int HH = R * NW;  // rows per half
  constexpr int TH = 2 * HH;
  using L = PairSmem<R, NW, H16>;
  constexpr int H16_W = L::H16_W;
  constexpr uint32_t TILE_BYTES = L::X_BYTES, H_BYTES = L::H_BYTES, MASK_BYTES = L::M_BYTES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *sx = reinterpret_cast<float *>(smem_raw + L::X_OFF);
  unsigned char *sh = smem_raw + L::H_OFF;
  uint32_t *sm = reinterpret_cast<uint32_t *>(smem_raw + L::M_OFF);
  PairRow(*mailbox)[2][NW][32] = reinterpret_cast<PairRow(*)[2][NW][32]>(smem_raw + L::MAIL_OFF);
  __shared__ uint64_t bars[3];  // [0] TMA landing, [1..2] edge exchange per sweep parity
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const pair64 q = pk2(0.25f, 0.25f);

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], NW);
    mbar_init(&bars[2], NW);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int2 d) {
    const TileRef r = unpack_tile(d);
    mbar_expect_tx(&bars[0], TILE_BYTES + H_BYTES + (r.full ? 0u : MASK_BYTES));
    tma_load_3d(sx, &tm_x, r.pcol, r.prow, r.plane, &bars[0]);
    tma_load_3d(sh, &tm_h, H16 ? (r.pcol & ~7) : r.pcol, r.prow, r.plane, &bars[0]);
    if (!r.full) tma_load_2d(sm, &tm_m, (r.pcol >> 5) & ~3, r.prow, &bars[0]);
  };

  int t = blockIdx.x;
  if (t >= ntiles) return;
  const int stride = gridDim.x;
  if (reverse) tiles += ntiles - 1;  // alternate passes walk the list backwards (L2 reuse, see the scalar kernel)
  const int dir = reverse ? -1 : 1;
  int2 cur = tiles[dir * t];
  int2 nxt = (t + stride < ntiles) ? tiles[dir * (t + stride)] : cur;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) issue(cur);
  int parity = 0;
  uint32_t phase = 0, mphase = 0;
  for (; t < ntiles; t += stride) {
    const int2 nxt2 = (t + 2 * stride < ntiles) ? tiles[dir * (t + 2 * stride)] : nxt;
    const TileRef td = unpack_tile(cur);
    const int pcol = td.pcol + 4 * lane;
    // global offsets of the thread's first row in the upper / lower half
    const long long base_lo = (long long)td.plane * g.plane + (long long)(td.prow + w * R) * g.pitch + pcol;
    PairRow x[R], h[R];
    PairMask<R> mb;
#pragma unroll
    for (int i = 0; i < PairMask<R>::W; ++i) mb.lo[i] = mb.hi[i] = td.full ? 0xffffffffu : 0u;

    mbar_wait(&bars[0], phase);
    phase ^= 1;
    {
      const float *xlo = sx + (w * R) * TILE_W + 4 * lane;
      const float *xhi = xlo + HH * TILE_W;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const float4 a = ld4(xlo + i * TILE_W), b = ld4(xhi + i * TILE_W);
        x[i].p[0] = pk2(a.x, b.x);
        x[i].p[1] = pk2(a.y, b.y);
        x[i].p[2] = pk2(a.z, b.z);
        x[i].p[3] = pk2(a.w, b.w);
      }
      if (H16) {
        const __half *hlo = reinterpret_cast<const __half *>(sh) + (w * R) * H16_W + (td.pcol & 7) + 4 * lane;
        const __half *hhi = hlo + HH * H16_W;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const uint2 ra = *reinterpret_cast<const uint2 *>(hlo + i * H16_W);
          const uint2 rb = *reinterpret_cast<const uint2 *>(hhi + i * H16_W);
          const float2 a0 = __half22float2(*reinterpret_cast<const __half2 *>(&ra.x));
          const float2 a1 = __half22float2(*reinterpret_cast<const __half2 *>(&ra.y));
          const float2 b0 = __half22float2(*reinterpret_cast<const __half2 *>(&rb.x));
          const float2 b1 = __half22float2(*reinterpret_cast<const __half2 *>(&rb.y));
          h[i].p[0] = pk2(a0.x, b0.x);
          h[i].p[1] = pk2(a0.y, b0.y);
          h[i].p[2] = pk2(a1.x, b1.x);
          h[i].p[3] = pk2(a1.y, b1.y);
        }
      } else {
        const float *hlo = reinterpret_cast<const float *>(sh) + (w * R) * TILE_W + 4 * lane;
        const float *hhi = hlo + HH * TILE_W;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float4 a = ld4(hlo + i * TILE_W), b = ld4(hhi + i * TILE_W);
          h[i].p[0] = pk2(a.x, b.x);
          h[i].p[1] = pk2(a.y, b.y);
          h[i].p[2] = pk2(a.z, b.z);
          h[i].p[3] = pk2(a.w, b.w);
        }
      }
      if (!td.full) {
        const uint32_t *mlo = sm + (w * R) * MASK_BOX_WORDS + ((pcol >> 5) - ((td.pcol >> 5) & ~3));
        const uint32_t *mhi = mlo + HH * MASK_BOX_WORDS;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          mb.lo[i / 8] |= ((mlo[i * MASK_BOX_WORDS] >> (pcol & 31)) & 0xFu) << ((i % 8) * 4);
          mb.hi[i / 8] |= ((mhi[i * MASK_BOX_WORDS] >> (pcol & 31)) & 0xFu) << ((i % 8) * 4);
        }
      }
    }
    __syncthreads();  // every thread has drained the staging buffers
    if (threadIdx.x == 0 && t + stride < ntiles) issue(nxt);

#pragma unroll 2
    for (int s = 0; s < nsweeps; ++s) {
      if (td.full)
        pair_sweep<R, NW, false>(x, h, mb, mailbox, &bars[1], parity, mphase, q);
      else
        pair_sweep<R, NW, true>(x, h, mb, mailbox, &bars[1], parity, mphase, q);
      parity ^= 1;
    }

    // store the inner region of both halves (branch-free, predicated 128-bit stores)
    {
      const bool lane_ok = (4 * lane >= halo_x) && (4 * lane < TILE_W - halo_x);
      const uint32_t pitch_bytes = (uint32_t)g.pitch * 4u;
      char *out_lo = reinterpret_cast<char *>(xout + base_lo);
      char *out_hi = out_lo + (size_t)HH * pitch_bytes;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int tr_lo = w * R + i, tr_hi = HH + w * R + i;
        const uint32_t nl = td.full ? 1u : mb.nib_lo(i) & 0xFu, nh = td.full ? 1u : mb.nib_hi(i) & 0xFu;
        const float4 a = make_float4(lo2(x[i].p[0]), lo2(x[i].p[1]), lo2(x[i].p[2]), lo2(x[i].p[3]));
        const float4 b = make_float4(hi2(x[i].p[0]), hi2(x[i].p[1]), hi2(x[i].p[2]), hi2(x[i].p[3]));
        st4_if(out_lo + (size_t)i * pitch_bytes, a, lane_ok && tr_lo >= halo_y && nl);
        st4_if(out_hi + (size_t)i * pitch_bytes, b, lane_ok && tr_hi < TH - halo_y && nh);
      }
    }
    cur = nxt;
    nxt = nxt2;
  }
}

}  // namespace fpie
