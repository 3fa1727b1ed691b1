// Cluster variant of the temporally blocked GridSolver kernel: CL vertically adjacent CTAs of one thread-block
// cluster sweep ONE tall tile (CL * R * NW rows x 128 columns) together.
//
// Why: one CTA per SM with a 168-row tile serialises its per-tile phases (wait for the TMA prefetch, shared
// memory -> registers, k sweeps, store) -- 28 % of the time the FMA pipes idle.  Two CTAs per SM overlap those
// phases but halve the tile height, and a short tile pays for its 2k halo rows in redundant work (84 rows at
// k = 8: 19 % of the rows are halo; 168 rows: 9.5 %).  A cluster removes the trade: each CTA keeps a short
// strip (two CTAs per SM, from different clusters, in independent phases), but the strips of one cluster form
// one tile with halo rows at its outer ends only.  Inside the tile the boundary warps exchange their edge rows
// every sweep through distributed shared memory: one st.async per lane straight into the neighbour CTA's
// mailbox, completion counted as transaction bytes on the neighbour's mbarrier -- a single one-way message,
// hidden behind the interior rows of the split-phase sweep exactly like the local mailbox (grid.cu).
#pragma once

#include "patch.cuh"  // cluster / DSMEM helpers

namespace fpie {

// one 16-byte row segment into another CTA's shared memory, signalling its mbarrier on arrival
__device__ __forceinline__ void st_async_row(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  st_async_cluster4(remote_addr, v, remote_bar);
}

// Shared-memory layout of one CTA of the cluster kernel.
template <int R, int NW, bool H16>
struct ClusterSmem {
  static constexpr uint32_t align128(uint32_t v) { return (v + 127u) & ~127u; }
  static constexpr int TH = R * NW;  // rows of ONE CTA's strip
  static constexpr int H16_W = TILE_W + 8;
  static constexpr uint32_t X_BYTES = TH * TILE_W * 4;
  static constexpr uint32_t H_BYTES = H16 ? TH * H16_W * 2 : X_BYTES;
  static constexpr uint32_t M_BYTES = TH * MASK_BOX_WORDS * 4;
  static constexpr uint32_t X_OFF = 0;
  static constexpr uint32_t H_OFF = align128(X_OFF + X_BYTES);
  static constexpr uint32_t M_OFF = align128(H_OFF + H_BYTES);
  static constexpr uint32_t MAIL_OFF = align128(M_OFF + M_BYTES);
  // mailbox [parity][slot][lane] of float4: slots 0..NW-1 top rows, NW..2NW-1 bottom rows of the warps,
  // 2NW = bottom row of the CTA above, 2NW+1 = top row of the CTA below (both written remotely)
  static constexpr int SLOTS = 2 * NW + 2;
  static constexpr uint32_t MAIL_BYTES = 2 * SLOTS * 32 * 16;
  static constexpr uint32_t TOTAL = MAIL_OFF + MAIL_BYTES;
};

template <int R, int NW, bool MIXED>
__device__ __forceinline__ void tile_sweep_cluster(float4 (&x)[R], const float4 (&h)[R], const uint32_t (&mb)[(R + 7) / 8],
                                                   float4 (*mail)[32], uint64_t *bar, uint32_t phase_bit, bool has_up,
                                                   bool has_dn, uint32_t rem_mail, uint32_t rem_bar,
                                                   uint32_t remote_bytes) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (rem_bar) {  // boundary warp of the strip: the edge row first goes to the neighbour CTA (longest latency)
    const bool to_up = (w == 0 && has_up);
    const float4 a = x[0], b = x[R - 1];
    st_async_row(rem_mail, make_float4(to_up ? a.x : b.x, to_up ? a.y : b.y, to_up ? a.z : b.z, to_up ? a.w : b.w),
                 rem_bar);
  }
  mail[w][lane] = x[0];
  mail[NW + w][lane] = x[R - 1];
  __syncwarp();
  if (lane == 0) {
    if (w == 0 && remote_bytes)
      mbar_expect_tx(bar, remote_bytes);  // (arrives and announces the neighbours' bytes of this phase)
    else
      mbar_arrive(bar);
  }
  const float4 first_old = x[1];
  float4 prev = x[0];
  float4 up, dn;
#ifndef FPIE_CL_PULL
#define FPIE_CL_PULL 8
#endif
  // (the neighbour CTA's row crosses the SM-to-SM network: pulled later than the local kernel's R - 8)
  constexpr int PULL_ROW = (R >= 18) ? R - FPIE_CL_PULL : (R >= 8) ? R - 4 : R - 2;
#pragma unroll
  for (int i = 1; i < R - 1; ++i) {
    if (i == PULL_ROW) {
      mbar_wait(bar, phase_bit);
      up = (w > 0) ? mail[NW + w - 1][lane] : (has_up ? mail[2 * NW][lane] : x[0]);
      dn = (w + 1 < NW) ? mail[w + 1][lane] : (has_dn ? mail[2 * NW + 1][lane] : x[R - 1]);
    }
    const float4 cur = x[i];
    row_update<MIXED>(x[i], h[i], prev, x[i + 1], mb[i / 8] >> ((i % 8) * 4));
    prev = cur;
  }
  row_update<MIXED>(x[0], h[0], up, first_old, mb[0]);
  row_update<MIXED>(x[R - 1], h[R - 1], prev, dn, mb[(R - 1) / 8] >> (((R - 1) % 8) * 4));
}

// Tile descriptors name the whole cluster tile (CL * TH rows); CTA `crank` of the cluster loads, sweeps and
// stores rows [crank * TH, (crank + 1) * TH) of it.  Clusters stride over the tile list.
template <int R, int NW, int OCC, bool H16>
__global__ void __launch_bounds__(NW * 32, OCC)
grid_sweepk_cluster_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h,
                           const __grid_constant__ CUtensorMap tm_m, PlaneGeom g, float *__restrict__ xout,
                           const int2 *__restrict__ tiles, int ntiles, int nsweeps, int halo_y, int halo_x, int reverse) {
  using L = ClusterSmem<R, NW, H16>;
  constexpr int TH = L::TH;
  constexpr int H16_W = L::H16_W;
  constexpr uint32_t TILE_BYTES = L::X_BYTES, H_BYTES = L::H_BYTES, MASK_BYTES = L::M_BYTES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *sx = reinterpret_cast<float *>(smem_raw + L::X_OFF);
  unsigned char *sh = smem_raw + L::H_OFF;
  uint32_t *sm = reinterpret_cast<uint32_t *>(smem_raw + L::M_OFF);
  float4(*mailbox)[L::SLOTS][32] = reinterpret_cast<float4(*)[L::SLOTS][32]>(smem_raw + L::MAIL_OFF);
  __shared__ uint64_t bars[3];  // [0] TMA landing, [1..2] edge exchange per sweep parity
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
  const bool has_up = crank > 0, has_dn = crank + 1 < csize;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], NW);
    mbar_init(&bars[2], NW);
    mbar_fence_init();
  }
  cluster_sync_all();  // every CTA's barriers exist before a neighbour's rows can arrive

  // where this warp's edge row goes in the neighbour CTA (boundary warps only)
  uint32_t rem_mail = 0, rem_bar = 0;
  {
    const uint32_t mail_base = smem_u32(&mailbox[0][0][0]);
    const uint32_t bar_base = smem_u32(&bars[1]);
    if (w == 0 && has_up) {
      rem_mail = map_to_cta(mail_base + ((2 * NW + 1) * 32 + lane) * 16, crank - 1);
      rem_bar = map_to_cta(bar_base, crank - 1);
    } else if (w == NW - 1 && has_dn) {
      rem_mail = map_to_cta(mail_base + ((2 * NW) * 32 + lane) * 16, crank + 1);
      rem_bar = map_to_cta(bar_base, crank + 1);
    }
  }
  constexpr uint32_t PARITY_BYTES = L::SLOTS * 32 * 16;
  const uint32_t remote_bytes = ((has_up ? 1u : 0u) + (has_dn ? 1u : 0u)) * 512u;

  auto issue = [&](int2 d) {  // one thread: arm the barrier and start the strip's bulk loads
    const TileRef r = unpack_tile(d);
    const int prow = r.prow + crank * TH;
    mbar_expect_tx(&bars[0], TILE_BYTES + H_BYTES + (r.full ? 0u : MASK_BYTES));
    tma_load_3d(sx, &tm_x, r.pcol, prow, r.plane, &bars[0]);
    tma_load_3d(sh, &tm_h, H16 ? (r.pcol & ~7) : r.pcol, prow, r.plane, &bars[0]);
    if (!r.full) tma_load_2d(sm, &tm_m, (r.pcol >> 5) & ~3, prow, &bars[0]);
  };

  int t = (int)cluster_id_x();
  const int stride = (int)cluster_count_x();
  if (t < ntiles) {
    if (reverse) tiles += ntiles - 1;
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
      const int strip_row = crank * TH + w * R;  // first row of this thread inside the cluster tile
      const long long base = (long long)td.plane * g.plane + (long long)(td.prow + strip_row) * g.pitch + pcol;
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

#pragma unroll 2
      for (int s = 0; s < nsweeps; ++s) {
        float4(*mail)[32] = mailbox[parity];
        const uint32_t rm = rem_mail + parity * PARITY_BYTES, rb = rem_bar + parity * 8;
        if (td.full)
          tile_sweep_cluster<R, NW, false>(x, h, mb, mail, &bars[1 + parity], (mphase >> parity) & 1u, has_up, has_dn,
                                           rem_bar ? rm : 0u, rem_bar ? rb : 0u, remote_bytes);
        else
          tile_sweep_cluster<R, NW, true>(x, h, mb, mail, &bars[1 + parity], (mphase >> parity) & 1u, has_up, has_dn,
                                          rem_bar ? rm : 0u, rem_bar ? rb : 0u, remote_bytes);
        mphase ^= 1u << parity;
        parity ^= 1;
      }
      // store the inner region of the CLUSTER tile (branch-free, predicated 128-bit stores)
      {
        const int total_h = csize * TH;
        const int lo = max(halo_y - strip_row, 0), hi = min(total_h - halo_y - strip_row, R);
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
  } else {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  cluster_sync_all();  // nobody leaves while a neighbour may still store into its mailbox
}

}  // namespace fpie
