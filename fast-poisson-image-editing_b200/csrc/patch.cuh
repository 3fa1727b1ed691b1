// Persistent small-image kernel (BASELINE config 5: batches of independent small edits; the GUI's
// reset + step per click, fpie/gui.py:96-99).
//
// A small patch does not need temporal blocking with redundant halos at all: one channel plane of a
// 256 x 256 patch is 256 KB of fp32 state -- it fits the register files of a few SMs.  A thread-block
// CLUSTER owns one (patch, plane) for the WHOLE step: the state and the quarter-gradient are read from
// HBM once, every sweep runs on registers, and the plane is written back once.  CTA c of the cluster holds
// rows [c * NW * R, (c + 1) * NW * R); inside a CTA the rows above / below a warp's strip travel through a
// shared-memory mailbox exactly as in the tiled kernel (grid.cu, tile_sweep_split); between CTAs the
// boundary warp stores its edge row straight into the neighbour CTA's mailbox through distributed shared
// memory (st.shared::cluster) and arrives on the neighbour's mbarrier (release / acquire at cluster
// scope).  One sweep = one split-phase barrier per CTA: publish the edge rows, update the interior rows,
// wait, update the two edge rows.  No halo, no recomputation, no HBM traffic between the first and the last
// sweep of a step; the arithmetic (and therefore every bit of the result) is that of the tiled kernel.
//
// A thread owns R rows x CPT columns (CPT = 4 or 8), a warp spans the whole patch width (<= 128 or 256
// columns), so there is no seam between column strips; left / right neighbours come from warp shuffles.
#pragma once

#include "tma.cuh"

// tuning knobs of the persistent kernel's sweep loop (measured defaults; -D to experiment)
#ifndef FPIE_PATCH_UNROLL
#define FPIE_PATCH_UNROLL 2
#endif
#ifndef FPIE_PATCH_PULL
#define FPIE_PATCH_PULL 3
#endif
#ifndef FPIE_PATCH_UNROLL_SHORT
#define FPIE_PATCH_UNROLL_SHORT 2
#endif

namespace fpie {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared-memory address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// 16 bytes into another CTA's shared memory; the store itself signals the remote mbarrier when it lands
// (complete_tx): data and notification are ONE one-way message -- no release fence, no separate arrive, i.e.
// none of the round trips a st + mbarrier.arrive.release.cluster pair costs (measured: 0.82 -> see DESIGN.md)
__device__ __forceinline__ void st_async_cluster4(uint32_t addr, float4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// local arrive / wait with cluster-scope ordering (remote CTAs' stores must be visible after the wait)
__device__ __forceinline__ void mbar_arrive_rel_cluster(uint64_t *bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PWAIT_DONE;\n"
      "bra PWAIT_LOOP;\n"
      "PWAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// CPT columns of one row
template <int CPT>
struct RowVec {
  float4 v[CPT / 4];
};

// One row of CPT pixels.  FRAME == false: bit j of `sel` set = pixel j is an unknown (updated), clear = it
// keeps its value.  FRAME == true: every pixel is an unknown except the thread's first one when `sel` bit 0
// is set and its last one when bit 1 is set (the patch's first / last column: thread-constant predicates).
template <int CPT, bool FRAME>
__device__ __forceinline__ void patch_row_update(RowVec<CPT> &xi, const RowVec<CPT> &hi, const RowVec<CPT> &prev,
                                                 const RowVec<CPT> &nxt, uint32_t sel) {
  const RowVec<CPT> cur = xi;
  const float lf = __shfl_up_sync(0xffffffffu, cur.v[CPT / 4 - 1].w, 1);
  const float rt = __shfl_down_sync(0xffffffffu, cur.v[0].x, 1);
#pragma unroll
  for (int q = 0; q < CPT / 4; ++q) {
    const float4 c = cur.v[q], h = hi.v[q], u = prev.v[q], d = nxt.v[q];
    const float left = (q == 0) ? lf : cur.v[q > 0 ? q - 1 : 0].w;
    const float right = (q == CPT / 4 - 1) ? rt : cur.v[q + 1 < CPT / 4 ? q + 1 : q].x;
    float4 o;
    o.x = jacobi_q(h.x, u.x, d.x, left, c.y);
    o.y = jacobi_q(h.y, u.y, d.y, c.x, c.z);
    o.z = jacobi_q(h.z, u.z, d.z, c.y, c.w);
    o.w = jacobi_q(h.w, u.w, d.w, c.z, right);
    if (!FRAME) {
      const uint32_t nib = sel >> (4 * q);
      o.x = (nib & 1u) ? o.x : c.x;
      o.y = (nib & 2u) ? o.y : c.y;
      o.z = (nib & 4u) ? o.z : c.z;
      o.w = (nib & 8u) ? o.w : c.w;
    } else {
      if (q == 0) o.x = (sel & 1u) ? c.x : o.x;
      if (q == CPT / 4 - 1) o.w = (sel & 2u) ? c.w : o.w;
    }
    xi.v[q] = o;
  }
}

// patch_row_update with the left / right neighbour values handed in: a sweep's shuffles depend on OLD values only,
// so the kernel issues all of them at the top of the sweep (one convergence check for the lot instead of one in
// front of every row's pair, and none behind the barrier)
template <int CPT, bool FRAME>
__device__ __forceinline__ void patch_row_update_lr(RowVec<CPT> &xi, const RowVec<CPT> &hi, const RowVec<CPT> &prev,
                                                    const RowVec<CPT> &nxt, uint32_t sel, float lf, float rt) {
  const RowVec<CPT> cur = xi;
#pragma unroll
  for (int q = 0; q < CPT / 4; ++q) {
    const float4 c = cur.v[q], h = hi.v[q], u = prev.v[q], d = nxt.v[q];
    const float left = (q == 0) ? lf : cur.v[q > 0 ? q - 1 : 0].w;
    const float right = (q == CPT / 4 - 1) ? rt : cur.v[q + 1 < CPT / 4 ? q + 1 : q].x;
    float4 o;
    o.x = jacobi_q(h.x, u.x, d.x, left, c.y);
    o.y = jacobi_q(h.y, u.y, d.y, c.x, c.z);
    o.z = jacobi_q(h.z, u.z, d.z, c.y, c.w);
    o.w = jacobi_q(h.w, u.w, d.w, c.z, right);
    if (!FRAME) {
      const uint32_t nib = sel >> (4 * q);
      o.x = (nib & 1u) ? o.x : c.x;
      o.y = (nib & 2u) ? o.y : c.y;
      o.z = (nib & 4u) ? o.z : c.z;
      o.w = (nib & 8u) ? o.w : c.w;
    } else {
      if (q == 0) o.x = (sel & 1u) ? c.x : o.x;
      if (q == CPT / 4 - 1) o.w = (sel & 2u) ? c.w : o.w;
    }
    xi.v[q] = o;
  }
}

// Mailbox of one CTA: [parity][slot][16-byte column group][lane] -- a warp's 128-bit accesses to one (slot, group)
// are 512 contiguous bytes, conflict-free (a [slot][lane] array of 32-byte rows costs two wavefronts per access).
// slot 0..NW-1 = top rows of the warps, NW..2NW-1 = bottom rows, 2NW = bottom row of the CTA above (written
// remotely), 2NW+1 = top row of the CTA below (written remotely).
template <int NW, int CPT>
struct PatchSmem {
  static constexpr int SLOTS = 2 * NW + 2;
  float4 mail[2][SLOTS][CPT / 4][32];
  uint64_t bar[2];
};

// FRAME: every pixel of every patch is an unknown except the 1-pixel frame (the full-square masks of
// config 5's throughput run and of any whole-image blend): the interior rows of a strip run a select-free
// stream with two thread-constant column predicates; the strip's first / last row -- which may be the patch's
// first / last row, never updated -- take per-pixel selects from a thread-constant mask.  Otherwise the
// per-pixel mask bits select everywhere (arbitrary masks, same results).
//
// The sweep loop is branch-free apart from the barrier spin: where a strip's neighbour rows come from (the
// next warp's slot, the slot a neighbour CTA writes, or -- at the patch's edge, where the row is never used -- the
// warp's own slot) is an address computed once, the rows for the neighbour CTAs leave behind ONE
// warp-uniform branch, and the loop is unrolled twice so that the register rotation at its back edge is paid every other
// sweep.  (Before: 399 issued instructions per warp and sweep for 256 FFMA, 8 % of the samples resolving
// branches, two-way bank conflicts on every mailbox access -- profiles/r02_patch_ncu_full_summary.txt.)
template <int R, int NW, int CPT, bool FRAME>
__global__ void __launch_bounds__(NW * 32, (R <= 4) ? 2 : 1)
grid_patch_kernel(PlaneGeom g, BatchMap bm, float *__restrict__ x, const float *__restrict__ hq,
                  const uint32_t *__restrict__ bits, int nsweeps, int nitems) {
  static_assert(R >= 2 && NW >= 2, "a strip has a first and a last row; a CTA has a first and a last warp");
  extern __shared__ __align__(16) unsigned char patch_smem_raw[];
  using S = PatchSmem<NW, CPT>;
  S &sm = *reinterpret_cast<S *>(patch_smem_raw);
  constexpr int Q = CPT / 4;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
  const bool has_up = crank > 0, has_dn = crank + 1 < csize;
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar[0], NW);  // one arrival per warp; the neighbour CTAs' rows count as transaction bytes
    mbar_init(&sm.bar[1], NW);
    mbar_fence_init();
  }
  cluster_sync_all();  // every CTA's barriers exist before anyone arrives remotely

  constexpr uint32_t SLOT_BYTES = Q * 32 * 16;
  constexpr uint32_t PARITY_BYTES = S::SLOTS * SLOT_BYTES;
  const uint32_t mail_base = smem_u32(&sm.mail[0][0][0][0]);
  const uint32_t bar_base = smem_u32(&sm.bar[0]);
  // my top row goes to the CTA above (its slot 2NW+1), my bottom row to the CTA below (slot 2NW)
  // (warp 0 sends up, warp NW - 1 down: never both, NW >= 2)
  const bool send_up = (w == 0 && has_up), send_dn = (w == NW - 1 && has_dn);
  uint32_t rem_mail = 0, rem_bar = 0;
  if (send_up) {
    rem_mail = map_to_cta(mail_base + (2 * NW + 1) * SLOT_BYTES + lane * 16, crank - 1);
    rem_bar = map_to_cta(bar_base, crank - 1);
  } else if (send_dn) {
    rem_mail = map_to_cta(mail_base + (2 * NW) * SLOT_BYTES + lane * 16, crank + 1);
    rem_bar = map_to_cta(bar_base, crank + 1);
  }
  // where the rows above / below this warp's strip are read from (byte offsets inside a parity block)
  const int up_slot = (w > 0) ? NW + w - 1 : (has_up ? 2 * NW : w);
  const int dn_slot = (w + 1 < NW) ? w + 1 : (has_dn ? 2 * NW + 1 : NW + w);
  const uint32_t my_top = mail_base + w * SLOT_BYTES + lane * 16, my_bot = mail_base + (NW + w) * SLOT_BYTES + lane * 16;
  const uint32_t up_src = mail_base + up_slot * SLOT_BYTES + lane * 16, dn_src = mail_base + dn_slot * SLOT_BYTES + lane * 16;
  const uint32_t remote_bytes = ((has_up ? 1u : 0u) + (has_dn ? 1u : 0u)) * SLOT_BYTES;
  const uint32_t arm_bytes = (w == 0) ? remote_bytes : 0u;
  auto sts4 = [](uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  };
  auto lds4 = [](uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
  };

  int parity = 0;
  uint32_t mphase = 0;
  const int rows_per_cta = NW * R;
  for (int item = (int)cluster_id_x(); item < nitems; item += (int)cluster_count_x()) {
    const int patch = item / 3, plane = item % 3;
    const int by = patch / bm.bcols, bx = patch % bm.bcols;
    const int r0 = crank * rows_per_cta + w * R;  // first patch row of this thread
    const int c0 = lane * CPT;                    // first patch column
    const long long base = (long long)plane * g.plane + (long long)(g.padr + by * bm.ph + r0) * g.pitch + g.padc +
                           bx * bm.pw + c0;
    RowVec<CPT> xr[R], hr[R];
    uint32_t sel[R];
    const bool col_in = c0 < bm.pw;  // (pw is a multiple of 4 and CPT | 8: a thread's columns are all in or handled by sel)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const bool in = col_in && (r0 + i) < bm.ph;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const bool inq = in && (c0 + 4 * q) < bm.pw;
        xr[i].v[q] = inq ? ld4(x + base + (long long)i * g.pitch + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        hr[i].v[q] = inq ? ld4(hq + base + (long long)i * g.pitch + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t s = 0;
      if (in) {
        const int pc = g.padc + bx * bm.pw + c0;
        const uint32_t *wp = bits + (long long)(g.padr + by * bm.ph + r0 + i) * g.wpitch + (pc >> 5);
        const uint32_t lo = wp[0];
        const uint32_t hi = ((pc & 31) + CPT > 32) ? wp[1] : 0u;
        s = __funnelshift_r(lo, hi, pc & 31) & ((1u << CPT) - 1u);
        // columns beyond the patch width belong to the neighbour patch of the mosaic: not ours
        if (c0 + CPT > bm.pw) s &= (1u << (bm.pw - c0)) - 1u;
      }
      sel[i] = s;
    }

    // FRAME (launcher guarantees ph == cluster rows, pw == 32 * CPT): thread-constant column predicates for the
    // interior rows of a strip; the strip's first / last row take the mask bits (all clear on the patch's first /
    // last row, the frame columns clear elsewhere -- exactly what was loaded into sel[0] / sel[R-1])
    const uint32_t fsel = (lane == 0 ? 1u : 0u) | (lane == 31 ? 2u : 0u);
    const bool top_frame = !has_up && w == 0, bottom_frame = !has_dn && w == NW - 1;
    constexpr int kPatchUnroll = (R >= 8) ? FPIE_PATCH_UNROLL : FPIE_PATCH_UNROLL_SHORT;
#pragma unroll kPatchUnroll
    for (int sw = 0; sw < nsweeps; ++sw) {
      const uint32_t poff = (uint32_t)parity * PARITY_BYTES;
      uint64_t *bar = &sm.bar[parity];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        sts4(my_top + poff + q * 512, xr[0].v[q]);
        sts4(my_bot + poff + q * 512, xr[R - 1].v[q]);
      }
      if (rem_bar) {  // boundary warps (one warp-uniform branch per sweep): the row also goes to the neighbour CTA
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float4 a = xr[0].v[q], b = xr[R - 1].v[q];
          const float4 e = make_float4(send_up ? a.x : b.x, send_up ? a.y : b.y, send_up ? a.z : b.z, send_up ? a.w : b.w);
          st_async_cluster4(rem_mail + poff + q * 512, e, rem_bar + parity * 8);
        }
      }
      __syncwarp();
      // one arrival per warp (lane 0, a predicated instruction rather than a branch); warp 0's arrival also
      // announces the bytes the neighbour CTAs store into this CTA's mailbox in this phase (0 bytes = a plain arrive)
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.eq.u32 p, %2, 0;\n"
          "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
          "}\n" ::"r"(smem_u32(bar)),
          "r"(arm_bytes), "r"((uint32_t)lane)
          : "memory");
      float lf[R], rt[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        lf[i] = __shfl_up_sync(0xffffffffu, xr[i].v[Q - 1].w, 1);
        rt[i] = __shfl_down_sync(0xffffffffu, xr[i].v[0].x, 1);
      }
      const RowVec<CPT> first_old = xr[1];
      RowVec<CPT> prev = xr[0];
      RowVec<CPT> up, dn;
      // tall strips pull the neighbours' rows a few interior rows before they are needed -- the barrier has
      // normally completed by then and the shared-memory latency hides behind the remaining rows (+2 % at 8 rows
      // per thread); with 4 rows per thread the barrier has NOT completed that early (-9 %): those wait last
      constexpr int PULL_ROW = (R >= 8) ? R - FPIE_PATCH_PULL : R - 1;
      auto pull = [&]() {
        mbar_wait(bar, (mphase >> parity) & 1u);
        mphase ^= 1u << parity;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          up.v[q] = lds4(up_src + poff + q * 512);
          dn.v[q] = lds4(dn_src + poff + q * 512);
        }
      };
#pragma unroll
      for (int i = 1; i < R - 1; ++i) {
        if (i == PULL_ROW) pull();
        const RowVec<CPT> cur = xr[i];
        patch_row_update_lr<CPT, FRAME>(xr[i], hr[i], prev, xr[i + 1], FRAME ? fsel : sel[i], lf[i], rt[i]);
        prev = cur;
      }
      if (PULL_ROW >= R - 1) pull();
      // (at the patch's first / last row `up` / `dn` is the strip's own row: never used, those rows have no unknowns)
      if (FRAME) {
        // the patch's first / last row (a warp-uniform condition) is never updated; every other strip edge is an
        // ordinary row of the select-free stream
        if (!top_frame) patch_row_update_lr<CPT, true>(xr[0], hr[0], up, first_old, fsel, lf[0], rt[0]);
        if (!bottom_frame) patch_row_update_lr<CPT, true>(xr[R - 1], hr[R - 1], prev, dn, fsel, lf[R - 1], rt[R - 1]);
      } else {
        patch_row_update_lr<CPT, false>(xr[0], hr[0], up, first_old, sel[0], lf[0], rt[0]);
        patch_row_update_lr<CPT, false>(xr[R - 1], hr[R - 1], prev, dn, sel[R - 1], lf[R - 1], rt[R - 1]);
      }
      parity ^= 1;
    }

#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const uint32_t nib = (sel[i] >> (4 * q)) & 0xFu;
        if (nib) st4(x + base + (long long)i * g.pitch + 4 * q, xr[i].v[q]);
      }
    }
  }
  cluster_sync_all();  // nobody leaves while a neighbour may still store into its mailbox
}

}  // namespace fpie
