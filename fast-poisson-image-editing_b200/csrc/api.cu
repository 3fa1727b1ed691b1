// extern "C" surface of libfpie_b200.so -- see include/fpie_b200.h.
#include "fpie_b200.h"

#include <cstring>
#include <string>

#include "equ_solver.cuh"
#include "grid_solver.cuh"

struct fpie_b200_grid {
  fpie::GridSolver impl;
  fpie_b200_grid(int d, cudaStream_t s, int k, int v) : impl(d, s, k, v) {}
};
struct fpie_b200_equ {
  fpie::EquSolver impl;
  fpie_b200_equ(int d, cudaStream_t s, int z) : impl(d, s, z) {}
};

namespace {
thread_local std::string g_last_error;

template <typename F>
int guarded(F &&body) {
  try {
    body();
    return 0;
  } catch (const fpie::Error &e) {
    g_last_error = e.what();
    return 1;
  } catch (const std::bad_alloc &) {
    g_last_error = "out of host memory";
    return 2;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return 3;
  }
}

#define NEED(h)                                             \
  if (!(h)) {                                               \
    g_last_error = "fpie_b200: null solver handle";         \
    return 4;                                               \
  }
}  // namespace

#define API extern "C" __attribute__((visibility("default")))

API int fpie_b200_abi_version(void) { return FPIE_B200_ABI_VERSION; }
API const char *fpie_b200_last_error(void) { return g_last_error.c_str(); }

API int fpie_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

API int fpie_b200_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor) {
  return guarded([&] {
    cudaDeviceProp prop{};
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (name && name_len > 0) {
      strncpy(name, prop.name, (size_t)name_len - 1);
      name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
  });
}

API int fpie_b200_host_alloc(int64_t bytes, void **out) {
  if (!out || bytes < 0) {
    g_last_error = "fpie_b200_host_alloc: bad arguments";
    return 4;
  }
  *out = nullptr;
  return guarded([&] { CUDA_CHECK(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable)); });
}

API int fpie_b200_host_free(void *ptr) {
  return guarded([&] {
    if (ptr) CUDA_CHECK(cudaFreeHost(ptr));
  });
}

// ---- GridSolver -----------------------------------------------------------
API int fpie_b200_grid_create(int device, void *stream, int block_k, int variant, fpie_b200_grid **out) {
  if (!out) {
    g_last_error = "fpie_b200_grid_create: null output";
    return 4;
  }
  *out = nullptr;
  return guarded([&] { *out = new fpie_b200_grid(device, (cudaStream_t)stream, block_k, variant); });
}
API int fpie_b200_grid_destroy(fpie_b200_grid *g) {
  return guarded([&] { delete g; });
}
API int fpie_b200_grid_reset(fpie_b200_grid *g, int n, int m, const int32_t *mask, int64_t mask_row_stride,
                             int64_t mask_col_stride, const float *tgt, const float *grad) {
  NEED(g);
  return guarded([&] { g->impl.reset(n, m, mask, mask_row_stride, mask_col_stride, tgt, grad); });
}
API int fpie_b200_grid_step(fpie_b200_grid *g, int iters, uint8_t *out_img, float *out_err3) {
  NEED(g);
  return guarded([&] { g->impl.step(iters, out_img, out_err3); });
}
API int fpie_b200_grid_step_into(fpie_b200_grid *g, int iters, uint8_t *dst, int64_t dst_row_stride, float *out_err3) {
  NEED(g);
  return guarded([&] { g->impl.step(iters, dst, out_err3, dst_row_stride); });
}
API int fpie_b200_grid_set_edge_rows(fpie_b200_grid *g, int rows) {
  NEED(g);
  return guarded([&] { g->impl.set_edge_rows(rows); });
}
API int fpie_b200_grid_pass_async(fpie_b200_grid *g, int nsweeps, int part) {
  NEED(g);
  return guarded([&] { g->impl.pass_async(nsweeps, part); });
}
API int fpie_b200_grid_flip(fpie_b200_grid *g) {
  NEED(g);
  return guarded([&] { g->impl.flip(); });
}
API int fpie_b200_grid_patch_info(fpie_b200_grid *g, int *usable, int *rows_per_thread, int *cols_per_thread, int *cluster,
                                  int64_t *launches) {
  NEED(g);
  return guarded([&] {
    int r = 0, c = 0, cl = 0;
    const bool ok = g->impl.patch_usable(&r, &c, &cl);
    if (usable) *usable = ok ? 1 : 0;
    if (rows_per_thread) *rows_per_thread = r;
    if (cols_per_thread) *cols_per_thread = c;
    if (cluster) *cluster = cl;
    if (launches) *launches = g->impl.patch_launches() | ((int64_t)g->impl.patch_clusters() << 40);
  });
}
API int fpie_b200_grid_halo_config(fpie_b200_grid *g, int band_lo, int band_hi, int force_rebuild, int *changed) {
  NEED(g);
  return guarded([&] {
    const bool c = g->impl.halo_config(band_lo, band_hi, force_rebuild != 0);
    if (changed) *changed = c ? 1 : 0;
  });
}
API int fpie_b200_grid_halo_export(fpie_b200_grid *g, int side, unsigned char *blob) {
  NEED(g);
  return guarded([&] { g->impl.halo_export(side, blob); });
}
API int fpie_b200_grid_halo_connect(fpie_b200_grid *g, int side, const unsigned char *blob, int same_process) {
  NEED(g);
  return guarded([&] { g->impl.halo_connect(side, blob, same_process != 0); });
}
API int fpie_b200_grid_band_sweeps_async(fpie_b200_grid *g, int iters) {
  NEED(g);
  return guarded([&] { g->impl.band_sweeps_async(iters); });
}
API int fpie_b200_grid_halo_stats(fpie_b200_grid *g, int64_t *exchanges) {
  NEED(g);
  return guarded([&] {
    if (exchanges) *exchanges = g->impl.halo_exchanges();
  });
}
API int fpie_b200_grid_halo_debug(fpie_b200_grid *g, int64_t *out16) {
  NEED(g);
  return guarded([&] { g->impl.halo_debug(out16); });
}
API int fpie_b200_grid_halo_trace_begin(fpie_b200_grid *g, int max_intervals) {
  NEED(g);
  return guarded([&] { g->impl.halo_trace_begin(max_intervals); });
}
API int fpie_b200_grid_halo_trace_read(fpie_b200_grid *g, float *out, int max_floats, int *written) {
  NEED(g);
  return guarded([&] {
    const int n = g->impl.halo_trace_read(out, max_floats);
    if (written) *written = n;
  });
}
API int fpie_b200_grid_set_formulation(fpie_b200_grid *g, int equ) {
  NEED(g);
  return guarded([&] { g->impl.set_formulation(equ != 0); });
}
API int fpie_b200_grid_solve(fpie_b200_grid *g, int max_iters, int check_every, float tol, float *out_err3,
                             int *iters_done) {
  NEED(g);
  return guarded([&] {
    const int done = g->impl.solve(max_iters, check_every, tol, out_err3);
    if (iters_done) *iters_done = done;
  });
}
API int fpie_b200_grid_state(fpie_b200_grid *g, float *out_state) {
  NEED(g);
  return guarded([&] { g->impl.state(out_state); });
}
API int fpie_b200_grid_sweeps_async(fpie_b200_grid *g, int iters) {
  NEED(g);
  return guarded([&] { g->impl.sweeps_async(iters); });
}
API int fpie_b200_grid_finish_async(fpie_b200_grid *g) {
  NEED(g);
  return guarded([&] { g->impl.finish_async(); });
}
API int fpie_b200_grid_sync(fpie_b200_grid *g) {
  NEED(g);
  return guarded([&] { g->impl.sync(); });
}
API int fpie_b200_grid_fetch(fpie_b200_grid *g, uint8_t *out_img, float *out_err3) {
  NEED(g);
  return guarded([&] { g->impl.fetch(out_img, out_err3); });
}
API int fpie_b200_grid_fetch_rows(fpie_b200_grid *g, int row_lo, int row_hi, uint8_t *out_img, float *out_err3) {
  NEED(g);
  return guarded([&] { g->impl.fetch(out_img, out_err3, 0, row_lo, row_hi); });
}
API int fpie_b200_grid_info(fpie_b200_grid *g, int64_t *unknowns, int64_t *launches, int *block_k,
                            int64_t *active_tiles, int64_t *total_tiles) {
  NEED(g);
  return guarded([&] {
    const fpie::GridStats &s = g->impl.stats();
    if (unknowns) *unknowns = s.unknowns;
    if (launches) *launches = s.launches;
    if (block_k) *block_k = g->impl.block_k();
    if (active_tiles) *active_tiles = s.active_tiles;
    if (total_tiles) *total_tiles = s.total_tiles;
  });
}
API int fpie_b200_grid_config(fpie_b200_grid *g, int *variant, int *rows_per_thread, int *warps, int *ctas_per_sm) {
  NEED(g);
  return guarded([&] { g->impl.config(variant, rows_per_thread, warps, ctas_per_sm); });
}
API int fpie_b200_grid_reset_from_images(fpie_b200_grid *g, const uint8_t *src, int sh, int sw, const uint8_t *mask,
                                         int mh, int mw, int mc, const uint8_t *tgt, int th, int tw, int h0, int w0,
                                         int h1, int w1, int grad_mode, int64_t *out_n, int32_t *out_box4) {
  NEED(g);
  return guarded([&] {
    g->impl.reset_from_images(src, sh, sw, mask, mh, mw, mc, tgt, th, tw, h0, w0, h1, w1, grad_mode, out_n, out_box4);
  });
}
API int fpie_b200_grid_reset_batch(fpie_b200_grid *g, const uint8_t *src, const uint8_t *mask, const uint8_t *tgt,
                                   int batch, int rows, int cols, int mask_channels, int grad_mode) {
  NEED(g);
  return guarded([&] { g->impl.reset_batch(src, mask, tgt, batch, rows, cols, mask_channels, grad_mode); });
}
API int fpie_b200_grid_reset_slab(fpie_b200_grid *g, const uint8_t *src, const uint8_t *mask, const uint8_t *tgt,
                                  int rows, int cols, int mask_channels, int grad_mode) {
  NEED(g);
  return guarded([&] {
    g->impl.reset_from_images(src, rows, cols, mask, rows, cols, mask_channels, tgt, rows, cols, 0, 0, 0, 0, grad_mode,
                              nullptr, nullptr, /*crop=*/false);
  });
}
API int fpie_b200_grid_band_view(fpie_b200_grid *g, int which_buffer, float **dev_base, int64_t *plane_stride,
                                 int64_t *row_pitch, int *pad_rows, int *pad_cols) {
  NEED(g);
  return guarded([&] { g->impl.band_view(which_buffer, dev_base, plane_stride, row_pitch, pad_rows, pad_cols); });
}
API int fpie_b200_grid_band_current(fpie_b200_grid *g, int *which_buffer) {
  NEED(g);
  return guarded([&] {
    if (which_buffer) *which_buffer = g->impl.current();
  });
}
API int fpie_b200_grid_set_row_window(fpie_b200_grid *g, int row_lo, int row_hi) {
  NEED(g);
  return guarded([&] { g->impl.set_row_window(row_lo, row_hi); });
}

// ---- EquSolver ------------------------------------------------------------
API int fpie_b200_equ_create(int device, void *stream, int block_size, fpie_b200_equ **out) {
  if (!out) {
    g_last_error = "fpie_b200_equ_create: null output";
    return 4;
  }
  *out = nullptr;
  return guarded([&] { *out = new fpie_b200_equ(device, (cudaStream_t)stream, block_size); });
}
API int fpie_b200_equ_destroy(fpie_b200_equ *e) {
  return guarded([&] { delete e; });
}
API int fpie_b200_equ_set_mode(fpie_b200_equ *e, int mode) {
  NEED(e);
  return guarded([&] { e->impl.set_mode(mode); });
}
API int fpie_b200_equ_partition(fpie_b200_equ *e, int n, int m, const int32_t *mask, int64_t mask_row_stride,
                                int64_t mask_col_stride, int32_t *out_ids) {
  NEED(e);
  return guarded([&] { e->impl.partition(n, m, mask, mask_row_stride, mask_col_stride, out_ids); });
}
API int fpie_b200_equ_reset(fpie_b200_equ *e, int64_t N, const int32_t *A, const float *X, const float *B) {
  NEED(e);
  return guarded([&] { e->impl.reset(N, A, X, B); });
}
API int fpie_b200_equ_step(fpie_b200_equ *e, int iters, uint8_t *out_img, float *out_err3) {
  NEED(e);
  return guarded([&] { e->impl.step(iters, out_img, out_err3); });
}
API int fpie_b200_equ_state(fpie_b200_equ *e, float *out_state) {
  NEED(e);
  return guarded([&] { e->impl.state(out_state); });
}
API int fpie_b200_equ_solve(fpie_b200_equ *e, int max_iters, int check_every, float tol, float *out_err3,
                            int *iters_done) {
  NEED(e);
  return guarded([&] {
    const int done = e->impl.solve(max_iters, check_every, tol, out_err3);
    if (iters_done) *iters_done = done;
  });
}
API int fpie_b200_equ_sweeps_async(fpie_b200_equ *e, int iters) {
  NEED(e);
  return guarded([&] { e->impl.sweeps_async(iters); });
}
API int fpie_b200_equ_finish_async(fpie_b200_equ *e) {
  NEED(e);
  return guarded([&] { e->impl.finish_async(); });
}
API int fpie_b200_equ_sync(fpie_b200_equ *e) {
  NEED(e);
  return guarded([&] { e->impl.sync(); });
}
API int fpie_b200_equ_fetch(fpie_b200_equ *e, uint8_t *out_img, float *out_err3) {
  NEED(e);
  return guarded([&] { e->impl.fetch(out_img, out_err3); });
}
API int fpie_b200_equ_info(fpie_b200_equ *e, int64_t *unknowns, int64_t *launches, int *path) {
  NEED(e);
  return guarded([&] {
    if (unknowns) *unknowns = e->impl.stats().unknowns;
    if (launches) *launches = e->impl.launches();
    if (path) *path = e->impl.path();
  });
}
API int fpie_b200_grid_on_box(fpie_b200_grid *g, fpie_b200_box_fn cb, void *user) {
  NEED(g);
  return guarded([&] { g->impl.set_box_callback(cb, user); });
}
API int fpie_b200_equ_on_box(fpie_b200_equ *e, fpie_b200_box_fn cb, void *user) {
  NEED(e);
  return guarded([&] { e->impl.set_box_callback(cb, user); });
}
API int fpie_b200_equ_set_window(fpie_b200_equ *e, int64_t lo, int64_t hi) {
  NEED(e);
  return guarded([&] { e->impl.set_window(lo, hi); });
}
API int fpie_b200_equ_fetch_rows(fpie_b200_equ *e, int64_t lo, int64_t hi, uint8_t *out_img, float *out_err3) {
  NEED(e);
  return guarded([&] { e->impl.fetch_rows(lo, hi, out_img, out_err3); });
}
API int fpie_b200_equ_gather_rows(fpie_b200_equ *e, const int32_t *dev_idx, int64_t n, float *dev_out) {
  NEED(e);
  return guarded([&] { e->impl.gather_rows(dev_idx, n, dev_out); });
}
API int fpie_b200_equ_scatter_rows(fpie_b200_equ *e, const int32_t *dev_idx, int64_t n, const float *dev_in) {
  NEED(e);
  return guarded([&] { e->impl.scatter_rows(dev_idx, n, dev_in); });
}
API int fpie_b200_equ_rows_checked(fpie_b200_equ *e, int on) {
  NEED(e);
  return guarded([&] { e->impl.set_rows_checked(on != 0); });
}
API int fpie_b200_equ_reset_from_images(fpie_b200_equ *e, const uint8_t *src, int sh, int sw, const uint8_t *mask,
                                        int mh, int mw, int mc, const uint8_t *tgt, int th, int tw, int h0, int w0,
                                        int h1, int w1, int grad_mode, int64_t *out_n, int32_t *out_box4) {
  NEED(e);
  return guarded([&] {
    e->impl.reset_from_images(src, sh, sw, mask, mh, mw, mc, tgt, th, tw, h0, w0, h1, w1, grad_mode, out_n, out_box4);
  });
}
API int fpie_b200_equ_step_paste(fpie_b200_equ *e, int iters, uint8_t *out_crop, float *out_err3) {
  NEED(e);
  return guarded([&] { e->impl.step_paste(iters, out_crop, out_err3); });
}
API int fpie_b200_equ_step_paste_into(fpie_b200_equ *e, int iters, uint8_t *dst, int64_t dst_row_stride,
                                      float *out_err3) {
  NEED(e);
  return guarded([&] { e->impl.step_paste(iters, dst, out_err3, dst_row_stride); });
}
API int fpie_b200_equ_system(fpie_b200_equ *e, int32_t *out_A, float *out_X, float *out_B) {
  NEED(e);
  return guarded([&] { e->impl.system(out_A, out_X, out_B); });
}
