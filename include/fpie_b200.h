/*
 * fpie_b200 -- C ABI of the B200-native Jacobi Poisson backend.
 *
 * This is the drop-in boundary for the one hot path of
 * Trinkle23897/Fast-Poisson-Image-Editing (fpie 0.3.2): the core-solver
 * interface `partition / reset / sync / step` that fpie's Processor layer
 * drives (fpie/process.py:138,187,270,275,385,390) and that the reference
 * binds with pybind11 per backend (fpie/core/cuda/solver.cc:3-15,
 * fpie/core/openmp/solver.cc:3-15; state and numpy->C copies in
 * fpie/core/base_solver.h:12-75 and :77-152).
 *
 * Conventions
 *   - plain C types only: pointers + sizes, no numpy/torch/pybind types;
 *   - every entry point returns 0 on success and a non-zero code on failure;
 *     `fpie_b200_last_error()` then returns a thread-local message (the
 *     reference checks no CUDA call at all: fpie/core/cuda/equ.cu:56-74);
 *   - "host" pointers may be pageable or pinned; "dev" pointers are CUDA
 *     device pointers on the solver's device;
 *   - all device work of a solver is enqueued on the `stream` given at
 *     creation (a `cudaStream_t` passed as `void*`; NULL = the legacy default
 *     stream), so a caller can bracket it with its own events;
 *   - image layouts at the boundary are the reference's: interleaved
 *     `[rows, cols, 3]`, float32 state, int32 mask / ids, uint8 images.
 *
 * There is no CPU fallback: without a CUDA device `*_create` fails.
 */
#ifndef FPIE_B200_H_
#define FPIE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPIE_B200_ABI_VERSION 1

/* gradient mixing modes of BaseProcessor.mixgrad (fpie/process.py:113-122) */
#define FPIE_B200_GRAD_SRC 0
#define FPIE_B200_GRAD_AVG 1
#define FPIE_B200_GRAD_MAX 2

typedef struct fpie_b200_grid fpie_b200_grid;
typedef struct fpie_b200_equ fpie_b200_equ;

/* ---- library ----------------------------------------------------------- */

int fpie_b200_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char *fpie_b200_last_error(void);
/* Number of CUDA devices (0 when there is no usable driver/GPU). */
int fpie_b200_device_count(void);
/* Name / SM count / compute capability of a device, for logs
 * (the reference prints a device table from every constructor:
 * fpie/core/cuda/utils.cu:5-22; here it is a query, not a side effect). */
int fpie_b200_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor);

/* Page-locked host memory for the images that cross the boundary: a `step` that lands its uint8
 * result in pinned memory runs at PCIe speed, a pageable destination costs several times that
 * (the reference returns a fresh pageable numpy array from every step, fpie/core/cuda/grid.cu:152).
 * The Processor layer keeps its full-size target canvas (fpie/process.py:268, 384) in such a buffer. */
int fpie_b200_host_alloc(int64_t bytes, void **out);
int fpie_b200_host_free(void *ptr);

/* ---- GridSolver ---------------------------------------------------------
 * Replaces CudaGridSolver (fpie/core/cuda/grid.cu:7-153) behind the
 * GridSolver interface of fpie/core/base_solver.h:77-152. */

/* GridSolver(grid_x, grid_y) constructor (fpie/process.py:312-313).
 * `block_k` = Jacobi sweeps fused per pass over HBM (temporal blocking depth,
 * 1..16; 0 = chosen from the grid size at reset); `variant` selects a kernel
 * family / register-tile shape for testing and tuning (0 = chosen from the
 * grid size at reset, 1 = one-sweep-per-launch kernels only; the table is in
 * csrc/grid.cu, variant_info). */
int fpie_b200_grid_create(int device, void *stream, int block_k, int variant, fpie_b200_grid **out);
int fpie_b200_grid_destroy(fpie_b200_grid *g);

/* GridSolver::reset(n, mask, tgt, grad) (base_solver.h:101-140 + grid.cu:33-52).
 * mask: int32 [n, m] with element strides (the Processor passes a
 * non-contiguous view, fpie/process.py:351); tgt, grad: float32 [n, m, 3]
 * C-contiguous.  Inputs are copied; mask pixels on the outer frame of the
 * grid are treated as unmasked (the Processor guarantees a zero frame,
 * process.py:342-351; the reference would read out of bounds otherwise). */
int fpie_b200_grid_reset(fpie_b200_grid *g, int n, int m, const int32_t *mask, int64_t mask_row_stride,
                         int64_t mask_col_stride, const float *tgt, const float *grad);

/* GridSolver::step(iteration) -> (img u8 [n, m, 3], err f32 [3])
 * (grid.cu:133-153).  Runs exactly `iters` more true-Jacobi sweeps on the
 * persistent state (np_solver.py:81-88 semantics), then the residual
 * (np_solver.py:90-96) and the clamp+truncate to uint8 (grid.cu:115-131).
 * Blocks until the results are in the host buffers. */
int fpie_b200_grid_step(fpie_b200_grid *g, int iters, uint8_t *out_img, float *out_err3);

/* step() writing the uint8 crop straight into a larger host image: row r of
 * the result goes to dst + r * dst_row_stride (bytes).  This is the Processor's
 * paste `tgt[x0:x1, y0:y1] = img` (fpie/process.py:393) done by the copy engine. */
int fpie_b200_grid_step_into(fpie_b200_grid *g, int iters, uint8_t *dst, int64_t dst_row_stride, float *out_err3);

/* Convergence-driven stepping (the reference only offers a fixed
 * iteration count, fpie/cli.py:46-61 prints err every -p sweeps and leaves the decision to the user):
 * run sweeps in chunks of `check_every` until every channel of err (the quantity step() returns) is
 * <= tol, or `max_iters` sweeps have run.  Never the default; the state stays on the device and a
 * following step(0) returns the image.  *iters_done = sweeps actually run. */
int fpie_b200_grid_solve(fpie_b200_grid *g, int max_iters, int check_every, float tol, float *out_err3,
                         int *iters_done);

/* The fp32 state [n, m, 3] (the reference never exposes it from native
 * cores; needed for the fp32 parity check). */
int fpie_b200_grid_state(fpie_b200_grid *g, float *out_state);

/* Device-resident pieces of step(), for measurement and for callers that
 * keep results on the device: enqueue `iters` sweeps (no sync, no copy);
 * enqueue residual + u8 conversion; wait for the stream; copy results. */
int fpie_b200_grid_sweeps_async(fpie_b200_grid *g, int iters);
int fpie_b200_grid_finish_async(fpie_b200_grid *g);
int fpie_b200_grid_sync(fpie_b200_grid *g);
int fpie_b200_grid_fetch(fpie_b200_grid *g, uint8_t *out_img, float *out_err3);
/* fetch() of rows [row_lo, row_hi) only -> out_img [row_hi - row_lo, m, 3]: a row band downloads its own
 * rows, not the halo rows it holds of its neighbours (the reference's MPI GridSolver gathers exactly the
 * band rows on the root, fpie/core/mpi/grid.cc:137-146). */
int fpie_b200_grid_fetch_rows(fpie_b200_grid *g, int row_lo, int row_hi, uint8_t *out_img, float *out_err3);

/* Number of masked pixels (unknowns) and kernel launches issued so far. */
int fpie_b200_grid_info(fpie_b200_grid *g, int64_t *unknowns, int64_t *launches, int *block_k,
                        int64_t *active_tiles, int64_t *total_tiles);

/* The kernel configuration in use (chosen at reset when `variant` / `block_k` were 0): the variant
 * number, the register tile (rows per thread x warps per CTA, 128 columns) and CTAs per SM. */
int fpie_b200_grid_config(fpie_b200_grid *g, int *variant, int *rows_per_thread, int *warps, int *ctas_per_sm);

/* Fused Processor-level reset (GridProcessor.reset, fpie/process.py:321-386)
 * executed on the device from uint8 images: mask threshold / frame clear /
 * bounding box, crop, mixed gradient, state upload.
 * src [sh, sw, 3], tgt [th, tw, 3], mask [mh, mw, mc] with any mc in 1..16 (thresholded on the channel mean);
 * (h0, w0) / (h1, w1) = position of mask pixel (0,0) in src / tgt.
 * out_box = {x0, x1, y0, y1} of the solved crop in target coordinates;
 * returns the crop size n*m in *out_n ("# of vars" of the reference). */
int fpie_b200_grid_reset_from_images(fpie_b200_grid *g, const uint8_t *src, int sh, int sw, const uint8_t *mask,
                                     int mh, int mw, int mc, const uint8_t *tgt, int th, int tw, int h0, int w0,
                                     int h1, int w1, int grad_mode, int64_t *out_n, int32_t *out_box4);

/* Batched small edits: `batch` independent blends of identical size, src / tgt
 * uint8 [batch, rows, cols, 3] and mask uint8 [batch, rows, cols, mask_channels]
 * (cols a multiple of 4).  Each patch is its own grid (no bounding-box crop, its
 * 1-pixel frame is fixed); the patches are laid out as one mosaic on the device
 * and swept together.  After this reset, fpie_b200_grid_step / _fetch return
 * out_img as uint8 [batch, rows, cols, 3] and out_err3 as float [batch, 3]. */
int fpie_b200_grid_reset_batch(fpie_b200_grid *g, const uint8_t *src, const uint8_t *mask, const uint8_t *tgt, int batch,
                               int rows, int cols, int mask_channels, int grad_mode);

/* The persistent small-image kernel (csrc/patch.cuh): when every patch of a batch is at most 512 rows x 256
 * columns, the batch holds at least 4 patches (12 channel planes to run side by side; FPIE_B200_PATCH=2 lifts that
 * limit, also for single images, FPIE_B200_PATCH=0 disables the kernel) and neither `variant` nor `block_k` was
 * fixed at creation, a step of >= 32 sweeps is ONE launch in which a thread-block cluster keeps a (patch, channel
 * plane) in registers for all sweeps.
 * *usable = 1 when the current problem qualifies; rows / cols per thread, CTAs per cluster; *launches = persistent
 * launches so far in the low 40 bits, clusters of the last launch above them. */
int fpie_b200_grid_patch_info(fpie_b200_grid *g, int *usable, int *rows_per_thread, int *cols_per_thread, int *cluster,
                              int64_t *launches);

/* ---- row-band sharding (multi-GPU GridSolver) ----------------------------
 * One solver per GPU holds one slab = a band of rows of the global grid plus
 * `halo` rows of its neighbours on each side (analogue of the reference's MPI
 * row bands, fpie/core/mpi/grid.cc:21-32, 108-147, but with deep halos so the
 * result equals single-device Jacobi bit for bit).  The caller (fpie_b200/band.py,
 * one process per GPU) runs at most `halo` sweeps, then overwrites the halo
 * rows of the CURRENT state buffer with the neighbour's band-edge rows
 * (NCCL send/recv or peer copies on the device pointers below), and repeats. */

/* Load one slab: src, mask, tgt are uint8 images of identical size rows x cols
 * (mask with 1 or 3 channels), already cut to the slab's rows and to the columns
 * of the global crop.  The whole slab is the grid (no bounding-box crop); its
 * outer frame is fixed (global frame rows, or halo rows the exchange refreshes). */
int fpie_b200_grid_reset_slab(fpie_b200_grid *g, const uint8_t *src, const uint8_t *mask, const uint8_t *tgt, int rows,
                              int cols, int mask_channels, int grad_mode);
/* Device view of state buffer `which_buffer` (0/1): grid pixel (r, c) of channel
 * p is the float at dev_base[p * plane_stride + (r + pad_rows) * row_pitch + c + pad_cols]. */
int fpie_b200_grid_band_view(fpie_b200_grid *g, int which_buffer, float **dev_base, int64_t *plane_stride,
                             int64_t *row_pitch, int *pad_rows, int *pad_cols);
/* Which of the two buffers holds the current state (it flips every pass). */
int fpie_b200_grid_band_current(fpie_b200_grid *g, int *which_buffer);
/* Restrict the residual of the following step / finish calls to grid rows
 * [row_lo, row_hi) -- a band counts only its own rows, not its halo. */
int fpie_b200_grid_set_row_window(fpie_b200_grid *g, int row_lo, int row_hi);

/* Split passes, so that the halo exchange of a band overlaps the bulk of a pass: after
 * set_edge_rows(rows) the tiles whose stored rows intersect the first / last `rows` grid rows are the
 * EDGE part (0) of the tile list, all others the INTERIOR part (1).  pass_async runs one pass
 * (1..block_k sweeps) over one part on the solver's stream WITHOUT flipping the state buffers; after
 * both parts of a pass have been enqueued (in either order) flip() makes its output the current state.
 * The reference's MPI solver has no such overlap: it exchanges between sweeps, blocking
 * (fpie/core/mpi/grid.cc:118-135). */
#define FPIE_B200_PART_EDGE 0
#define FPIE_B200_PART_INTERIOR 1
int fpie_b200_grid_set_edge_rows(fpie_b200_grid *g, int rows);
int fpie_b200_grid_pass_async(fpie_b200_grid *g, int nsweeps, int part);
int fpie_b200_grid_flip(fpie_b200_grid *g);

/* The halo exchange as a data plane behind this ABI (csrc/halo.cu): copy-engine peer copies over NVLink and
 * stream memory operations instead of a communication library's SM kernels.  Call sequence per slab, after
 * every reset:
 *   halo_config(band_lo, band_hi)     the slab's rows [band_lo, band_hi) are its own band, the rows above /
 *                                     below are halo rows (0 rows on a side = no neighbour there); also sets the
 *                                     residual row window and the edge / interior split of the tile list.
 *                                     Returns *changed = 1 when the link had to be rebuilt (first call, new
 *                                     geometry, or force_rebuild): only then are export / connect needed again --
 *                                     on BOTH ends of a link, so a caller whose neighbour reports a rebuild calls
 *                                     again with force_rebuild = 1.
 *   halo_export(side, blob[128])      an opaque description of this slab's receive box for that side
 *                                     (side 0 = up, 1 = down), to be handed to the neighbour on that side
 *   halo_connect(side, blob, same_process)   the NEIGHBOUR's blob for the box that takes this slab's rows;
 *                                     between processes the box is mapped with cudaIpcOpenMemHandle
 *   band_sweeps_async(iters)          `iters` sweeps with a halo exchange every `halo` sweeps, overlapped with
 *                                     the interior tiles of the passes around it; every band of the problem must
 *                                     make the same call (the neighbours wait for each other's rows)
 * Replaces the blocking one-row MPI_Sendrecv of fpie/core/mpi/grid.cc:118-135. */
#define FPIE_B200_HALO_BLOB_BYTES 128
int fpie_b200_grid_halo_config(fpie_b200_grid *g, int band_lo, int band_hi, int force_rebuild, int *changed);
int fpie_b200_grid_halo_export(fpie_b200_grid *g, int side, unsigned char *blob);
int fpie_b200_grid_halo_connect(fpie_b200_grid *g, int side, const unsigned char *blob, int same_process);
int fpie_b200_grid_band_sweeps_async(fpie_b200_grid *g, int iters);
/* Exchanges started so far; and a phase trace of the first `max_intervals` exchange intervals of the next
 * band_sweeps_async call: pairs (tag, milliseconds since the first mark) -- tags 1-6 = solver stream before /
 * after the edge and interior parts of a pass, 10/11 = halo stream around the peer copies, 20/21 = solver
 * stream around the wait for the neighbours' rows.  trace_read synchronises; returns the floats written. */
int fpie_b200_grid_halo_stats(fpie_b200_grid *g, int64_t *exchanges);
/* Link counters for diagnostics, readable while the solver's streams are blocked on a neighbour: out16 =
 * {rows[2], sent[2], received[2], flag words [side][parity] (4), current buffer, block_k, variant, edge tile
 * entries, interior tile entries, exchange pending} (-1 = not available). */
int fpie_b200_grid_halo_debug(fpie_b200_grid *g, int64_t *out16);
int fpie_b200_grid_halo_trace_begin(fpie_b200_grid *g, int max_intervals);
int fpie_b200_grid_halo_trace_read(fpie_b200_grid *g, float *out, int max_floats, int *written);

/* Formulation built by the image-level resets (reset_from_images / reset_slab) that follow:
 * 0 (default) = GridSolver's (unmasked pixels hold the target, fpie/process.py:354-378);
 * 1 = EquSolver's, laid out on the grid: unknowns carry X = target and B = grad + the targets of
 * their neighbours outside the mask (fpie/process.py:227-266), every other pixel the constant 0 --
 * the arithmetic of np_solver.py:33-41 on row-major ids, i.e. what EquSolver computes, in a form
 * that shards by row bands. */
int fpie_b200_grid_set_formulation(fpie_b200_grid *g, int equ);

/* ---- EquSolver ----------------------------------------------------------
 * Replaces CudaEquSolver (fpie/core/cuda/equ.cu:7-211) behind the EquSolver
 * interface of fpie/core/base_solver.h:12-75. */

/* EquSolver(block_size) constructor (fpie/process.py:173-174). */
int fpie_b200_equ_create(int device, void *stream, int block_size, fpie_b200_equ **out);
int fpie_b200_equ_destroy(fpie_b200_equ *e);

/* Iteration scheme: FPIE_B200_EQU_JACOBI (default; true Jacobi, fpie/np_solver.py:33-41) or
 * FPIE_B200_EQU_REDBLACK -- the reference OpenMP backend's deterministic red-black Gauss-Seidel:
 * partition labels odd pixels first, then even ones (fpie/core/openmp/equ.cc:22-56) and a sweep is
 * two in-place half-sweeps (equ.cc:107-118).  Set before partition / reset; red-black needs the
 * ids of this solver's own partition (or of reset_from_images). */
#define FPIE_B200_EQU_JACOBI 0
#define FPIE_B200_EQU_REDBLACK 1
/* Jacobi through the index-mapped gather kernels only.  In FPIE_B200_EQU_JACOBI the solver may
 * instead run the temporally blocked grid kernel when it can prove that the system is the 4-neighbour
 * structure of a mask it labelled itself (partition / reset_from_images): same bits, several times
 * faster; FPIE_B200_EQU_GATHER switches that promotion off. */
#define FPIE_B200_EQU_GATHER 2
int fpie_b200_equ_set_mode(fpie_b200_equ *e, int mode);

/* EquSolver::partition(mask) -> ids (equ.cu:36-54; np_solver.py:14-16):
 * row-major inclusive count of mask > 0, computed by a device prefix scan.
 * mask int32 [n, m] with element strides; ids int32 [n, m] C-contiguous.
 * ids on unmasked pixels hold the running count (the caller zeroes them,
 * process.py:188). */
int fpie_b200_equ_partition(fpie_b200_equ *e, int n, int m, const int32_t *mask, int64_t mask_row_stride,
                            int64_t mask_col_stride, int32_t *out_ids);

/* EquSolver::reset(N, A, X, B) (base_solver.h:27-59 + equ.cu:56-74).
 * A int32 [N, 4] (up, down, left, right; 0 = constant-zero row), X, B float32
 * [N, 3], C-contiguous; row 0 is the zero constant.  Inputs are copied.
 * Entries of A outside [0, N) are rejected. */
int fpie_b200_equ_reset(fpie_b200_equ *e, int64_t N, const int32_t *A, const float *X, const float *B);

/* EquSolver::step(iteration) -> (img u8 [N, 3], err f32 [3]) (equ.cu:189-211),
 * true Jacobi (np_solver.py:33-50 semantics). */
int fpie_b200_equ_step(fpie_b200_equ *e, int iters, uint8_t *out_img, float *out_err3);
int fpie_b200_equ_state(fpie_b200_equ *e, float *out_state);
/* Convergence-driven stepping, as fpie_b200_grid_solve (red-black mode: one sweep = both half sweeps). */
int fpie_b200_equ_solve(fpie_b200_equ *e, int max_iters, int check_every, float tol, float *out_err3, int *iters_done);

int fpie_b200_equ_sweeps_async(fpie_b200_equ *e, int iters);
int fpie_b200_equ_finish_async(fpie_b200_equ *e);
int fpie_b200_equ_sync(fpie_b200_equ *e);
int fpie_b200_equ_fetch(fpie_b200_equ *e, uint8_t *out_img, float *out_err3);
/* path & 7: 0 = generic int4 gather, 1 = compact-table gather (bit 3: the 4-byte distance table, bit 4: fp16 B stream --
 * 34 instead of 52 bytes per unknown and sweep, same bits), 2 = promoted to the tiled grid kernel, 3 = red-black */
int fpie_b200_equ_info(fpie_b200_equ *e, int64_t *unknowns, int64_t *launches, int *path);

/* Fused Processor-level reset (EquProcessor.reset, fpie/process.py:192-271)
 * on the device: mask canonicalisation, partition scan, index compaction and
 * the A / X / B build for the three gradient modes.  Returns N = K + 1 in
 * *out_n and the crop box in target coordinates.  The scatter list
 * (process.py:269) stays on the device: see fpie_b200_equ_step_paste. */
int fpie_b200_equ_reset_from_images(fpie_b200_equ *e, const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh,
                                    int mw, int mc, const uint8_t *tgt, int th, int tw, int h0, int w0, int h1,
                                    int w1, int grad_mode, int64_t *out_n, int32_t *out_box4);

/* step() followed by the Processor's scatter of the K solved pixels into the
 * crop box (process.py:273-280): out_crop is uint8 [x1-x0, y1-y0, 3], holding
 * the target's pixels outside the mask.  Only valid after
 * fpie_b200_equ_reset_from_images. */
int fpie_b200_equ_step_paste(fpie_b200_equ *e, int iters, uint8_t *out_crop, float *out_err3);
/* Same, writing the crop straight into a larger host image (row r at dst + r * dst_row_stride bytes):
 * the Processor's `tgt[tgt_index] = x[1:]` (process.py:278) without a host-side pass. */
int fpie_b200_equ_step_paste_into(fpie_b200_equ *e, int iters, uint8_t *dst, int64_t dst_row_stride, float *out_err3);
/* Read back the system built by reset_from_images (parity checks). */
int fpie_b200_equ_system(fpie_b200_equ *e, int32_t *out_A, float *out_X, float *out_B);

/* Early notice of the blend's bounding box.  `*_reset_from_images` uploads the mask first; as soon as its bounding box
 * is known -- before the source / target rows travel and the system is built -- `cb(user, box4)` is called on the
 * calling thread with (x0, x1, y0, y1) in TARGET coordinates (what out_box4 returns at the end).  The Processor uses it
 * to start its private copy of the target (fpie/process.py:268, 384) for what lies outside the box while the device
 * works.  cb == NULL removes it. */
typedef void (*fpie_b200_box_fn)(void *user, const int32_t *box4);
int fpie_b200_grid_on_box(fpie_b200_grid *g, fpie_b200_box_fn cb, void *user);
int fpie_b200_equ_on_box(fpie_b200_equ *e, fpie_b200_box_fn cb, void *user);

/* Id-range sharding of a general system across devices (fpie_b200/shard.py; the reference's analogue is the MPI
 * EquSolver, fpie/core/mpi/equ.cc:50-59 offsets and 123-146 the exchange -- which moves ALL of X through rank 0 every
 * `min_interval` sweeps; here a rank holds its id range plus `depth` layers of ghost unknowns and only those move).
 *   set_window  the residual of finish / step sums rows [lo, hi) only (the rows this rank owns);
 *   fetch_rows  the uint8 rows [lo, hi) of the last finish (+ err);
 *   gather_rows / scatter_rows  rows of X by index, packed [n, 3] fp32, from / into DEVICE buffers (idx: int32 on the
 *               device), enqueued on the solver's stream.  Indices are validated on the device; the first call after a
 *               reset reports a bad index synchronously, rows_checked(1) turns the per-call read-back off. */
int fpie_b200_equ_set_window(fpie_b200_equ *e, int64_t lo, int64_t hi);
int fpie_b200_equ_fetch_rows(fpie_b200_equ *e, int64_t lo, int64_t hi, uint8_t *out_img, float *out_err3);
int fpie_b200_equ_gather_rows(fpie_b200_equ *e, const int32_t *dev_idx, int64_t n, float *dev_out);
int fpie_b200_equ_scatter_rows(fpie_b200_equ *e, const int32_t *dev_idx, int64_t n, const float *dev_in);
int fpie_b200_equ_rows_checked(fpie_b200_equ *e, int on);

#ifdef __cplusplus
}
#endif
#endif /* FPIE_B200_H_ */
