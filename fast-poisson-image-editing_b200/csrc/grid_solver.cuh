// GridSolver: masked 5-point Jacobi on a padded planar fp32 layout (sm_100a).
// Host-side class; the kernels live in grid_kernels.cu / grid_tiles.cu.
#pragma once

#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "prep.cuh"

namespace fpie {

// Register-tile geometry of the temporally blocked kernel: a CTA of NW warps
// owns a TILE_H x TILE_W pixel tile of one channel plane; a thread owns
// ROWS_PER_THREAD rows x 4 columns of it in registers.
constexpr int TILE_W = 128;  // 32 lanes x float4
constexpr int MAX_BLOCK_K = 16;
constexpr int PAD_ROWS = 16;  // >= MAX_BLOCK_K
constexpr int MASK_BOX_WORDS = 8;  // mask words staged per tile row (a 128-px row spans <= 5)
constexpr int PAD_COLS = 32;  // >= round_up(MAX_BLOCK_K, 4), multiple of 32

struct TileShape {
  int rows_per_thread;
  int warps;
  int cluster = 1;  // CTAs of one thread-block cluster stacked vertically on one tile (grid_cluster.cuh)
  int tile_h() const { return rows_per_thread * warps * cluster; }
  int box_h() const { return rows_per_thread * warps; }  // rows one CTA stages per TMA box
};

// An EquSolver system whose unknowns are the masked pixels of an n x m crop in row-major order,
// handed over on the device: mask[p] > 0 marks unknowns, ids[p] their row, X / B are channel planes.
struct EquEmbed {
  int n, m;
  const int32_t *mask;
  const int32_t *ids;
  float *X;        // [3][pitch]
  const float *B;  // [3][pitch]
  long long pitch;
};

struct GridStats {
  int64_t unknowns = 0;
  int64_t launches = 0;
  int64_t active_tiles = 0;
  int64_t total_tiles = 0;
};

class GridSolver {
 public:
  GridSolver(int device, cudaStream_t stream, int block_k, int variant);
  ~GridSolver();

  void reset(int n, int m, const int32_t *mask, int64_t mask_rs, int64_t mask_cs, const float *tgt,
             const float *grad);
  void reset_from_images(const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh, int mw, int mc,
                         const uint8_t *tgt, int th, int tw, int h0, int w0, int h1, int w1, int grad_mode,
                         int64_t *out_n, int32_t *out_box4, bool crop = true);
  void set_box_callback(void (*cb)(void *, const int32_t *), void *user) { upload_.set_box_callback(cb, user); }
  void reset_batch(const uint8_t *src, const uint8_t *mask, const uint8_t *tgt, int batch, int ph, int pw, int mc,
                   int grad_mode);
  // EquSolver promotion: state = X on masked pixels and 0 elsewhere, gradient = B (see equ.cu)
  void reset_from_equ(const EquEmbed &e);
  void export_to_equ(const EquEmbed &e);
  void sweeps_async(int iters);
  void finish_async();
  void sync();
  void fetch(uint8_t *out_img, float *out_err3, int64_t row_stride = 0, int row_lo = 0, int row_hi = -1);
  void step(int iters, uint8_t *out_img, float *out_err3, int64_t row_stride = 0);
  int solve(int max_iters, int check_every, float tol, float *out_err3);
  void state(float *out);
  void set_row_window(int lo, int hi);
  void set_edge_rows(int rows);
  void pass_async(int nsweeps, int part);
  void flip();
  // image-level resets build the EquSolver's system on the grid (boundary folded into B, zero outside the mask)
  void set_formulation(bool equ) { equ_form_ = equ; }
  void band_view(int which, float **base, int64_t *plane_stride, int64_t *row_pitch, int *pad_rows, int *pad_cols);

  // ---- row-band halo link: peer copies + stream memory operations behind the C ABI (halo.cu) ----
  static constexpr int kHaloBlobBytes = 128;
  // this slab holds grid rows [band_lo, band_hi) of its own plus the rows above / below them as halos
  bool halo_config(int band_lo, int band_hi, bool force = false);  // true when the link was (re)built: export / connect again
  // receive box of side (0 = up, 1 = down) as an opaque blob for the neighbour on that side
  void halo_export(int side, unsigned char *blob);
  // the neighbour's receive box for MY rows; same_process: the blob carries a raw pointer, not an IPC handle
  void halo_connect(int side, const unsigned char *blob, bool same_process);
  void halo_disconnect();
  // `iters` sweeps; halo rows refreshed every `halo` sweeps (halo = rows held of each neighbour)
  void band_sweeps_async(int iters);
  int64_t halo_exchanges() const { return halo_exchanges_; }
  void halo_debug(int64_t *out16);
  // phase trace of the first intervals of the next band_sweeps_async call (milliseconds since its start)
  void halo_trace_begin(int max_intervals);
  int halo_trace_read(float *out, int max_floats);

  int device() const { return device_; }
  int block_k() const { return block_k_; }
  int current() const { return cur_; }
  void config(int *variant, int *rows, int *warps, int *occ) const;
  const GridStats &stats() const { return stats_; }
  int64_t patch_launches() const { return patch_launches_; }
  int patch_clusters() const { return patch_clusters_; }
  bool patch_usable(int *r, int *c, int *cl) const { return ready_ && patch_shape(r, c, cl); }
  const PlaneGeom &geom() const { return geom_; }

 private:
  void require_ready() const;
  void layout(int n, int m);
  void build_tiles(const uint32_t *flags);
  void after_state_loaded();
  void make_tensor_maps();
  void configure(int variant, int block_k);
  void auto_configure(int n, int m);
  void drop_graphs();
  void build_from_upload();
  static TileShape shape_for(int variant);

  int device_;
  cudaStream_t stream_;
  int block_k_;
  int halo_x_;
  int variant_;
  bool equ_form_ = false;
  bool resid_equ_ = false;  // the loaded state is in the EquSolver's formulation: its residual expression applies
  bool auto_tune_ = false;  // tile shape chosen at reset
  bool auto_k_ = false;     // blocking depth chosen at reset
  int sm_count_ = 0;
  TileShape shape_{16, 12};

  bool ready_ = false;
  PlaneGeom geom_{};
  PlaneGeom zeroed_{};             // geometry whose padding is known to be zero in the buffers below
  const float *zeroed_ptr_[4] = {nullptr, nullptr, nullptr, nullptr};
  bool zeroed_batch_ = false;
  int cur_ = 0;
  int win_lo_ = 0, win_hi_ = 0;
  DeviceBuffer<float> x_[2];
  DeviceBuffer<float> hq_;
  DeviceBuffer<__half> hq16_;
  DeviceBuffer<int> flag_;
  bool h16_ok_ = false;
  bool force_h32_ = false;
  DeviceBuffer<uint32_t> bits_;
  DeviceBuffer<float> stage_;
  BlendUpload upload_;
  UploadChunks chunks_;  // row chunks of the last crop-mode upload (consumed by build_from_upload)
  BatchMap batch_{0, 0, 0, 0};
  DeviceBuffer<double> batch_err_;
  DeviceBuffer<int32_t> mask_stage_;
  DeviceBuffer<uint8_t> img_;
  DeviceBuffer<double> err_;  // [3] residual sums + [1] unknown count (as double)
  DeviceBuffer<int2> tiles_;
  DeviceBuffer<uint32_t> tile_flags_;
  // CUDA graph of kGraphPasses full passes per starting buffer, replayed by long sweeps_async calls
  static constexpr int kGraphPasses = 16;
  cudaStream_t cap_stream_ = nullptr;
  cudaGraphExec_t graph_[2] = {nullptr, nullptr};
  bool graph_off_ = false, graph_warm_ = false;
  bool patch_off_ = false;      // persistent small-image kernel disabled
  int patch_rows_ = 0;          // rows per thread override (0 = automatic)
  int patch_min_iters_ = 32;    // shorter runs stay on the tiled kernel (the persistent launch reads and writes the planes once)
  int64_t patch_launches_ = 0;
  int patch_clusters_ = 0;      // clusters of the last persistent launch
  bool serpentine_ = true;  // alternate passes walk the tile list in opposite directions
  // edge / interior partition of the tile list (set_edge_rows)
  std::vector<int2> host_tiles_;
  std::vector<int> host_tile_row_;
  DeviceBuffer<int2> tiles_part_[2];
  int n_part_[2] = {0, 0};
  int edge_rows_ = 0;
  CUtensorMap tm_x_[2];
  CUtensorMap tm_h_;
  CUtensorMap tm_h16_;
  CUtensorMap tm_m_;
  int n_tile_entries_ = 0;
  double *host_err_ = nullptr;  // pinned [4]
  GridStats stats_;

  struct HaloSide {
    int rows = 0;                       // halo rows held on this side (0 = no neighbour)
    int row_send = 0, row_recv = 0;     // first grid row sent to / refreshed from the neighbour
    unsigned char *inbox = nullptr;     // device: [2 parities][3 planes][rows][m] floats, then 2 flag words
    size_t inbox_bytes = 0, parity_bytes = 0;
    unsigned char *remote = nullptr;    // the neighbour's inbox for MY rows, mapped into this process
    bool remote_ipc = false;
    uint32_t sent = 0, received = 0;    // exchange intervals so far (the flag words carry these counters)
  };
  void halo_free_side(HaloSide &s);
  void halo_send(int which);
  void halo_recv(int which);
  void run_pass(int nsweeps, const int2 *tiles, int ntiles);
  void preload_kernels();
  bool patch_shape(int *rows_per_thread, int *cols_per_thread, int *cluster) const;
  void patch_sweeps(int iters);
  HaloSide halo_[2];
  int band_lo_ = 0, band_hi_ = 0;
  bool halo_pending_ = false;
  int64_t halo_exchanges_ = 0;
  cudaStream_t halo_stream_ = nullptr;
  cudaEvent_t ev_edge_ = nullptr, ev_sent_ = nullptr;
  uint32_t *halo_seq_ = nullptr;  // device: [2] flag values in flight (source of the 4-byte flag copies)
  // trace
  std::vector<cudaEvent_t> trace_ev_;
  std::vector<int> trace_tag_;
  int trace_left_ = 0;
  void trace_mark(int tag, cudaStream_t s);
};

}  // namespace fpie
