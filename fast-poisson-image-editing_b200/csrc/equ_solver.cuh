// EquSolver: index-mapped (gather) Jacobi on a compacted list of unknowns.
#pragma once

#include <cuda_fp16.h>

#include <memory>

#include "common.cuh"
#include "grid_solver.cuh"

namespace fpie {

struct EquStats {
  int64_t unknowns = 0;
  int64_t launches = 0;
};

class EquSolver {
 public:
  EquSolver(int device, cudaStream_t stream, int block_size);
  ~EquSolver();

  void set_mode(int mode);
  void partition(int n, int m, const int32_t *mask, int64_t mask_rs, int64_t mask_cs, int32_t *out_ids);
  void reset(int64_t N, const int32_t *A, const float *X, const float *B);
  void reset_from_images(const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh, int mw, int mc,
                         const uint8_t *tgt, int th, int tw, int h0, int w0, int h1, int w1, int grad_mode,
                         int64_t *out_n, int32_t *out_box4);
  void set_box_callback(void (*cb)(void *, const int32_t *), void *user) { upload_.set_box_callback(cb, user); }
  void sweeps_async(int iters);
  void finish_async();
  void sync();
  void fetch(uint8_t *out_img, float *out_err3);
  void step(int iters, uint8_t *out_img, float *out_err3);
  int solve(int max_iters, int check_every, float tol, float *out_err3);
  void step_paste(int iters, uint8_t *out_crop, float *out_err3, int64_t row_stride = 0);
  void state(float *out);
  void system(int32_t *out_A, float *out_X, float *out_B);
  // id-range sharding (fpie_b200/shard.py): residual over rows [lo, hi) only; uint8 rows [lo, hi); rows of X by
  // index, packed [n, 3], to / from DEVICE buffers on the solver's stream
  void set_window(int64_t lo, int64_t hi);
  void fetch_rows(int64_t lo, int64_t hi, uint8_t *out_img, float *out_err3);
  void gather_rows(const int32_t *dev_idx, int64_t n, float *dev_out);
  void scatter_rows(const int32_t *dev_idx, int64_t n, const float *dev_in);
  void set_rows_checked(bool on) { rows_checked_ = on; }

  const EquStats &stats() const { return stats_; }
  int64_t launches() const { return stats_.launches + (tiled_ ? tiled_->stats().launches : 0); }
  bool structured() const { return structured_; }
  // 0 = generic int4 gather, 1 = compact-table gather (bit 3 set: 4-byte distance table; bit 4: fp16 B stream),
  // 2 = promoted to the tiled grid kernel, 3 = red-black
  int path() const {
    if (mode_ == 1) return 3;
    if (promoted_) return 2;
    if (!structured_) return 0;
    return 1 | (delta16_ ? 8 : 0) | (b16_ok_ ? 16 : 0);
  }

 private:
  void require_ready() const;
  void allocate(int64_t N);
  // device-side inclusive scan of (mask > 0) over a contiguous device mask
  void scan_ids(const int32_t *dev_mask, int64_t count, int32_t *dev_ids);
  void label(const int32_t *dev_mask, int n, int m, int32_t *dev_ids);
  void compact_tables();
  void try_promote(int n, int m);
  void pull_tiled_state();
  void drop_graphs();
  void check_rows_flag(const char *who);

  int device_;
  cudaStream_t stream_;
  int block_;
  bool ready_ = false;
  int mode_ = 0;        // 0 = Jacobi, 1 = red-black Gauss-Seidel
  int64_t n_mid_ = 0;   // first even id (red-black mode)
  DeviceBuffer<int32_t> rb_tmp_;
  int64_t win_lo_ = 0, win_hi_ = 0;  // rows the residual sums over (the whole system unless set_window)
  bool rows_checked_ = false;        // gather / scatter index lists already validated (skip the per-call sync)
  int64_t N_ = 0;      // rows including the constant row 0
  int64_t pitch_ = 0;  // floats per channel plane of X / B
  int cur_ = 0;
  DeviceBuffer<int4> A_;
  DeviceBuffer<int2> ud_;      // compact table (up, down, left/right presence bits)
  DeviceBuffer<uint32_t> d16_; // 4-byte table: 15-bit distances to up / down + presence bits
  DeviceBuffer<__half> b16_;   // fp16 copy of B (streamed when every value is exactly representable)
  bool delta16_ = false, b16_ok_ = false, no_delta16_ = false;
  bool d16_pipe_ = true;  // persistent, software-pipelined form of the 4-byte-table kernel
  int d16_ctas_per_sm_ = 4;
  int sm_count_ = 148;
  long long delta16_min_ = 1ll << 21;  // unknowns from which the 4-byte table is used (working set beyond the L2)
  bool structured_ = false;    // left/right are always i-1 / i+1 or absent
  bool force_generic_ = false;
  // promotion to the temporally blocked grid kernel (row-major ids on a known crop)
  std::unique_ptr<GridSolver> tiled_;
  bool promoted_ = false, tiled_dirty_ = false, no_promote_ = false;
  int part_n_ = 0, part_m_ = 0;  // geometry of the last partition() / fused reset (0 = unknown)
  DeviceBuffer<float> X_[2];
  DeviceBuffer<float> B_;
  DeviceBuffer<float> stage_;
  DeviceBuffer<int32_t> istage_;
  BlendUpload upload_;
  DeviceBuffer<int32_t> ids_;
  DeviceBuffer<uint32_t> block_sums_;
  DeviceBuffer<uint8_t> img_;
  DeviceBuffer<double> err_;
  DeviceBuffer<int> flag_;
  double *host_err_ = nullptr;
  int *host_flag_ = nullptr;
  // state of the fused Processor-level reset (scatter list + crop canvas)
  bool fused_ = false;
  int crop_n_ = 0, crop_m_ = 0;
  DeviceBuffer<int32_t> pix_;     // [K] linear crop index of unknown i+1
  DeviceBuffer<uint8_t> canvas_;  // [crop_n, crop_m, 3] target pixels of the crop
  // CUDA graph of kGraphSweeps gather sweeps per starting buffer (long runs on small systems)
  static constexpr int kGraphSweeps = 64;
  cudaStream_t cap_stream_ = nullptr;
  cudaGraphExec_t graph_[2] = {nullptr, nullptr};
  bool graph_off_ = false, graph_warm_ = false;
  EquStats stats_;
};

}  // namespace fpie
