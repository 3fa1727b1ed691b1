// Row-band halo link of the GridSolver: the data plane of the multi-GPU path, behind the C ABI.
//
// The reference's multi-worker GridSolver swaps ONE halo row per neighbour with blocking
// MPI_Sendrecv between sweeps (fpie/core/mpi/grid.cc:118-135).  Here a band keeps `halo` rows of
// each neighbour, sweeps `halo` times between exchanges (deep halo: bit-identical to one device)
// and moves its band-edge rows with COPY-ENGINE peer copies over NVLink:
//
//   sender (halo stream):   wait for the edge tiles of the interval's last pass (event)
//                           cudaMemcpy2DAsync  my rows  ->  the neighbour's inbox[parity]   (3 planes)
//                           4-byte copy        my counter ->  the neighbour's flag[parity]
//   receiver (solver stream, before the edge tiles of the next interval's first pass):
//                           cuStreamWaitValue32(flag[parity] >= interval)       (no SM involved)
//                           cudaMemcpy2DAsync  inbox[parity] -> my halo rows                 (3 planes)
//
// No kernel of the exchange occupies an SM (round 1 used NCCL send/recv, whose kernels compete with a
// persistent 148-CTA sweep grid and advance all ranks in lock-step), neighbours synchronise pairwise
// through the flag words only, and the interior tiles of the passes on either side of an exchange run
// while the rows are in flight.  The inbox is double-buffered by interval parity; that is enough:
// a sender can only produce interval i + 2 after it has consumed the receiver's interval i + 1, which
// the receiver sent after it had emptied inbox[i % 2] (stream order on both sides).
//
// Between processes the inbox is shared through cudaIpcGetMemHandle / cudaIpcOpenMemHandle (one
// process per GPU, fpie_b200/band.py exchanges the 128-byte blobs over torch.distributed); inside
// one process (tests: several bands on one box) the blob carries the raw pointer.

#include <cuda.h>

#include <cstring>

#include "grid_solver.cuh"

namespace fpie {

namespace {

typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WriteValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

template <typename Fn>
Fn driver_fn(const char *name) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CUDA_CHECK(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q));
  FPIE_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "fpie_b200: stream memory operations are not available in this driver");
  return reinterpret_cast<Fn>(fn);
}

void stream_wait_geq(cudaStream_t s, const uint32_t *flag, uint32_t value) {
  static WaitValueFn fn = driver_fn<WaitValueFn>("cuStreamWaitValue32");
  const CUresult rc = fn(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(flag), value, CU_STREAM_WAIT_VALUE_GEQ);
  FPIE_REQUIRE(rc == CUDA_SUCCESS, "cuStreamWaitValue32 failed");
}

void stream_write(cudaStream_t s, uint32_t *word, uint32_t value) {
  static WriteValueFn fn = driver_fn<WriteValueFn>("cuStreamWriteValue32");
  const CUresult rc = fn(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(word), value, CU_STREAM_WRITE_VALUE_DEFAULT);
  FPIE_REQUIRE(rc == CUDA_SUCCESS, "cuStreamWriteValue32 failed");
}

struct Blob {  // what halo_export hands to the neighbour (<= kHaloBlobBytes)
  uint32_t magic;
  int32_t rows, cols, device;
  uint64_t raw;  // the inbox address in the exporting process
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(Blob) <= GridSolver::kHaloBlobBytes, "halo blob too large");
constexpr uint32_t kBlobMagic = 0xF91EB200u;

}  // namespace

void GridSolver::halo_free_side(HaloSide &s) {
  if (s.remote && s.remote_ipc) cudaIpcCloseMemHandle(s.remote);
  s.remote = nullptr;
  s.remote_ipc = false;
  if (s.inbox) cudaFree(s.inbox);
  s.inbox = nullptr;
  s.inbox_bytes = s.parity_bytes = 0;
  s.rows = 0;
  s.sent = s.received = 0;
}

void GridSolver::halo_disconnect() {
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device_);
  if (halo_stream_) cudaStreamSynchronize(halo_stream_);
  for (auto &s : halo_) halo_free_side(s);
  if (halo_seq_) cudaFree(halo_seq_);
  halo_seq_ = nullptr;
  if (ev_edge_) cudaEventDestroy(ev_edge_);
  if (ev_sent_) cudaEventDestroy(ev_sent_);
  ev_edge_ = ev_sent_ = nullptr;
  if (halo_stream_) cudaStreamDestroy(halo_stream_);
  halo_stream_ = nullptr;
  for (auto e : trace_ev_) cudaEventDestroy(e);
  trace_ev_.clear();
  trace_tag_.clear();
  halo_pending_ = false;
  if (prev >= 0 && prev != device_) cudaSetDevice(prev);
}

// Geometry of the link.  Keeps an existing link (inboxes, counters, mappings) when the geometry is unchanged,
// so that repeated resets of the same problem (the end-to-end loop, the GUI) reconnect nothing.
bool GridSolver::halo_config(int band_lo, int band_hi, bool force) {
  require_ready();
  FPIE_REQUIRE(0 <= band_lo && band_lo < band_hi && band_hi <= geom_.n, "halo_config: band outside the slab");
  DeviceGuard guard(device_);
  const int rows[2] = {band_lo, geom_.n - band_hi};
  FPIE_REQUIRE(band_hi - band_lo >= std::max(rows[0], rows[1]), "halo_config: the band is shorter than its halo");
  const bool same = !force && halo_stream_ && band_lo == band_lo_ && band_hi == band_hi_ && halo_[0].rows == rows[0] &&
                    halo_[1].rows == rows[1] &&
                    halo_[0].parity_bytes == (size_t)3 * rows[0] * geom_.m * sizeof(float) &&
                    halo_[1].parity_bytes == (size_t)3 * rows[1] * geom_.m * sizeof(float);
  band_lo_ = band_lo;
  band_hi_ = band_hi;
  set_row_window(band_lo, band_hi);
  const int depth = std::max(rows[0], rows[1]);
  // edge tiles = those that write the rows we send (2 * halo from the slab edge) or read the halo rows we
  // receive (their load region reaches block_k rows beyond what they store)
  if (depth > 0) set_edge_rows(2 * depth + block_k_);
  preload_kernels();  // (before any stream of this process can be waiting for a neighbour)
  if (same) return false;
  halo_disconnect();
  if (depth == 0) return true;
  CUDA_CHECK(cudaStreamCreateWithFlags(&halo_stream_, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_edge_, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_sent_, cudaEventDisableTiming));
  CUDA_CHECK(cudaMalloc(&halo_seq_, 2 * sizeof(uint32_t)));
  CUDA_CHECK(cudaMemset(halo_seq_, 0, 2 * sizeof(uint32_t)));
  for (int side = 0; side < 2; ++side) {
    HaloSide &s = halo_[side];
    s.rows = rows[side];
    if (!s.rows) continue;
    s.parity_bytes = (size_t)3 * s.rows * geom_.m * sizeof(float);
    s.inbox_bytes = 2 * s.parity_bytes + 256;
    CUDA_CHECK(cudaMalloc(&s.inbox, s.inbox_bytes));
    CUDA_CHECK(cudaMemset(s.inbox + 2 * s.parity_bytes, 0, 256));
    s.row_send = side == 0 ? band_lo : band_hi - s.rows;
    s.row_recv = side == 0 ? 0 : band_hi;
  }
  CUDA_CHECK(cudaDeviceSynchronize());
  return true;
}

void GridSolver::halo_export(int side, unsigned char *blob_out) {
  FPIE_REQUIRE(side == 0 || side == 1, "halo_export: side must be 0 (up) or 1 (down)");
  FPIE_REQUIRE(blob_out, "halo_export: null output");
  DeviceGuard guard(device_);
  Blob b{};
  const HaloSide &s = halo_[side];
  b.magic = kBlobMagic;
  b.rows = s.rows;
  b.cols = geom_.m;
  b.device = device_;
  b.raw = reinterpret_cast<uint64_t>(s.inbox);
  if (s.inbox) CUDA_CHECK(cudaIpcGetMemHandle(&b.ipc, s.inbox));
  memset(blob_out, 0, kHaloBlobBytes);
  memcpy(blob_out, &b, sizeof(b));
}

void GridSolver::halo_connect(int side, const unsigned char *blob_in, bool same_process) {
  FPIE_REQUIRE(side == 0 || side == 1, "halo_connect: side must be 0 (up) or 1 (down)");
  FPIE_REQUIRE(blob_in, "halo_connect: null blob");
  DeviceGuard guard(device_);
  HaloSide &s = halo_[side];
  FPIE_REQUIRE(s.rows > 0, "halo_connect: this slab has no neighbour on that side (halo_config first)");
  Blob b;
  memcpy(&b, blob_in, sizeof(b));
  FPIE_REQUIRE(b.magic == kBlobMagic, "halo_connect: not a halo blob");
  FPIE_REQUIRE(b.rows == s.rows && b.cols == geom_.m, "halo_connect: the neighbour's halo geometry differs from this slab's");
  if (s.remote && s.remote_ipc) cudaIpcCloseMemHandle(s.remote);
  s.remote = nullptr;
  if (same_process) {
    if (b.device != device_) {
      int can = 0;
      CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device_, b.device));
      FPIE_REQUIRE(can, "halo_connect: the neighbour's device is not peer-accessible");
      const cudaError_t rc = cudaDeviceEnablePeerAccess(b.device, 0);
      if (rc != cudaSuccess && rc != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(rc);
      cudaGetLastError();
    }
    s.remote = reinterpret_cast<unsigned char *>(b.raw);
    s.remote_ipc = false;
  } else {
    void *p = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
    s.remote = static_cast<unsigned char *>(p);
    s.remote_ipc = true;
  }
}

// Counters of the link for diagnostics, readable while the solver's streams are blocked:
// out[0..1] rows, [2..3] sent, [4..5] received, [6..9] flag words (side 0 parity 0/1, side 1 parity 0/1),
// [10] current buffer, [11] block_k, [12] variant, [13..14] edge / interior tile entries, [15] pending
void GridSolver::halo_debug(int64_t *out) {
  DeviceGuard guard(device_);
  for (int i = 0; i < 16; ++i) out[i] = -1;
  cudaStream_t tmp = nullptr;
  CUDA_CHECK(cudaStreamCreateWithFlags(&tmp, cudaStreamNonBlocking));
  for (int side = 0; side < 2; ++side) {
    const HaloSide &s = halo_[side];
    out[side] = s.rows;
    out[2 + side] = s.sent;
    out[4 + side] = s.received;
    if (s.inbox) {
      uint32_t f[2] = {0, 0};
      cudaMemcpyAsync(f, s.inbox + 2 * s.parity_bytes, sizeof(f), cudaMemcpyDeviceToHost, tmp);
      cudaStreamSynchronize(tmp);
      out[6 + 2 * side] = f[0];
      out[7 + 2 * side] = f[1];
    }
  }
  cudaStreamDestroy(tmp);
  out[10] = cur_;
  out[11] = block_k_;
  out[12] = variant_;
  out[13] = n_part_[0];
  out[14] = n_part_[1];
  out[15] = halo_pending_ ? 1 : 0;
}

void GridSolver::trace_mark(int tag, cudaStream_t s) {
  if (trace_left_ <= 0) return;
  cudaEvent_t e = nullptr;
  CUDA_CHECK(cudaEventCreate(&e));
  CUDA_CHECK(cudaEventRecord(e, s));
  trace_ev_.push_back(e);
  trace_tag_.push_back(tag);
}

void GridSolver::halo_trace_begin(int max_intervals) {
  for (auto e : trace_ev_) cudaEventDestroy(e);
  trace_ev_.clear();
  trace_tag_.clear();
  trace_left_ = max_intervals;
}

// (tag, milliseconds since the first mark) pairs of the traced intervals; returns the number of floats written
int GridSolver::halo_trace_read(float *out, int max_floats) {
  DeviceGuard guard(device_);
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (halo_stream_) CUDA_CHECK(cudaStreamSynchronize(halo_stream_));
  int n = 0;
  for (size_t i = 0; i < trace_ev_.size() && n + 2 <= max_floats; ++i) {
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, trace_ev_[0], trace_ev_[i]));
    out[n++] = (float)trace_tag_[i];
    out[n++] = ms;
  }
  return n;
}

// Band-edge rows of state buffer `which` -> the neighbours' inboxes, on the halo stream, ordered after
// everything the solver's stream holds so far (the edge tiles that produced those rows).
void GridSolver::halo_send(int which) {
  const PlaneGeom &g = geom_;
  CUDA_CHECK(cudaEventRecord(ev_edge_, stream_));
  CUDA_CHECK(cudaStreamWaitEvent(halo_stream_, ev_edge_, 0));
  trace_mark(10, halo_stream_);
  const size_t row_bytes = (size_t)g.m * sizeof(float), pitch_bytes = (size_t)g.pitch * sizeof(float);
  for (int side = 0; side < 2; ++side) {
    HaloSide &s = halo_[side];
    if (!s.rows) continue;
    FPIE_REQUIRE(s.remote, "band_sweeps: the halo link is not connected (halo_connect)");
    const int parity = (int)(s.sent & 1u);
    for (int p = 0; p < 3; ++p) {
      const float *src = x_[which].ptr + (size_t)p * g.plane + (size_t)(g.padr + s.row_send) * g.pitch + g.padc;
      unsigned char *dst = s.remote + (size_t)parity * s.parity_bytes + (size_t)p * s.rows * row_bytes;
      CUDA_CHECK(cudaMemcpy2DAsync(dst, row_bytes, src, pitch_bytes, row_bytes, s.rows, cudaMemcpyDefault, halo_stream_));
    }
    s.sent += 1;
    // the flag: a counter word written locally by the stream, then copied like the data (ordered after it)
    stream_write(halo_stream_, halo_seq_ + side, s.sent);
    uint32_t *flag = reinterpret_cast<uint32_t *>(s.remote + 2 * s.parity_bytes) + parity;
    CUDA_CHECK(cudaMemcpyAsync(flag, halo_seq_ + side, sizeof(uint32_t), cudaMemcpyDefault, halo_stream_));
  }
  trace_mark(11, halo_stream_);
  CUDA_CHECK(cudaEventRecord(ev_sent_, halo_stream_));
  halo_pending_ = true;
  halo_exchanges_ += 1;
}

// The neighbours' rows of the interval in flight -> the halo rows of state buffer `which`, on the solver's
// stream: a stream-level wait on the flag word (no SM spins), then local copies out of the inbox.
void GridSolver::halo_recv(int which) {
  if (!halo_pending_) return;
  const PlaneGeom &g = geom_;
  const size_t row_bytes = (size_t)g.m * sizeof(float), pitch_bytes = (size_t)g.pitch * sizeof(float);
  trace_mark(20, stream_);
  for (int side = 0; side < 2; ++side) {
    HaloSide &s = halo_[side];
    if (!s.rows) continue;
    const int parity = (int)(s.received & 1u);
    const uint32_t *flag = reinterpret_cast<const uint32_t *>(s.inbox + 2 * s.parity_bytes) + parity;
    stream_wait_geq(stream_, flag, s.received + 1);
    for (int p = 0; p < 3; ++p) {
      float *dst = x_[which].ptr + (size_t)p * g.plane + (size_t)(g.padr + s.row_recv) * g.pitch + g.padc;
      const unsigned char *src = s.inbox + (size_t)parity * s.parity_bytes + (size_t)p * s.rows * row_bytes;
      CUDA_CHECK(cudaMemcpy2DAsync(dst, pitch_bytes, src, row_bytes, row_bytes, s.rows, cudaMemcpyDeviceToDevice, stream_));
    }
    s.received += 1;
  }
  // my own rows have left before the edge tiles two passes on overwrite them
  CUDA_CHECK(cudaStreamWaitEvent(stream_, ev_sent_, 0));
  trace_mark(21, stream_);
  halo_pending_ = false;
}

// `iters` Jacobi sweeps on a band: passes of block_k sweeps, halo rows refreshed every `halo` sweeps.  The
// LAST pass of an interval sweeps its edge tiles first and hands the finished band-edge rows to the copy
// engines while its interior tiles are still being swept; the FIRST pass of the next interval sweeps its
// interior before it needs the received rows.  Same arithmetic, same bits as one device.
void GridSolver::band_sweeps_async(int iters) {
  require_ready();
  FPIE_REQUIRE(iters >= 0, "band_sweeps: negative iteration count");
  const int depth = std::max(halo_[0].rows, halo_[1].rows);
  if (depth == 0) {  // a single band: nothing to exchange
    sweeps_async(iters);
    return;
  }
  FPIE_REQUIRE(variant_ != 1, "band_sweeps: not available for the one-sweep-per-launch kernels");
  FPIE_REQUIRE(edge_rows_ > 0, "band_sweeps needs halo_config");
  DeviceGuard guard(device_);
  int left = iters;
  const int k = block_k_;
  int interval = 0;
  while (left > 0) {
    const int s = std::min(depth, left);
    const int npass = (int)ceil_div(s, k);
    for (int idx = 0; idx < npass; ++idx) {
      const int ns = std::min(k, s - idx * k);
      const bool first = idx == 0, last = idx == npass - 1;
      if (first && last) {
        halo_recv(cur_);
        trace_mark(1, stream_);
        run_pass(ns, tiles_part_[0].ptr, n_part_[0]);
        trace_mark(2, stream_);
        halo_send(cur_ ^ 1);
        run_pass(ns, tiles_part_[1].ptr, n_part_[1]);
        trace_mark(3, stream_);
      } else if (first) {
        trace_mark(4, stream_);
        run_pass(ns, tiles_part_[1].ptr, n_part_[1]);
        trace_mark(5, stream_);
        halo_recv(cur_);
        run_pass(ns, tiles_part_[0].ptr, n_part_[0]);
        trace_mark(6, stream_);
      } else if (last) {
        run_pass(ns, tiles_part_[0].ptr, n_part_[0]);
        trace_mark(2, stream_);
        halo_send(cur_ ^ 1);
        run_pass(ns, tiles_part_[1].ptr, n_part_[1]);
        trace_mark(3, stream_);
      } else {
        run_pass(ns, tiles_part_[0].ptr, n_part_[0]);
        run_pass(ns, tiles_part_[1].ptr, n_part_[1]);
      }
      cur_ ^= 1;
    }
    left -= s;
    if (trace_left_ > 0 && ++interval >= trace_left_) trace_left_ = 0;
  }
  halo_recv(cur_);  // leave the halo rows fresh for whoever runs next
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace fpie
