// Upload + mask canonicalisation + bounding box for the fused resets.
// Follows fpie/process.py:209-224 (Equ) == :338-351 (Grid).

#include <algorithm>
#include <climits>

#include "prep.cuh"

namespace fpie {

// box = {min row, max row, min col, max col} of the canonical mask.
// Grid-stride over pixels; warp shuffle reduction, then one set of atomics per CTA.
__global__ void __launch_bounds__(256) mask_bbox_kernel(BlendImages b, int *__restrict__ box) {
  const long long total = (long long)b.mh * b.mw;
  int rmin = INT_MAX, rmax = INT_MIN, cmin = INT_MAX, cmax = INT_MIN;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / b.mw), c = (int)(idx % b.mw);
    if (canonical_mask_at(b, r, c)) {
      rmin = min(rmin, r);
      rmax = max(rmax, r);
      cmin = min(cmin, c);
      cmax = max(cmax, c);
    }
  }
  rmin = __reduce_min_sync(0xffffffffu, rmin);
  rmax = __reduce_max_sync(0xffffffffu, rmax);
  cmin = __reduce_min_sync(0xffffffffu, cmin);
  cmax = __reduce_max_sync(0xffffffffu, cmax);
  __shared__ int part[4][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    part[0][w] = rmin;
    part[1][w] = rmax;
    part[2][w] = cmin;
    part[3][w] = cmax;
  }
  __syncthreads();
  if (w == 0) {
    const bool in = lane < (int)(blockDim.x >> 5);
    rmin = __reduce_min_sync(0xffffffffu, in ? part[0][lane] : INT_MAX);
    rmax = __reduce_max_sync(0xffffffffu, in ? part[1][lane] : INT_MIN);
    cmin = __reduce_min_sync(0xffffffffu, in ? part[2][lane] : INT_MAX);
    cmax = __reduce_max_sync(0xffffffffu, in ? part[3][lane] : INT_MIN);
    if (lane == 0 && rmin <= rmax) {
      atomicMin(&box[0], rmin);
      atomicMax(&box[1], rmax);
      atomicMin(&box[2], cmin);
      atomicMax(&box[3], cmax);
    }
  }
}

void BlendUpload::release() {
  src_.release();
  mask_.release();
  tgt_.release();
}

// (called by the owning solver's destructor while its device is current; idempotent)
void BlendUpload::destroy_stream() {
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (start_ev_) cudaEventDestroy(start_ev_);
  for (auto &e : chunk_ev_) {
    if (e) cudaEventDestroy(e);
    e = nullptr;
  }
  copy_stream_ = nullptr;
  start_ev_ = nullptr;
}

BlendUpload::~BlendUpload() { destroy_stream(); }

void BlendUpload::upload_batch(cudaStream_t stream, const uint8_t *src, const uint8_t *mask, const uint8_t *tgt,
                               int batch, int ph, int pw, int mc, int mode, int bcols) {
  FPIE_REQUIRE(src && mask && tgt, "reset_batch: null image");
  FPIE_REQUIRE(batch > 0 && ph > 0 && pw > 0 && bcols > 0, "reset_batch: empty batch");
  FPIE_REQUIRE(mc >= 1 && mc <= 16, "reset_batch: mask must have 1..16 channels");
  FPIE_REQUIRE(mode >= 0 && mode <= 2, "reset_batch: unknown gradient mode");
  const size_t px = (size_t)batch * ph * pw;
  src_.resize(px * 3);
  mask_.resize(px * mc);
  tgt_.resize(px * 3);
  CUDA_CHECK(cudaMemcpyAsync(src_.ptr, src, px * 3, cudaMemcpyHostToDevice, stream));
  CUDA_CHECK(cudaMemcpyAsync(mask_.ptr, mask, px * mc, cudaMemcpyHostToDevice, stream));
  CUDA_CHECK(cudaMemcpyAsync(tgt_.ptr, tgt, px * 3, cudaMemcpyHostToDevice, stream));
  BlendImages b{};
  b.src = src_.ptr;
  b.mask = mask_.ptr;
  b.tgt = tgt_.ptr;
  b.sh = b.mh = b.th = ph;
  b.sw = b.mw = b.tw = pw;
  b.mc = mc;
  b.mode = mode;
  b.batch = batch;
  b.bcols = bcols;
  const int brows = (int)ceil_div(batch, bcols);
  b.n = brows * ph;
  b.m = bcols * pw;
  img_ = b;
}

void BlendUpload::upload(cudaStream_t stream, const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh, int mw,
                         int mc, const uint8_t *tgt, int th, int tw, int h0, int w0, int h1, int w1, int mode,
                         bool crop, UploadChunks *chunks) {
  FPIE_REQUIRE(src && mask && tgt, "reset_from_images: null image");
  FPIE_REQUIRE(sh > 0 && sw > 0 && mh > 0 && mw > 0 && th > 0 && tw > 0, "reset_from_images: empty image");
  FPIE_REQUIRE(mc >= 1 && mc <= 16, "reset_from_images: mask must have 1..16 channels");
  FPIE_REQUIRE(mode >= 0 && mode <= 2, "reset_from_images: unknown gradient mode");
  const size_t sbytes = (size_t)sh * sw * 3, mbytes = (size_t)mh * mw * mc, tbytes = (size_t)th * tw * 3;
  src_.resize(sbytes);
  mask_.resize(mbytes);
  tgt_.resize(tbytes);
  box_.resize(4);
  if (chunks) chunks->count = 0;
  CUDA_CHECK(cudaMemcpyAsync(mask_.ptr, mask, mbytes, cudaMemcpyHostToDevice, stream));
  const int init[4] = {INT_MAX, INT_MIN, INT_MAX, INT_MIN};
  CUDA_CHECK(cudaMemcpyAsync(box_.ptr, init, sizeof(init), cudaMemcpyHostToDevice, stream));

  BlendImages b{};
  b.src = src_.ptr;
  b.mask = mask_.ptr;
  b.tgt = tgt_.ptr;
  b.sh = sh; b.sw = sw; b.mh = mh; b.mw = mw; b.mc = mc; b.th = th; b.tw = tw;
  b.h0 = h0; b.w0 = w0; b.h1 = h1; b.w1 = w1;
  b.mode = mode;
  b.slab = crop ? 0 : 1;
  if (!crop) {
    // slab mode (row-band sharding): the whole mask image is the grid; its own frame acts as the
    // fixed boundary (global frame rows, or halo rows refreshed by the neighbour band)
    b.x0 = 0;
    b.y0 = 0;
    b.n = mh;
    b.m = mw;
    FPIE_REQUIRE(h0 >= 0 && w0 >= 0 && h0 + mh <= sh && w0 + mw <= sw, "reset: the slab falls outside the source image");
    FPIE_REQUIRE(h1 >= 0 && w1 >= 0 && h1 + mh <= th && w1 + mw <= tw, "reset: the slab falls outside the target image");
    CUDA_CHECK(cudaMemcpyAsync(src_.ptr, src, sbytes, cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(tgt_.ptr, tgt, tbytes, cudaMemcpyHostToDevice, stream));
    img_ = b;
    return;
  }
  const long long total = (long long)mh * mw;
  mask_bbox_kernel<<<(int)std::min<long long>(ceil_div(total, 256), 148 * 16), 256, 0, stream>>>(b, box_.ptr);
  CUDA_CHECK(cudaGetLastError());
  int box[4];
  CUDA_CHECK(cudaMemcpyAsync(box, box_.ptr, sizeof(box), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  // the reference dies with "zero-size array to reduction operation minimum" (process.py:220)
  FPIE_REQUIRE(box[0] <= box[1], "reset: the mask is empty after thresholding and clearing its 1-pixel frame");
  b.x0 = box[0] - 1;
  b.y0 = box[2] - 1;
  b.n = box[1] + 2 - b.x0;
  b.m = box[3] + 2 - b.y0;
  // the reference leaves these checks commented out (process.py:203-207) and
  // then wraps around or throws from numpy; refuse instead
  FPIE_REQUIRE(h0 + b.x0 >= 0 && w0 + b.y0 >= 0 && h0 + b.x0 + b.n <= sh && w0 + b.y0 + b.m <= sw,
               "reset: the mask bounding box falls outside the source image");
  FPIE_REQUIRE(h1 + b.x0 >= 0 && w1 + b.y0 >= 0 && h1 + b.x0 + b.n <= th && w1 + b.y0 + b.m <= tw,
               "reset: the mask bounding box falls outside the target image");
  if (box_cb_) {
    const int32_t bx[4] = {h1 + b.x0, h1 + b.x0 + b.n, w1 + b.y0, w1 + b.y0 + b.m};
    box_cb_(box_user_, bx);
  }
  // Only the crop's rows of the source and the target are read on the device (every gradient term of a masked
  // pixel stays inside the crop: its frame is unmasked): copy those rows, in chunks, on the uploader's own stream.
  if (!copy_stream_) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&start_ev_, cudaEventDisableTiming));
    for (auto &e : chunk_ev_) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  // (whatever still reads the previous images was enqueued on `stream`)
  CUDA_CHECK(cudaEventRecord(start_ev_, stream));
  CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, start_ev_, 0));
  const size_t srow = (size_t)sw * 3, trow = (size_t)tw * 3;
  const size_t crop_bytes = (size_t)b.n * (srow + trow);
  const int parts = (int)std::min<size_t>(UploadChunks::kMax, std::max<size_t>(1, crop_bytes / (12u << 20)));
  const uint8_t *s0 = src + (size_t)(h0 + b.x0) * srow, *t0 = tgt + (size_t)(h1 + b.x0) * trow;
  uint8_t *ds0 = src_.ptr + (size_t)(h0 + b.x0) * srow, *dt0 = tgt_.ptr + (size_t)(h1 + b.x0) * trow;
  int done = 0;
  for (int k = 0; k < parts; ++k) {
    const int hi = (int)((long long)b.n * (k + 1) / parts);
    if (hi > done) {
      CUDA_CHECK(cudaMemcpyAsync(ds0 + done * srow, s0 + done * srow, (size_t)(hi - done) * srow, cudaMemcpyHostToDevice, copy_stream_));
      CUDA_CHECK(cudaMemcpyAsync(dt0 + done * trow, t0 + done * trow, (size_t)(hi - done) * trow, cudaMemcpyHostToDevice, copy_stream_));
    }
    CUDA_CHECK(cudaEventRecord(chunk_ev_[k], copy_stream_));
    if (chunks) {
      chunks->row_hi[k] = hi;
      chunks->ev[k] = chunk_ev_[k];
    }
    done = hi;
  }
  if (chunks)
    chunks->count = parts;
  else
    CUDA_CHECK(cudaStreamWaitEvent(stream, chunk_ev_[parts - 1], 0));
  img_ = b;
}

}  // namespace fpie
