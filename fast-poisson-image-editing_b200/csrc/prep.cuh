// Device-side Processor preprocessing shared by the fused resets
// (fpie/process.py:209-224 / 338-351 mask canonicalisation, :113-122 mixgrad,
// :241-246 / :362-376 gradient).
#pragma once

#include "common.cuh"

namespace fpie {

// The three uint8 images of one blend, resident on the device, plus the crop.
struct BlendImages {
  const uint8_t *src;
  const uint8_t *mask;
  const uint8_t *tgt;
  int sh, sw;      // source rows / cols
  int mh, mw, mc;  // mask rows / cols / channels (any count >= 1; the reference takes mean(-1))
  int th, tw;      // target rows / cols
  // crop box in mask coordinates and the offsets of mask (0,0) in src / tgt
  int x0, y0, n, m;
  int h0, w0, h1, w1;
  int mode;  // FPIE_B200_GRAD_*
  // batch mode (batch > 0): src / mask / tgt hold `batch` patches of mh x mw pixels back to back;
  // the grid is a mosaic of bcols patches per row, each patch keeping its own fixed frame
  int batch, bcols;
  // slab mode (row-band sharding): the mask image is already the canonical crop of the global mask and
  // the whole image is the grid; its first / last rows are halo rows, not a real frame
  int slab;
};

// Batch mosaic geometry shared by the output kernels (batch == 0: plain image).
struct BatchMap {
  int batch, ph, pw, bcols;
};

struct CropBox {
  int x0, x1, y0, y1;  // in mask coordinates; crop = mask[x0:x1, y0:y1]
};

// Upload the images (host -> device buffers owned by the caller), find the
// bounding box of the canonical mask and validate it against src / tgt.
// Throws fpie::Error for an empty mask or a box that leaves an image.
// Row chunks of a crop-mode upload: crop rows [0, row_hi[k]) of the source and the target are on the device once
// ev[k] has completed (the copies run on the uploader's own stream, beside whatever the caller enqueues).
struct UploadChunks {
  static constexpr int kMax = 8;
  int count = 0;
  int row_hi[kMax] = {};
  cudaEvent_t ev[kMax] = {};
};

class BlendUpload {
 public:
  ~BlendUpload();
  // The mask goes first: its bounding box decides which rows of the source and the target are needed at all
  // (only those are copied), and the caller learns the crop geometry while they are still on their way.  With
  // `chunks` the caller gets one event per row chunk and orders its own work behind them; without, `stream` itself
  // waits for the last copy.
  void upload(cudaStream_t stream, const uint8_t *src, int sh, int sw, const uint8_t *mask, int mh, int mw, int mc,
              const uint8_t *tgt, int th, int tw, int h0, int w0, int h1, int w1, int mode, bool crop = true,
              UploadChunks *chunks = nullptr);
  // `batch` patches of ph x pw pixels each (src / tgt [batch, ph, pw, 3], mask [batch, ph, pw, mc])
  void upload_batch(cudaStream_t stream, const uint8_t *src, const uint8_t *mask, const uint8_t *tgt, int batch, int ph,
                    int pw, int mc, int mode, int bcols);
  const BlendImages &images() const { return img_; }
  void release();
  void destroy_stream();
  // block until the row copies of the last upload have left the caller's buffers (error paths: the caller gets its
  // buffers back when the reset call returns, whatever happened)
  void wait_copies() {
    if (copy_stream_) cudaStreamSynchronize(copy_stream_);
  }
  // called (on the calling thread, from inside `upload`) as soon as the blend's bounding box in TARGET coordinates
  // (x0, x1, y0, y1) is known -- before the source / target rows travel: lets the host prepare its side meanwhile
  void set_box_callback(void (*cb)(void *, const int32_t *), void *user) {
    box_cb_ = cb;
    box_user_ = user;
  }

 private:
  DeviceBuffer<uint8_t> src_, mask_, tgt_;
  DeviceBuffer<int> box_;
  BlendImages img_{};
  void (*box_cb_)(void *, const int32_t *) = nullptr;
  void *box_user_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;  // created on first use, on the device that is current then
  cudaEvent_t start_ev_ = nullptr, chunk_ev_[UploadChunks::kMax] = {};
};

#ifdef __CUDACC__
// canonical mask bit of mask-image pixel (r, c): threshold on the channel mean
// (mean(-1) >= 128  <=>  sum >= 128 * channels, exact) and a cleared 1-px frame.
__device__ __forceinline__ bool canonical_mask_at(const BlendImages &b, int r, int c) {
  if (r <= 0 || c <= 0 || r >= b.mh - 1 || c >= b.mw - 1) return false;
  const uint8_t *p = b.mask + ((long long)r * b.mw + c) * b.mc;
  int s = 0;
  for (int k = 0; k < b.mc; ++k) s += (int)p[k];
  return s >= 128 * b.mc;
}

// the thresholded mask bit without the cleared frame (inside the image only): in slab mode this is what
// the pixel is in the GLOBAL mask, also on the slab's halo frame rows
__device__ __forceinline__ bool raw_mask_at(const BlendImages &b, int r, int c) {
  if (r < 0 || c < 0 || r >= b.mh || c >= b.mw) return false;
  const uint8_t *p = b.mask + ((long long)r * b.mw + c) * b.mc;
  int s = 0;
  for (int k = 0; k < b.mc; ++k) s += (int)p[k];
  return s >= 128 * b.mc;
}

// In batch mode, re-base the image pointers on the patch that holds mosaic pixel
// (i, j) and return the patch-local coordinates; false if (i, j) is in no patch.
__device__ __forceinline__ bool patch_view(const BlendImages &b, int i, int j, BlendImages &pb, int &pi, int &pj) {
  pb = b;
  pi = i;
  pj = j;
  if (b.batch <= 0) return true;
  const int by = i / b.mh, bx = j / b.mw;
  const int id = by * b.bcols + bx;
  if (bx >= b.bcols || id >= b.batch) return false;
  const long long base = (long long)id * b.mh * b.mw;
  pb.src += base * 3;
  pb.tgt += base * 3;
  pb.mask += base * b.mc;
  pi = i - by * b.mh;
  pj = j - bx * b.mw;
  return true;
}

__device__ __forceinline__ float mix_one(int mode, float a, float b) {
  if (mode == 0) return a;               // src
  if (mode == 1) return (a + b) * 0.5f;  // avg (exact: integers)
  return (fabsf(a) < fabsf(b)) ? b : a;  // max: strict <, ties keep the source difference
}

// grad(p) for crop pixel (i, j), channel ch: sum over the four neighbours q of
// mix(src(p) - src(q), tgt(p) - tgt(q)).  All operands are small integers or
// halves, so the fp32 sum is exact in any order.
__device__ __forceinline__ float pixel_gradient(const BlendImages &b, int i, int j, int ch) {
  const long long sp = ((long long)(b.h0 + b.x0 + i) * b.sw + (b.w0 + b.y0 + j)) * 3 + ch;
  const long long tp = ((long long)(b.h1 + b.x0 + i) * b.tw + (b.w1 + b.y0 + j)) * 3 + ch;
  const float sc = (float)b.src[sp], tc = (float)b.tgt[tp];
  const long long so[4] = {-(long long)b.sw * 3, (long long)b.sw * 3, -3, 3};
  const long long to[4] = {-(long long)b.tw * 3, (long long)b.tw * 3, -3, 3};
  float g = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) g += mix_one(b.mode, sc - (float)b.src[sp + so[k]], tc - (float)b.tgt[tp + to[k]]);
  return g;
}

__device__ __forceinline__ float target_at(const BlendImages &b, int i, int j, int ch) {
  return (float)b.tgt[((long long)(b.h1 + b.x0 + i) * b.tw + (b.w1 + b.y0 + j)) * 3 + ch];
}
#endif

}  // namespace fpie
