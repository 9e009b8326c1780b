// glue.cu -- upstream glue of the hot path on the GPU: image_crop.
//
//   image_crop::ImageCrop::cameraCallback      image_crop/src/image_crop.cpp:20-43
// crops the camera image to target_width x target_height around the image centre after cv_bridge::toCvShare(msg,
// "bgr8") (a channel shuffle for rgb8 / bgra8 / rgba8 / mono8 sources) and shifts the principal point of the
// CameraInfo.  On the GPU it is one pass: every thread produces four output pixels (12 bytes, three aligned 32-bit
// stores) of the dense bgr8 result from their source bytes; reads of neighbouring threads are contiguous, so both
// sides move whole sectors.  Pure byte moves: HBM-bound, bit-exact by construction.
#include "common.cuh"

namespace mod {

namespace {

struct CropParams {
  const uint8_t* src;
  uint8_t* out;
  size_t src_frame_stride;  // bytes between consecutive source frames
  int step;                 // source row stride in bytes
  int x0, y0, tw, th, batch;
  int bpp, ib, ig, ir;      // source bytes per pixel and the byte index of B, G, R inside a source pixel
};

__global__ void __launch_bounds__(256) image_crop_kernel(const CropParams p) {
  const size_t per_frame = static_cast<size_t>(p.tw) * p.th;
  const size_t total = per_frame * p.batch;
  const size_t groups = (total + 3) / 4;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint8_t px[12];
    const size_t first = 4 * g;
    const int n = static_cast<int>(total - first < 4 ? total - first : 4);
    size_t frame = first / per_frame;
    size_t in_frame = first - frame * per_frame;
    int r = static_cast<int>(in_frame / p.tw);
    int c = static_cast<int>(in_frame - static_cast<size_t>(r) * p.tw);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k < n) {
        const uint8_t* s = p.src + frame * p.src_frame_stride + static_cast<size_t>(p.y0 + r) * p.step +
                           static_cast<size_t>(p.x0 + c) * p.bpp;
        px[3 * k] = __ldg(s + p.ib);
        px[3 * k + 1] = __ldg(s + p.ig);
        px[3 * k + 2] = __ldg(s + p.ir);
        if (++c == p.tw) {
          c = 0;
          if (++r == p.th) {
            r = 0;
            frame++;
          }
        }
      } else {
        px[3 * k] = px[3 * k + 1] = px[3 * k + 2] = 0;
      }
    }
    uint8_t* o = p.out + first * 3;
    if (n == 4 && (reinterpret_cast<uintptr_t>(p.out) & 3u) == 0) {
      uint32_t* o32 = reinterpret_cast<uint32_t*>(o);  // first * 3 = 12 g: 4-byte aligned whenever the base is
      o32[0] = px[0] | (px[1] << 8) | (px[2] << 16) | (static_cast<uint32_t>(px[3]) << 24);
      o32[1] = px[4] | (px[5] << 8) | (px[6] << 16) | (static_cast<uint32_t>(px[7]) << 24);
      o32[2] = px[8] | (px[9] << 8) | (px[10] << 16) | (static_cast<uint32_t>(px[11]) << 24);
    } else {
      for (int i = 0; i < 3 * n; i++) o[i] = px[i];
    }
  }
}

}  // namespace

cudaError_t launch_image_crop(cudaStream_t s, const uint8_t* src, size_t src_frame_stride, int step, int encoding,
                              int x0, int y0, int tw, int th, int batch, uint8_t* out) {
  static const int bpp_of[5] = {3, 3, 4, 4, 1};
  static const int idx_of[5][3] = {{0, 1, 2}, {2, 1, 0}, {0, 1, 2}, {2, 1, 0}, {0, 0, 0}};
  const size_t total = static_cast<size_t>(tw) * th * batch;
  if (total == 0) return cudaSuccess;
  CropParams p{src, out, src_frame_stride, step, x0, y0, tw, th, batch, bpp_of[encoding],
               idx_of[encoding][0], idx_of[encoding][1], idx_of[encoding][2]};
  size_t blocks = ((total + 3) / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride beyond 16 CTAs per SM
  image_crop_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace mod
