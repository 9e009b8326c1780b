// glue.cu -- upstream glue of the hot path on the GPU: image_crop, MONO8 conversion.
//
//   image_crop::ImageCrop::cameraCallback      image_crop/src/image_crop.cpp:20-43
// crops the camera image to target_width x target_height around the image centre after cv_bridge::toCvShare(msg,
// "bgr8") (a channel shuffle for rgb8 / bgra8 / rgba8 / mono8 sources) and shifts the principal point of the
// CameraInfo.  On the GPU it is one pass: every thread produces four output pixels (12 bytes, three aligned 32-bit
// stores) of the dense bgr8 result; reads of neighbouring threads are contiguous, so both sides move whole sectors.
// Two kernels: word loads + funnel shift + byte permute when a thread's four pixels share a row (target_w % 4 == 0),
// a byte-wise kernel otherwise.  Pure byte moves: HBM-bound, bit-exact by construction.
//
//   cv_bridge::toCvCopy(image, MONO8)           scene_flow_constructor/src/scene_flow_constructor.cpp:220-221
// converts both camera images to 8-bit grey in front of the viso2 odometer: cv::cvtColor with COLOR_BGR2GRAY /
// RGB2GRAY / BGRA2GRAY / RGBA2GRAY, which for 8-bit images is fixed-point arithmetic,
//     grey = (cB * B + cG * G + cR * R + (1 << (shift - 1))) >> shift,
// with (cR, cG, cB, shift) = (4899, 9617, 1868, 14) in the OpenCV 3.2 the reference is built against (ROS Melodic;
// imgproc/src/color.cpp: R2Y, G2Y, B2Y, yuv_shift) and (9798, 19235, 3735, 15) from OpenCV 4 on (RY15, GY15, BY15,
// gray_shift).  Both rules are integer-exact; the caller names the one its OpenCV uses (MOD_GRAY_*).  A mono8 source
// is copied.  One pass, four output pixels (one aligned 32-bit store) per thread on the fast path.
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

// Fast path (target_w % 4 == 0, 4-byte aligned output): a thread's four pixels lie in one row, so their source bytes are
// one contiguous run of 4 * kBpp bytes.  The run is fetched as aligned 32-bit words (one more than kBpp when it starts
// inside a word) and shifted into place with funnel shifts; the channel shuffle of the encoding is then a fixed byte
// permutation (PRMT) into three output words.  ~3x fewer instructions than the byte-wise kernel.
template <int kBpp, int kB, int kG, int kR>
__global__ void __launch_bounds__(256) image_crop_rows_kernel(const CropParams p) {
  // A CTA owns kU consecutive output rows (of the batch's crops laid end to end) and a thread the same group of four
  // columns in each of them: kU independent runs are in flight per thread, and the only division is one per CTA (a
  // 64-bit division per group made the first version instruction-bound: 76-84 % of the measured HBM peak).
  constexpr int kU = 4;
  const int groups_per_row = p.tw >> 2;
  const unsigned rows_total = static_cast<unsigned>(p.th) * p.batch;
  const unsigned row0 = blockIdx.x * kU;
  const unsigned frame0 = row0 / p.th;
  const int r0 = static_cast<int>(row0 - frame0 * p.th);
  for (int cg = threadIdx.x; cg < groups_per_row; cg += blockDim.x) {
    const int c = cg << 2;
    uint32_t wsrc[kU][kBpp + 1];
    unsigned shift[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) {
      shift[u] = 0;
#pragma unroll
      for (int k = 0; k <= kBpp; k++) wsrc[u][k] = 0u;
      if (row0 + u < rows_total) {
        int r = r0 + u;
        unsigned frame = frame0;
        if (r >= p.th) {
          frame += r / p.th;
          r = r % p.th;
        }
        const uint8_t* s = p.src + frame * p.src_frame_stride + static_cast<size_t>(p.y0 + r) * p.step +
                           static_cast<size_t>(p.x0 + c) * kBpp;
        shift[u] = static_cast<unsigned>(reinterpret_cast<uintptr_t>(s) & 3u);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s - shift[u]);
#pragma unroll
        for (int k = 0; k < kBpp; k++) wsrc[u][k] = __ldg(wp + k);
        if (shift[u]) wsrc[u][kBpp] = __ldg(wp + kBpp);  // only touched when the run really extends into it
      }
    }
#pragma unroll
    for (int u = 0; u < kU; u++) {
      if (row0 + u >= rows_total) break;
      uint32_t a[kBpp];
#pragma unroll
      for (int k = 0; k < kBpp; k++) a[k] = __funnelshift_r(wsrc[u][k], wsrc[u][k + 1], 8 * shift[u]);
      // output byte j = 3 * pixel + channel comes from source byte pixel * kBpp + {kB, kG, kR}[channel]
      uint32_t o[3];
#pragma unroll
      for (int wd = 0; wd < 3; wd++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int j = 4 * wd + b, px = j / 3, ch = j - 3 * px;
          const int sb = px * kBpp + (ch == 0 ? kB : (ch == 1 ? kG : kR));
          v |= ((a[sb >> 2] >> (8 * (sb & 3))) & 0xffu) << (8 * b);
        }
        o[wd] = v;
      }
      uint32_t* o32 = reinterpret_cast<uint32_t*>(p.out + (static_cast<size_t>(row0 + u) * p.tw + c) * 3);
      o32[0] = o[0];
      o32[1] = o[1];
      o32[2] = o[2];
    }
  }
}

struct GrayParams {
  const uint8_t* src;
  uint8_t* out;
  size_t src_frame_stride;  // bytes between consecutive source frames
  int step;                 // source row stride in bytes
  int w, h, batch;
  int bpp;                  // source bytes per pixel (1: copy)
  int c0, c1, c2, shift;    // weights of source bytes 0, 1, 2 of a pixel
};

__device__ __forceinline__ uint32_t grey_of(uint32_t b0, uint32_t b1, uint32_t b2, const GrayParams& p) {
  return (b0 * p.c0 + b1 * p.c1 + b2 * p.c2 + (1u << (p.shift - 1))) >> p.shift;
}

// general path: one output pixel per thread, byte loads
__global__ void __launch_bounds__(256) to_mono8_kernel(const GrayParams p) {
  const size_t per_frame = static_cast<size_t>(p.w) * p.h;
  const size_t total = per_frame * p.batch;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t frame = g / per_frame;
    const size_t in_frame = g - frame * per_frame;
    const int r = static_cast<int>(in_frame / p.w);
    const int c = static_cast<int>(in_frame - static_cast<size_t>(r) * p.w);
    const uint8_t* s = p.src + frame * p.src_frame_stride + static_cast<size_t>(r) * p.step + static_cast<size_t>(c) * p.bpp;
    p.out[g] = p.bpp == 1 ? __ldg(s) : static_cast<uint8_t>(grey_of(__ldg(s), __ldg(s + 1), __ldg(s + 2), p));
  }
}

// fast path (w % 4 == 0, 4-byte aligned output): a thread's four pixels are one contiguous run of 4 * kBpp source
// bytes, fetched as aligned 32-bit words and funnel-shifted into place (as in image_crop_rows_kernel)
template <int kBpp>
__global__ void __launch_bounds__(256) to_mono8_rows_kernel(const GrayParams p) {
  // A CTA owns kU consecutive image rows (of the batch's frames laid end to end); a thread takes the same group of four
  // columns in each of them, so kU independent runs are loaded before the first one is used (12 .. 16 bytes per thread
  // leave too few bytes in flight per SM to fill HBM) and the only division is one per CTA.
  constexpr int kU = 4;
  const int groups_per_row = p.w >> 2;
  const unsigned rows_total = static_cast<unsigned>(p.h) * p.batch;
  const unsigned row0 = blockIdx.x * kU;
  unsigned frame0 = row0 / p.h;
  const int r0 = static_cast<int>(row0 - frame0 * p.h);
  for (int cg = threadIdx.x; cg < groups_per_row; cg += blockDim.x) {
    const int c = cg << 2;
    uint32_t wsrc[kU][kBpp + 1];
    unsigned shift[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) {
      shift[u] = 0;
#pragma unroll
      for (int k = 0; k <= kBpp; k++) wsrc[u][k] = 0u;
      if (row0 + u < rows_total) {
        int r = r0 + u;
        unsigned frame = frame0;
        if (r >= p.h) {  // kU <= h is not required: wrap as often as needed
          frame += r / p.h;
          r = r % p.h;
        }
        const uint8_t* s = p.src + frame * p.src_frame_stride + static_cast<size_t>(r) * p.step + static_cast<size_t>(c) * kBpp;
        shift[u] = static_cast<unsigned>(reinterpret_cast<uintptr_t>(s) & 3u);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s - shift[u]);
#pragma unroll
        for (int k = 0; k < kBpp; k++) wsrc[u][k] = __ldg(wp + k);
        if (shift[u]) wsrc[u][kBpp] = __ldg(wp + kBpp);  // only touched when the run really extends into it
      }
    }
#pragma unroll
    for (int u = 0; u < kU; u++) {
      if (row0 + u >= rows_total) break;
      uint32_t a[kBpp];
#pragma unroll
      for (int k = 0; k < kBpp; k++) a[k] = __funnelshift_r(wsrc[u][k], wsrc[u][k + 1], 8 * shift[u]);
      uint32_t o = 0;
      if constexpr (kBpp == 1) {
        o = a[0];
      } else {
#pragma unroll
        for (int px = 0; px < 4; px++) {
          uint32_t b[3];
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            const int sb = px * kBpp + ch;
            b[ch] = (a[sb >> 2] >> (8 * (sb & 3))) & 0xffu;
          }
          o |= grey_of(b[0], b[1], b[2], p) << (8 * px);
        }
      }
      *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(row0 + u) * p.w + c) = o;
    }
  }
}

}  // namespace

cudaError_t launch_to_mono8(cudaStream_t s, const uint8_t* src, size_t src_frame_stride, int step, int encoding, int w,
                            int h, int batch, int rule, uint8_t* out) {
  static const int bpp_of[5] = {3, 3, 4, 4, 1};
  // (cR, cG, cB, shift): OpenCV 3.x (R2Y, G2Y, B2Y, yuv_shift) / OpenCV 4.x (RY15, GY15, BY15, gray_shift)
  static const int rules[2][4] = {{4899, 9617, 1868, 14}, {9798, 19235, 3735, 15}};
  const size_t total = static_cast<size_t>(w) * h * batch;
  if (total == 0) return cudaSuccess;
  const bool blue_first = encoding == MOD_ENC_BGR8 || encoding == MOD_ENC_BGRA8;
  const int* k = rules[rule];
  GrayParams p{src, out, src_frame_stride, step, w, h, batch, bpp_of[encoding], blue_first ? k[2] : k[0], k[1],
               blue_first ? k[0] : k[2], k[3]};
  if (w % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
    const size_t rows = static_cast<size_t>(h) * batch;
    if (rows > 0xffffffffull) return cudaErrorInvalidValue;
    const unsigned nb = static_cast<unsigned>((rows + 3) / 4);  // four rows per CTA
    switch (p.bpp) {
      case 3: to_mono8_rows_kernel<3><<<nb, 256, 0, s>>>(p); break;
      case 4: to_mono8_rows_kernel<4><<<nb, 256, 0, s>>>(p); break;
      default: to_mono8_rows_kernel<1><<<nb, 256, 0, s>>>(p); break;
    }
    return cudaGetLastError();
  }
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  to_mono8_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
  return cudaGetLastError();
}

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
  unsigned nb = static_cast<unsigned>(blocks);
  if (tw % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0 && static_cast<size_t>(th) * batch <= 0xffffffffull) {
    nb = static_cast<unsigned>((static_cast<size_t>(th) * batch + 3) / 4);  // four output rows per CTA
    switch (encoding) {
      case 0: image_crop_rows_kernel<3, 0, 1, 2><<<nb, 256, 0, s>>>(p); break;  // bgr8
      case 1: image_crop_rows_kernel<3, 2, 1, 0><<<nb, 256, 0, s>>>(p); break;  // rgb8
      case 2: image_crop_rows_kernel<4, 0, 1, 2><<<nb, 256, 0, s>>>(p); break;  // bgra8
      case 3: image_crop_rows_kernel<4, 2, 1, 0><<<nb, 256, 0, s>>>(p); break;  // rgba8
      default: image_crop_rows_kernel<1, 0, 0, 0><<<nb, 256, 0, s>>>(p); break; // mono8
    }
    return cudaGetLastError();
  }
  image_crop_kernel<<<nb, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace mod
