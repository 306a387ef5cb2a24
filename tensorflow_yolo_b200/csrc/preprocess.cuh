// preprocess.cuh -- image preprocessing on the device: cv2.resize(INTER_LINEAR) of 8-bit BGR images to the network
// size, BGR -> RGB, leaving uint8 NHWC for the first conv (which applies the /255 table).
//
// Replaces the per-image host work of the reference's preprocess_image (net/base.py:115-155):
//     net_image = cv2.resize(image, tuple(new_shape[0:2]));  net_image = net_image[:, :, ::-1];  net_image / 255.
// bit for bit.  cv2's INTER_LINEAR for 8-bit images is fixed point (modules/imgproc/src/resize.cpp):
//   * per destination column: fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx, with sx clamped to
//     [0, src_w - 1] (fx = 0 when clamped); weights short(round((1 - fx) * 2048)), short(round(fx * 2048));
//     rows likewise but without clamping the weights (the row INDEX is clamped when it is used);
//     scale = 1. / ((double)dst / src)  -- the tables are built on the host in exactly these types (engine.cu);
//   * horizontal pass in int:  D = S[sx] * a0 + S[sx + 1] * a1            (up to 255 * 2048)
//   * vertical pass:           dst = (((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16) + 2) >> 2
//   * an exact 2x decimation in both directions is rerouted to INTER_AREA: (s00 + s01 + s10 + s11 + 2) >> 2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace yb {

struct ResizeImage {
  const unsigned char* src;   // device, BGR, row stride `stride` bytes
  int sh, sw, stride;
  int area2x;                 // 1: exact 2x decimation (INTER_AREA path)
  int xtab, ytab;             // offsets (in entries) of this image's column / row tables
};

struct ResizeTab {            // one destination column or row
  int ofs;                    // source index (columns: clamped; rows: may be -1 or src_h - 1 + ...)
  short c0, c1;               // fixed-point weights (11 fractional bits)
};

// one thread per destination pixel; out: [n, dh, dw, 3] uint8, channel order reversed (BGR -> RGB)
__global__ void __launch_bounds__(256) resize_bgr2rgb_kernel(const ResizeImage* __restrict__ imgs, const ResizeTab* __restrict__ tabs,
                                                             int dh, int dw, unsigned char* __restrict__ out) {
  const int img = blockIdx.z;
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  if (dx >= dw) return;
  const ResizeImage im = imgs[img];
  unsigned char* o = out + (((size_t)img * dh + dy) * dw + dx) * 3;
  if (im.area2x) {
    const unsigned char* r0 = im.src + (size_t)(2 * dy) * im.stride + (size_t)(2 * dx) * 3;
    const unsigned char* r1 = r0 + im.stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[2 - c] = (unsigned char)((r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2);
    return;
  }
  const ResizeTab tx = tabs[im.xtab + dx];
  const ResizeTab ty = tabs[im.ytab + dy];
  int y0 = ty.ofs, y1 = ty.ofs + 1;
  y0 = y0 < 0 ? 0 : (y0 < im.sh ? y0 : im.sh - 1);
  y1 = y1 < 0 ? 0 : (y1 < im.sh ? y1 : im.sh - 1);
  const int x0 = tx.ofs;
  const int x1 = (x0 + 1 < im.sw) ? x0 + 1 : im.sw - 1;     // weight c1 is 0 whenever x0 is the last column
  const unsigned char* p00 = im.src + (size_t)y0 * im.stride + (size_t)x0 * 3;
  const unsigned char* p01 = im.src + (size_t)y0 * im.stride + (size_t)x1 * 3;
  const unsigned char* p10 = im.src + (size_t)y1 * im.stride + (size_t)x0 * 3;
  const unsigned char* p11 = im.src + (size_t)y1 * im.stride + (size_t)x1 * 3;
  const int a0 = tx.c0, a1 = tx.c1, b0 = ty.c0, b1 = ty.c1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int d0 = p00[c] * a0 + p01[c] * a1;
    const int d1 = p10[c] * a0 + p11[c] * a1;
    const int v = (((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2;
    o[2 - c] = (unsigned char)v;
  }
}

}  // namespace yb
