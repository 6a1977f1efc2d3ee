// Overlay rendering on the device: the two layers the reference draws on top of a frame.
//   vector layer  pathfinder_viewer.py:179-191  cv2.polylines of the kept vectors (0,0,255), cv2.circle radius 1 in
//                 (255,0,255) at their starts, then the rejected ones in (255,255,0) when draw_bad_flow is set
//   lamp layer    pathfinder_viewer.py:210-222  HSV (0,255,V) at each kept point -> cv2.cvtColor(HSV2BGR) -> filled
//                 cv2.circle of radius 6 in the point's own colour
// The rasterisation is third-party (opencv imgproc/src/drawing.cpp), restated in oracle/overlay.py and checked there
// against cv2.line / cv2.circle: clipLine (64-bit integers, double quotient truncated), LineIterator with
// connectivity 8 walking left to right, the midpoint circle of radius 1 (four pixels) and the filled disc of radius 6
// (row half-widths 6,5,5,5,4,3,0).  Draw order = the reference's: a later primitive overwrites an earlier one, so
// every class of primitive is its own launch; inside a class all primitives have one colour (vectors) or do not
// overlap at the reference's 30-pixel grid (lamps).
#include "common.cuh"

namespace b2of {

__device__ __forceinline__ void put_bgr(uint8_t* img, int cols, int x, int y, uchar3 c) {
  uint8_t* p = img + ((size_t)y * cols + x) * 3;
  p[0] = c.x; p[1] = c.y; p[2] = c.z;
}

// cv::clipLine
__device__ __forceinline__ bool clip_line(long long right, long long bottom, long long& x1, long long& y1,
                                          long long& x2, long long& y2) {
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// one thread per vector of the wanted class (mask == want): cv2.line(layer, p, q, colour, 1)
__global__ void __launch_bounds__(128) overlay_lines(const int32_t* __restrict__ pts, const int32_t* __restrict__ nxt,
                                                      const uint8_t* __restrict__ mask, int n, int want, int rows,
                                                      int cols, uchar3 colour, uint8_t* __restrict__ layer) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n || mask[(size_t)b * n + i] != want) return;
  const size_t o = ((size_t)b * n + i) * 2;
  long long x1 = pts[o], y1 = pts[o + 1], x2 = nxt[o], y2 = nxt[o + 1];
  if (!clip_line(cols - 1, rows - 1, x1, y1, x2, y2)) return;
  uint8_t* img = layer + (size_t)b * rows * cols * 3;
  int dx = (int)(x2 - x1), dy = (int)(y2 - y1), sy = 1;
  int x = (int)x1, y = (int)y1;
  if (dx < 0) { dx = -dx; dy = -dy; x = (int)x2; y = (int)y2; }     // LineIterator(..., leftToRight = true)
  if (dy < 0) { dy = -dy; sy = -1; }
  const bool vert = dy > dx;
  if (vert) { const int t = dx; dx = dy; dy = t; }
  int err = dx - 2 * dy;
  const int plus = 2 * dx, minus = -2 * dy;
  for (int k = 0; k <= dx; ++k) {
    put_bgr(img, cols, x, y, colour);
    const bool m = err < 0;
    err += minus + (m ? plus : 0);
    if (vert) { y += sy; x += m ? 1 : 0; }
    else { x += 1; y += m ? sy : 0; }
  }
}

// cv2.circle(layer, p, radius 1, colour, thickness 1): the four 4-neighbours, clipped to the frame
__global__ void __launch_bounds__(128) overlay_dots(const int32_t* __restrict__ pts, const uint8_t* __restrict__ mask,
                                                     int n, int want, int rows, int cols, uchar3 colour,
                                                     uint8_t* __restrict__ layer) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n || mask[(size_t)b * n + i] != want) return;
  const size_t o = ((size_t)b * n + i) * 2;
  const int cx = pts[o], cy = pts[o + 1];
  uint8_t* img = layer + (size_t)b * rows * cols * 3;
  const int ox[4] = {0, -1, 1, 0}, oy[4] = {-1, 0, 0, 1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = cx + ox[k], y = cy + oy[k];
    if ((unsigned)x < (unsigned)cols && (unsigned)y < (unsigned)rows) put_bgr(img, cols, x, y, colour);
  }
}

// cv2.cvtColor(HSV2BGR) of (h = 0, s = 255, v): float32, v / 255 * 255 truncated (oracle/pathfinder.py::hsv2bgr_u8)
__device__ __forceinline__ uchar3 lamp_colour(uint8_t v) {
  const float vf = __fmul_rn((float)v, (float)(1.0 / 255.0));
  const float s = __fmul_rn(255.f, (float)(1.0 / 255.0));
  // sector 0, f = 0: b = v (1 - s), g = v (1 - s (1 - f)), r = v
  const float pb = __fmul_rn(vf, __fsub_rn(1.f, s));
  const float pg = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, 0.f))));
  auto q = [](float x) { return (uint8_t)min(max((int)__fmul_rn(x, 255.f), 0), 255); };
  return make_uchar3(q(pb), q(pg), q(vf));
}

// one warp per kept point: the filled disc of radius 6 in the point's own colour (113 pixels)
__global__ void __launch_bounds__(128) overlay_lamps(const int32_t* __restrict__ kept_pts,
                                                      const uint8_t* __restrict__ danger_v,
                                                      const int32_t* __restrict__ n_kept, int n, int rows, int cols,
                                                      uint8_t* __restrict__ bgr) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
  if (i >= n_kept[b]) return;
  const size_t o = ((size_t)b * n + i) * 2;
  const int cx = kept_pts[o], cy = kept_pts[o + 1];
  if ((unsigned)cx >= (unsigned)cols || (unsigned)cy >= (unsigned)rows) return;   // the reference would raise
  const uchar3 c = lamp_colour(danger_v[(size_t)b * n + i]);
  uint8_t* img = bgr + (size_t)b * rows * cols * 3;
  // 13 rows x 13 columns of candidates, the disc's row half-widths select
  for (int k = lane; k < 169; k += 32) {
    const int dy = k / 13 - 6, dx = k - (k / 13) * 13 - 6;
    const int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
    const int hw = ady == 0 ? 6 : ady <= 3 ? 5 : ady == 4 ? 4 : ady == 5 ? 3 : 0;
    const int x = cx + dx, y = cy + dy;
    if (adx <= hw && (unsigned)x < (unsigned)cols && (unsigned)y < (unsigned)rows) put_bgr(img, cols, x, y, c);
  }
}

int overlay_vectors_dev(const int32_t* all_pts, const int32_t* all_next, const uint8_t* mask, int n_pts, int batch,
                        int rows, int cols, int draw_bad, uint8_t* layer, cudaStream_t st) {
  const char* fn = "overlay_vectors";
  B2OF_ASSERT(n_pts >= 0 && batch >= 0 && rows > 0 && cols > 0, fn);
  if (batch == 0) return B2OF_OK;
  B2OF_ASSERT(layer != nullptr && (n_pts == 0 || (all_pts && all_next && mask)), fn);
  B2OF_CUDA(cudaMemsetAsync(layer, 0, (size_t)batch * rows * cols * 3, st));
  if (n_pts == 0) return B2OF_OK;
  dim3 g(cdiv(n_pts, 128), batch);
  overlay_lines<<<g, 128, 0, st>>>(all_pts, all_next, mask, n_pts, 1, rows, cols, make_uchar3(0, 0, 255), layer);
  B2OF_LAUNCH_CHECK();
  overlay_dots<<<g, 128, 0, st>>>(all_pts, mask, n_pts, 1, rows, cols, make_uchar3(255, 0, 255), layer);
  B2OF_LAUNCH_CHECK();
  if (draw_bad) {
    overlay_lines<<<g, 128, 0, st>>>(all_pts, all_next, mask, n_pts, 0, rows, cols, make_uchar3(255, 255, 0), layer);
    B2OF_LAUNCH_CHECK();
    overlay_dots<<<g, 128, 0, st>>>(all_pts, mask, n_pts, 0, rows, cols, make_uchar3(255, 255, 0), layer);
    B2OF_LAUNCH_CHECK();
  }
  return B2OF_OK;
}

int overlay_lamps_dev(const int32_t* kept_pts, const uint8_t* danger_v, const int32_t* n_kept, int n_pts, int batch,
                      int rows, int cols, uint8_t* bgr, cudaStream_t st) {
  const char* fn = "draw_sparse_lamps";
  B2OF_ASSERT(n_pts >= 0 && batch >= 0 && rows > 0 && cols > 0, fn);
  if (batch == 0) return B2OF_OK;
  B2OF_ASSERT(bgr != nullptr && (n_pts == 0 || (kept_pts && danger_v && n_kept)), fn);
  B2OF_CUDA(cudaMemsetAsync(bgr, 0, (size_t)batch * rows * cols * 3, st));
  if (n_pts == 0) return B2OF_OK;
  overlay_lamps<<<dim3(cdiv(n_pts, 4), batch), 128, 0, st>>>(kept_pts, danger_v, n_kept, n_pts, rows, cols, bgr);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

}  // namespace b2of
