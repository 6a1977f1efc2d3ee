// K7-K9: Shi-Tomasi corners, replaces cv2.goodFeaturesToTrack as called at SparseOF.py:69 (feature_params
// SparseOF.py:10-13, disc mask SparseOF.py:61-66).  Arithmetic spec: SURVEY.md App. A.5 (opencv corner.cpp +
// featureselect.cpp -- third-party, restated in oracle/gftt.py).
//   K7 gftt_mineig      : Sobel 3x3 -> products -> blockSize^2 box (REFLECT_101) -> min eigenvalue, + masked max
//   K8 gftt_nms_compact : quality threshold + 3x3 non-max suppression + mask, warp-aggregated compaction
//   K9 gftt_select      : sort (value desc, address desc) + greedy min-distance pass (sequential semantics)
#include <math.h>

#include "common.cuh"

namespace b2of {

constexpr int GF_T = 32;  // output tile edge

__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Sobel taps of cv2.getDerivKernels for ksize 5 and 7 (derivative, smoothing)
__constant__ float c_sobel_d[2][7] = {{-1, -2, 0, 2, 1, 0, 0}, {-1, -4, -5, 0, 5, 4, 1}};
__constant__ float c_sobel_s[2][7] = {{1, 4, 6, 4, 1, 0, 0}, {1, 6, 15, 20, 15, 6, 1}};

// KS = 3: the reference's path (3x3 Sobel straight from the image).  KS = 5 / 7 (cv2's gradientSize): the tile's source
// pixels are staged in shared memory, the float32 row filters of cv2's separable Sobel run once per tile row and the
// column filters once per halo pixel -- same taps, same fused multiply-add order as opencv's sepFilter2D
// (oracle/gftt.py::sobel_f32, probed against cv2.Sobel bit for bit).
template <int KS>
__global__ void __launch_bounds__(256) gftt_mineig(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                                    size_t step, size_t frame_stride, int w, int h, int bs,
                                                    float scale, int harris, float hk, float* __restrict__ eig,
                                                    unsigned int* __restrict__ max_key) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int r = bs / 2;
  const int E = GF_T + bs - 1;  // halo tile edge (anchor at bs/2, window [-r, bs-1-r])
  float* sC = (float*)gsm;                       // [3][E][E]
  double* sH = (double*)(sC + ((3 * E * E + 1) & ~1));  // [3][E][GF_T]
  const int x0 = blockIdx.x * GF_T, y0 = blockIdx.y * GF_T;
  const uint8_t* ib = img + blockIdx.z * frame_stride;
  const int t = threadIdx.x;
  if (KS != 3) {
    constexpr int RS = KS / 2;
    // halo pixel (iy, ix) sits at image position (reflect(y0 - r + iy), reflect(x0 - r + ix)); its Sobel window is
    // taken around THAT position (cv2 filters the whole image, then box-filters with REFLECT_101), so the staged
    // source rows / columns are the image range those positions and their windows touch
    int px_lo = w, px_hi = -1, py_lo = h, py_hi = -1;
    for (int i = 0; i < E; ++i) {
      const int px = reflect101(x0 - r + i, w), py = reflect101(y0 - r + i, h);
      px_lo = min(px_lo, px); px_hi = max(px_hi, px); py_lo = min(py_lo, py); py_hi = max(py_hi, py);
    }
    const int sx_lo = max(px_lo - RS, 0), sx_hi = min(px_hi + RS, w - 1);
    const int sy_lo = max(py_lo - RS, 0), sy_hi = min(py_hi + RS, h - 1);
    const int SW = sx_hi - sx_lo + 1, SH = sy_hi - sy_lo + 1;          // <= E + 2 RS each
    float* sT = (float*)(sH + 3 * E * GF_T);       // [SH][E] derivative row filter (exact integers)
    float* sR = sT + (E + 2 * RS) * E;             // [SH][E] scaled smoothing row filter
    uint8_t* sU = (uint8_t*)(sR + (E + 2 * RS) * E);   // [SH][SW] source pixels
    for (int i = t; i < SH * SW; i += 256) {
      const int yy = i / SW, xx = i - yy * SW;
      sU[i] = ib[(size_t)(sy_lo + yy) * step + sx_lo + xx];
    }
    __syncthreads();
    float ks_[KS], kd_[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      kd_[j] = c_sobel_d[KS == 5 ? 0 : 1][j];
      ks_[j] = __fmul_rn(c_sobel_s[KS == 5 ? 0 : 1][j], scale);
    }
    for (int i = t; i < SH * E; i += 256) {
      const int yy = i / E, ix = i - yy * E;
      const int px = reflect101(x0 - r + ix, w);
      const uint8_t* row = sU + yy * SW - sx_lo;
      float tt = 0.f, rr = 0.f;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const float v = (float)row[reflect101(px + j - RS, w)];
        tt = __fmaf_rn(kd_[j], v, tt);                      // integers: exact in any order
        rr = j == 0 ? __fmul_rn(v, ks_[0]) : __fmaf_rn(v, ks_[j], rr);
      }
      sT[i] = tt; sR[i] = rr;
    }
    __syncthreads();
    for (int i = t; i < E * E; i += 256) {
      const int iy = i / E, ix = i - iy * E;
      const int py = reflect101(y0 - r + iy, h);
      auto rowi = [&](int j) { return (reflect101(py + j, h) - sy_lo) * E + ix; };
      float fx = __fmul_rn(sT[rowi(0)], ks_[RS]);
#pragma unroll
      for (int j = 1; j <= RS; ++j) fx = __fmaf_rn(__fadd_rn(sT[rowi(j)], sT[rowi(-j)]), ks_[RS + j], fx);
      float fy = __fmul_rn(__fsub_rn(sR[rowi(1)], sR[rowi(-1)]), kd_[RS + 1]);
#pragma unroll
      for (int j = 2; j <= RS; ++j) fy = __fmaf_rn(__fsub_rn(sR[rowi(j)], sR[rowi(-j)]), kd_[RS + j], fy);
      sC[i] = __fmul_rn(fx, fx);
      sC[E * E + i] = __fmul_rn(fx, fy);
      sC[2 * E * E + i] = __fmul_rn(fy, fy);
    }
  } else
  for (int i = t; i < E * E; i += 256) {
    int iy = i / E, ix = i - iy * E;
    int x = reflect101(x0 - r + ix, w), y = reflect101(y0 - r + iy, h);
    int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const uint8_t* r0 = ib + (size_t)reflect101(y - 1, h) * step;
    const uint8_t* r1 = ib + (size_t)y * step;
    const uint8_t* r2 = ib + (size_t)reflect101(y + 1, h) * step;
    // cv2.Sobel(CV_32F, scale) as opencv's separable filter rounds it in its SIMD body (probed bit for bit against
    // cv2 4.13, oracle/gftt.py::sobel3_f32): the scaled smoothing taps [s, 2s, s] go through fused multiply-adds
    const float s2 = 2.f * scale;
    float t0 = (float)(r0[xp] - r0[xm]), t1 = (float)(r1[xp] - r1[xm]), t2 = (float)(r2[xp] - r2[xm]);
    float fx = __fmaf_rn(t0 + t2, scale, __fmul_rn(t1, s2));
    float ra = __fmaf_rn((float)r0[xp], scale, __fmaf_rn((float)r0[x], s2, __fmul_rn((float)r0[xm], scale)));
    float rb = __fmaf_rn((float)r2[xp], scale, __fmaf_rn((float)r2[x], s2, __fmul_rn((float)r2[xm], scale)));
    float fy = __fsub_rn(rb, ra);
    sC[i] = __fmul_rn(fx, fx);
    sC[E * E + i] = __fmul_rn(fx, fy);
    sC[2 * E * E + i] = __fmul_rn(fy, fy);
  }
  __syncthreads();
  for (int i = t; i < 3 * E * GF_T; i += 256) {
    int c = i / (E * GF_T), rem = i - c * E * GF_T;
    int iy = rem / GF_T, x = rem - iy * GF_T;
    const float* row = sC + c * E * E + iy * E + x;
    double s = 0;
    for (int k = 0; k < bs; ++k) s += (double)row[k];
    sH[i] = s;
  }
  __syncthreads();
  unsigned int local_max = 0;  // ordered key; 0 is below every real float
  const uint8_t* mb = mask ? mask + blockIdx.z * frame_stride : nullptr;
  for (int i = t; i < GF_T * GF_T; i += 256) {
    int y = i / GF_T, x = i - y * GF_T;
    int gx = x0 + x, gy = y0 + y;
    if (gx >= w || gy >= h) continue;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int k = 0; k < bs; ++k) {
      s0 += sH[(y + k) * GF_T + x];
      s1 += sH[E * GF_T + (y + k) * GF_T + x];
      s2 += sH[2 * E * GF_T + (y + k) * GF_T + x];
    }
    float c0 = (float)s0, c1 = (float)s1, c2 = (float)s2, v;
    if (harris) {
      // cv2's calcHarris as its SIMD body rounds it (probed bit for bit): (a c - b b) - k ((a + c)(a + c)), float32
      const float tr = __fadd_rn(c0, c2);
      v = __fsub_rn(__fsub_rn(__fmul_rn(c0, c2), __fmul_rn(c1, c1)), __fmul_rn(hk, __fmul_rn(tr, tr)));
    } else {
      float a = c0 * 0.5f, b = c1, c = c2 * 0.5f;
      float d = __fsub_rn(a, c);
      v = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
    }
    eig[blockIdx.z * (size_t)w * h + (size_t)gy * w + gx] = v;
    if (!mb || mb[(size_t)gy * step + gx]) {
      unsigned int k = f2ord(v);
      local_max = k > local_max ? k : local_max;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned int other = __shfl_xor_sync(0xffffffffu, local_max, o);
    local_max = other > local_max ? other : local_max;
  }
  if ((t & 31) == 0 && local_max) atomicMax(max_key + blockIdx.z, local_max);
}

// K7, shipped form for the 3 x 3 Sobel (the reference's gradientSize) and blockSize 3 / 5 / 7: the same arithmetic as
// gftt_mineig<3> with every loop bound a compile-time constant and the work laid out so that each value is converted
// and loaded once (gftt_mineig<3> spends 80 % of its instructions on index divisions, float -> double conversions
// inside run-time loops and 7 + 7 shared-memory reads per output and channel: 795 us per 16 frames at 1080p).
//   A  products fx fx, fx fy, fy fy of the (32 + BS - 1)^2 halo pixels.  Tiles whose halo and its +-1 neighbours lie
//      inside the frame (all but the outer ring): thread = (halo column, run of five rows), 7 x 3 bytes loaded and
//      converted once for five pixels, a warp's loads on one image row each.  Border tiles: per pixel with reflected
//      coordinates, as in gftt_mineig.
//      (Staging the source tile in shared memory first -- reflected coordinates resolved once per byte, one path for
//      every tile -- was built and measured 25 % slower: byte-wide shared-memory reads of four rows per warp conflict.)
//   B  horizontal BS-sums in double: thread = (channel, halo row, group of four outputs): three 16-byte reads, the
//      first sum direct, the next three by sliding (cv2's own box filter slides in double), two 16-byte writes.
//   C  vertical BS-sums in double + eigenvalue / Harris response: thread = (column, run of four rows), first sum
//      direct, three slides; coalesced stores, masked maximum as an ordered key.
template <int BS>
__global__ void __launch_bounds__(256) gftt_mineig3_fast(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                                          size_t step, size_t frame_stride, int w, int h, float scale,
                                                          int harris, float hk, float* __restrict__ eig,
                                                          unsigned int* __restrict__ max_key) {
  constexpr int R = BS / 2, E = GF_T + BS - 1, EP = (E + 3) & ~3;      // halo edge; product row pitch (16-byte rows)
  __shared__ __align__(16) float sP[3][E][EP];
  __shared__ __align__(16) double sH[3][E][GF_T];
  const int x0 = blockIdx.x * GF_T, y0 = blockIdx.y * GF_T;
  const uint8_t* ib = img + blockIdx.z * frame_stride;
  const int t = threadIdx.x;
  const float s2 = 2.f * scale;
  // cv2.Sobel(CV_32F, scale) as opencv's separable filter rounds it in its SIMD body (probed bit for bit against
  // cv2 4.13, oracle/gftt.py::sobel3_f32): the scaled smoothing taps [s, 2s, s] go through fused multiply-adds
  auto products = [&](float a0, float a1, float a2, float b0, float b2, float c0, float c1, float c2, int iy, int ix) {
    // rows a (above), b (this), c (below); columns 0 (left), 1, 2 (right)
    const float t0 = a2 - a0, t1 = b2 - b0, t2 = c2 - c0;              // small integers: exact
    const float fx = __fmaf_rn(t0 + t2, scale, __fmul_rn(t1, s2));
    const float ra = __fmaf_rn(a2, scale, __fmaf_rn(a1, s2, __fmul_rn(a0, scale)));
    const float rb = __fmaf_rn(c2, scale, __fmaf_rn(c1, s2, __fmul_rn(c0, scale)));
    const float fy = __fsub_rn(rb, ra);
    sP[0][iy][ix] = __fmul_rn(fx, fx);
    sP[1][iy][ix] = __fmul_rn(fx, fy);
    sP[2][iy][ix] = __fmul_rn(fy, fy);
  };
  const bool interior = x0 - R - 1 >= 0 && y0 - R - 1 >= 0 && x0 - R + E < w && y0 - R + E < h;
  if (interior) {
    // thread = (halo column, run of five rows): the lanes of a warp read consecutive bytes of ONE image row per load
    // (five-pixel row runs touched four rows per load: four L1 wavefronts each); the row's horizontal difference and
    // smoothed value are computed once and serve the three pixel rows that use them
    constexpr int RUN = 5, NRUN = (E + RUN - 1) / RUN;
    for (int i = t; i < E * NRUN; i += 256) {
      const int q = i / E, ix = i - q * E;
      const int iy0 = q * RUN;
      const uint8_t* p = ib + (size_t)(y0 - R + iy0 - 1) * step + (x0 - R + ix - 1);
      float td[RUN + 2], hs[RUN + 2];
#pragma unroll
      for (int k = 0; k < RUN + 2; ++k) {
        const bool in = iy0 + k - 1 <= E;                              // (the last run of a column is shorter)
        const float l = in ? (float)p[k * step] : 0.f, c = in ? (float)p[k * step + 1] : 0.f,
                    r = in ? (float)p[k * step + 2] : 0.f;
        td[k] = r - l;                                                 // small integers: exact
        hs[k] = __fmaf_rn(r, scale, __fmaf_rn(c, s2, __fmul_rn(l, scale)));
      }
#pragma unroll
      for (int k = 0; k < RUN; ++k) {
        if (iy0 + k >= E) break;
        const float fx = __fmaf_rn(td[k] + td[k + 2], scale, __fmul_rn(td[k + 1], s2));
        const float fy = __fsub_rn(hs[k + 2], hs[k]);
        sP[0][iy0 + k][ix] = __fmul_rn(fx, fx);
        sP[1][iy0 + k][ix] = __fmul_rn(fx, fy);
        sP[2][iy0 + k][ix] = __fmul_rn(fy, fy);
      }
    }
  } else {
    for (int i = t; i < E * E; i += 256) {
      const int iy = i / E, ix = i - iy * E;
      const int x = reflect101(x0 - R + ix, w), y = reflect101(y0 - R + iy, h);
      const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      const uint8_t* r0 = ib + (size_t)reflect101(y - 1, h) * step;
      const uint8_t* r1 = ib + (size_t)y * step;
      const uint8_t* r2 = ib + (size_t)reflect101(y + 1, h) * step;
      products((float)r0[xm], (float)r0[x], (float)r0[xp], (float)r1[xm], (float)r1[xp], (float)r2[xm], (float)r2[x],
               (float)r2[xp], iy, ix);
    }
  }
  __syncthreads();
  for (int i = t; i < 3 * E * (GF_T / 4); i += 256) {
    const int c = i / (E * (GF_T / 4)), rem = i - c * (E * (GF_T / 4));
    const int iy = rem / (GF_T / 4), g = rem - iy * (GF_T / 4);
    const float* row = &sP[c][iy][4 * g];
    constexpr int NV = (BS + 3 + 3) / 4;                               // float4 reads covering BS + 3 values
    double d[4 * NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 v = *(const float4*)(row + 4 * k);
      d[4 * k] = (double)v.x; d[4 * k + 1] = (double)v.y; d[4 * k + 2] = (double)v.z; d[4 * k + 3] = (double)v.w;
    }
    double o0 = d[0];
#pragma unroll
    for (int k = 1; k < BS; ++k) o0 += d[k];
    const double o1 = o0 - d[0] + d[BS], o2 = o1 - d[1] + d[BS + 1], o3 = o2 - d[2] + d[BS + 2];
    // row layout of sH: outputs 4g, 4g + 1 at [2g, 2g + 1]; outputs 4g + 2, 4g + 3 at 16 + ((2g + 8) mod 16): a thread's
    // two 16-byte stores are contiguous across a quarter warp, and the sixteen doubles a half warp of step C reads
    // fall on sixteen different bank pairs (the plain layout [4g .. 4g + 3] costs the stores a two-way conflict)
    *(double2*)&sH[c][iy][2 * g] = make_double2(o0, o1);
    *(double2*)&sH[c][iy][GF_T / 2 + ((2 * g + 8) & 15)] = make_double2(o2, o3);
  }
  __syncthreads();
  unsigned int local_max = 0;  // ordered key; 0 is below every real float
  const uint8_t* mb = mask ? mask + blockIdx.z * frame_stride : nullptr;
  {
    const int x = t & 31, yq = (t >> 5) * 4;
    const int xs = (x & 2) ? GF_T / 2 + ((2 * (x >> 2) + 8) & 15) + (x & 1) : 2 * (x >> 2) + (x & 1);   // step B's layout
    double sum[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double hv[BS + 3];
#pragma unroll
      for (int k = 0; k < BS + 3; ++k) hv[k] = sH[c][yq + k][xs];
      double o = hv[0];
#pragma unroll
      for (int k = 1; k < BS; ++k) o += hv[k];
      sum[c][0] = o;
#pragma unroll
      for (int j = 1; j < 4; ++j) { o = o - hv[j - 1] + hv[j + BS - 1]; sum[c][j] = o; }
    }
    const int gx = x0 + x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gy = y0 + yq + j;
      if (gx >= w || gy >= h) continue;
      const float c0 = (float)sum[0][j], c1 = (float)sum[1][j], c2 = (float)sum[2][j];
      float v;
      if (harris) {
        // cv2's calcHarris as its SIMD body rounds it (probed bit for bit): (a c - b b) - k ((a + c)(a + c)), float32
        const float tr = __fadd_rn(c0, c2);
        v = __fsub_rn(__fsub_rn(__fmul_rn(c0, c2), __fmul_rn(c1, c1)), __fmul_rn(hk, __fmul_rn(tr, tr)));
      } else {
        const float a = c0 * 0.5f, b = c1, c = c2 * 0.5f;
        const float d = __fsub_rn(a, c);
        v = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
      }
      eig[blockIdx.z * (size_t)w * h + (size_t)gy * w + gx] = v;
      if (!mb || mb[(size_t)gy * step + gx]) {
        const unsigned int k = f2ord(v);
        local_max = k > local_max ? k : local_max;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned int other = __shfl_xor_sync(0xffffffffu, local_max, o);
    local_max = other > local_max ? other : local_max;
  }
  if ((t & 31) == 0 && local_max) atomicMax(max_key + blockIdx.z, local_max);
}

// thread = four adjacent pixels of a row (one 16-byte load of the score map when rows are 16-byte aligned): almost
// every pixel fails the quality threshold, so the kernel is a stream over the map -- one pixel per thread left a
// 4-byte load per thread in flight and ran at 1.1 TB/s
__global__ void __launch_bounds__(256) gftt_nms_compact(const float* __restrict__ eig,
                                                         const uint8_t* __restrict__ mask, size_t step,
                                                         size_t frame_stride, int w, int h, double quality,
                                                         const unsigned int* __restrict__ max_key,
                                                         unsigned long long* __restrict__ cand, int cand_cap,
                                                         int* __restrict__ cand_count) {
  const int lane = threadIdx.x & 31;
  const int x4 = (blockIdx.x * 32 + lane) * 4;
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  const float* e = eig + (size_t)b * w * h;
  float v4[4] = {0.f, 0.f, 0.f, 0.f};
  float thr = 0.f;
  const bool row_ok = y >= 1 && y < h - 1 && x4 < w;
  if (row_ok) {
    const unsigned int mk = max_key[b];
    const float max_val = mk ? ord2f(mk) : 0.f;
    thr = (float)((double)max_val * quality);
    const float* r = e + (size_t)y * w + x4;
    if ((w & 3) == 0) {
      const float4 f = __ldg((const float4*)r);
      v4[0] = f.x; v4[1] = f.y; v4[2] = f.z; v4[3] = f.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x4 + k < w) v4[k] = r[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    const float v = v4[k];
    bool is_cand = false;
    if (row_ok && x >= 1 && x < w - 1 && v > thr && v != 0.f &&
        (!mask || mask[b * frame_stride + (size_t)y * step + x])) {
      is_cand = true;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
          if (e[(size_t)(y + dy) * w + x + dx] > v) is_cand = false;
    }
    const unsigned int ballot = __ballot_sync(0xffffffffu, is_cand);
    if (ballot) {
      int base = 0;
      if (lane == 0) base = atomicAdd(cand_count + b, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (is_cand) {
        const int slot = base + __popc(ballot & ((1u << lane) - 1));
        if (slot < cand_cap)
          cand[(size_t)b * cand_cap + slot] = ((unsigned long long)f2ord(v) << 32) | (unsigned int)(y * w + x);
      }
    }
  }
}

// sort (value desc, address desc) + greedy min-distance pass of one frame's candidates.  `keys`, `head`, `nxt` all
// live in the same memory (shared or global): the two instantiations are inlined with their address space known.
__device__ __forceinline__ void gftt_select_body(unsigned long long* keys, int* head, int* nxt, int n, int np2, int w,
                                                 int h, int max_corners, int cell, double md2, float* out,
                                                 int corners_cap, int* n_out) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        int l = i ^ j;
        if (l > i) {
          unsigned long long a = keys[i], c = keys[l];
          bool desc = (i & k) == 0;
          if (desc ? a < c : a > c) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  // cell == 0: cv2's minDistance < 1 branch (no distance test).  The host decides with cv2's own double
  // arithmetic (minDistance >= 1, cvRound) and passes the cell size and the squared distance down.
  if (cell == 0) {
    int total = (max_corners > 0 && n > max_corners) ? max_corners : n;
    for (int i = threadIdx.x; i < total && i < corners_cap; i += blockDim.x) {
      unsigned int idx = (unsigned int)(keys[i] & 0xffffffffull);
      out[2 * i] = (float)(idx % w);
      out[2 * i + 1] = (float)(idx / w);
    }
    if (threadIdx.x == 0) *n_out = total;
    return;
  }
  const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
  for (int i = threadIdx.x; i < gw * gh; i += blockDim.x) head[i] = -1;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int accepted = 0;
  for (int i = 0; i < n; ++i) {
    unsigned int idx = (unsigned int)(keys[i] & 0xffffffffull);
    int x = idx % w, y = idx / w;
    int cx = x / cell, cy = y / cell;
    bool bad = false;
    if (lane < 9) {
      int xx = cx + lane % 3 - 1, yy = cy + lane / 3 - 1;
      if (xx >= 0 && xx < gw && yy >= 0 && yy < gh) {
        for (int q = head[yy * gw + xx]; q >= 0; q = nxt[q]) {
          unsigned int qi = (unsigned int)(keys[q] & 0xffffffffull);
          float dx = (float)(x - (int)(qi % w)), dy = (float)(y - (int)(qi / w));
          if ((double)(dx * dx + dy * dy) < md2) { bad = true; break; }
        }
      }
    }
    if (__any_sync(0xffffffffu, bad)) continue;
    if (lane == 0) {
      nxt[i] = head[cy * gw + cx];
      head[cy * gw + cx] = i;
      if (accepted < corners_cap) { out[2 * accepted] = (float)x; out[2 * accepted + 1] = (float)y; }
    }
    __syncwarp();
    ++accepted;
    if (max_corners > 0 && accepted == max_corners) break;
  }
  if (lane == 0) *n_out = accepted;
}

// One CTA per frame.  The sort passes and the sequential greedy pass are chains of dependent accesses: when the
// frame's candidates (and the cell grid) fit the kernel's shared memory -- smem_keys keys + links, smem_cells heads --
// they are sorted and linked there; a frame with more candidates works in the global workspace.
__global__ void __launch_bounds__(1024) gftt_select(unsigned long long* __restrict__ cand, int cand_cap,
                                                     const int* __restrict__ cand_count, int w, int h,
                                                     int max_corners, int cell, double md2,
                                                     int* __restrict__ cell_head,
                                                     int* __restrict__ next_in_cell, float* __restrict__ corners,
                                                     int corners_cap, int* __restrict__ n_corners, int smem_keys,
                                                     int smem_cells) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  const int b = blockIdx.x;
  unsigned long long* keys = cand + (size_t)b * cand_cap;
  int n = cand_count[b];
  if (n > cand_cap) n = cand_cap;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  if (np2 > cand_cap) np2 = cand_cap;  // cand_cap is a power of two
  float* out = corners + (size_t)b * corners_cap * 2;
  const int cells = cell > 0 ? ((w + cell - 1) / cell) * ((h + cell - 1) / cell) : 0;
  unsigned long long* skeys = (unsigned long long*)sel_smem;
  int* snxt = (int*)(skeys + smem_keys);
  int* shead = snxt + smem_keys;
  if (np2 <= smem_keys && cells <= smem_cells) {          // (uniform over the CTA)
    for (int i = threadIdx.x; i < np2; i += blockDim.x) skeys[i] = i < n ? keys[i] : 0ull;
    __syncthreads();
    gftt_select_body(skeys, shead, snxt, n, np2, w, h, max_corners, cell, md2, out, corners_cap, n_corners + b);
    return;
  }
  if (smem_keys > 0 && cells <= smem_cells && max_corners > 0 && cell > 0) {
    // More candidates than shared memory holds, and a bounded number of corners wanted: the greedy pass walks the
    // candidates in descending order and stops at max_corners, so it almost always needs only the strongest few.
    // Take the candidates of the top bins of a 4096-bin histogram of the score's leading bits -- as many whole bins
    // as fit -- sort those in shared memory and run the pass on them: they are exactly a prefix of the full order.
    // If the pass runs out of them before max_corners are accepted, the general path below redoes the frame.
    int* hist = shead + smem_cells;                        // 4096 bins + the compaction counter
    __shared__ int s_thr_bin, s_m;
    for (int i = threadIdx.x; i <= 4096; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[(int)(keys[i] >> 52)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, tb = 4096;
      for (int bin = 4095; bin >= 0; --bin) {
        if (cum + hist[bin] > smem_keys) break;
        cum += hist[bin]; tb = bin;
      }
      s_thr_bin = tb; s_m = cum;
    }
    __syncthreads();
    const int tb = s_thr_bin, m = s_m;
    if (m >= max_corners && m < n) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = keys[i];
        if ((int)(k >> 52) >= tb) skeys[atomicAdd(&hist[4096], 1)] = k;
      }
      int mp2 = 1;
      while (mp2 < m) mp2 <<= 1;
      __syncthreads();
      for (int i = m + threadIdx.x; i < mp2; i += blockDim.x) skeys[i] = 0ull;
      __syncthreads();
      gftt_select_body(skeys, shead, snxt, m, mp2, w, h, max_corners, cell, md2, out, corners_cap, n_corners + b);
      __syncthreads();                                     // (threads >= 32 left the body early; all meet here)
      if (n_corners[b] == max_corners) return;             // written by lane 0 before the barrier above
    }
    __syncthreads();
  }
  {
    for (int i = n + threadIdx.x; i < np2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    const int gw = cell > 0 ? (w + cell - 1) / cell : 0, gh = cell > 0 ? (h + cell - 1) / cell : 0;
    gftt_select_body(keys, cell_head + (size_t)b * gw * gh, next_in_cell + (size_t)b * cand_cap, n, np2, w, h,
                     max_corners, cell, md2, out, corners_cap, n_corners + b);
  }
}

static int gftt_check(int rows, int cols, const b2of_gftt_params* p) {
  const char* fn = "goodFeaturesToTrack";
  B2OF_ASSERT(p != nullptr, fn);
  B2OF_ASSERT(rows > 0 && cols > 0, fn);
  B2OF_ASSERT(p->quality_level > 0 && p->min_distance >= 0 && p->max_corners >= 0, fn);
  B2OF_ASSERT(p->block_size >= 1, fn);
  if (p->gradient_size != 3 && p->gradient_size != 5 && p->gradient_size != 7)
    return fail(B2OF_E_BADARG, "(-215:Assertion failed) ksize == 3 || ksize == 5 || ksize == 7 in function 'Sobel'");
  if (p->block_size > 31) return fail(B2OF_E_UNSUPPORTED, "blockSize > 31 is not supported");
  return B2OF_OK;
}

struct GfLayout {
  float* eig; unsigned long long* cand; int* next_in_cell; int* cell_head; unsigned int* max_key; int* cand_count;
  int cand_cap; int cell; size_t cells; size_t bytes;
};

static void gf_layout(int rows, int cols, const b2of_gftt_params* p, int batch, void* base, size_t cap, GfLayout* L) {
  Arena ar(base, cap);
  size_t n = (size_t)rows * cols;
  int cc = 1024;
  while ((size_t)cc < n / 2 + 1) cc <<= 1;
  L->cand_cap = cc;
  int cell = p->min_distance >= 1 ? cv_round(p->min_distance) : 1;
  if (cell < 1) cell = 1;
  L->cell = p->min_distance >= 1 ? cell : 0;
  L->cells = (size_t)cdiv(cols, cell) * cdiv(rows, cell);
  L->eig = ar.take<float>(n * batch);
  L->cand = ar.take<unsigned long long>((size_t)cc * batch);
  L->next_in_cell = ar.take<int>((size_t)cc * batch);
  L->cell_head = ar.take<int>(p->min_distance >= 1 ? L->cells * batch : 1);
  L->max_key = ar.take<unsigned int>(batch);
  L->cand_count = ar.take<int>(batch);
  L->bytes = align_up(ar.off, 256);
}

size_t gftt_workspace_bytes(int rows, int cols, const b2of_gftt_params* p, int batch) {
  if (gftt_check(rows, cols, p)) return 0;
  if (batch < 1) batch = 1;
  GfLayout L;
  gf_layout(rows, cols, p, batch, nullptr, 0, &L);
  return L.bytes;
}

int gftt_dev(const uint8_t* img, const uint8_t* mask, size_t step, size_t frame_stride, int batch, int rows, int cols,
             const b2of_gftt_params* p, float* corners, int cap, int* n_corners, void* ws, size_t ws_bytes,
             cudaStream_t st) {
  int rc = gftt_check(rows, cols, p);
  if (rc) return rc;
  const char* fn = "goodFeaturesToTrack";
  B2OF_ASSERT(img != nullptr && n_corners != nullptr && step >= (size_t)cols, fn);
  B2OF_ASSERT(cap >= 0 && (corners != nullptr || cap == 0), fn);
  if (batch <= 0) return B2OF_OK;
  GfLayout L;
  gf_layout(rows, cols, p, batch, ws, ws_bytes, &L);
  if (ws == nullptr || ws_bytes < L.bytes)
    return fail(B2OF_E_NOMEM, "gftt workspace too small: %zu B given, %zu B needed", ws_bytes, L.bytes);
  B2OF_CUDA(cudaMemsetAsync(L.max_key, 0, sizeof(unsigned int) * batch, st));
  B2OF_CUDA(cudaMemsetAsync(L.cand_count, 0, sizeof(int) * batch, st));
  const int bs = p->block_size;
  const int E = GF_T + bs - 1;
  const int ks = p->gradient_size;
  size_t smem = (size_t)((3 * E * E + 1) & ~1) * sizeof(float) + (size_t)3 * E * GF_T * sizeof(double);
  if (ks != 3) smem += (size_t)2 * (E + ks - 1) * E * sizeof(float) + (size_t)(E + ks - 1) * (E + ks - 1) + 16;
  static PerDeviceMax max_set[3];
  if (smem > 48 * 1024 && max_set[ks == 3 ? 0 : ks == 5 ? 1 : 2].raise(smem)) {
    if (ks == 3) B2OF_CUDA(cudaFuncSetAttribute(gftt_mineig<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (ks == 5) B2OF_CUDA(cudaFuncSetAttribute(gftt_mineig<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else B2OF_CUDA(cudaFuncSetAttribute(gftt_mineig<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  float scale = (float)(1.0 / ((double)(1 << (p->gradient_size - 1)) * bs * 255.0));
  dim3 g1(cdiv(cols, GF_T), cdiv(rows, GF_T), batch);
  const bool fast_ok = ks == 3 && cols >= 16 && rows >= 16;
  if (fast_ok && bs == 3)
    gftt_mineig3_fast<3><<<g1, 256, 0, st>>>(img, mask, step, frame_stride, cols, rows, scale, p->use_harris,
                                             (float)p->k, L.eig, L.max_key);
  else if (fast_ok && bs == 5)
    gftt_mineig3_fast<5><<<g1, 256, 0, st>>>(img, mask, step, frame_stride, cols, rows, scale, p->use_harris,
                                             (float)p->k, L.eig, L.max_key);
  else if (fast_ok && bs == 7)
    gftt_mineig3_fast<7><<<g1, 256, 0, st>>>(img, mask, step, frame_stride, cols, rows, scale, p->use_harris,
                                             (float)p->k, L.eig, L.max_key);
  else if (ks == 3)
    gftt_mineig<3><<<g1, 256, smem, st>>>(img, mask, step, frame_stride, cols, rows, bs, scale, p->use_harris,
                                          (float)p->k, L.eig, L.max_key);
  else if (ks == 5)
    gftt_mineig<5><<<g1, 256, smem, st>>>(img, mask, step, frame_stride, cols, rows, bs, scale, p->use_harris,
                                          (float)p->k, L.eig, L.max_key);
  else
    gftt_mineig<7><<<g1, 256, smem, st>>>(img, mask, step, frame_stride, cols, rows, bs, scale, p->use_harris,
                                          (float)p->k, L.eig, L.max_key);
  B2OF_LAUNCH_CHECK();
  dim3 g2(cdiv(cols, 128), cdiv(rows, 8), batch);
  gftt_nms_compact<<<g2, 256, 0, st>>>(L.eig, mask, step, frame_stride, cols, rows, p->quality_level, L.max_key, L.cand,
                                       L.cand_cap, L.cand_count);
  B2OF_LAUNCH_CHECK();
  // shared memory of the selection kernel: up to 4096 keys + links (48 KB) and the cell grid when it fits beside them
#ifdef B2OF_SEL_NOSMEM                                     // developer A/B: everything in the global workspace
  const int sel_keys = 0;
#else
  const int sel_keys = L.cand_cap < 4096 ? L.cand_cap : 4096;
#endif
  const long long cells = L.cell > 0 ? (long long)cdiv(cols, L.cell) * cdiv(rows, L.cell) : 0;
  const size_t sel_hist = sel_keys > 0 ? (size_t)4097 * 4 : 0;          // histogram of the prefix selection
  const int sel_cells = (size_t)sel_keys * 12 + (size_t)cells * 4 + sel_hist <= 200 * 1024 ? (int)cells : 0;
  const size_t sel_smem = (size_t)sel_keys * 12 + (size_t)sel_cells * 4 + sel_hist;
  static PerDeviceMax sel_max;
  if (sel_smem > 48 * 1024 && sel_max.raise(sel_smem))
    B2OF_CUDA(cudaFuncSetAttribute(gftt_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  gftt_select<<<batch, 1024, sel_smem, st>>>(L.cand, L.cand_cap, L.cand_count, cols, rows, p->max_corners, L.cell,
                                             p->min_distance * p->min_distance, L.cell_head, L.next_in_cell, corners,
                                             cap, n_corners, sel_keys, sel_cells);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

}  // namespace b2of
