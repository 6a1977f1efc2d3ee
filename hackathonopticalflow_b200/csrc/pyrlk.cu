// K10-K11: pyramidal Lucas-Kanade, replaces cv2.calcOpticalFlowPyrLK as called at pathfinder_viewer.py:156-158 /
// DenseOF.py:183-185 (45x45 grid form, prev = current frame, next = previous frame) and SparseOF.py:35-36
// (15x15 track form).  Arithmetic spec: SURVEY.md App. A.4 (opencv lkpyramid.cpp -- third-party, restated in
// oracle/pyrlk.py): u8 pyrDown chain, int16 Scharr derivatives, 14-bit fixed-point bilinear patches, float32
// 2x2 normal equations, <= maxCount Newton steps.
//
// LK_SPLIT warps per feature (each owns a share of the window's rows, of the template patch and of every window sum),
// all pyramid levels inside one launch; the template patch (Iw, Ixw, Iyw as int16) lives in shared memory, window sums
// are exact 64-bit integers reduced with warp shuffles (and, for LK_SPLIT > 1, one shared-memory exchange between the
// feature's warps), so the result does not depend on the split.  Shipped: one warp per feature.  Two warps per feature
// -- twice the resident warps on the same 12 KB patches -- was measured 5-9 % SLOWER (profiles/README.md): half
// windows of 22 rows leave the four-rows-per-step walk less to overlap and every Newton step pays a named barrier.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b2of {

int pyrdown_dev(const uint8_t*, int, int, size_t, size_t, uint8_t*, size_t, size_t, int, cudaStream_t);

constexpr int LK_MAX_LEVELS = 12;
#ifndef B2OF_LK_WARPS
#define B2OF_LK_WARPS 4
#endif
constexpr int LK_WARPS = B2OF_LK_WARPS;          // warps per CTA
#ifndef B2OF_LK_SPLIT
#define B2OF_LK_SPLIT 1
#endif
constexpr int LK_SPLIT = B2OF_LK_SPLIT;   // warps per feature
constexpr int LK_FEATS = LK_WARPS / LK_SPLIT;   // features per CTA

struct LkLevels {
  int n;  // number of levels (effective maxLevel + 1)
  int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
  size_t step[LK_MAX_LEVELS];
  size_t off[LK_MAX_LEVELS];         // byte offset of level l inside one image's pyramid block
  size_t doff[LK_MAX_LEVELS];        // short2 offset of level l inside one image's derivative block
  size_t pyr_bytes, deriv_elems;     // per image
};

static int lk_plan(int rows, int cols, const b2of_lk_params* p, LkLevels* L) {
  memset(L, 0, sizeof *L);
  int w = cols, h = rows, n = 0;
  size_t off = 0, doff = 0;
  for (int level = 0; level <= p->max_level && level < LK_MAX_LEVELS; ++level) {
    L->w[n] = w; L->h[n] = h;
    L->step[n] = align_up(w, 16);
    L->off[n] = off; L->doff[n] = doff;
    off += align_up(L->step[n] * h, 256);
    doff += align_up((size_t)w * h, 64);
    ++n;
    w = (w + 1) / 2; h = (h + 1) / 2;
    if (w <= p->win_w || h <= p->win_h) break;
  }
  L->n = n;
  L->pyr_bytes = off;
  L->deriv_elems = doff;
  return B2OF_OK;
}

// ---- K10: Scharr 3x3 -> (Ix, Iy) int16, REFLECT_101 (calcScharrDeriv) ----
__device__ __forceinline__ int reflect101_pm1(int i, int n) {   // BORDER_REFLECT_101 for i in [-1, n]
  if (n == 1) return 0;
  return i < 0 ? 1 : (i >= n ? n - 2 : i);
}

// four pixels per thread: the 3 x 6 neighbourhood comes in as bytes from L1, the four (Ix, Iy) pairs go out as one
// 16-byte store when the row start is aligned
__global__ void __launch_bounds__(256) lk_scharr(const uint8_t* __restrict__ img, size_t step, size_t img_bstride,
                                                  int w, int h, short2* __restrict__ d, size_t d_bstride) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x0 >= w) return;
  const uint8_t* b = img + blockIdx.z * img_bstride;
  const uint8_t* r0 = b + (size_t)reflect101_pm1(y - 1, h) * step;
  const uint8_t* r1 = b + (size_t)y * step;
  const uint8_t* r2 = b + (size_t)reflect101_pm1(y + 1, h) * step;
  int t0[6], t1[6];                      // vertical smooth (3, 10, 3) and difference at columns x0 - 1 .. x0 + 4
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int xx = reflect101_pm1(min(x0 - 1 + k, w), w);
    const int a = r0[xx], c = r1[xx], e = r2[xx];
    t0[k] = (a + e) * 3 + c * 10;
    t1[k] = e - a;
  }
  short2 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    o[k] = make_short2((short)(t0[k + 2] - t0[k]), (short)((t1[k] + t1[k + 2]) * 3 + t1[k + 1] * 10));
  short2* out = d + blockIdx.z * d_bstride + (size_t)y * w + x0;
  if (x0 + 3 < w && (((size_t)out) & 15) == 0) {
    *(int4*)out = make_int4(*(int*)&o[0], *(int*)&o[1], *(int*)&o[2], *(int*)&o[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x0 + k < w) out[k] = o[k];
  }
}

struct LkArgs {
  const uint8_t* pyr_i;   // [batch] pyramid blocks of the first image
  const uint8_t* pyr_j;   // [batch] pyramid blocks of the second image
  const short2* deriv;    // [batch] derivative blocks of the first image
  LkLevels L;
  const float* prev_pts; size_t pts_bstride;  // in points
  float* next_pts; uint8_t* status; float* err;
  int n_pts;
  int ww, wh;
  int max_count; float eps2d_hi, eps2d_lo;  // eps^2 as a double split (hi + lo)
  int flags; float min_eig_thr;
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __shfl_xor_sync(0xffffffffu, (int)(v & 0xffffffffll), o);
    int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), o);
    v += ((long long)hi << 32) | (unsigned int)lo;
  }
  return v;
}

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
  // cvRound((1-a)(1-b)*2^14) ...: float products rounded half-to-even
  w00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
  w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
  w10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
  w11 = 16384 - w00 - w01 - w10;
}

// Patch layout in shared memory: rows of wwp = ww rounded up to even elements, so that two horizontally adjacent
// elements are one aligned 4-byte (Iw pair) / 8-byte (derivative pair) access.
__host__ __device__ inline int lk_wwp(int ww) { return (ww + 1) & ~1; }
__host__ __device__ inline size_t lk_patch_bytes(int ww, int wh) {
  size_t n = (size_t)lk_wwp(ww) * wh;
  return ((((n + 3) & ~(size_t)3) * 2 + n * 4) + 15) & ~(size_t)15;
}

// Work split of a window inside the image: the warp's lanes own strips two columns wide (S = wwp / 2 strips) and,
// when the window is narrow, G = 32 / S groups of rows.  A lane walks down its strip: the three bytes of the next
// image row are loaded once and serve as bottom neighbours of this row and top neighbours of the next one
// (1.5 loads per element instead of 4), addresses advance by one row step, and the template patch comes in as
// aligned pairs.  Window sums are exact 64-bit integers, so the result does not depend on the split.
struct LkStrips {
  int S, G, RG;
  __device__ __forceinline__ LkStrips(int ww, int nrows) {   // nrows = rows of this warp's share of the window
    S = lk_wwp(ww) >> 1;
    G = S <= 16 ? 32 / S : 1;
    RG = (nrows + G - 1) / G;
  }
};

// sum over the window of (bilinear(J) >> 9 - Iw) * {Ixw, Iyw}  (or |diff| when ABS).
// Windows that stick out of the image read BORDER_REFLECT_101 pixels: the three reflected column indices of a
// strip are computed once, the reflected row index once per row.
template <bool ABS, bool INSIDE>
__device__ __forceinline__ void lk_window_strips(const uint8_t* __restrict__ J, size_t step, int w, int h, int ix,
                                                 int iy, int w00, int w01, int w10, int w11, const short* sI,
                                                 const short2* sD, int ww, int ya, int yb, int lane, long long& s1,
                                                 long long& s2) {
  const int wwp = lk_wwp(ww);
  const LkStrips st(ww, yb - ya);
  for (int u = lane; u < st.S * st.G; u += 32) {
    const int g = u / st.S, sx = u - g * st.S;
    const int x0 = 2 * sx;
    const bool two = x0 + 1 < ww;
    const int y0 = ya + g * st.RG, y1 = min(yb, y0 + st.RG);
    int c0 = ix + x0, c1 = c0 + 1, c2 = c0 + 2;
    if (!INSIDE) { c0 = reflect101(c0, w); c1 = reflect101(c1, w); c2 = reflect101(c2, w); }
    auto rowp = [&](int y) { return J + (size_t)(INSIDE ? iy + y : reflect101(iy + y, h)) * step; };
    const uint8_t* r = rowp(y0);
    int t0 = r[c0], t1 = r[c1], t2 = two ? r[c2] : 0;
    const short* pI = sI + y0 * wwp + x0;
    const short2* pD = sD + y0 * wwp + x0;
    auto one_row = [&](int b0, int b1, int b2, int ipair, int2 dp) {
      const int v0 = t0 * w00 + t1 * w01 + b0 * w10 + b1 * w11;
      const int v1 = t1 * w00 + t2 * w01 + b1 * w10 + b2 * w11;
      const int diff0 = ((v0 + 256) >> 9) - (int)(short)(ipair & 0xffff);
      const int diff1 = ((v1 + 256) >> 9) - (ipair >> 16);
      if (ABS) {
        s1 += abs(diff0);
        if (two) s1 += abs(diff1);
      } else {
        s1 += (long long)diff0 * (int)(short)(dp.x & 0xffff);
        s2 += (long long)diff0 * (dp.x >> 16);
        if (two) {
          s1 += (long long)diff1 * (int)(short)(dp.y & 0xffff);
          s2 += (long long)diff1 * (dp.y >> 16);
        }
      }
      t0 = b0; t1 = b1; t2 = b2;
    };
    int y = y0;
    // four rows per step, every load of the step issued before the first use
    for (; y + 4 <= y1; y += 4) {
      int b[4][3], ip[4];
      int2 dp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint8_t* rk = INSIDE ? r + (size_t)(k + 1) * step : rowp(y + k + 1);
        b[k][0] = rk[c0]; b[k][1] = rk[c1]; b[k][2] = two ? rk[c2] : 0;
        ip[k] = *(const int*)(pI + k * wwp);
        dp[k] = ABS ? make_int2(0, 0) : *(const int2*)(pD + k * wwp);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) one_row(b[k][0], b[k][1], b[k][2], ip[k], dp[k]);
      if (INSIDE) r += 4 * step;
      pI += 4 * wwp; pD += 4 * wwp;
    }
    for (; y < y1; ++y) {
      const uint8_t* rk = INSIDE ? r + step : rowp(y + 1);
      if (INSIDE) r = rk;
      one_row(rk[c0], rk[c1], two ? rk[c2] : 0, *(const int*)pI, ABS ? make_int2(0, 0) : *(const int2*)pD);
      pI += wwp; pD += wwp;
    }
  }
}

template <bool ABS>
__device__ __forceinline__ void lk_window_pass(const uint8_t* __restrict__ J, size_t step, int w, int h, int ix,
                                               int iy, int w00, int w01, int w10, int w11, const short* sI,
                                               const short2* sD, int ww, int wh, int ya, int yb, int lane,
                                               long long& o1, long long& o2) {
  const bool inside = ix >= 0 && iy >= 0 && ix + ww < w && iy + wh < h;
  long long s1 = 0, s2 = 0;
  if (inside) lk_window_strips<ABS, true>(J, step, w, h, ix, iy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, s1, s2);
  else lk_window_strips<ABS, false>(J, step, w, h, ix, iy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, s1, s2);
  o1 = warp_sum_ll(s1);
  o2 = ABS ? 0 : warp_sum_ll(s2);
}

// template patch (Iw, Ixw, Iyw) of the window at (ipx, ipy) + the three sums of the normal matrix.  Image pixels
// outside the frame are BORDER_REFLECT_101, derivatives outside the frame are zero.
template <bool INSIDE>
__device__ __forceinline__ void lk_patch_strips(const uint8_t* __restrict__ I, const short2* __restrict__ D, size_t step,
                                                int W, int H, int ipx, int ipy, int w00, int w01, int w10, int w11,
                                                short* sI, short2* sD, int ww, int ya, int yb, int lane,
                                                long long& sA11, long long& sA12, long long& sA22) {
  const int wwp = lk_wwp(ww);
  const LkStrips st(ww, yb - ya);
  for (int u = lane; u < st.S * st.G; u += 32) {
    const int g = u / st.S, sx = u - g * st.S;
    const int x0 = 2 * sx;
    const bool two = x0 + 1 < ww;
    const int y0 = ya + g * st.RG, y1 = min(yb, y0 + st.RG);
    const int a0 = ipx + x0;                     // unreflected columns a0, a0 + 1, a0 + 2
    int c0 = a0, c1 = a0 + 1, c2 = a0 + 2;
    bool v0 = true, v1 = true, v2 = two;         // derivative columns inside the frame
    if (!INSIDE) {
      v0 = (unsigned)c0 < (unsigned)W; v1 = (unsigned)c1 < (unsigned)W; v2 = two && (unsigned)c2 < (unsigned)W;
      c0 = reflect101(c0, W); c1 = reflect101(c1, W); c2 = reflect101(c2, W);
    }
    const short2 z = make_short2(0, 0);
    auto load_row = [&](int y, int& i0, int& i1, int& i2, short2& d0, short2& d1, short2& d2) {
      const int ya = ipy + y;
      const uint8_t* r = I + (size_t)(INSIDE ? ya : reflect101(ya, H)) * step;
      i0 = r[c0]; i1 = r[c1]; i2 = two ? r[c2] : 0;
      const bool vy = INSIDE || (unsigned)ya < (unsigned)H;
      const short2* dr = D + (size_t)(vy ? ya : 0) * W;
      d0 = (vy && v0) ? dr[a0] : z;
      d1 = (vy && v1) ? dr[a0 + 1] : z;
      d2 = (vy && v2) ? dr[a0 + 2] : z;
    };
    int t0, t1, t2;
    short2 e0, e1, e2;
    load_row(y0, t0, t1, t2, e0, e1, e2);
    short* pI = sI + y0 * wwp + x0;
    short2* pD = sD + y0 * wwp + x0;
#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
      int b0, b1, b2;
      short2 f0, f1, f2;
      load_row(y + 1, b0, b1, b2, f0, f1, f2);
      const int iv0 = t0 * w00 + t1 * w01 + b0 * w10 + b1 * w11;
      const int iv1 = t1 * w00 + t2 * w01 + b1 * w10 + b2 * w11;
      const int dx0 = e0.x * w00 + e1.x * w01 + f0.x * w10 + f1.x * w11;
      const int dy0 = e0.y * w00 + e1.y * w01 + f0.y * w10 + f1.y * w11;
      const int dx1 = e1.x * w00 + e2.x * w01 + f1.x * w10 + f2.x * w11;
      const int dy1 = e1.y * w00 + e2.y * w01 + f1.y * w10 + f2.y * w11;
      const int ival0 = (iv0 + 256) >> 9, ixv0 = (dx0 + 8192) >> 14, iyv0 = (dy0 + 8192) >> 14;
      int ival1 = (iv1 + 256) >> 9, ixv1 = (dx1 + 8192) >> 14, iyv1 = (dy1 + 8192) >> 14;
      if (!two) ival1 = ixv1 = iyv1 = 0;             // padding column of an odd-width window
      *(int*)pI = (ival0 & 0xffff) | (ival1 << 16);
      *(int2*)pD = make_int2((ixv0 & 0xffff) | (iyv0 << 16), (ixv1 & 0xffff) | (iyv1 << 16));
      sA11 += (long long)ixv0 * ixv0 + (long long)ixv1 * ixv1;
      sA12 += (long long)ixv0 * iyv0 + (long long)ixv1 * iyv1;
      sA22 += (long long)iyv0 * iyv0 + (long long)iyv1 * iyv1;
      t0 = b0; t1 = b1; t2 = b2; e0 = f0; e1 = f1; e2 = f2;
      pI += wwp; pD += wwp;
    }
  }
}

__global__ void __launch_bounds__(LK_WARPS * 32) lk_track(LkArgs a) {
  extern __shared__ __align__(16) unsigned char lk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int feat = warp / LK_SPLIT, half = warp - feat * LK_SPLIT;
  const int pt = blockIdx.x * LK_FEATS + feat;
  const int b = blockIdx.y;
  if (pt >= a.n_pts) return;                                       // both warps of a feature leave together
  const int ww = a.ww, wh = a.wh;
  const int wwp = lk_wwp(ww);
  const size_t per_warp = lk_patch_bytes(ww, wh);
  // this warp's rows of the window (and of the patch: a warp only ever reads the patch rows it wrote itself)
  const int ya = (wh * half) / LK_SPLIT, yb = (wh * (half + 1)) / LK_SPLIT;
  // exchange of the partial window sums between the feature's warps: [parity][warp of the feature][3]
  long long* xch = (long long*)(lk_smem + LK_FEATS * per_warp) + feat * (2 * LK_SPLIT * 3);
  int xpar = 0;
  auto feature_sum = [&](long long& v0, long long& v1, long long& v2, int n) {
    if (LK_SPLIT == 1) return;
    long long* slot = xch + xpar * (LK_SPLIT * 3);
    if (lane == 0) {
      slot[half * 3] = v0;
      if (n > 1) slot[half * 3 + 1] = v1;
      if (n > 2) slot[half * 3 + 2] = v2;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + feat), "n"(LK_SPLIT * 32) : "memory");
    v0 = 0; v1 = 0; v2 = 0;
#pragma unroll
    for (int k = 0; k < LK_SPLIT; ++k) {
      v0 += slot[k * 3];
      if (n > 1) v1 += slot[k * 3 + 1];
      if (n > 2) v2 += slot[k * 3 + 2];
    }
    xpar ^= 1;        // the next exchange uses the other slot: this one is re-written only after another barrier
  };
  short* sI = (short*)(lk_smem + feat * per_warp);                 // [wh][wwp] int16
  short2* sD = (short2*)(sI + (((size_t)wwp * wh + 3) & ~(size_t)3));   // [wh][wwp] (Ixw, Iyw), 8-byte aligned
  const uint8_t* PI = a.pyr_i + (size_t)b * a.L.pyr_bytes;
  const uint8_t* PJ = a.pyr_j + (size_t)b * a.L.pyr_bytes;
  const short2* DV = a.deriv + (size_t)b * a.L.deriv_elems;
  const size_t pidx = (size_t)b * a.pts_bstride + pt;
  const size_t oidx = (size_t)b * a.n_pts + pt;
  const float ppx = a.prev_pts[2 * pidx], ppy = a.prev_pts[2 * pidx + 1];
  const float hx = (ww - 1) * 0.5f, hy = (wh - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  const double eps2 = (double)a.eps2d_hi + (double)a.eps2d_lo;
  float nx = 0.f, ny = 0.f;  // nextPts[i]
  if (a.flags & B2OF_OPTFLOW_USE_INITIAL_FLOW) { nx = a.next_pts[2 * oidx]; ny = a.next_pts[2 * oidx + 1]; }
  int status = 1;
  float err = 0.f;
  const int max_level = a.L.n - 1;
  for (int level = max_level; level >= 0; --level) {
    const int W = a.L.w[level], H = a.L.h[level];
    const size_t step = a.L.step[level];
    const uint8_t* I = PI + a.L.off[level];
    const uint8_t* J = PJ + a.L.off[level];
    const short2* D = DV + a.L.doff[level];
    const float sc = 1.f / (float)(1 << level);
    float px = ppx * sc, py = ppy * sc;
    float qx, qy;  // nextPt
    if (level == max_level) {
      if (a.flags & B2OF_OPTFLOW_USE_INITIAL_FLOW) { qx = nx * sc; qy = ny * sc; }
      else { qx = px; qy = py; }
    } else {
      qx = nx * 2.f; qy = ny * 2.f;
    }
    nx = qx; ny = qy;
    px -= hx; py -= hy;
    int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -ww || ipx >= W || ipy < -wh || ipy >= H) {
      if (level == 0) { status = 0; err = 0.f; }
      continue;
    }
    int w00, w01, w10, w11;
    lk_weights(px - ipx, py - ipy, w00, w01, w10, w11);
    // ---- template patch + normal matrix ----
    long long sA11 = 0, sA12 = 0, sA22 = 0;
    {
      const bool inside = ipx >= 0 && ipy >= 0 && ipx + ww < W && ipy + wh < H;
      __syncwarp();
      if (inside) lk_patch_strips<true>(I, D, step, W, H, ipx, ipy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, sA11, sA12, sA22);
      else lk_patch_strips<false>(I, D, step, W, H, ipx, ipy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, sA11, sA12, sA22);
      sA11 = warp_sum_ll(sA11); sA12 = warp_sum_ll(sA12); sA22 = warp_sum_ll(sA22);
      __syncwarp();
      feature_sum(sA11, sA12, sA22, 3);
    }
    float A11 = __fmul_rn((float)sA11, FLT_SCALE), A12 = __fmul_rn((float)sA12, FLT_SCALE),
          A22 = __fmul_rn((float)sA22, FLT_SCALE);
    float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    float dd = __fsub_rn(A11, A22);
    float min_eig = __fdiv_rn(
        __fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
        (float)(2 * ww * wh));
    if (a.flags & B2OF_OPTFLOW_LK_GET_MIN_EIGENVALS) err = min_eig;
    if (min_eig < a.min_eig_thr || Dt < 1.1920929e-07f) {
      if (level == 0) status = 0;
      continue;
    }
    Dt = __fdiv_rn(1.f, Dt);
    qx -= hx; qy -= hy;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < a.max_count; ++j) {
      int ix = (int)floorf(qx), iy = (int)floorf(qy);
      if (ix < -ww || ix >= W || iy < -wh || iy >= H) {
        if (level == 0) status = 0;
        break;
      }
      lk_weights(qx - ix, qy - iy, w00, w01, w10, w11);
      long long s1, s2, s3 = 0;
      lk_window_pass<false>(J, step, W, H, ix, iy, w00, w01, w10, w11, sI, sD, ww, wh, ya, yb, lane, s1, s2);
      feature_sum(s1, s2, s3, 2);
      float b1 = __fmul_rn((float)s1, FLT_SCALE), b2 = __fmul_rn((float)s2, FLT_SCALE);
      float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt);
      float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt);
      qx += dx; qy += dy;
      nx = qx + hx; ny = qy + hy;
      if ((double)dx * dx + (double)dy * dy <= eps2) break;
      if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
        nx -= dx * 0.5f; ny -= dy * 0.5f;
        break;
      }
      pdx = dx; pdy = dy;
    }
    if (status && level == 0 && !(a.flags & B2OF_OPTFLOW_LK_GET_MIN_EIGENVALS)) {
      float ex = nx - hx, ey = ny - hy;
      int ix = (int)floorf(ex), iy = (int)floorf(ey);
      if (ix < -ww || ix >= W || iy < -wh || iy >= H) {
        status = 0;
      } else {
        lk_weights(ex - ix, ey - iy, w00, w01, w10, w11);
        long long s1, s2, s3 = 0;
        lk_window_pass<true>(J, step, W, H, ix, iy, w00, w01, w10, w11, sI, sD, ww, wh, ya, yb, lane, s1, s2);
        feature_sum(s1, s2, s3, 1);
        err = __fmul_rn((float)s1, 1.f / (float)(32 * ww * wh));
      }
    }
  }
  if (lane == 0 && half == 0) {
    a.next_pts[2 * oidx] = nx;
    a.next_pts[2 * oidx + 1] = ny;
    a.status[oidx] = (uint8_t)status;
    a.err[oidx] = err;
  }
}

static int lk_check(int rows, int cols, const b2of_lk_params* p) {
  const char* fn = "calcOpticalFlowPyrLK";
  B2OF_ASSERT(p != nullptr, fn);
  B2OF_ASSERT(rows > 0 && cols > 0, fn);
  B2OF_ASSERT(p->max_level >= 0 && p->win_w > 2 && p->win_h > 2, fn);
  size_t per_warp = lk_patch_bytes(p->win_w, p->win_h);
  if (per_warp * LK_FEATS + 1024 > 200 * 1024)
    return fail(B2OF_E_UNSUPPORTED, "winSize %dx%d needs more shared memory than one SM has", p->win_w, p->win_h);
  return B2OF_OK;
}

size_t pyrlk_workspace_bytes(int rows, int cols, const b2of_lk_params* p, int batch) {
  if (lk_check(rows, cols, p)) return 0;
  LkLevels L;
  lk_plan(rows, cols, p, &L);
  if (batch < 1) batch = 1;
  return (size_t)batch * (2 * L.pyr_bytes + L.deriv_elems * sizeof(short2)) + 1024;
}

// level 0 of a pyramid block: the caller's frames copied into the padded-step layout, the whole batch in one launch
// (16-byte units when the source allows it; the padded destination rows always do: the tail of the last unit lands in
// the row padding)
__global__ void __launch_bounds__(256) lk_stage_level0(const uint8_t* __restrict__ src, size_t step, size_t frame_stride,
                                                        uint8_t* __restrict__ dst, size_t dstep, size_t dstride, int cols,
                                                        int vec) {
  const int u = blockIdx.x * 256 + threadIdx.x;
  const uint8_t* s = src + blockIdx.z * frame_stride + blockIdx.y * step;
  uint8_t* d = dst + blockIdx.z * dstride + blockIdx.y * dstep;
  if (vec) {
    if (16 * u >= cols) return;
    if (16 * u + 16 <= cols) {
      *(uint4*)(d + 16 * u) = __ldg((const uint4*)(s + 16 * u));
    } else {
      for (int c = 16 * u; c < cols; ++c) d[c] = s[c];
    }
  } else if (u < cols) {
    d[u] = s[u];
  }
}

int pyrlk_dev(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride, int batch, int rows, int cols,
              const float* prev_pts, size_t pts_bstride, int n_pts, float* next_pts, uint8_t* status, float* err,
              const b2of_lk_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = lk_check(rows, cols, p);
  if (rc) return rc;
  const char* fn = "calcOpticalFlowPyrLK";
  B2OF_ASSERT(prev != nullptr && next != nullptr && step >= (size_t)cols, fn);
  B2OF_ASSERT(n_pts >= 0 && batch >= 0, fn);
  if (n_pts == 0 || batch == 0) return B2OF_OK;
  B2OF_ASSERT(prev_pts != nullptr && next_pts != nullptr && status != nullptr && err != nullptr, fn);
  LkArgs a{};
  lk_plan(rows, cols, p, &a.L);
  size_t need = (size_t)batch * (2 * a.L.pyr_bytes + a.L.deriv_elems * sizeof(short2));
  if (ws == nullptr || ws_bytes < need)
    return fail(B2OF_E_NOMEM, "pyrlk workspace too small: %zu B given, %zu B needed", ws_bytes, need);
  uint8_t* pi = (uint8_t*)ws;
  uint8_t* pj = pi + (size_t)batch * a.L.pyr_bytes;
  short2* dv = (short2*)(pj + (size_t)batch * a.L.pyr_bytes);
  // level 0: copy into the padded-step pyramid block, then the pyrDown chain
  for (int which = 0; which < 2; ++which) {
    uint8_t* dst = which ? pj : pi;
    const uint8_t* src = which ? next : prev;
    if (batch == 1 || frame_stride != 0) {
      // one launch for the whole batch (blockIdx.z = batch item) instead of one 2-D copy per image
      const bool vec = (((uintptr_t)src | step | frame_stride | (uintptr_t)dst | a.L.pyr_bytes) & 15) == 0;   // (destination steps are 16-byte multiples)
      const int units = vec ? cdiv(cols, 16) : cols;
      lk_stage_level0<<<dim3(cdiv(units, 256), rows, batch), 256, 0, st>>>(src, step, frame_stride, dst, a.L.step[0],
                                                                         a.L.pyr_bytes, cols, vec ? 1 : 0);
      B2OF_LAUNCH_CHECK();
    } else {
      return fail(B2OF_E_BADARG, "frame_stride == 0 with batch > 1");
    }
    for (int l = 1; l < a.L.n; ++l) {
      {
        ProfScope ps(PT_PYRDOWN, st, batch * 1.25 * a.L.w[l - 1] * a.L.h[l - 1]);
        rc = pyrdown_dev(dst + a.L.off[l - 1], a.L.h[l - 1], a.L.w[l - 1], a.L.step[l - 1], a.L.pyr_bytes,
                         dst + a.L.off[l], a.L.step[l], a.L.pyr_bytes, batch, st);
      }
      if (rc) return rc;
    }
  }
  for (int l = 0; l < a.L.n; ++l) {
    dim3 grid(cdiv(cdiv(a.L.w[l], 4), 256), a.L.h[l], batch);
    {
      ProfScope ps(PT_LK_SCHARR, st, batch * 5.0 * a.L.w[l] * a.L.h[l]);
      lk_scharr<<<grid, 256, 0, st>>>(pi + a.L.off[l], a.L.step[l], a.L.pyr_bytes, a.L.w[l], a.L.h[l], dv + a.L.doff[l],
                                      a.L.deriv_elems);
    }
    B2OF_LAUNCH_CHECK();
  }
  a.pyr_i = pi; a.pyr_j = pj; a.deriv = dv;
  a.prev_pts = prev_pts; a.pts_bstride = pts_bstride;
  a.next_pts = next_pts; a.status = status; a.err = err;
  a.n_pts = n_pts;
  a.ww = p->win_w; a.wh = p->win_h;
  int mc = 30;
  double eps = 0.001;
  if (p->crit_type & B2OF_TERM_COUNT) mc = p->crit_max_count < 0 ? 0 : (p->crit_max_count > 100 ? 100 : p->crit_max_count);
  if (p->crit_type & B2OF_TERM_EPS) eps = p->crit_eps < 0 ? 0. : (p->crit_eps > 10. ? 10. : p->crit_eps);
  double eps2 = eps * eps;
  a.max_count = mc;
  a.eps2d_hi = (float)eps2;
  a.eps2d_lo = (float)(eps2 - (double)a.eps2d_hi);
  a.flags = p->flags;
  a.min_eig_thr = (float)p->min_eig_threshold;
  size_t per_warp = lk_patch_bytes(p->win_w, p->win_h);
  size_t smem = per_warp * LK_FEATS + (size_t)LK_FEATS * 2 * LK_SPLIT * 3 * sizeof(long long);
  static PerDeviceMax max_set;
  if (smem > 48 * 1024 && max_set.raise(smem))
    B2OF_CUDA(cudaFuncSetAttribute(lk_track, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(n_pts, LK_FEATS), batch);
  {
    // algorithmic bytes: both pyramids once (u8) + derivatives of the first (4 B/px) + 21 B per point
    double px = 0;
    for (int l = 0; l < a.L.n; ++l) px += (double)a.L.w[l] * a.L.h[l];
    ProfScope ps(PT_LK_TRACK, st, batch * (6.0 * px + 21.0 * n_pts));
    lk_track<<<grid, LK_WARPS * 32, smem, st>>>(a);
  }
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

}  // namespace b2of
