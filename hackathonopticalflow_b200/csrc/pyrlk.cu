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


// One image's pyramid block: every level is stored with a border of pad_x columns and pad_y rows on each side, filled
// with BORDER_REFLECT_101 pixels (derivatives: zeros -- cv2 pads them with BORDER_CONSTANT) by lk_pad_borders.  A
// window may stick out of the frame by its own size (cv2 tracks such points), and with the border in memory the walk
// over a window at the frame edge is the same straight-line code as anywhere else: the reflecting walk (an index
// reflection per row and five byte loads per strip row) took a third of lk_track's time for a tenth of the windows.
struct LkLevels {
  int n;  // number of levels (effective maxLevel + 1)
  int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
  int pad_x, pad_y;
  size_t step[LK_MAX_LEVELS];        // row pitch in bytes (w + 2 pad_x rounded up to 16)
  size_t dstep[LK_MAX_LEVELS];       // derivative row pitch in elements (= step)
  size_t off[LK_MAX_LEVELS];         // byte offset of pixel (0, 0) of level l inside one image's pyramid block
  size_t base[LK_MAX_LEVELS];        // byte offset of the level's padded block (row -pad_y, column -pad_x)
  size_t doff[LK_MAX_LEVELS];        // short2 offset of derivative (0, 0) of level l inside one image's block
  size_t dbase[LK_MAX_LEVELS];
  size_t pyr_bytes, deriv_elems;     // per image
};

static int lk_plan(int rows, int cols, const b2of_lk_params* p, LkLevels* L) {
  memset(L, 0, sizeof *L);
  int w = cols, h = rows, n = 0;
  size_t off = 0, doff = 0;
  // columns read: window column -win_w .. w - 1 + win_w, plus the over-read of the last four-column strip and of its
  // aligned word pair (<= 10 columns); rows: -win_h .. h - 1 + win_h, plus the rows requested ahead
  L->pad_x = (int)align_up(p->win_w + 10, 16);
  L->pad_y = p->win_h + 5;            // (+4: lk_track requests the image words of a strip one four-row step ahead)
  for (int level = 0; level <= p->max_level && level < LK_MAX_LEVELS; ++level) {
    L->w[n] = w; L->h[n] = h;
    L->step[n] = align_up(w + 2 * L->pad_x, 16);
    L->dstep[n] = L->step[n];
    L->base[n] = off; L->dbase[n] = doff;
    L->off[n] = off + (size_t)L->pad_y * L->step[n] + L->pad_x;
    L->doff[n] = doff + (size_t)L->pad_y * L->dstep[n] + L->pad_x;
    off += align_up(L->step[n] * (h + 2 * L->pad_y), 256);
    doff += align_up(L->dstep[n] * (h + 2 * L->pad_y), 64);
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

// four pixels per thread.  Rows of a pyramid level start 16-byte aligned (LkLevels), so the 3 x 6 neighbourhood of
// the pixels x0 .. x0 + 3 (x0 a multiple of four) is three aligned words per row -- the one before (its last byte), the
// thread's own, the one after (its first byte) -- instead of eighteen byte loads; at the frame's left / right edge the
// missing neighbour is the REFLECT_101 pixel, one of the thread's own bytes.  The four (Ix, Iy) pairs go out as one
// 16-byte store when the row start is aligned.
__global__ void __launch_bounds__(256) lk_scharr(const uint8_t* __restrict__ img, size_t step, size_t img_bstride,
                                                  int w, int h, short2* __restrict__ d, size_t d_bstride,
                                                  size_t d_pitch) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x0 >= w) return;
  const uint8_t* b = img + blockIdx.z * img_bstride;
  const uint8_t* r0 = b + (size_t)reflect101_pm1(y - 1, h) * step;
  const uint8_t* r1 = b + (size_t)y * step;
  const uint8_t* r2 = b + (size_t)reflect101_pm1(y + 1, h) * step;
  int t0[6], t1[6];                      // vertical smooth (3, 10, 3) and difference at columns x0 - 1 .. x0 + 4
  if (((step | (size_t)b) & 3) == 0 && x0 + 4 <= w && w >= 2) {
    auto row6 = [&](const uint8_t* r, int* v) {
      const uint32_t own = __ldg((const uint32_t*)(r + x0));
      v[1] = own & 255u; v[2] = (own >> 8) & 255u; v[3] = (own >> 16) & 255u; v[4] = own >> 24;
      v[0] = x0 > 0 ? (int)(__ldg((const uint32_t*)(r + x0 - 4)) >> 24) : v[2];              // column -1 is column 1
      // column x0 + 4: the next word's first byte (the word may hang over w inside the 16-byte-padded row), or, at
      // the right edge, column w -> w - 2
      v[5] = x0 + 4 < w ? (int)(__ldg((const uint32_t*)(r + x0 + 4)) & 255u) : v[3];
    };
    int a[6], c[6], e[6];
    row6(r0, a); row6(r1, c); row6(r2, e);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      t0[k] = (a[k] + e[k]) * 3 + c[k] * 10;
      t1[k] = e[k] - a[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int xx = reflect101_pm1(min(x0 - 1 + k, w), w);
      const int a = r0[xx], c = r1[xx], e = r2[xx];
      t0[k] = (a + e) * 3 + c * 10;
      t1[k] = e - a;
    }
  }
  short2 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    o[k] = make_short2((short)(t0[k + 2] - t0[k]), (short)((t1[k] + t1[k + 2]) * 3 + t1[k + 1] * 10));
  short2* out = d + blockIdx.z * d_bstride + (size_t)y * d_pitch + x0;
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

// Patch layout in shared memory, per warp: Iw as int16 [wh][wwp] with wwp = ww rounded up to a multiple of four, then
// one 16-byte record per (row, group of four columns): (Ix0 | Ix1 << 16, Ix2 | Ix3 << 16, Iy0 | Iy1 << 16,
// Iy2 | Iy3 << 16).  Padding columns hold zeros, so whatever the window walk computes for them multiplies to nothing.
__host__ __device__ inline int lk_wwp(int ww) { return (ww + 3) & ~3; }
__host__ __device__ inline size_t lk_iw_bytes(int ww, int wh) { return (((size_t)lk_wwp(ww) * wh * 2) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t lk_patch_bytes(int ww, int wh) {
  return lk_iw_bytes(ww, wh) + (size_t)lk_wwp(ww) * wh * 4;
}

// Work split of a window: the warp's lanes own strips NC columns wide (S strips) and, when the window is narrow,
// G = 32 / S groups of rows.  A lane walks down its strip: the image row below is loaded once and serves as the
// bottom neighbours of this row and the top neighbours of the next one, addresses advance by one row step, and the
// template patch comes in as aligned records.  Window sums are exact integers, so the result does not depend on the
// split.  The window walk uses NC = 4 (12 strips x 2 row groups for the 45-wide grid window), the patch pass NC = 2.
template <int NC>
struct LkStrips {
  int S, G, RG;
  int g0, sx0;               // (row group, strip) of unit u = lane: computed once per kernel, not once per window pass
  __device__ __forceinline__ LkStrips(int ww, int nrows, int lane) {   // nrows = rows of this warp's share of the window
    S = lk_wwp(ww) / NC;
    G = S <= 16 ? 32 / S : 1;
    RG = (nrows + G - 1) / G;
    g0 = lane / S; sx0 = lane - g0 * S;
  }
  __device__ __forceinline__ void unit(int u, int lane, int& g, int& sx) const {
    if (u == lane) { g = g0; sx = sx0; }
    else { g = u / S; sx = u - g * S; }
  }
};

// d = c + a.h0 * b.b0 + a.h1 * b.b1 (LO: bytes 0, 1 of b; HI: bytes 2, 3), a signed 16-bit halves, b unsigned bytes
__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// sum over the window of (bilinear(J) >> 9 - Iw) * {Ixw, Iyw}  (or |diff| when ABS).  J points at pixel (0, 0) of a
// level stored with its reflected border (LkLevels), so a window anywhere is walked with the same code.
//
// A lane's strip is four columns wide.  Its five source bytes of an image row (columns c .. c + 4) are two aligned
// words and two funnel shifts: R = bytes c .. c + 3, R2 = bytes c + 1 .. c + 4.  The 14-bit bilinear weights fit 16
// bits, so one IDP.2A does two of the four taps of an element (the bottom pair as SIGNED halves: w11 is what the three
// rounded weights leave of 2^14 and is -1 when all three round up and the true weight is below one half -- about one
// window position in ten thousand; cv2 carries it as a signed short too, and 256 - 255 keeps the sum positive):
//   v0 = 256 + (w00, w01) . R.lo + (w10, w11) . Rbelow.lo,   v1: R2.lo,   v2: R.hi,   v3: R2.hi        (8 IDP.2A)
// against sixteen IMADs and four adds.  The four differences stay packed: (v >> 9) + 0x2000 - Iw is positive and below
// 2^15, so two of them are subtracted from the packed Iw pair with ONE 32-bit add; a byte permute sorts the low and
// the high bytes of such a pair together and four more IDP.2A per pair multiply them by the packed (Ix0, Ix1) and
// (Iy0, Iy1) of the patch -- sum(e * Ix) = sum(el * Ix) + 256 sum(eh * Ix) with e = diff + 0x2000, and the offset
// 0x2000 * sum(Ix) is a constant of the patch, taken off by the caller.  All partial sums fit 32 bits for a strip
// (|Ix| <= 4080, el <= 255: 2^20 per product, at most a few hundred products); they are widened per strip.
template <bool ABS>
__device__ __forceinline__ void lk_window_strips(const uint8_t* __restrict__ J, size_t step, int ix, int iy,
                                                 unsigned wt, unsigned wb, const short* sI, const int4* sD, int ww,
                                                 int ya, int yb, int lane, const LkStrips<4>& st, long long& s1,
                                                 long long& s2) {
  const int wwp = lk_wwp(ww);
  const ptrdiff_t stepw = (ptrdiff_t)(step >> 2);          // (pyramid steps are multiples of 16 bytes)
  for (int u = lane; u < st.S * st.G; u += 32) {
    int g, sx;
    st.unit(u, lane, g, sx);
    const int x0 = 4 * sx;
    const int y0 = ya + g * st.RG, y1 = min(yb, y0 + st.RG);
    const uint8_t* p0 = J + (ptrdiff_t)(iy + y0) * (ptrdiff_t)step + (ix + x0);
    const unsigned sh = ((unsigned)(uintptr_t)p0 & 3u) * 8u;
    const uint32_t* q = (const uint32_t*)(p0 - ((uintptr_t)p0 & 3));   // the aligned word holding column ix + x0
    auto shifted = [&](unsigned lo, unsigned hi, unsigned& R, unsigned& R2) {
      R = __funnelshift_r(lo, hi, sh);
      R2 = __funnelshift_rc(lo, hi, sh + 8u);
    };
    unsigned Rt, Rt2;
    shifted(q[0], q[1], Rt, Rt2);
    const short* pI = sI + y0 * wwp + x0;
    const int4* pD = sD + y0 * st.S + sx;
    int aL1 = 0, aH1 = 0, aL2 = 0, aH2 = 0;
    unsigned sabs = 0;
    auto one_row = [&](unsigned blo, unsigned bhi, int2 iq, int4 dq) {
      unsigned Rb, Rb2;
      shifted(blo, bhi, Rb, Rb2);
      const unsigned v0 = (unsigned)dp2a_lo_su((int)wb, Rb, (int)__dp2a_lo(wt, Rt, 256u));
      const unsigned v1 = (unsigned)dp2a_lo_su((int)wb, Rb2, (int)__dp2a_lo(wt, Rt2, 256u));
      const unsigned v2 = (unsigned)dp2a_hi_su((int)wb, Rb, (int)__dp2a_hi(wt, Rt, 256u));
      const unsigned v3 = (unsigned)dp2a_hi_su((int)wb, Rb2, (int)__dp2a_hi(wt, Rt2, 256u));
      // e = (v >> 9) + 0x2000 - Iw, two to a register
      const unsigned e01 = __byte_perm(v0 >> 9, v1 >> 9, 0x5410) + 0x20002000u - (unsigned)iq.x;
      const unsigned e23 = __byte_perm(v2 >> 9, v3 >> 9, 0x5410) + 0x20002000u - (unsigned)iq.y;
      if (ABS) {
        // padding columns hold Iw = 0 and must not count: the caller's window is ww wide
        const int d0 = (int)(e01 & 0xffffu) - 0x2000, d1 = (int)(e01 >> 16) - 0x2000;
        const int d2 = (int)(e23 & 0xffffu) - 0x2000, d3 = (int)(e23 >> 16) - 0x2000;
        sabs += abs(d0);
        if (x0 + 1 < ww) sabs += abs(d1);
        if (x0 + 2 < ww) sabs += abs(d2);
        if (x0 + 3 < ww) sabs += abs(d3);
      } else {
        const unsigned q01 = __byte_perm(e01, 0, 0x3120), q23 = __byte_perm(e23, 0, 0x3120);   // (el0, el1, eh0, eh1)
        aL1 = dp2a_lo_su(dq.x, q01, aL1); aH1 = dp2a_hi_su(dq.x, q01, aH1);
        aL2 = dp2a_lo_su(dq.z, q01, aL2); aH2 = dp2a_hi_su(dq.z, q01, aH2);
        aL1 = dp2a_lo_su(dq.y, q23, aL1); aH1 = dp2a_hi_su(dq.y, q23, aH1);
        aL2 = dp2a_lo_su(dq.w, q23, aL2); aH2 = dp2a_hi_su(dq.w, q23, aH2);
      }
      Rt = Rb; Rt2 = Rb2;
    };
    int y = y0;
    // four rows per step.  The image words of the NEXT step's rows are requested before this step's arithmetic (a
    // whole step of work hides their latency; the rows past the strip's last that this reads lie in the level's
    // border); the patch records come from shared memory at the head of the step that uses them.
    unsigned nl[4], nh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { nl[k] = q[(k + 1) * stepw]; nh[k] = q[(k + 1) * stepw + 1]; }
    for (; y + 4 <= y1; y += 4) {
      unsigned bl[4], bh[4];
      int2 ip[4];
      int4 dp[4];
      q += 4 * stepw;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bl[k] = nl[k]; bh[k] = nh[k];
        nl[k] = q[(k + 1) * stepw]; nh[k] = q[(k + 1) * stepw + 1];
        ip[k] = *(const int2*)(pI + k * wwp);
        dp[k] = ABS ? make_int4(0, 0, 0, 0) : pD[k * st.S];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) one_row(bl[k], bh[k], ip[k], dp[k]);
      pI += 4 * wwp; pD += 4 * st.S;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {                  // the up to three rows left: their words are already here
      if (y + k < y1) {
        one_row(nl[k], nh[k], *(const int2*)pI, ABS ? make_int4(0, 0, 0, 0) : *pD);
        pI += wwp; pD += st.S;
      }
    }
    if (ABS) {
      s1 += sabs;
    } else {
      s1 += (long long)aL1 + 256ll * aH1;
      s2 += (long long)aL2 + 256ll * aH2;
    }
  }
}

// sIx, sIy: sums of the patch's Ixw / Iyw over this warp's rows (the 0x2000 offset of the packed differences)
template <bool ABS>
__device__ __forceinline__ void lk_window_pass(const uint8_t* __restrict__ J, size_t step, int ix, int iy, int w00,
                                               int w01, int w10, int w11, const short* sI, const int4* sD, int ww,
                                               int ya, int yb, int lane, const LkStrips<4>& st, long long sIx,
                                               long long sIy, long long& o1, long long& o2) {
  const unsigned wt = (unsigned)w00 | ((unsigned)w01 << 16), wb = (unsigned)w10 | ((unsigned)w11 << 16);
  long long s1 = 0, s2 = 0;
  lk_window_strips<ABS>(J, step, ix, iy, wt, wb, sI, sD, ww, ya, yb, lane, st, s1, s2);
  o1 = warp_sum_ll(s1);
  o2 = 0;
  if (!ABS) {
    o1 -= 0x2000ll * sIx;
    o2 = warp_sum_ll(s2) - 0x2000ll * sIy;
  }
}

// template patch (Iw, Ixw, Iyw) of the window at (ipx, ipy) + the three sums of the normal matrix and the plain sums
// of Ixw, Iyw.  I and D point at element (0, 0) of levels stored with their borders (reflected pixels, zero
// derivatives).  Same four-column strips as the window walk: the image row comes in as two aligned words and Iw takes
// eight IDP.2A; the five (Ix, Iy) pairs of a derivative row are unpacked once and serve four elements (the 16 x 16-bit
// products have no packed form: 32 IMADs per strip row, which is what bounds this loop).  One 8-byte and one 16-byte
// store per strip row.
__device__ __forceinline__ void lk_patch_strips(const uint8_t* __restrict__ I, const short2* __restrict__ D, size_t step,
                                                size_t dstep, int ipx, int ipy, int w00, int w01, int w10, int w11,
                                                short* sI, int4* sD, int ww, int ya, int yb, int lane,
                                                const LkStrips<4>& st, long long& sA11, long long& sA12,
                                                long long& sA22, long long& sIx, long long& sIy) {
  const int wwp = lk_wwp(ww);
  const unsigned wt = (unsigned)w00 | ((unsigned)w01 << 16), wb = (unsigned)w10 | ((unsigned)w11 << 16);
  const ptrdiff_t stepw = (ptrdiff_t)(step >> 2);
  int accx = 0, accy = 0;                          // |Ixw| <= 4080: 32 bits hold any window that fits shared memory
  for (int u = lane; u < st.S * st.G; u += 32) {
    int g, sx;
    st.unit(u, lane, g, sx);
    const int x0 = 4 * sx;
    const int nv = ww - x0;                        // elements x0 + k with k >= nv are padding and are stored as zeros
    const int y0 = ya + g * st.RG, y1 = min(yb, y0 + st.RG);
    const uint8_t* p0 = I + (ptrdiff_t)(ipy + y0) * (ptrdiff_t)step + (ipx + x0);
    const unsigned sh = ((unsigned)(uintptr_t)p0 & 3u) * 8u;
    const uint32_t* q = (const uint32_t*)(p0 - ((uintptr_t)p0 & 3));
    const int* dr = (const int*)(D + (ptrdiff_t)(ipy + y0) * (ptrdiff_t)dstep + (ipx + x0));
    unsigned Rt, Rt2;
    int ex[5], ey[5];                              // the row above: unpacked derivatives of columns x0 .. x0 + 4
    auto unpack = [](int v, int& x, int& y) { x = (int)(short)(v & 0xffff); y = v >> 16; };
    {
      const unsigned lo = q[0], hi = q[1];
      Rt = __funnelshift_r(lo, hi, sh); Rt2 = __funnelshift_rc(lo, hi, sh + 8u);
#pragma unroll
      for (int k = 0; k < 5; ++k) unpack(dr[k], ex[k], ey[k]);
    }
    int2* pI = (int2*)(sI + y0 * wwp + x0);
    int4* pD = sD + y0 * st.S + sx;
    int q11 = 0, q12 = 0, q22 = 0;                 // widened every four rows
    auto flush = [&]() { sA11 += q11; sA12 += q12; sA22 += q22; q11 = q12 = q22 = 0; };
    // the loads of the row after next are issued before this row's arithmetic: a whole row of work hides them (the
    // borders make the one row read past the strip's last harmless)
    unsigned nlo, nhi;
    int ndv[5];
    q += stepw; dr += dstep;
    nlo = q[0]; nhi = q[1];
#pragma unroll
    for (int k = 0; k < 5; ++k) ndv[k] = dr[k];
#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
      const unsigned lo = nlo, hi = nhi;
      int dv[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) dv[k] = ndv[k];
      q += stepw; dr += dstep;
      nlo = q[0]; nhi = q[1];
#pragma unroll
      for (int k = 0; k < 5; ++k) ndv[k] = dr[k];
      const unsigned Rb = __funnelshift_r(lo, hi, sh), Rb2 = __funnelshift_rc(lo, hi, sh + 8u);
      const unsigned v0 = (unsigned)dp2a_lo_su((int)wb, Rb, (int)__dp2a_lo(wt, Rt, 256u));
      const unsigned v1 = (unsigned)dp2a_lo_su((int)wb, Rb2, (int)__dp2a_lo(wt, Rt2, 256u));
      const unsigned v2 = (unsigned)dp2a_hi_su((int)wb, Rb, (int)__dp2a_hi(wt, Rt, 256u));
      const unsigned v3 = (unsigned)dp2a_hi_su((int)wb, Rb2, (int)__dp2a_hi(wt, Rt2, 256u));
      unsigned iw01 = __byte_perm(v0 >> 9, v1 >> 9, 0x5410), iw23 = __byte_perm(v2 >> 9, v3 >> 9, 0x5410);
      int fx[5], fy[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) unpack(dv[k], fx[k], fy[k]);
      int ix[4], iy[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ix[k] = (ex[k] * w00 + ex[k + 1] * w01 + fx[k] * w10 + fx[k + 1] * w11 + 8192) >> 14;
        iy[k] = (ey[k] * w00 + ey[k + 1] * w01 + fy[k] * w10 + fy[k + 1] * w11 + 8192) >> 14;
      }
      if (nv < 4) {                                // the window's last strip
        if (nv < 2) { ix[1] = iy[1] = 0; iw01 &= 0xffffu; }
        if (nv < 3) { ix[2] = iy[2] = 0; iw23 = 0; }
        ix[3] = iy[3] = 0; iw23 &= 0xffffu;
      }
      *pI = make_int2((int)iw01, (int)iw23);
      *pD = make_int4((int)__byte_perm(ix[0], ix[1], 0x5410), (int)__byte_perm(ix[2], ix[3], 0x5410),
                      (int)__byte_perm(iy[0], iy[1], 0x5410), (int)__byte_perm(iy[2], iy[3], 0x5410));
#pragma unroll
      for (int k = 0; k < 4; ++k) {                // |Ixw| <= 4080: sixteen products fit 32 bits with room to spare
        q11 += ix[k] * ix[k]; q12 += ix[k] * iy[k]; q22 += iy[k] * iy[k];
      }
      accx += (ix[0] + ix[1]) + (ix[2] + ix[3]); accy += (iy[0] + iy[1]) + (iy[2] + iy[3]);
      Rt = Rb; Rt2 = Rb2;
#pragma unroll
      for (int k = 0; k < 5; ++k) { ex[k] = fx[k]; ey[k] = fy[k]; }
      pI += wwp >> 2; pD += st.S;
      if (((y - y0) & 3) == 3) flush();
    }
    flush();
  }
  sIx += accx; sIy += accy;
}

__global__ void __launch_bounds__(LK_WARPS * 32, 512 / (LK_WARPS * 32)) lk_track(LkArgs a) {
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
  const LkStrips<4> st(ww, yb - ya, lane);
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
  int4* sD = (int4*)(lk_smem + feat * per_warp + lk_iw_bytes(ww, wh));   // [wh][wwp / 4] packed (Ixw, Iyw) records
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
    const size_t step = a.L.step[level], dstep = a.L.dstep[level];
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
    long long sA11 = 0, sA12 = 0, sA22 = 0, sIx = 0, sIy = 0;
    {
      __syncwarp();
      lk_patch_strips(I, D, step, dstep, ipx, ipy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, st, sA11, sA12, sA22,
                      sIx, sIy);
      sA11 = warp_sum_ll(sA11); sA12 = warp_sum_ll(sA12); sA22 = warp_sum_ll(sA22);
      sIx = warp_sum_ll(sIx); sIy = warp_sum_ll(sIy);
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
      lk_window_pass<false>(J, step, ix, iy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, st, sIx, sIy, s1, s2);
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
        lk_window_pass<true>(J, step, ix, iy, w00, w01, w10, w11, sI, sD, ww, ya, yb, lane, st, sIx, sIy, s1, s2);
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

// Border of one level of every pyramid block (LkLevels): word (4 columns) per thread over the top band, the bottom
// band and the two side bands of the padded block.  Image words are assembled from BORDER_REFLECT_101 pixels (a side
// word that straddles the frame edge rewrites its interior bytes with themselves); derivative borders are zeros.
__global__ void __launch_bounds__(256) lk_pad_borders(uint8_t* __restrict__ img, size_t img_bstride, short2* __restrict__ der,
                                                       size_t der_bstride, int n_der, size_t base, size_t dbase,
                                                       size_t step, int w, int h, int pad_x, int pad_y) {
  const int wpr = (int)(step >> 2);                       // words per padded row
  const int lw = pad_x >> 2, rw0 = (pad_x + w) >> 2;      // side bands: words [0, lw) and [rw0, wpr)
  const int side = lw + (wpr - rw0);
  const int band = 2 * pad_y * wpr;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= band + h * side) return;
  int Y, Xw;                                              // padded coordinates (row, word)
  if (t < band) {
    Y = t / wpr; Xw = t - Y * wpr;
    if (Y >= pad_y) Y += h;
  } else {
    const int u = t - band;
    const int r = u / side, c = u - r * side;
    Y = pad_y + r; Xw = c < lw ? c : rw0 + (c - lw);
  }
  uint8_t* ib = img + blockIdx.y * img_bstride + base;
  const uint8_t* src = ib + (size_t)pad_y * step + pad_x + (size_t)reflect101(Y - pad_y, h) * step;
  unsigned v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) v |= (unsigned)src[reflect101(4 * Xw + k - pad_x, w)] << (8 * k);
  *(unsigned*)(ib + (size_t)Y * step + 4 * Xw) = v;
  if ((int)blockIdx.y < n_der) {
    short2* db = der + blockIdx.y * der_bstride + dbase + (size_t)Y * step + 4 * Xw;
    const bool row_in = Y >= pad_y && Y < pad_y + h;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = 4 * Xw + k - pad_x;
      if (!(row_in && x >= 0 && x < w)) db[k] = make_short2(0, 0);
    }
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
      lk_stage_level0<<<dim3(cdiv(units, 256), rows, batch), 256, 0, st>>>(src, step, frame_stride, dst + a.L.off[0], a.L.step[0],
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
                                      a.L.deriv_elems, a.L.dstep[l]);
    }
    B2OF_LAUNCH_CHECK();
  }
  // borders of every level: both pyramids (pi and pj are adjacent: 2 x batch blocks) and the derivatives of the first
  for (int l = 0; l < a.L.n; ++l) {
    const int wpr = (int)(a.L.step[l] >> 2);
    const int side = (a.L.pad_x >> 2) + (wpr - ((a.L.pad_x + a.L.w[l]) >> 2));
    const int total = 2 * a.L.pad_y * wpr + a.L.h[l] * side;
    ProfScope ps(PT_LK_SCHARR, st, 2.0 * batch * 4.0 * total);
    lk_pad_borders<<<dim3(cdiv(total, 256), 2 * batch), 256, 0, st>>>(pi, a.L.pyr_bytes, dv, a.L.deriv_elems, batch,
                                                                      a.L.base[l], a.L.dbase[l], a.L.step[l], a.L.w[l],
                                                                      a.L.h[l], a.L.pad_x, a.L.pad_y);
    B2OF_LAUNCH_CHECK();
  }
  a.pyr_i = pi; a.pyr_j = pj; a.deriv = dv;
  a.prev_pts = prev_pts; a.pts_bstride = pts_bstride;
  a.next_pts = next_pts; a.status = status; a.err = err;
  a.n_pts = n_pts;
  a.ww = p->win_w; a.wh = p->win_h;
  int mc = 30;
  double eps = 0.01;     // cv2's value when the criteria carry no EPS (probed against the wheel: COUNT-only == (COUNT|EPS, 0.01))
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
