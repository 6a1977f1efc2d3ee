// K5/K6, strip form (the reference's window: winsize 15, box): one fused kernel per (level, iteration), built
// around the unit that bounds this kernel on B200 -- the L1 / shared-memory data pipe (one 128-byte wavefront per
// cycle per SM; every global or shared access costs one wavefront per 128 bytes it touches).
//
//   * a CTA owns a strip of 112 output columns (128 halo columns = 4 whole warps per row: no partial warps, every
//     row load falls on whole 128-byte lines) and walks down it in blocks of 16 rows;
//   * M lives in a ring of 16 + 14 rows as three planes (float2 ch0/1, float2 ch2/3, float ch4) whose odd row
//     stride makes both the row-parallel and the column-parallel accesses conflict-free;
//   * step A (all warps): UpdateMatrices for the new rows, thread = (column, run of rows), bottom gather corners
//     carried down the column in registers (see fb_iter);
//   * step B (3 warps, lane = row x half): horizontal 15-sums in place, each half-row walked once;
//   * step C (4 warps, thread = output column): the vertical 15-sum is a running sum carried in registers down the
//     whole strip (one add and one subtract per row: two reads of each summed row instead of three), refreshed
//     from the ring every FBS_REFRESH blocks so rounding does not accumulate; 2x2 solve; coalesced 8-byte stores.
// Wavefronts per output pixel: about 1.9 (A 1.0, B 0.6, C 0.3) against 3.6 for the square-tile kernel.
#pragma once

constexpr int FBS_NT = 384;               // threads per CTA
constexpr int FBS_EW = 128;               // halo columns per strip
constexpr int FBS_TW = 112;               // output columns per strip
constexpr int FBS_M = 7;                  // window radius
constexpr int FBS_PADL = 8;               // halo column 0 sits at image column x0 - 8 (8-pixel aligned)
constexpr int FBS_RB = 15;                // rows per block
constexpr int FBS_NR = FBS_RB + 2 * FBS_M;   // ring rows
constexpr int FBS_ES = 129;               // plane row stride in elements (odd)
constexpr int FBS_HL = FBS_TW / 2;        // outputs [0,HL) are summed left->right, [HL,TW) right->left
constexpr int FBS_RUNS = FBS_NT / FBS_EW; // row runs per block in step A
#ifndef FBS_BULK_PF
#define FBS_BULK_PF 1
#endif
#ifndef FBS_EXP
#define FBS_EXP 0
#endif
#ifndef FBS_REFRESH
#define FBS_REFRESH 4
#endif
constexpr size_t FBS_SMEM = (size_t)FBS_NR * FBS_ES * 20;

__device__ __forceinline__ void sts_f2(unsigned addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }

struct FbRow {               // one row of a thread's run in flight through step A
  float2 d;                  // flow vector
  float4 q; float q4;        // R0 record
  float fx, fy;              // fractional warp position
  int y, ot;                 // image row; R1 element offset of the top-left corner
  bool inside;
  FbCorner bot;              // bottom corner pair
};

template <int MODE>
__global__ void __launch_bounds__(FBS_NT, 2) fb_iter_strip(IterArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int EW = FBS_EW, TW = FBS_TW, M = FBS_M, RB = FBS_RB, NR = FBS_NR, ES = FBS_ES, HL = FBS_HL;
  float2* Pxy = (float2*)smem;
  float2* Pzw = Pxy + NR * ES;
  float* Pe = (float*)(Pzw + NR * ES);
  const unsigned s_xy = smem_u32(smem);                 // float2 [NR][ES]  (M0, M1)
  const unsigned s_zw = s_xy + NR * ES * 8;             // float2 [NR][ES]  (M2, M3)
  const unsigned s_e = s_zw + NR * ES * 8;              // float  [NR][ES]  (M4)
  const int pair = blockIdx.x;                          // pair index fastest (frame p + 1 shared through L2)
  const int x0 = blockIdx.y * TW;
  const int w = a.w, h = a.h, pitch = a.pitch;
  const int ys = blockIdx.z * a.nb * RB;                // rows [ys, ye) are this CTA's outputs
  const int ye = min(ys + a.nb * RB, h);
  const float* base0 = a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride;
  const float4* __restrict__ R0a = (const float4*)base0;
  const float* __restrict__ R0b = base0 + 4 * a.plane_stride;
  const float4* __restrict__ R1a = (const float4*)(base0 + a.r_frame_stride);
  const float* __restrict__ R1b = base0 + a.r_frame_stride + 4 * a.plane_stride;
  const float2* __restrict__ fin = MODE ? a.flow_in + (size_t)pair * a.flow_in_pair_stride : nullptr;
  float2* __restrict__ fo = a.flow_out + (size_t)pair * a.flow_out_pair_stride;
  const int t = threadIdx.x;

  // ---- step A constants: thread = (halo column cx, run) ----
  const int cx = t & (EW - 1), run = t >> 7;            // 3 runs of 128 columns
  const int x = clampi(x0 - FBS_PADL + cx, 0, w - 1);
  const float xf = (float)x;
  const bool xb_border = (unsigned)(x - 5) >= (unsigned)(w - 10);   // cv2's own (unsigned) test
  const float bwx = border_w(x, w);
  int uxa = 0, uxb = 0;
  float ufx = 0.f;
  if (MODE == 2) { uxa = a.ux0[x]; uxb = a.ux1[x]; ufx = a.ufx[x]; }

  // ---- step C state: running vertical sums of output column t (t < TW) ----
  float2 vxy = make_float2(0.f, 0.f), vzw = vxy;
  float ve = 0.f;

  int off = 0;                                          // physical ring row of logical row 0
  for (int s = 0;; ++s) {
    const int yb = ys + s * RB;                         // first output row of this block
    if (yb >= ye) break;
    const int lstart = s == 0 ? 0 : 2 * M;              // logical rows [lstart, NR) are new; row l = image row yb - M + l
    const int nrows = NR - lstart;

    // ---- L2 prefetch of the next block's new rows by the copy engine (no LSU wavefronts): one bulk prefetch per
    // (row, stream), issued by one thread each; R1 is prefetched at the undisplaced position (the flow moves the
    // gather by a few rows, which the neighbouring blocks' prefetches cover)
    if (FBS_BULK_PF) {
      const int nrow_pf = s == 0 ? NR + RB : RB;        // first block: its own rows too
      const int np = nrow_pf * 5;
      for (int u = t; u < np; u += FBS_NT) {
        const int r = u / 5, k = u - r * 5;
        const int yy = min(max(yb - M + (s == 0 ? r : NR + r), 0), h - 1);
        if (yy >= ye + M + RB) continue;
        const int xs = max(x0 - FBS_PADL, 0);
        const int cols = min(EW, pitch - xs);
        const size_t o = (size_t)yy * pitch + xs;
        const void* p;
        int bytes;
        if (k == 0) { p = R0a + o; bytes = cols * 16; }
        else if (k == 1) { p = R1a + o; bytes = cols * 16; }
        else if (k == 2) { p = R0b + o; bytes = cols * 4; }
        else if (k == 3) { p = R1b + o; bytes = cols * 4; }
        else { if (MODE != 1) continue; p = fin + o; bytes = cols * 8; }
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
      }
    }

    // ---- step A: M on the new rows.  thread = (halo column, run of consecutive rows); software-pipelined one row
    // deep: the gather and R0 loads of row k + 1 (and the flow vector of row k + 2) are in flight while row k is
    // consumed, so memory latency overlaps the arithmetic of the same warp.  Walking down a column the bottom
    // corners of row k are the top corners of row k + 1 whenever the integer part of the warp advanced by exactly
    // one row (almost always); they are folded into a 5-value partial sum for row k + 1 as soon as they arrive, so
    // only that partial sum is carried.
    {
      const int RS = (nrows + FBS_RUNS - 1) / FBS_RUNS;
      const int l0 = lstart + run * RS;
      const int n = min(l0 + RS, NR) - l0;              // rows of this thread's run
      if (FBS_EXP != 4 && n > 0) {
        int pr = l0 + off;
        if (pr >= NR) pr -= NR;
        unsigned sa = (unsigned)(pr * ES + cx);         // element index into the planes
        const unsigned sa_end = (unsigned)(NR * ES + cx);
        int yu = yb - M + l0;                           // unclamped image row of the row being set up
        const float4* r0a = R0a; const float* r0b = R0b; const float4* r1a = R1a; const float* r1b = R1b;
        const float2* fi = fin;
        int wm1 = w - 1, hm1 = h - 1, pit = pitch;
        int pitb = h > 1 ? pitch : 0;                   // keeps the unused bottom-corner loads in bounds
        int thr = xb_border ? 0 : h - 10;               // (unsigned)(y - 5) >= thr  <=>  border pixel
        pin(r0a); pin(r0b); pin(r1a); pin(r1b); pin(fi); pin(wm1); pin(hm1); pin(pit); pin(pitb); pin(thr);

        // stage 1: everything of a row that only needs its flow vector: addresses, weights, loads in flight
        auto stage1 = [&](FbRow& R, int o) {
          R.q = ldg_f4<0>(r0a + o);
          R.q4 = ldg_f1<0>(r0b + o);
          float fx = xf + R.d.x, fy = (float)R.y + R.d.y;
          const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
          R.fx = fx - (float)x1; R.fy = fy - (float)y1;
          R.inside = (unsigned)x1 < (unsigned)wm1 && (unsigned)y1 < (unsigned)hm1;
          R.ot = R.inside ? y1 * pit + x1 : 0;
          const int ob = R.ot + pitb;
          R.bot.a0 = ldg_f4<0>(r1a + ob); R.bot.a1 = ldg_f4<16>(r1a + ob);
          R.bot.e0 = ldg_f1<0>(r1b + ob); R.bot.e1 = ldg_f1<4>(r1b + ob);
        };
        // top half of the bilinear sum of row R from the corner pair c
        auto top_partial = [&](const FbRow& R, const FbCorner& c, float (&tp)[5]) {
          const float gy = 1.f - R.fy;
          const float a00 = (1.f - R.fx) * gy, a01 = R.fx * gy;
          tp[0] = fmaf(a01, c.a1.x, a00 * c.a0.x);
          tp[1] = fmaf(a01, c.a1.y, a00 * c.a0.y);
          tp[2] = fmaf(a01, c.a1.z, a00 * c.a0.z);
          tp[3] = fmaf(a01, c.a1.w, a00 * c.a0.w);
          tp[4] = fmaf(a01, c.e1, a00 * c.e0);
        };
        auto load_top = [&](const FbRow& R, float (&tp)[5]) {
          FbCorner c;
          c.a0 = ldg_f4<0>(r1a + R.ot); c.a1 = ldg_f4<16>(r1a + R.ot);
          c.e0 = ldg_f1<0>(r1b + R.ot); c.e1 = ldg_f1<4>(r1b + R.ot);
          top_partial(R, c, tp);
        };
        // stage 2: consume a row whose loads have landed
        auto stage2 = [&](const FbRow& R, const float (&tp)[5]) {
          const float a10 = (1.f - R.fx) * R.fy, a11 = R.fx * R.fy;
          float r2 = fmaf(a11, R.bot.a1.x, fmaf(a10, R.bot.a0.x, tp[0]));
          float r3 = fmaf(a11, R.bot.a1.y, fmaf(a10, R.bot.a0.y, tp[1]));
          float r4 = fmaf(a11, R.bot.a1.z, fmaf(a10, R.bot.a0.z, tp[2]));
          float r5 = fmaf(a11, R.bot.a1.w, fmaf(a10, R.bot.a0.w, tp[3]));
          float r6 = fmaf(a11, R.bot.e1, fmaf(a10, R.bot.e0, tp[4]));
          const float4 q = R.q;
          const float q4 = R.q4;
          r2 = R.inside ? r2 : 0.f;
          r3 = R.inside ? r3 : 0.f;
          r4 = R.inside ? r4 : q.z;               // (q + q) * 0.5 = q, (q4 + q4) * 0.25 = q4 * 0.5: exact
          r5 = R.inside ? r5 : q.w;
          r6 = R.inside ? r6 : q4;
          r4 = (q.z + r4) * 0.5f;
          r5 = (q.w + r5) * 0.5f;
          r6 = (q4 + r6) * 0.25f;
          r2 = (q.x - r2) * 0.5f;
          r3 = (q.y - r3) * 0.5f;
          r2 += r4 * R.d.y + r6 * R.d.x;
          r3 += r6 * R.d.y + r5 * R.d.x;
          if ((unsigned)(R.y - 5) >= (unsigned)thr) {
            const float sc = bwx * border_w(R.y, h);
            r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
          }
          sts_f2(s_xy + sa * 8, r4 * r4 + r6 * r6, (r4 + r5) * r6);
          sts_f2(s_zw + sa * 8, r5 * r5 + r6 * r6, r4 * r2 + r6 * r3);
          sts_f1(s_e + sa * 4, r6 * r2 + r5 * r3);
          sa += ES;
          if (sa == sa_end) sa -= NR * ES;
        };
        auto next_row = [&](int& y, int& o) {          // image row / element offset of the next row to set up
          y = min(max(yu, 0), hm1);
          o = y * pit + x;
          ++yu;
        };

        FbRow A, B;
        float tp[5];
        int oA, oB = 0, yn, on;
        float2 dn = make_float2(0.f, 0.f);
        next_row(A.y, oA);
        A.d = fetch_flow_m<MODE>(a, fi, oA, A.y, uxa, uxb, ufx);
        B.y = A.y; B.d = A.d;
        if (n > 1) { next_row(B.y, oB); B.d = fetch_flow_m<MODE>(a, fi, oB, B.y, uxa, uxb, ufx); }
        stage1(A, oA);
        load_top(A, tp);
        int k = 0;
        for (; k + 1 < n; k += 2) {
          // rows k (in A, loads in flight) and k + 1 (in B, flow vector in flight)
          stage1(B, oB);
          if (k + 2 < n) { next_row(yn, on); dn = fetch_flow_m<MODE>(a, fi, on, yn, uxa, uxb, ufx); }
          stage2(A, tp);
          if (B.ot == A.ot + pitb) top_partial(B, A.bot, tp); else load_top(B, tp);
          A.d = dn; A.y = yn; oA = on;
          if (k + 2 < n) stage1(A, oA);
          if (k + 3 < n) { next_row(yn, on); dn = fetch_flow_m<MODE>(a, fi, on, yn, uxa, uxb, ufx); }
          stage2(B, tp);
          if (k + 2 < n) { if (A.ot == B.ot + pitb) top_partial(A, B.bot, tp); else load_top(A, tp); }
          B.d = dn; B.y = yn; oB = on;
        }
        if (k < n) stage2(A, tp);
      }
    }
    __syncthreads();

    // ---- step B: horizontal 15-sums in place.  Window positions p = cx - 1: output xo sums p in [xo, xo + 14].
    //   left half : outputs [0,HL)  walked left->right, result stored at p = xo       (reads p >= xo)
    //   right half: outputs [HL,TW) walked right->left, result stored at p = xo + 14  (reads p <= xo + 14)
    // positions [HL, HL + 14) are written by neither half.  thread = (plane, row, half); a warp holds 16 rows x 2
    // halves, which the odd row stride spreads over all banks.
    if (FBS_EXP != 3 && FBS_EXP != 5 && t < 192) {
      const int plane = t >> 6, rg = (t >> 5) & 1, lane = t & 31;
      const int r = rg * 16 + (lane & 15);
      const bool right = (lane >> 4) != 0;
      if (r < nrows) {
        int pr = lstart + r + off;
        if (pr >= NR) pr -= NR;
        if (plane < 2) {
          float2* rowp = (plane ? Pzw : Pxy) + pr * ES + 1;        // &P[pr][p = 0]
          if (!right) {
            float2 sm = rowp[0];
#pragma unroll
            for (int k = 1; k < 2 * M; ++k) sm = add2(sm, rowp[k]);
#pragma unroll 8
            for (int xo = 0; xo < HL; ++xo) {
              sm = add2(sm, rowp[xo + 2 * M]);
              const float2 old = rowp[xo];
              rowp[xo] = sm;
              sm = sub2(sm, old);
            }
          } else {
            float2 sm = rowp[TW];
#pragma unroll
            for (int k = 1; k < 2 * M; ++k) sm = add2(sm, rowp[TW + k]);
#pragma unroll 8
            for (int xo = TW - 1; xo >= HL; --xo) {
              sm = add2(sm, rowp[xo]);
              const float2 old = rowp[xo + 2 * M];
              rowp[xo + 2 * M] = sm;
              sm = sub2(sm, old);
            }
          }
        } else {
          float* rowp = Pe + pr * ES + 1;
          if (!right) {
            float sm = rowp[0];
#pragma unroll
            for (int k = 1; k < 2 * M; ++k) sm += rowp[k];
#pragma unroll 8
            for (int xo = 0; xo < HL; ++xo) {
              sm += rowp[xo + 2 * M];
              const float old = rowp[xo];
              rowp[xo] = sm;
              sm -= old;
            }
          } else {
            float sm = rowp[TW];
#pragma unroll
            for (int k = 1; k < 2 * M; ++k) sm += rowp[TW + k];
#pragma unroll 8
            for (int xo = TW - 1; xo >= HL; --xo) {
              sm += rowp[xo];
              const float old = rowp[xo + 2 * M];
              rowp[xo + 2 * M] = sm;
              sm -= old;
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- step C: vertical running 15-sums + 2x2 solve; thread = output column ----
    if (FBS_EXP != 3 && FBS_EXP != 6 && t < TW && x0 + t < w) {
      const int col = 1 + (t < HL ? t : t + 2 * M);                      // where step B left this column's sums
      int pn = off;                                                      // physical row of logical row 0
      if (s == 0 || ((yb / RB) % FBS_REFRESH) == 0) {   // absolute block index: results do not depend on the segmentation
        // (re)start the running sums from the 14 carried rows: logical rows 0..13
        vxy = make_float2(0.f, 0.f); vzw = vxy; ve = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * M; ++k) {
          const int e = pn * ES + col;
          vxy = add2(vxy, Pxy[e]);
          vzw = add2(vzw, Pzw[e]);
          ve += Pe[e];
          if (++pn == NR) pn = 0;
        }
      } else {
        pn += 2 * M;
        if (pn >= NR) pn -= NR;
      }
      int po = off;                                                      // oldest row of the window
      const int gx = x0 + t;
      const int nr = min(RB, ye - yb);
      float2* orow = fo + (size_t)yb * a.out_pitch + gx;
      // (g11*g22 - g12^2 + 1e-3) with g = v * inv_area: the common factor inv_area^2 is moved into the constant
      const float eps = 1e-3f / (a.inv_area * a.inv_area);
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        const int en = pn * ES + col, eo = po * ES + col;
        vxy = add2(vxy, Pxy[en]);
        vzw = add2(vzw, Pzw[en]);
        ve += Pe[en];
        const float g11 = vxy.x, g12 = vxy.y, g22 = vzw.x, h1 = vzw.y, h2 = ve;
        const float idet = __frcp_rn(g11 * g22 - g12 * g12 + eps);
        *orow = make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
        orow += a.out_pitch;
        vxy = sub2(vxy, Pxy[eo]);
        vzw = sub2(vzw, Pzw[eo]);
        ve -= Pe[eo];
        if (++pn == NR) pn = 0;
        if (++po == NR) po = 0;
      }
    }
    off += RB;
    if (off >= NR) off -= NR;
    __syncthreads();
  }
}
