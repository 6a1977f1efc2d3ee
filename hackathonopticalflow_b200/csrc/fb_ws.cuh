// K5/K6 for the reference's window (winsize 15, box): ONE fused kernel per (level, iteration) --
//   flow_in (zero | previous iteration | bilinear x(1/pyr_scale) upsample of the coarser level)
//   -> UpdateMatrices (bilinear warp of R1, border attenuation) -> 15x15 box sum -> 2x2 solve -> flow_out
// M (the 5-channel matrix field) never leaves the SM.
//
// Shape: one CTA per SM walks down a strip of 112 output columns (128 halo columns = 4 whole warps per row, so
// every row load falls on whole 128-byte lines) in blocks of FBW_RB rows.  The three steps of a block run on
// DIFFERENT warps, on different blocks, at the same time:
//
//   A warps (16): UpdateMatrices for block s + 2 -> ring rows.  thread = (halo column, run of 4 rows); walking
//                 down a column the bottom corners of one pixel's bilinear gather are the top corners of the next
//                 one whenever the integer part of the warp advanced by exactly one row (almost always), so they
//                 are carried in registers and only two new corners are loaded; the flow vector is fetched one
//                 row ahead.  Memory-latency bound.
//   B warps  (4): horizontal 15-sums in place on block s + 1.  lane = (row, half-row); each half-row is walked
//                 once with the 15 most recent inputs in a register window (one read and one write per element).
//   C warps  (4): vertical 15-sums for block s as running sums carried in registers down the whole strip (one
//                 add and one subtract per row), restarted from the ring every FBW_REFRESH blocks so rounding does
//                 not accumulate; 2x2 solve; coalesced 8-byte stores.  The C warps also feed the copy engine: bulk
//                 L2 prefetches (cp.async.bulk.prefetch.L2) of the rows step A needs FBW_PF_BLOCKS blocks later.
//
// so the serial row / column walks of B and C hide behind the gathers of A instead of stalling the whole CTA at a
// barrier (with barrier-separated phases 44 % of all warp time was barrier wait).  Blocks are handed from stage to
// stage through mbarriers (full_a[], full_b[], empty_c[]; stage = block mod FBW_STAGES).  The M ring holds
// 14 + FBW_STAGES * FBW_RB rows as three planes (float2 ch0/1, float2 ch2/3, float ch4) whose odd row stride makes the
// row-parallel accesses of B and the column-parallel accesses of A and C conflict-free.
//
// What bounds it (ncu, profiles/README.md): the L1 / shared-memory data pipe.  Every global or shared access costs
// one wavefront per 128 bytes it touches and the pipe sustains about 0.6 wavefronts per cycle per SM; this kernel
// needs 2.9 wavefronts per output pixel (1.4 global: the unaligned 16-byte corner gathers; 1.5 shared), its
// predecessor with 56x56 tiles and barrier-separated phases needed 3.6.
#pragma once

constexpr int FBS_EW = 128;               // halo columns per strip
constexpr int FBS_TW = 112;               // output columns per strip
constexpr int FBS_M = 7;                  // window radius
constexpr int FBS_PADL = 8;               // halo column 0 sits at image column x0 - 8 (8-pixel aligned)
constexpr int FBS_ES = 129;               // plane row stride in elements (odd)
constexpr int FBS_HL = FBS_TW / 2;        // outputs [0,HL) are summed left->right, [HL,TW) right->left
constexpr int FBW_REFRESH = 4;            // blocks between restarts of the vertical running sums
constexpr int FBW_A_WARPS = 16;
#ifndef FBW_PF_WARP
#define FBW_PF_WARP 5   // who issues the L2 prefetches: 5 one lane per C warp, uniform addresses (shipped); 0 five lanes per C
                        // warp (round 1: a per-lane serialisation loop, 45 instructions per prefetch); 1 / 2 the spare B warp
                        // (bulk / per line); 3 nobody;
#endif                  // 4 a 25th warp of its own (bulk), paced by step A's progress word
constexpr int FBW_NT = (FBW_A_WARPS + 8) * 32 + (FBW_PF_WARP == 4 ? 32 : 0);   // 16 A warps, 4 B warps, 4 C warps [+ 1]
constexpr int FBW_RUNS = FBW_A_WARPS / 4;               // row runs per block in step A (128 columns = 4 warps each)
constexpr int FBW_RB = FBW_A_WARPS;                     // rows per block: 4 rows per A thread
constexpr int FBW_PF_BLOCKS = 4;                        // L2 prefetch distance in blocks
#ifndef FBW_SKIP
#define FBW_SKIP 0   // timing experiments only: bit 0 / 1 / 2 switches step A / B / C off
#endif
#ifndef FBW_STAGES_N
#define FBW_STAGES_N 3
#endif
#ifndef FBW_A2
#define FBW_A2 0        // 1: step A software-pipelined two rows deep (register budget moved from B / C warps with setmaxnreg)
#endif
#ifndef FBW_REGS_A
#define FBW_REGS_A 88   // FBW_A2: registers per thread after setmaxnreg (A inc, B / C dec); 512 A + 128 B + 128 C
#define FBW_REGS_B 56   // must add up to the 768 x 80 the CTA is launched with
#define FBW_REGS_C 72
#endif
#ifndef FBW_TRYWAIT_HINT_NS
#define FBW_TRYWAIT_HINT_NS 0   // != 0: mbarrier.try_wait with this suspend-time hint instead of nanosleep polling
#endif
#ifndef FBW_SLEEP_NS
#define FBW_SLEEP_NS 200   // back-off between mbarrier polls
#endif
#ifndef FBW_PF_LEAD
#define FBW_PF_LEAD (FBW_PF_WARP == 4 ? 3 : 5)   // the prefetching warp, once block s is in the ring, fetches block s + LEAD
#endif
constexpr int FBW_STAGES = FBW_STAGES_N;                // blocks in flight between step A and step C
constexpr int FBW_NR = 2 * FBS_M + FBW_STAGES * FBW_RB; // ring rows
constexpr size_t FBW_PLANES = (size_t)FBW_NR * FBS_ES * 20;
constexpr size_t FBW_SMEM = FBW_PLANES + 3 * FBW_STAGES * 8 + 16;   // + mbarriers + step A's progress word

__device__ __forceinline__ void sts_f2(unsigned addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }

__device__ __forceinline__ void mbar_init(unsigned addr, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
#if FBW_TRYWAIT_HINT_NS
  // hardware-suspended wait: try_wait parks the warp until the phase completes or the hint (ns) expires
  asm volatile(
      "{ .reg .pred p;\n"
      "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@!p bra W_%=;\n"
      "}" ::"r"(addr), "r"(parity), "r"(FBW_TRYWAIT_HINT_NS) : "memory");
#else
  // poll with back-off: a spinning warp would take issue slots from the warps it is waiting for
  asm volatile(
      "{ .reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "W_%=: nanosleep.u32 %2;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W_%=;\n"
      "D_%=: }" ::"r"(addr), "r"(parity), "n"(FBW_SLEEP_NS) : "memory");
#endif
}

__device__ __forceinline__ float rcp_approx(float x) {     // 1 ulp; the determinant is >= 1e-3 / inv_area^2 > 0
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- step B: horizontal 15-sums of one half ring row of one plane, in place.  Window positions p = cx - 1:
// output xo sums p in [xo, xo + 14].  The left half (outputs [0,HL)) is walked left->right and stored at p = xo
// (reads p >= xo); the right half (outputs [HL,TW)) right->left, stored at p = xo + 14 (reads p <= xo + 14);
// positions [HL, HL + 14) are written by neither, so the two halves never race.
template <typename T>
__device__ __forceinline__ T hp_add(T a, T b);
template <> __device__ __forceinline__ float2 hp_add<float2>(float2 a, float2 b) { return add2(a, b); }
template <> __device__ __forceinline__ float hp_add<float>(float a, float b) { return a + b; }
template <typename T>
__device__ __forceinline__ T hp_sub(T a, T b);
template <> __device__ __forceinline__ float2 hp_sub<float2>(float2 a, float2 b) { return sub2(a, b); }
template <> __device__ __forceinline__ float hp_sub<float>(float a, float b) { return a - b; }

// The 15 most recent inputs stay in a register window (circular, statically indexed: the walk is unrolled by the
// window length), so every element is read from shared memory once and written once: 2 accesses per element
// instead of 3 on the unit that bounds the kernel.
template <typename T, bool FULL>
__device__ __forceinline__ void hpass_half_row(T* rowp, bool right, int nv) {
  // nv = outputs this strip really has (FULL: the whole strip, compile-time bounds; otherwise the last strip of a
  // row: nothing beyond its outputs is summed)
  constexpr int M = FBS_M, HL = FBS_HL, WN = 2 * M + 1;
  const int TW = FULL ? FBS_TW : nv;
  T win[WN];
  if (!right) {
    const int HLv = FULL ? HL : min(HL, nv);
    // outputs xo = 0 .. HL-1: sum of positions [xo, xo + 14], stored at position xo
    T sm = rowp[0];
    win[0] = sm;
#pragma unroll
    for (int k = 1; k < 2 * M; ++k) { win[k] = rowp[k]; sm = hp_add(sm, win[k]); }
    for (int xb = 0; xb < HLv; xb += WN) {
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int xo = xb + j;
        if (xo < HLv) {
          // slot (j + 14) % 15 receives position xo + 14; slot j holds position xo (the one leaving the window)
          const T nw = rowp[xo + 2 * M];
          sm = hp_add(sm, nw);
          rowp[xo] = sm;
          sm = hp_sub(sm, win[j]);
          win[(j + 2 * M) % WN] = nw;
        }
      }
    }
  } else if (TW > HL) {
    // outputs xo = TW-1 .. HL: sum of positions [xo, xo + 14], stored at position xo + 14; walk right to left:
    // mirrored index u = TW - 1 - xo, position q(u, k) = TW - 1 + 14 - u - k
    T* top = rowp + (TW - 1 + 2 * M);                   // position of the right-most input
    T sm = top[0];
    win[0] = sm;
#pragma unroll
    for (int k = 1; k < 2 * M; ++k) { win[k] = top[-k]; sm = hp_add(sm, win[k]); }
    for (int ub = 0; ub < TW - HL; ub += WN) {
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int u = ub + j;
        if (u < TW - HL) {
          const T nw = top[-(u + 2 * M)];
          sm = hp_add(sm, nw);
          top[-u] = sm;
          sm = hp_sub(sm, win[j]);
          win[(j + 2 * M) % WN] = nw;
        }
      }
    }
  }
}

template <int MODE, bool STATS>
__global__ void __launch_bounds__(FBW_NT, 1) fb_iter_ws(IterArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int EW = FBS_EW, TW = FBS_TW, M = FBS_M, RB = FBW_RB, NR = FBW_NR, ES = FBS_ES, HL = FBS_HL;
  float2* Pxy = (float2*)smem;                          // float2 [NR][ES]  (M0, M1)
  float2* Pzw = Pxy + NR * ES;                          // float2 [NR][ES]  (M2, M3)
  float* Pe = (float*)(Pzw + NR * ES);                  // float  [NR][ES]  (M4)
  const unsigned s_xy = smem_u32(smem);
  const unsigned s_zw = s_xy + NR * ES * 8;
  const unsigned s_e = s_zw + NR * ES * 8;
  const unsigned s_bar = s_xy + (unsigned)FBW_PLANES;   // full_a[3], full_b[3], empty_c[3]
  const unsigned s_prog = s_bar + 3 * FBW_STAGES * 8;   // blocks step A (its first warp) has put into the ring
  const int pair = blockIdx.x;                          // pair index fastest (frame p + 1 shared through L2)
  const int x0 = blockIdx.y * TW;
  const int w = a.w, h = a.h, pitch = a.pitch;
  const int ys = blockIdx.z * a.nb * RB;                // rows [ys, ye) are this CTA's outputs
  const int ye = min(ys + a.nb * RB, h);
  const int nblk = (ye - ys + RB - 1) / RB;
  const int t = threadIdx.x;
#ifndef FBW_TRIM
#define FBW_TRIM 1
#endif
  const int nv = FBW_TRIM ? min(TW, w - x0) : TW;       // outputs of this strip (the last strip of a row is narrower)

  if (t == 0) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(s_prog), "r"(0) : "memory");
    for (int i = 0; i < FBW_STAGES; ++i) {
      mbar_init(s_bar + i * 8, FBW_A_WARPS * 32);       // full_a: every A thread arrives
      mbar_init(s_bar + (FBW_STAGES + i) * 8, 128);     // full_b
      mbar_init(s_bar + (2 * FBW_STAGES + i) * 8, 128); // empty_c
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (t < FBW_A_WARPS * 32) {
    // =========================== A warps: UpdateMatrices ===========================
    // R of the pair's first frame: {float4 plane ch0..3, float plane ch4}; the second frame follows at r_frame_stride
    const float* base0 = a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride;
    const float2* __restrict__ fin = MODE ? a.flow_in + (size_t)pair * a.flow_in_pair_stride : nullptr;
    const int cx = t & (EW - 1), run = t >> 7;
    // halo columns [PADL - M, PADL + nv + M) feed this strip's outputs: a warp whose 32 columns lie beyond them only
    // passes the barriers on (last strip of a row: 16 of 128 columns at 1080p, a quarter of the work at 240 x 135)
    const bool a_on = (cx & ~31) < FBS_PADL + nv + FBS_M;
    const int x = clampi(x0 - FBS_PADL + cx, 0, w - 1);
    const float xf = (float)x;
    const bool xb_border = (unsigned)(x - 5) >= (unsigned)(w - 10);   // cv2's own (unsigned) test
    const float bwx = border_w(x, w);
    int uxa = 0, uxb = 0;
    float ufx = 0.f;
    if (MODE == 2) { uxa = a.ux0[x]; uxb = a.ux1[x]; ufx = a.ufx[x]; }
    // one 64-bit base (the pair's R0 record plane) and 32-bit byte offsets to the other three planes; loop
    // invariants are pinned in registers (ptxas otherwise re-derives them from the constant bank every row)
    const char* rb = (const char*)base0;
    unsigned c_r0b = (unsigned)(16 * a.plane_stride);                       // R0 ch4 plane
    unsigned c_r1a = (unsigned)(4 * a.r_frame_stride);                      // R1 record plane
    unsigned c_r1b = c_r1a + c_r0b;                                         // R1 ch4 plane
    const float2* fi = fin;
    int wm1 = w - 1, hm1 = h - 1, pit = pitch;
    int pitb = h > 1 ? pitch : 0;                       // keeps the unused bottom-corner loads in bounds
    int thr = xb_border ? 0 : h - 10;                   // (unsigned)(y - 5) >= thr  <=>  border pixel
    pin(rb); pin(c_r0b); pin(c_r1a); pin(c_r1b); pin(fi); pin(wm1); pin(hm1); pin(pit); pin(pitb); pin(thr);

#if FBW_A2
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FBW_REGS_A));
    // Software pipeline over the rows of a thread's run, in groups of up to four rows:
    //   F  the group's flow vectors (loaded while the previous group is still being consumed)
    //   G  row k: warp address, loads of the two bottom corners (top corners only for the first row of a group or
    //      when the carry fails), of the R0 record and of its fifth channel
    //   S  row k: bilinear blend, UpdateMatrices, three shared-memory stores
    // issued as G0 G1 S0 G2 S1 G3 S2 S3: the loads of row k + 1 are in flight while row k is consumed.  Three corner
    // sets rotate (top of row k + 1 = bottom of row k whenever the warp moved down by exactly one row); when the
    // carry fails the top corners are reloaded into the set row k has just released.
    int j0 = 0;
    float2 dn[4];                                       // flow vectors of the next group
    auto load_flows = [&](int yu0, int cnt) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yk = min(max(yu0 + min(k, cnt - 1), 0), hm1);
        dn[k] = fetch_flow_m<MODE>(a, fi, yk * pit + x, yk, uxa, uxb, ufx);
      }
    };
    {
      const int nrows0 = 2 * M + RB, RS0 = (nrows0 + FBW_RUNS - 1) / FBW_RUNS;
      const int n0 = min(run * RS0 + RS0, nrows0) - run * RS0;
      if (n0 > 0) load_flows(ys - M + run * RS0, min(4, n0));
    }
    for (int s = 0; s < nblk; ++s) {
      const int yb = ys + s * RB;
      const int nrows = s == 0 ? 2 * M + RB : RB;
      const int y_first = s == 0 ? yb - M : yb + M;
      const int RS = (nrows + FBW_RUNS - 1) / FBW_RUNS;
      const int l0 = run * RS;
      const int n = min(l0 + RS, nrows) - l0;
      bool waited = s < FBW_STAGES;                     // C must be done with block s - FBW_STAGES before the first store
      if (!(FBW_SKIP & 1) && n > 0 && a_on) {
        int pr = j0 + l0;
        if (pr >= NR) pr -= NR;
        unsigned sa = (unsigned)(pr * ES + cx);
        const unsigned sa_end = (unsigned)(NR * ES + cx);
        int yu = y_first + l0;
        for (int g0 = 0; g0 < n; g0 += 4) {
          const int gn = min(4, n - g0);
          float2 d0 = dn[0], d1 = dn[1], d2 = dn[2], d3 = dn[3];
          const int y0r = min(max(yu, 0), hm1), y1r = min(max(yu + min(1, gn - 1), 0), hm1);
          const int y2r = min(max(yu + min(2, gn - 1), 0), hm1), y3r = min(max(yu + min(3, gn - 1), 0), hm1);
          yu += gn;
          FbCorner c0, c1, c2;
          float4 qa, qb;
          float qa4, qb4;
          float fxa, fya, fxb, fyb;                     // bilinear fractions of the two rows in flight
          bool insa, insb, needb = false;
          int ot_b = 0, o_carry;
          auto ldc = [&](FbCorner& c, int o) {
            const float4* pa = (const float4*)(rb + ((unsigned)o * 16u + c_r1a));
            const float* pe = (const float*)(rb + ((unsigned)o * 4u + c_r1b));
            c.a0 = ldg_f4<0>(pa); c.a1 = ldg_f4<16>(pa); c.e0 = ldg_f1<0>(pe); c.e1 = ldg_f1<4>(pe);
          };
          // G: returns the R1 offset of the top-left corner; loads the bottom corners, q, q4
          auto G = [&](const float2 d, const int y, FbCorner& bot, float4& q, float& q4, float& fxo, float& fyo,
                       bool& ins) -> int {
            const int o = y * pit + x;
            q = ldg_f4<0>((const float4*)(rb + (unsigned)o * 16u));
            q4 = ldg_f1<0>((const float*)(rb + ((unsigned)o * 4u + c_r0b)));
            float fx = xf + d.x, fy = (float)y + d.y;
            const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
            fxo = fx - (float)x1; fyo = fy - (float)y1;
            ins = (unsigned)x1 < (unsigned)wm1 && (unsigned)y1 < (unsigned)hm1;
            const int ot = ins ? y1 * pit + x1 : 0;
            ldc(bot, ot + pitb);
            return ot;
          };
          auto S = [&](const float2 d, const int y, const FbCorner& top, const FbCorner& bot, const float4 q,
                       const float q4, float fx, float fy, const bool inside) {
            const float gx = 1.f - fx, gy = 1.f - fy;
            const float a00 = gx * gy, a01 = fx * gy, a10 = gx * fy, a11 = fx * fy;
            float r2 = fmaf(a11, bot.a1.x, fmaf(a10, bot.a0.x, fmaf(a01, top.a1.x, a00 * top.a0.x)));
            float r3 = fmaf(a11, bot.a1.y, fmaf(a10, bot.a0.y, fmaf(a01, top.a1.y, a00 * top.a0.y)));
            float r4 = fmaf(a11, bot.a1.z, fmaf(a10, bot.a0.z, fmaf(a01, top.a1.z, a00 * top.a0.z)));
            float r5 = fmaf(a11, bot.a1.w, fmaf(a10, bot.a0.w, fmaf(a01, top.a1.w, a00 * top.a0.w)));
            float r6 = fmaf(a11, bot.e1, fmaf(a10, bot.e0, fmaf(a01, top.e1, a00 * top.e0)));
            r2 = inside ? r2 : 0.f;
            r3 = inside ? r3 : 0.f;
            r4 = inside ? r4 : q.z;
            r5 = inside ? r5 : q.w;
            r6 = inside ? r6 : q4;
            r4 = (q.z + r4) * 0.5f;
            r5 = (q.w + r5) * 0.5f;
            r6 = (q4 + r6) * 0.25f;
            r2 = (q.x - r2) * 0.5f;
            r3 = (q.y - r3) * 0.5f;
            r2 += r4 * d.y + r6 * d.x;
            r3 += r6 * d.y + r5 * d.x;
            if ((unsigned)(y - 5) >= (unsigned)thr) {
              const float sc = bwx * border_w(y, h);
              r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
            }
            if (!waited) {
              mbar_wait(s_bar + (2 * FBW_STAGES + s % FBW_STAGES) * 8, (unsigned)((s / FBW_STAGES - 1) & 1));
              waited = true;
            }
            sts_f2(s_xy + sa * 8, r4 * r4 + r6 * r6, (r4 + r5) * r6);
            sts_f2(s_zw + sa * 8, r5 * r5 + r6 * r6, r4 * r2 + r6 * r3);
            sts_f1(s_e + sa * 4, r6 * r2 + r5 * r3);
            sa += ES;
            if (sa == sa_end) sa -= NR * ES;
          };
          // G0 (top corners too), G1
          {
            const int ot = G(d0, y0r, c1, qa, qa4, fxa, fya, insa);
            ldc(c0, ot);
            o_carry = ot + pitb;
          }
          if (gn > 1) {
            ot_b = G(d1, y1r, c2, qb, qb4, fxb, fyb, insb);
            needb = ot_b != o_carry;
            o_carry = ot_b + pitb;
          }
          // the next group's flow vectors (this run, or the first group of this thread's run in the next block)
          if (g0 + 4 < n) load_flows(yu, min(4, n - g0 - 4));
          else if (s + 1 < nblk) load_flows(ys + (s + 1) * RB + M + run * (RB / FBW_RUNS), RB / FBW_RUNS);
          S(d0, y0r, c0, c1, qa, qa4, fxa, fya, insa);
          if (gn > 1) {
            if (needb) ldc(c1, ot_b);
            if (gn > 2) {
              ot_b = G(d2, y2r, c0, qa, qa4, fxa, fya, insa);
              needb = ot_b != o_carry;
              o_carry = ot_b + pitb;
            }
            S(d1, y1r, c1, c2, qb, qb4, fxb, fyb, insb);
            if (gn > 2) {
              if (needb) ldc(c2, ot_b);
              if (gn > 3) {
                ot_b = G(d3, y3r, c1, qb, qb4, fxb, fyb, insb);
                needb = ot_b != o_carry;
              }
              S(d2, y2r, c2, c0, qa, qa4, fxa, fya, insa);
              if (gn > 3) {
                if (needb) ldc(c0, ot_b);
                S(d3, y3r, c0, c1, qb, qb4, fxb, fyb, insb);
              }
            }
          }
        }
      } else if (s + 1 < nblk && n <= 0) {
        load_flows(ys + (s + 1) * RB + M + run * (RB / FBW_RUNS), RB / FBW_RUNS);
      }
      if (!waited) mbar_wait(s_bar + (2 * FBW_STAGES + s % FBW_STAGES) * 8, (unsigned)((s / FBW_STAGES - 1) & 1));
      mbar_arrive(s_bar + (s % FBW_STAGES) * 8);        // full_a[stage]: block s is in the ring
      if (FBW_PF_WARP == 4 && t == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(s_prog), "r"(s + 1) : "memory");
      j0 += nrows;
      if (j0 >= NR) j0 -= NR;
    }
#else
    int j0 = 0;                                         // ring row (mod NR) of the first new M row of the block
    for (int s = 0; s < nblk; ++s) {
      const int yb = ys + s * RB;
      const int nrows = s == 0 ? 2 * M + RB : RB;       // M rows [y_first, y_first + nrows), y_first below
      const int y_first = s == 0 ? yb - M : yb + M;     // image row of the first new M row
      if (s >= FBW_STAGES)                              // C must be done with block s - FBW_STAGES
        mbar_wait(s_bar + (2 * FBW_STAGES + s % FBW_STAGES) * 8, (unsigned)((s / FBW_STAGES - 1) & 1));

      const int RS = (nrows + FBW_RUNS - 1) / FBW_RUNS;
      const int l0 = run * RS;
      const int n = min(l0 + RS, nrows) - l0;           // rows of this thread's run
      if (!(FBW_SKIP & 1) && n > 0 && a_on) {
        int pr = j0 + l0;
        if (pr >= NR) pr -= NR;
        unsigned sa = (unsigned)(pr * ES + cx);         // element index into the planes
        const unsigned sa_end = (unsigned)(NR * ES + cx);
        int yu = y_first + l0;                          // unclamped image row of the row being set up

        auto next_row = [&](int& y, int& o) {
          y = min(max(yu, 0), hm1);
          o = y * pit + x;
          ++yu;
        };

        // plain form (64 registers): a row's loads are issued and consumed in the same iteration; only the flow
        // vector is fetched one row ahead; bottom corners carried as the next row's top corners
        FbCorner cA, cB;
        cA.a0 = cA.a1 = make_float4(0.f, 0.f, 0.f, 0.f);
        cA.e0 = cA.e1 = 0.f;
        cB = cA;
        int o_carry = -1 << 30;
        int yA, oA, yB = 0, oB = 0;
        next_row(yA, oA);
        float2 dA = fetch_flow_m<MODE>(a, fi, oA, yA, uxa, uxb, ufx), dB = dA;
        auto rowf = [&](const float2 d, float2& dn, const int y, int& yn, const int o, int& on, bool has_next,
                        FbCorner& top, FbCorner& bot) {
          const float4 q = ldg_f4<0>((const float4*)(rb + (unsigned)o * 16u));
          const float q4 = ldg_f1<0>((const float*)(rb + ((unsigned)o * 4u + c_r0b)));
          if (has_next) { next_row(yn, on); dn = fetch_flow_m<MODE>(a, fi, on, yn, uxa, uxb, ufx); }
          float fx = xf + d.x, fy = (float)y + d.y;
          const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
          fx -= (float)x1; fy -= (float)y1;
          const bool inside = (unsigned)x1 < (unsigned)wm1 && (unsigned)y1 < (unsigned)hm1;
          const int ot = inside ? y1 * pit + x1 : 0;
          if (ot != o_carry) {
            const float4* pa = (const float4*)(rb + ((unsigned)ot * 16u + c_r1a));
            const float* pe = (const float*)(rb + ((unsigned)ot * 4u + c_r1b));
            top.a0 = ldg_f4<0>(pa); top.a1 = ldg_f4<16>(pa); top.e0 = ldg_f1<0>(pe); top.e1 = ldg_f1<4>(pe);
          }
          const int ob = ot + pitb;
          {
            const float4* pa = (const float4*)(rb + ((unsigned)ob * 16u + c_r1a));
            const float* pe = (const float*)(rb + ((unsigned)ob * 4u + c_r1b));
            bot.a0 = ldg_f4<0>(pa); bot.a1 = ldg_f4<16>(pa); bot.e0 = ldg_f1<0>(pe); bot.e1 = ldg_f1<4>(pe);
          }
          o_carry = ob;
          const float gx = 1.f - fx, gy = 1.f - fy;
          const float a00 = gx * gy, a01 = fx * gy, a10 = gx * fy, a11 = fx * fy;
          float r2 = fmaf(a11, bot.a1.x, fmaf(a10, bot.a0.x, fmaf(a01, top.a1.x, a00 * top.a0.x)));
          float r3 = fmaf(a11, bot.a1.y, fmaf(a10, bot.a0.y, fmaf(a01, top.a1.y, a00 * top.a0.y)));
          float r4 = fmaf(a11, bot.a1.z, fmaf(a10, bot.a0.z, fmaf(a01, top.a1.z, a00 * top.a0.z)));
          float r5 = fmaf(a11, bot.a1.w, fmaf(a10, bot.a0.w, fmaf(a01, top.a1.w, a00 * top.a0.w)));
          float r6 = fmaf(a11, bot.e1, fmaf(a10, bot.e0, fmaf(a01, top.e1, a00 * top.e0)));
          r2 = inside ? r2 : 0.f;
          r3 = inside ? r3 : 0.f;
          r4 = inside ? r4 : q.z;
          r5 = inside ? r5 : q.w;
          r6 = inside ? r6 : q4;
          r4 = (q.z + r4) * 0.5f;
          r5 = (q.w + r5) * 0.5f;
          r6 = (q4 + r6) * 0.25f;
          r2 = (q.x - r2) * 0.5f;
          r3 = (q.y - r3) * 0.5f;
          r2 += r4 * d.y + r6 * d.x;
          r3 += r6 * d.y + r5 * d.x;
          if ((unsigned)(y - 5) >= (unsigned)thr) {
            const float sc = bwx * border_w(y, h);
            r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
          }
          sts_f2(s_xy + sa * 8, r4 * r4 + r6 * r6, (r4 + r5) * r6);
          sts_f2(s_zw + sa * 8, r5 * r5 + r6 * r6, r4 * r2 + r6 * r3);
          sts_f1(s_e + sa * 4, r6 * r2 + r5 * r3);
          sa += ES;
          if (sa == sa_end) sa -= NR * ES;
        };
        int k = 0;
        for (; k + 1 < n; k += 2) {
          rowf(dA, dB, yA, yB, oA, oB, true, cA, cB);
          rowf(dB, dA, yB, yA, oB, oA, k + 2 < n, cB, cA);
        }
        if (k < n) rowf(dA, dB, yA, yB, oA, oB, false, cA, cB);
      }
      mbar_arrive(s_bar + (s % FBW_STAGES) * 8);        // full_a[stage]: block s is in the ring
      if (FBW_PF_WARP == 4 && t == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(s_prog), "r"(s + 1) : "memory");
      j0 += nrows;
      if (j0 >= NR) j0 -= NR;
    }
#endif
  } else if (t < FBW_A_WARPS * 32 + 128) {
#if FBW_A2
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FBW_REGS_B));
#endif
    // =========================== B warps: horizontal sums in place ===========================
    const int bt = t - FBW_A_WARPS * 32, bw = bt >> 5, lane = bt & 31;
    const int rr = lane & 15;
    const bool right = (lane >> 4) != 0;
    int j0 = 0;
#if FBW_PF_WARP >= 1 && FBW_PF_WARP <= 3
    // The fourth B warp has no rows to sum in a steady-state block (3 planes -> 3 units): it feeds the L2 instead,
    // FBW_PF_LEAD blocks ahead of the block it is waiting for (step A is at most two blocks further on).
    const char* pf_r0 = (const char*)(a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride);
    const char* pf_fl = MODE == 1 ? (const char*)(a.flow_in + (size_t)pair * a.flow_in_pair_stride) : nullptr;
    const int pf_xs = max(x0 - FBS_PADL, 0);
    const int pf_cols = min(EW, pitch - pf_xs);
    const size_t pf_r1 = (size_t)4 * a.r_frame_stride, pf_e = (size_t)16 * a.plane_stride;
    auto prefetch_block_b = [&](int sb) {
      if (FBW_PF_WARP == 3 || bw != 3 || sb >= nblk) return;   // 3: no prefetch at all (timing experiments)
      const int nrows = sb == 0 ? 2 * M + RB : RB;
      const int y0 = sb == 0 ? ys - M : ys + sb * RB + M;
#if FBW_PF_WARP == 1
      if (lane == 0) {
        for (int r = 0; r < nrows; ++r) {
          const int yy = min(max(y0 + r, 0), h - 1);
          const size_t o = (size_t)yy * pitch + pf_xs;
          const char* p0 = pf_r0 + o * 16;
          const char* pe = pf_r0 + pf_e + o * 4;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"(pf_cols * 16) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0 + pf_r1), "r"(pf_cols * 16) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pe), "r"(pf_cols * 4) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pe + pf_r1), "r"(pf_cols * 4) : "memory");
          if (MODE == 1)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf_fl + o * 8), "r"(pf_cols * 8) : "memory");
        }
      }
#else
      // one 128-byte line per lane: instruction 1 = the two record planes (16 lines each), instruction 2 = the two
      // ch4 planes (4 lines each) and the flow row (8 lines)
      const int half = lane >> 4, l16 = lane & 15;
      const bool on1 = l16 * 8 < pf_cols;
      const char* q1 = pf_r0 + (half ? pf_r1 : 0) + (size_t)l16 * 128;
      const char* q2;
      bool on2;
      if (lane < 8) { q2 = pf_r0 + pf_e + (size_t)(lane & 3) * 128 + (lane >> 2 ? pf_r1 : 0); on2 = (lane & 3) * 32 < pf_cols; }
      else if (lane < 16) { q2 = pf_fl + (size_t)(lane - 8) * 128; on2 = MODE == 1 && (lane - 8) * 16 < pf_cols; }
      else { q2 = pf_r0; on2 = false; }
      for (int r = 0; r < nrows; ++r) {
        const int yy = min(max(y0 + r, 0), h - 1);
        const size_t o = (size_t)yy * pitch + pf_xs;
        if (on1) prefetch_l2(q1 + o * 16);
        if (on2) prefetch_l2(q2 + o * (lane < 8 ? 4 : 8));
      }
#endif
    };
    for (int sb = 0; sb < FBW_PF_LEAD; ++sb) prefetch_block_b(sb);
#endif
    for (int s = 0; s < nblk; ++s) {
      const int nrows = s == 0 ? 2 * M + RB : RB;
#if FBW_PF_WARP >= 1 && FBW_PF_WARP <= 3
      prefetch_block_b(s + FBW_PF_LEAD);
#endif
      mbar_wait(s_bar + (s % FBW_STAGES) * 8, (unsigned)((s / FBW_STAGES) & 1));   // full_a[stage]
      // work units = (group of 16 rows, plane), dealt round-robin to the four B warps; lane = row x half
      const int units = ((nrows + 15) >> 4) * 3;
      for (int u = bw; u < units; u += 4) {
        const int grp = u / 3, plane = u - grp * 3;
        const int r = grp * 16 + rr;
        if (!(FBW_SKIP & 2) && r < nrows) {
          int pr = j0 + r;
          if (pr >= NR) pr -= NR;
          if (nv == TW) {
            if (plane == 0) hpass_half_row<float2, true>(Pxy + pr * ES + 1, right, nv);
            else if (plane == 1) hpass_half_row<float2, true>(Pzw + pr * ES + 1, right, nv);
            else hpass_half_row<float, true>(Pe + pr * ES + 1, right, nv);
          } else {
            if (plane == 0) hpass_half_row<float2, false>(Pxy + pr * ES + 1, right, nv);
            else if (plane == 1) hpass_half_row<float2, false>(Pzw + pr * ES + 1, right, nv);
            else hpass_half_row<float, false>(Pe + pr * ES + 1, right, nv);
          }
        }
      }
      mbar_arrive(s_bar + (FBW_STAGES + s % FBW_STAGES) * 8);   // full_b[stage]
      j0 += nrows;
      if (j0 >= NR) j0 -= NR;
    }
  } else if (t < FBW_A_WARPS * 32 + 256) {
    // =========================== C warps: vertical running sums + solve ===========================
#if FBW_A2
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FBW_REGS_C));
#endif
    const int ct = t - FBW_A_WARPS * 32 - 128;
    const bool act = ct < TW && x0 + ct < w;
    const int col = 1 + (ct < HL ? ct : ct + 2 * M);    // where step B left this column's sums
    float2* __restrict__ fo = a.flow_out + (size_t)pair * a.flow_out_pair_stride;
    float2 vxy = make_float2(0.f, 0.f), vzw = vxy;
    float ve = 0.f;
    const float eps = 1e-3f / (a.inv_area * a.inv_area);
    // per-pair flow statistics folded into the last iteration (a.stats_acc != nullptr): every C thread sums its own
    // column, the CTA adds one set of fixed-point partials (same layout as flow_stats_accum, pathfinder.cu)
    constexpr bool do_stats = STATS;
    float st_m = 0.f, st_x = 0.f, st_y = 0.f, st_mx = 0.f;
    // the C warps also feed the copy engine: L2 prefetch of the rows step A will need FBW_PF_BLOCKS blocks from now,
    // one bulk prefetch per (row, stream), issued by lanes 0..4 of each C warp for RB / 4 rows
    const float* pbase0 = a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride;
    const int pf_xs = max(x0 - FBS_PADL, 0);
    const int pf_cols = min(EW, pitch - pf_xs);
#if FBW_PF_WARP == 5
    // one lane per C warp issues the five bulk prefetches of a row straight after each other (uniform addresses:
    // no per-lane serialisation loop around the copy-engine instruction)
    const char* pf_r0 = (const char*)pbase0;
    const char* pf_fl = MODE == 1 ? (const char*)(a.flow_in + (size_t)pair * a.flow_in_pair_stride) : nullptr;
    const size_t pf_r1 = (size_t)4 * a.r_frame_stride, pf_e = (size_t)16 * a.plane_stride;
    const int cwu = __shfl_sync(0xffffffffu, ct >> 5, 0);
    auto prefetch_block = [&](int sb) {
      if (sb >= nblk) return;
      const int nrows = sb == 0 ? 2 * M + RB : RB;
      const int y0 = sb == 0 ? ys - M : ys + sb * RB + M;
      if ((ct & 31) != 0) return;
      for (int r = cwu; r < nrows; r += 4) {
        const int yy = min(max(y0 + r, 0), h - 1);
        const size_t o = (size_t)yy * pitch + pf_xs;
        const char* p0 = pf_r0 + o * 16;
        const char* pe = pf_r0 + pf_e + o * 4;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"(pf_cols * 16) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0 + pf_r1), "r"(pf_cols * 16) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pe), "r"(pf_cols * 4) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pe + pf_r1), "r"(pf_cols * 4) : "memory");
        if (MODE == 1)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf_fl + o * 8), "r"(pf_cols * 8) : "memory");
      }
    };
#else
    auto prefetch_block = [&](int sb) {                 // rows of M that block sb adds: image rows [y0, y0 + nrows)
      if (FBW_PF_WARP != 0 || sb >= nblk) return;
      const int lane = ct & 31, cw = ct >> 5;
      if (lane >= 5 || (lane == 4 && MODE != 1)) return;
      const int nrows = sb == 0 ? 2 * M + RB : RB;
      const int y0 = sb == 0 ? ys - M : ys + sb * RB + M;
      for (int r = cw; r < nrows; r += 4) {
        const int yy = min(max(y0 + r, 0), h - 1);
        const size_t o = (size_t)yy * pitch + pf_xs;
        const void* p;
        int bytes;
        if (lane == 0) { p = (const float4*)pbase0 + o; bytes = pf_cols * 16; }
        else if (lane == 1) { p = (const float4*)(pbase0 + a.r_frame_stride) + o; bytes = pf_cols * 16; }
        else if (lane == 2) { p = pbase0 + 4 * a.plane_stride + o; bytes = pf_cols * 4; }
        else if (lane == 3) { p = pbase0 + a.r_frame_stride + 4 * a.plane_stride + o; bytes = pf_cols * 4; }
        else { p = a.flow_in + (size_t)pair * a.flow_in_pair_stride + o; bytes = pf_cols * 8; }
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
      }
    };
#endif
    for (int sb = 0; sb < FBW_PF_BLOCKS; ++sb) prefetch_block(sb);
    int po = 0;                                         // ring row of the oldest row of the window (row yb - M)
    for (int s = 0; s < nblk; ++s) {
      const int yb = ys + s * RB;
      mbar_wait(s_bar + (FBW_STAGES + s % FBW_STAGES) * 8, (unsigned)((s / FBW_STAGES) & 1));   // full_b[stage]
      if (!(FBW_SKIP & 4) && act) {
        int pn = po;
        if (s == 0 || ((yb / RB) % FBW_REFRESH) == 0) {
          // (re)start the running sums from the 14 rows above the window's newest row
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
        int pold = po;
        const int nr = min(RB, ye - yb);
        float2* orow = fo + (size_t)yb * a.out_pitch + (x0 + ct);
        // This serial walk is the pipeline's critical stage (one warp per scheduler, in-order issue), so it is
        // written for instruction-level parallelism: four rows per step, all 24 shared-memory reads first, then
        //   W0 = V + n0,  W1 = W0 + (n1 - o0),  W2 = W1 + (n2 - o1),  W3 = W2 + (n3 - o2),  V = W3 - o3
        // (n = row entering the window, o = row leaving it): the only serial chain is four adds, the differences
        // and the four 2x2 solves are independent of it.
        auto solve_store = [&](float2 wxy, float2 wzw, float we) {
          const float g11 = wxy.x, g12 = wxy.y, g22 = wzw.x, h1 = wzw.y, h2 = we;
          const float idet = rcp_approx(g11 * g22 - g12 * g12 + eps);
          const float2 f = make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
          *orow = f;
          orow += a.out_pitch;
          if (do_stats) {
            const float mg = sqrtf(f.x * f.x + f.y * f.y);
            st_m += mg; st_x += f.x; st_y += f.y; st_mx = fmaxf(st_mx, mg);
          }
        };
        int r = 0;
        for (; r + 4 <= nr; r += 4) {
          float2 nxy[4], nzw[4], oxy[4], ozw[4];
          float ne[4], oe[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int en = pn * ES + col, eo = pold * ES + col;
            nxy[k] = Pxy[en]; nzw[k] = Pzw[en]; ne[k] = Pe[en];
            oxy[k] = Pxy[eo]; ozw[k] = Pzw[eo]; oe[k] = Pe[eo];
            if (++pn == NR) pn = 0;
            if (++pold == NR) pold = 0;
          }
          float2 wxy[4], wzw[4];
          float we[4];
          wxy[0] = add2(vxy, nxy[0]); wzw[0] = add2(vzw, nzw[0]); we[0] = ve + ne[0];
#pragma unroll
          for (int k = 1; k < 4; ++k) {
            wxy[k] = add2(wxy[k - 1], sub2(nxy[k], oxy[k - 1]));
            wzw[k] = add2(wzw[k - 1], sub2(nzw[k], ozw[k - 1]));
            we[k] = we[k - 1] + (ne[k] - oe[k - 1]);
          }
          vxy = sub2(wxy[3], oxy[3]); vzw = sub2(wzw[3], ozw[3]); ve = we[3] - oe[3];
#pragma unroll
          for (int k = 0; k < 4; ++k) solve_store(wxy[k], wzw[k], we[k]);
        }
        for (; r < nr; ++r) {                           // last rows of the image
          const int en = pn * ES + col, eo = pold * ES + col;
          vxy = add2(vxy, Pxy[en]); vzw = add2(vzw, Pzw[en]); ve += Pe[en];
          solve_store(vxy, vzw, ve);
          vxy = sub2(vxy, Pxy[eo]); vzw = sub2(vzw, Pzw[eo]); ve -= Pe[eo];
          if (++pn == NR) pn = 0;
          if (++pold == NR) pold = 0;
        }
      }
      mbar_arrive(s_bar + (2 * FBW_STAGES + s % FBW_STAGES) * 8);   // empty_c[stage]: its oldest rows may be reused
      prefetch_block(s + FBW_PF_BLOCKS);
      po += RB;
      if (po >= NR) po -= NR;
    }
    if (do_stats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        st_m += __shfl_xor_sync(0xffffffffu, st_m, o);
        st_x += __shfl_xor_sync(0xffffffffu, st_x, o);
        st_y += __shfl_xor_sync(0xffffffffu, st_y, o);
        st_mx = fmaxf(st_mx, __shfl_xor_sync(0xffffffffu, st_mx, o));
      }
      // the ring is dead by now (every block has been consumed): its first bytes hold the four warps' partials
      unsigned long long* s_acc = (unsigned long long*)smem;      // [3][4]
      unsigned int* s_mx = (unsigned int*)(s_acc + 12);           // [4]
      asm volatile("bar.sync 1, 128;" ::: "memory");             // all four C warps are past their last ring read
      const double Q = 1048576.0;
      if ((ct & 31) == 0) {
        const int wdx = ct >> 5;
        s_acc[0 * 4 + wdx] = (unsigned long long)llrint((double)st_m * Q);
        s_acc[1 * 4 + wdx] = (unsigned long long)llrint((double)st_x * Q);
        s_acc[2 * 4 + wdx] = (unsigned long long)llrint((double)st_y * Q);
        s_mx[wdx] = __float_as_uint(st_mx);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (ct < 4) {
        unsigned long long* acc = a.stats_acc + (size_t)pair * 4;
        if (ct < 3) atomicAdd(acc + ct, s_acc[ct * 4] + s_acc[ct * 4 + 1] + s_acc[ct * 4 + 2] + s_acc[ct * 4 + 3]);
        else atomicMax((unsigned int*)(acc + 3), max(max(s_mx[0], s_mx[1]), max(s_mx[2], s_mx[3])));
      }
    }
  }
#if FBW_PF_WARP == 4
  else {
    // =========================== prefetch warp: feeds the L2 through the copy engine ===========================
    // One bulk prefetch per (row, stream) of the rows step A needs FBW_PF_LEAD blocks after the block that has just
    // entered the ring.  It takes part in no barrier: it polls the progress word step A publishes, so it can neither stall the
    // pipeline nor wait for a phase that has already gone by.
    const char* pf_r0 = (const char*)(a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride);
    const char* pf_fl = MODE == 1 ? (const char*)(a.flow_in + (size_t)pair * a.flow_in_pair_stride) : nullptr;
    const int pf_xs = max(x0 - FBS_PADL, 0);
    const int pf_cols = min(min(EW, pitch - pf_xs), FBS_PADL + nv + M + 8);   // nothing beyond the strip's own columns
    const size_t pf_r1 = (size_t)4 * a.r_frame_stride, pf_e = (size_t)16 * a.plane_stride;
    auto prefetch_block_p = [&](int sb) {
      if (sb >= nblk) return;
      const int nrows = sb == 0 ? 2 * M + RB : RB;
      const int y0 = sb == 0 ? ys - M : ys + sb * RB + M;
      if ((t & 31) == 0) {
        for (int r = 0; r < nrows; ++r) {
          const int yy = min(max(y0 + r, 0), h - 1);
          const size_t o = (size_t)yy * pitch + pf_xs;
          const char* p0 = pf_r0 + o * 16;
          const char* pe = pf_r0 + pf_e + o * 4;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"(pf_cols * 16) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0 + pf_r1), "r"(pf_cols * 16) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pe), "r"(pf_cols * 4) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pe + pf_r1), "r"(pf_cols * 4) : "memory");
          if (MODE == 1)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf_fl + o * 8), "r"(pf_cols * 8) : "memory");
        }
      }
    };
    int done = 0;                                       // blocks [0, done) have been requested
    for (;;) {
      unsigned p;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(p) : "r"(s_prog) : "memory");
      const int target = min((int)p + FBW_PF_LEAD, nblk);
      for (int sb = max(done, (int)p); sb < target; ++sb) prefetch_block_p(sb);   // what step A already did is skipped
      done = max(done, target);
      if ((int)p >= nblk || done >= nblk) break;
      asm volatile("nanosleep.u32 200;" ::: "memory");
    }
  }
#endif
}
