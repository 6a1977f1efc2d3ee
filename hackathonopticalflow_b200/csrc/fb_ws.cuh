// K5/K6 for the reference's window (winsize 15, box): ONE fused kernel per (level, iteration) --
//   flow_in (zero | previous iteration | bilinear x(1/pyr_scale) upsample of the coarser level)
//   -> UpdateMatrices (bilinear warp of R1, border attenuation) -> 15x15 box sum -> 2x2 solve -> flow_out
// M (the 5-channel matrix field) never leaves the SM.
//
// Shape: one CTA per SM walks down a strip of TW output columns (EW = TW + 16 halo columns = whole warps per row,
// starting on an 8-pixel boundary, so every row load falls on whole 128-byte lines) in blocks of RB rows.  The three
// steps of a block run on DIFFERENT warps, on different blocks, at the same time:
//
//   A warps: UpdateMatrices for block s + 2 -> ring rows.  thread = (halo column, run of 4 rows); walking down a
//            column the bottom corners of one pixel's bilinear gather are the top corners of the next one whenever
//            the integer part of the warp advanced by exactly one row (almost always), so they are carried in
//            registers and only two new corners are loaded; the flow vector is fetched one row ahead.
//   B warps: horizontal 15-sums in place on block s + 1.  lane = (plane, row, half-row); each half-row is walked
//            once with the 15 most recent inputs in a register window (one read and one write per element).
//   C warps: vertical 15-sums for block s as running sums carried in registers down the whole strip (one add and
//            one subtract per row), restarted from the ring every FBW_REFRESH_ROWS rows so rounding does not
//            accumulate; 2x2 solve; coalesced 8-byte stores; on the last iteration the per-pair flow statistics.
//            One lane per C warp also feeds the copy engine: bulk L2 prefetches (cp.async.bulk.prefetch.L2) of the
//            rows step A needs FBW_PF_BLOCKS blocks later.
//
// so the serial row / column walks of B and C hide behind the gathers of A instead of stalling the whole CTA at a
// barrier.  Blocks are handed from stage to stage through mbarriers (full_a[], full_b[], empty_c[]; stage = block
// mod FBW_STAGES).  The M ring holds 14 + FBW_STAGES * RB rows as three planes (float2 ch0/1, float2 ch2/3, float
// ch4) whose odd row stride makes the row-parallel accesses of B and the column-parallel accesses of A and C
// conflict-free.
//
// Two geometries (FBW_WIDE):
//   0  EW 128 / TW 112, RB 16, 16 A + 4 B + 4 C warps (768 threads, 80 registers), ring 160 KB   (shipped)
//   1  EW 192 / TW 176, RB  8, 12 A + 2 B + 6 C warps (640 threads, 96 registers), ring 147 KB: 8 % fewer halo columns
//      per output and 11 strips instead of 18 at 1920 columns -- but measured 50 % slower (step B becomes the stage
//      everything waits for), kept as a compile-time experiment only.
//
// What bounds it (ncu, profiles/README.md): the L1 / shared-memory data pipe (l1tex__data_pipe_lsu_wavefronts 55-60 %
// of peak in every variant, time proportional to the wavefront count); experiments that add memory-level parallelism
// to step A (software pipelining with setmaxnreg), move the prefetch issue to other warps or change the barrier
// waits do not help (profiles/README.md, round 2 table).
#pragma once

#ifndef FBW_WIDE
#define FBW_WIDE 0      // 1 was measured 50 % SLOWER (profiles/README.md): with 8-row blocks the two B warps have half the
#endif                  // row parallelism for a walk 1.6x as long and become the stage everything else waits for
#if FBW_WIDE
constexpr int FBS_EW = 192;               // halo columns per strip
constexpr int FBW_A_WARPS = 12;
constexpr int FBW_B_WARPS = 2;
constexpr int FBW_C_WARPS = 6;
#else
#ifndef FBW_EW_N
#define FBW_EW_N 128
#endif
constexpr int FBS_EW = FBW_EW_N;
#ifndef FBW_A_WARPS_N
#define FBW_A_WARPS_N (FBW_EW_N / 8)       // four row runs of EW columns
#endif
constexpr int FBW_A_WARPS = FBW_A_WARPS_N;
// Register budgets per role (setmaxnreg at the head of each role's branch).  With the shipped 16 + 4 + 4 warps every
// role stays at the 80 registers the CTA is launched with: the three instructions move nothing, but they end ptxas'
// allocation regions at the role boundaries, and that build measures 1.2 % faster (5.34 vs 5.41 ms, same bits).
// 20 A warps need 80 / 48 / 56 (A / B / C, launched at 72): step A alone then runs 17 % faster (its throughput is
// proportional to the number of A warps: 12 / 16 / 20 warps = 5.88 / 4.54 / 3.75 ms), the whole kernel only 1 %
// (5.36 ms): the B and C warps take issue slots from a stage that is bound by the latency of each warp's dependent
// chain, and with 56 registers step C is slower (profiles/README.md).
#if !defined(FBW_REGS_A) && !FBW_WIDE && FBW_A_WARPS_N == 16 && !defined(FBW_MAXNREG)
#define FBW_REGS_A 80
#define FBW_REGS_B 80
#define FBW_REGS_C 80
#endif
constexpr int FBW_B_WARPS = 4;
constexpr int FBW_C_WARPS = 4;
#endif
constexpr int FBS_M = 7;                  // window radius
constexpr int FBS_PADL = 8;               // halo column 0 sits at image column x0 - 8 (8-pixel aligned)
#ifndef FBW_PAIR
#define FBW_PAIR 0
#endif
// FBW_PAIR: two CTAs of a cluster share a 240-column strip.  Each computes UpdateMatrices for 128 columns (left CTA:
// image columns [X0 - 8, X0 + 120), right CTA: [X0 + 120, X0 + 248)) and owns 120 output columns; the 7 columns of M
// its horizontal sums need from the other side arrive in its ring through st.async (DSMEM) from the neighbour's A
// warps, whose completion bytes are part of the block's full_a barrier.  1920 columns = 8 pairs = 16 CTAs per row of
// strips instead of 18, and step A computes 256 columns per 240 outputs instead of 128 per 112.
#if FBW_PAIR
constexpr int FBS_TW = 120;               // output columns per CTA
constexpr int FBS_HALO = 7;               // columns received from the neighbour
constexpr int FBS_ES = FBS_EW + FBS_HALO; // 135: own columns at [7 * rank, 7 * rank + 128), the neighbour's 7 beside them
static_assert(FBS_EW == 128 && (FBS_ES & 1), "pair geometry");
#else
constexpr int FBS_TW = FBS_EW - 16;       // output columns per strip
constexpr int FBS_ES = FBS_EW + 1;        // plane row stride in elements (odd)
#endif
#ifndef FBS_HL_N
#define FBS_HL_N (FBS_TW / 2)
#endif
// outputs [0,HL) are summed left->right, [HL,TW) right->left (stored 14 positions further right).  With HL = 56 the
// 14-column jump in "where step B left column ct's sums" sits inside the second C warp, whose lanes 24-31 then share
// banks with lanes 0-23 (one extra wavefront per ring read of that warp, 1.96 M of 13.8 M step-C wavefronts per
// 16-pair launch); HL = 64 puts the jump between two warps and removes them -- and times the same (5.453 vs 5.455 ms
// per 64 pairs x 3 launches), so the shorter left-half walk stays
constexpr int FBS_HL = FBS_HL_N;
constexpr int FBW_RUNS = FBW_A_WARPS * 32 / FBS_EW;     // row runs per block in step A
constexpr int FBW_RB = 4 * FBW_RUNS;                    // rows per block: 4 rows per A thread
constexpr int FBW_REFRESH_ROWS = 64;                    // rows between restarts of the vertical running sums
constexpr int FBW_REFRESH = FBW_REFRESH_ROWS / FBW_RB;  // ... in blocks
constexpr int FBW_NT = (FBW_A_WARPS + FBW_B_WARPS + FBW_C_WARPS) * 32;
constexpr int FBW_PF_BLOCKS = 64 / FBW_RB;              // L2 prefetch distance in blocks (64 rows)
#ifndef FBW_ADDR_WIDE
#define FBW_ADDR_WIDE 1
#endif
#ifndef FBW_UP_CARRY
#define FBW_UP_CARRY 0  // 1: carry the interpolated coarse rows down a column in upsample mode -- measured 10 % SLOWER for
#endif                  // those launches (the conditional loads no longer overlap with the previous row)
#ifndef FBW_SKIP
#define FBW_SKIP 0   // timing experiments only: bit 0 / 1 / 2 switches step A / B / C off
#endif
#ifndef FBW_STAGES_N
#define FBW_STAGES_N 3
#endif
constexpr int FBW_STAGES = FBW_STAGES_N;                // blocks in flight between step A and step C
constexpr int FBW_NR = 2 * FBS_M + FBW_STAGES * FBW_RB; // ring rows
constexpr size_t FBW_PLANES = (size_t)FBW_NR * FBS_ES * 20;
#ifndef FBW_SMEM_PAD
#define FBW_SMEM_PAD 0   // experiment: extra dynamic shared memory (pushes the carve-out to the next step: 64 -> 32 KB of L1)
#endif
constexpr size_t FBW_SMEM = FBW_PLANES + 3 * FBW_STAGES * 8 + 16 + FBW_SMEM_PAD;   // + mbarriers
static_assert(FBW_A_WARPS * 32 % FBS_EW == 0, "step A: whole row runs");
static_assert(FBS_TW <= FBW_C_WARPS * 32, "step C: one thread per output column");

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
#ifndef FBW_SLEEP_NS
#define FBW_SLEEP_NS 200
#endif
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  // poll with back-off (a bare try_wait loop, a try_wait with a suspend-time hint and longer sleeps all time the same)
  asm volatile(
      "{ .reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "W_%=: nanosleep.u32 %2;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W_%=;\n"
      "D_%=: }" ::"r"(addr), "r"(parity), "n"(FBW_SLEEP_NS) : "memory");
}

#if FBW_PAIR
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// this thread's arrival on a local mbarrier, announcing `bytes` of st.async traffic for the current phase
__device__ __forceinline__ void mbar_arrive_expect(unsigned addr, unsigned bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(addr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// store into the neighbour's shared memory; the bytes are counted on the neighbour's mbarrier when they land
__device__ __forceinline__ void st_async_f2(unsigned cluster_addr, float x, float y, unsigned cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1,%2}, [%3];" ::"r"(
                   cluster_addr), "f"(x), "f"(y), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_f1(unsigned cluster_addr, float x, unsigned cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(cluster_addr),
               "f"(x), "r"(cluster_bar)
               : "memory");
}
#endif

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

// Sums are taken per block of 15 positions, not as a sliding sum: output xo = 15 b + j is
//   (suffix of block b from j) + (prefix of block b + 1 up to j - 1),
// i.e. exactly its own 15 inputs, added in a fixed order.  A sliding sum (s += new; s -= old) is one add cheaper per
// output but carries the absolute rounding error of whatever passed through the window: behind bright texture a dark
// wall's sums keep errors of the texture's magnitude, and on real footage that was the path's largest source of
// disagreement with cv2 at near-singular pixels (restated in numpy: 102 of 32,400 sampled pixels beyond 0.5 px with a
// sliding horizontal pass, 10 with direct sums -- cv2 against its own plain build: 12).  Cost: 42 adds per 15 outputs
// instead of 30; the register window is the same (a slot holds a block's suffix sum until its output is out, then the
// next block's input), and every element is still read from shared memory once and written once.
// The sliding walk stays the default (BLOCKED = false: 3 % more pairs/s end to end, and on well-conditioned pixels the
// two agree to 1e-3 px); the blocked walk is chosen per call with B2OF_FARNEBACK_BLOCKED_SUMS in the flags.
template <typename T, int DIR>
__device__ __forceinline__ void hpass_blocked(T* __restrict__ base, int n_out, T (&win)[2 * FBS_M + 1]) {
  constexpr int WN = 2 * FBS_M + 1;
  if (n_out <= 0) return;
#pragma unroll
  for (int k = 0; k < WN; ++k) win[k] = base[DIR * k];
  for (int eb = 0; eb < n_out; eb += WN) {
    // win[] holds block eb / 15 raw: turn it into suffix sums in place
#pragma unroll
    for (int k = WN - 2; k >= 0; --k) win[k] = hp_add(win[k], win[k + 1]);
    T* b = base + DIR * eb;
    const int left = n_out - eb;                       // outputs from this block on
    T pre;
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      if (j < left) {
        b[DIR * j] = j == 0 ? win[0] : hp_add(win[j], pre);
        if (j + 1 < left) {                            // element e + 15: the last input of output e + 1
          const T nw = b[DIR * (j + WN)];
          pre = j == 0 ? nw : hp_add(pre, nw);
          win[j] = nw;
        }
      }
    }
  }
}
template <typename T, bool FULL, bool BLOCKED>
__device__ __forceinline__ void hpass_half_row(T* rowp, bool right, int nv) {
  // nv = outputs this strip really has (FULL: the whole strip, compile-time bounds; otherwise the last strip of a
  // row: nothing beyond its outputs is summed)
  constexpr int M = FBS_M, HL = FBS_HL, WN = 2 * M + 1;
  const int TW = FULL ? FBS_TW : nv;
  T win[WN];
  if constexpr (BLOCKED) {
    // One walk for both halves: element e of the walk is position e (left half, left -> right, stored at its own
    // position) or position TW - 1 + 14 - e counted down from the right-most input (right half, right -> left).
    // Output e sums elements [e, e + 14] and is stored at element e.  The direction is a compile-time constant of each
    // call, so that the loads and stores of an unrolled block have provably different offsets and the loads can be
    // issued ahead.
    if (!right) hpass_blocked<T, 1>(rowp, FULL ? HL : min(HL, nv), win);
    else hpass_blocked<T, -1>(rowp + (TW - 1 + 2 * M), TW - HL, win);
    return;
  }
  if (!right) {
    const int HLv = FULL ? HL : min(HL, nv);
    T sm = rowp[0];
    win[0] = sm;
#pragma unroll
    for (int k = 1; k < 2 * M; ++k) { win[k] = rowp[k]; sm = hp_add(sm, win[k]); }
    for (int xb = 0; xb < HLv; xb += WN) {
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int xo = xb + j;
        if (xo < HLv) {
          const T nw = rowp[xo + 2 * M];
          sm = hp_add(sm, nw);
          rowp[xo] = sm;
          sm = hp_sub(sm, win[j]);
          win[(j + 2 * M) % WN] = nw;
        }
      }
    }
  } else if (TW > HL) {
    T* top = rowp + (TW - 1 + 2 * M);
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

#ifdef FBW_MAXNREG
#define FBW_BOUNDS __maxnreg__(FBW_MAXNREG)
#elif FBW_PAIR || defined(FBW_CLUSTER_ONLY)
#define FBW_BOUNDS __cluster_dims__(1, 2, 1) __launch_bounds__(FBW_NT, 1)
#else
#define FBW_BOUNDS __launch_bounds__(FBW_NT, 1)
#endif
template <int MODE, bool STATS, bool HBLOCK>
__global__ void FBW_BOUNDS fb_iter_ws(IterArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int EW = FBS_EW, TW = FBS_TW, M = FBS_M, RB = FBW_RB, NR = FBW_NR, ES = FBS_ES, HL = FBS_HL;
  constexpr int NA = FBW_A_WARPS * 32, NB = FBW_B_WARPS * 32, NC = FBW_C_WARPS * 32;
  float2* Pxy = (float2*)smem;                          // float2 [NR][ES]  (M0, M1)
  float2* Pzw = Pxy + NR * ES;                          // float2 [NR][ES]  (M2, M3)
  float* Pe = (float*)(Pzw + NR * ES);                  // float  [NR][ES]  (M4)
  const unsigned s_xy = smem_u32(smem);
  const unsigned s_zw = s_xy + NR * ES * 8;
  const unsigned s_e = s_zw + NR * ES * 8;
  const unsigned s_bar = s_xy + (unsigned)FBW_PLANES;   // full_a[], full_b[], empty_c[]
  const int pair = blockIdx.x;                          // pair index fastest (frame p + 1 shared through L2)
#if FBW_PAIR
  const int rank = (int)cluster_rank();                 // 0: left CTA of the pair, 1: right (== blockIdx.y & 1)
  const int x0 = (blockIdx.y >> 1) * (2 * TW) + rank * TW;          // first output column
  const int xa = rank ? x0 : x0 - FBS_PADL;                         // image column of halo column 0
  const int uoff = rank * FBS_HALO;                                 // ring index of halo column 0
  const int woff = 1 - rank;                            // ring index where output 0's window starts
#else
  const int x0 = blockIdx.y * TW;
  const int xa = x0 - FBS_PADL;
  constexpr int uoff = 0, woff = 1;
#endif
  const int w = a.w, h = a.h, pitch = a.pitch;
  const int ys = blockIdx.z * a.nb * RB;                // rows [ys, ye) are this CTA's outputs
  const int ye = min(ys + a.nb * RB, h);
  const int nblk = (ye - ys + RB - 1) / RB;
  const int t = threadIdx.x;
  const int nv = max(min(TW, w - x0), 0);               // outputs of this strip (the last strip of a row is narrower)

  if (t == 0) {
    for (int i = 0; i < FBW_STAGES; ++i) {
      mbar_init(s_bar + i * 8, NA);                     // full_a: every A thread arrives
      mbar_init(s_bar + (FBW_STAGES + i) * 8, NB);      // full_b
#if FBW_PAIR
#ifdef FBW_PAIR_NOREMOTE   // timing experiment: the neighbour's step C is not waited for (racy)
      mbar_init(s_bar + (2 * FBW_STAGES + i) * 8, NC);
#else
      mbar_init(s_bar + (2 * FBW_STAGES + i) * 8, NC + FBW_C_WARPS);   // empty_c: + one arrival per C warp next door
#endif
#else
      mbar_init(s_bar + (2 * FBW_STAGES + i) * 8, NC);  // empty_c
#endif
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#if FBW_PAIR
  cluster_sync_all();                                   // both CTAs' barriers exist before anything is sent across
#else
  __syncthreads();
#endif

  if (t < NA) {
    // =========================== A warps: UpdateMatrices ===========================
#ifdef FBW_REGS_A   // experiment: registers moved between the roles (the three counts must add up to the launch allocation)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FBW_REGS_A));
#endif
    // R of the pair's first frame: {float4 plane ch0..3, float plane ch4}; the second frame follows at r_frame_stride
    const float* base0 = a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride;
    const float2* __restrict__ fin = MODE ? a.flow_in + (size_t)pair * a.flow_in_pair_stride : nullptr;
    const int run = t / EW, cx = t - run * EW;
    // halo columns [PADL - M, PADL + nv + M) feed this strip's outputs: a warp whose 32 columns lie beyond them only
    // passes the barriers on (last strip of a row)
#if FBW_PAIR
    // columns 121..127 of the left CTA / 0..6 of the right one are also the neighbour's halo: their warps always work
#ifdef FBW_PAIR_NOPUSH   // timing experiment: nothing is sent (wrong results)
    const bool a_push = false;
#else
    const bool a_push = rank ? cx < FBS_HALO : cx >= EW - FBS_HALO;
#endif
    const bool a_on = rank ? (cx < 32 || (cx & ~31) < nv + FBS_M) : (cx >= EW - 32 || (cx & ~31) < FBS_PADL + nv + FBS_M);
    const unsigned nbr = (unsigned)(rank ^ 1);
    // neighbour's ring element = mine +- 121; its shared-memory window is laid out like this CTA's, so one mapped
    // base serves the three planes and the barriers
    const unsigned r_xy = mapa_u32(s_xy, nbr) + (rank ? 8 * (EW - FBS_HALO) : -8 * (EW - FBS_HALO));
    const unsigned r_e = mapa_u32(s_e, nbr) + (rank ? 4 * (EW - FBS_HALO) : -4 * (EW - FBS_HALO));
    const unsigned r_bar = mapa_u32(s_bar, nbr);
#else
    const bool a_on = (cx & ~31) < FBS_PADL + nv + FBS_M;
#endif
    const int x = clampi(xa + cx, 0, w - 1);
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
#if FBW_ADDR_WIDE
    // four 64-bit plane bases: every address is one IMAD.WIDE (FMA pipe) instead of LEA + IADD3 + IADD3.X (ALU pipe)
    const float4* R0A = (const float4*)base0;
    const float* R0E = base0 + 4 * a.plane_stride;
    const float4* R1A = (const float4*)(base0 + a.r_frame_stride);
    const float* R1E = base0 + a.r_frame_stride + 4 * a.plane_stride;
    pin(R0A); pin(R0E); pin(R1A); pin(R1E);
#define FB_R0A(o) (R0A + (unsigned)(o))
#define FB_R0E(o) (R0E + (unsigned)(o))
#define FB_R1A(o) (R1A + (unsigned)(o))
#define FB_R1E(o) (R1E + (unsigned)(o))
#else
#define FB_R0A(o) ((const float4*)(rb + (unsigned)(o) * 16u))
#define FB_R0E(o) ((const float*)(rb + ((unsigned)(o) * 4u + c_r0b)))
#define FB_R1A(o) ((const float4*)(rb + ((unsigned)(o) * 16u + c_r1a)))
#define FB_R1E(o) ((const float*)(rb + ((unsigned)(o) * 4u + c_r1b)))
#endif

#ifndef FBW_GSHFL
#define FBW_GSHFL 0
#endif
#ifndef FBW_PREF_NEXT
#define FBW_PREF_NEXT 0   // 1: fetch the flow vector of the next block's first row before handing this block over -- measured 0.8 % SLOWER
#endif
    float2 d_pref = make_float2(0.f, 0.f);
    bool have_pref = false;
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
        unsigned sa = (unsigned)(pr * ES + cx + uoff);  // element index into the planes
        const unsigned sa_end = (unsigned)(NR * ES + cx + uoff);
#if FBW_PAIR
        const unsigned r_full = r_bar + (s % FBW_STAGES) * 8;   // the neighbour's full_a[stage]
#endif
        int yu = y_first + l0;                          // unclamped image row of the row being set up

        auto next_row = [&](int& y, int& o) {
          y = min(max(yu, 0), hm1);
          o = y * pit + x;
          ++yu;
        };

        // a row's loads are issued and consumed in the same iteration; only the flow vector is fetched one row
        // ahead; bottom corners carried as the next row's top corners
        FbCorner cA, cB;
        cA.a0 = cA.a1 = make_float4(0.f, 0.f, 0.f, 0.f);
        cA.e0 = cA.e1 = 0.f;
        cB = cA;
        int o_carry = -1 << 30;
        int yA, oA, yB = 0, oB = 0;
        next_row(yA, oA);
#if FBW_UP_CARRY
        // MODE 2 (flow = bilinear x(1/pyr_scale) upsample of the coarser level): walking down a column, consecutive
        // rows read the same pair of coarse rows, or the pair one further down -- the horizontally interpolated
        // coarse rows are carried (same arithmetic, same bits: 1 to 1.5 gathers per row instead of 4)
        int u_ya = -1, u_yb = -1;
        float2 u_ha = make_float2(0.f, 0.f), u_hb = u_ha;
        auto coarse_row = [&](int r) {
          const float2 p0 = fi[r * a.in_pitch + uxa], p1 = fi[r * a.in_pitch + uxb];
          const float gx = 1.f - ufx;
          return make_float2(lerp_nc(p0.x, gx, p1.x, ufx), lerp_nc(p0.y, gx, p1.y, ufx));
        };
        auto fetch = [&](int o, int y) -> float2 {
          if (MODE != 2) return fetch_flow_m<MODE>(a, fi, o, y, uxa, uxb, ufx);
          const int ya = a.uy0[y], yb = a.uy1[y];
          const float fy = a.ufy[y];
          if (ya != u_ya) {
            u_ha = ya == u_yb ? u_hb : coarse_row(ya);
            u_ya = ya;
          }
          if (yb != u_yb) {
            u_hb = yb == ya ? u_ha : coarse_row(yb);
            u_yb = yb;
          }
          const float gy = 1.f - fy;
          return make_float2(__fmul_rn(lerp_nc(u_ha.x, gy, u_hb.x, fy), a.up_mult),
                             __fmul_rn(lerp_nc(u_ha.y, gy, u_hb.y, fy), a.up_mult));
        };
#else
        auto fetch = [&](int o, int y) -> float2 { return fetch_flow_m<MODE>(a, fi, o, y, uxa, uxb, ufx); };
#endif
        // (the first row's flow vector was requested at the end of the previous block: its latency passes while
        // the block is handed over instead of at the head of this run's dependent chain)
        float2 dA = (FBW_PREF_NEXT && have_pref) ? d_pref : fetch(oA, yA), dB = dA;
        auto rowf = [&](const float2 d, float2& dn, const int y, int& yn, const int o, int& on, bool has_next,
                        FbCorner& top, FbCorner& bot) {
          const float4 q = lds_f4<0>(FB_R0A(o));
          const float q4 = lds_f1<0>(FB_R0E(o));
          if (has_next) { next_row(yn, on); dn = fetch(on, yn); }
          float fx = xf + d.x, fy = (float)y + d.y;
          const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
          fx -= (float)x1; fy -= (float)y1;
          const bool inside = (unsigned)x1 < (unsigned)wm1 && (unsigned)y1 < (unsigned)hm1;
          const int ot = inside ? y1 * pit + x1 : 0;
          if (ot != o_carry) {
            const float4* pa = FB_R1A(ot);
            const float* pe = FB_R1E(ot);
            top.a0 = ldgat_f4<0>(pa); top.a1 = ldgat_f4<16>(pa); top.e0 = ldgat_f1<0>(pe); top.e1 = ldgat_f1<4>(pe);
          }
          const int ob = ot + pitb;
#if FBW_GSHFL
          {
            // the right-hand corner of lane l is the left-hand corner of lane l + 1 whenever the two pixels' integer
            // displacements agree (almost always): it comes over by shuffle, and only lane 31 and the lanes at a
            // displacement step load it themselves -- one unaligned 16-byte load per lane and row instead of two
            const float4* pa = FB_R1A(ob);
            const float* pe = FB_R1E(ob);
            bot.a0 = ldg_f4<0>(pa); bot.e0 = ldg_f1<0>(pe);
            const int obn = __shfl_down_sync(0xffffffffu, ob, 1);
            const bool own = obn != ob + 1 || (t & 31) == 31;
            if (own) { bot.a1 = ldg_f4<16>(pa); bot.e1 = ldg_f1<4>(pe); }
            const float nx = __shfl_down_sync(0xffffffffu, bot.a0.x, 1), ny = __shfl_down_sync(0xffffffffu, bot.a0.y, 1);
            const float nz = __shfl_down_sync(0xffffffffu, bot.a0.z, 1), nw = __shfl_down_sync(0xffffffffu, bot.a0.w, 1);
            const float ne = __shfl_down_sync(0xffffffffu, bot.e0, 1);
            if (!own) { bot.a1 = make_float4(nx, ny, nz, nw); bot.e1 = ne; }
          }
#else
          {
            const float4* pa = FB_R1A(ob);
            const float* pe = FB_R1E(ob);
            bot.a0 = ldgat_f4<0>(pa); bot.a1 = ldgat_f4<16>(pa); bot.e0 = ldgat_f1<0>(pe); bot.e1 = ldgat_f1<4>(pe);
          }
#endif
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
          const float m0 = r4 * r4 + r6 * r6, m1 = (r4 + r5) * r6, m2 = r5 * r5 + r6 * r6, m3 = r4 * r2 + r6 * r3;
          const float m4 = r6 * r2 + r5 * r3;
          sts_f2(s_xy + sa * 8, m0, m1);
          sts_f2(s_zw + sa * 8, m2, m3);
          sts_f1(s_e + sa * 4, m4);
#if FBW_PAIR
          if (a_push) {
            st_async_f2(r_xy + sa * 8, m0, m1, r_full);
            st_async_f2(r_xy + sa * 8 + NR * ES * 8, m2, m3, r_full);
            st_async_f1(r_e + sa * 4, m4, r_full);
          }
#endif
          sa += ES;
          if (sa == sa_end) sa -= NR * ES;
        };
        int k = 0;
        for (; k + 1 < n; k += 2) {
          rowf(dA, dB, yA, yB, oA, oB, true, cA, cB);
          rowf(dB, dA, yB, yA, oB, oA, k + 2 < n, cB, cA);
        }
        if (k < n) rowf(dA, dB, yA, yB, oA, oB, false, cA, cB);
        if (FBW_PREF_NEXT && s + 1 < nblk) {            // block s + 1 >= 1: RB rows, runs of RB / FBW_RUNS
          const int yn = min(max(ys + (s + 1) * RB + M + run * (RB / FBW_RUNS), 0), hm1);
          d_pref = fetch(yn * pit + x, yn);
          have_pref = true;
        }
      }
#if FBW_PAIR
      // thread 0's arrival also announces the neighbour's 7 columns x nrows x 20 bytes
#ifdef FBW_PAIR_NOPUSH
      if (t == 0) mbar_arrive(s_bar + (s % FBW_STAGES) * 8);
#else
      if (t == 0) mbar_arrive_expect(s_bar + (s % FBW_STAGES) * 8, (unsigned)(FBS_HALO * 20 * nrows));
#endif
      else
#endif
      mbar_arrive(s_bar + (s % FBW_STAGES) * 8);        // full_a[stage]: block s is in the ring
      j0 += nrows;
      if (j0 >= NR) j0 -= NR;
    }
  } else if (t < NA + NB) {
    // =========================== B warps: horizontal sums in place ===========================
    // items of a block: (plane, half, row) -- 4 * nrows of float2 type (planes 0, 1), then 2 * nrows of float type
    // (plane 2); a warp takes 32 items of one type at a time (no divergence between the two element types), rows
    // fastest across lanes (the odd row stride spreads them over the banks)
#ifdef FBW_REGS_A
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FBW_REGS_B));
#endif
    const int bt = t - NA, bw = bt >> 5, lane = bt & 31;
    int j0 = 0;
    for (int s = 0; s < nblk; ++s) {
      const int nrows = s == 0 ? 2 * M + RB : RB;
      mbar_wait(s_bar + (s % FBW_STAGES) * 8, (unsigned)((s / FBW_STAGES) & 1));   // full_a[stage]
      // float-type items go 16 to a warp (one half-row walker per row, left or right half): 32 of them in one
      // 4-byte access would put the left and the right walkers of the same rows on the same banks every other step
      const int n2 = 4 * nrows, n1 = 2 * nrows;
      const int u2 = (n2 + 31) >> 5, u1 = (n1 + 15) >> 4;
      for (int u = bw; u < u2 + u1; u += FBW_B_WARPS) {
        const bool wide = u < u2;
        const int i = wide ? u * 32 + lane : (u - u2) * 16 + lane;
        if (!(FBW_SKIP & 2) && i < (wide ? n2 : n1) && (wide || lane < 16)) {
          const int q = i / nrows, r = i - q * nrows;   // q = 2 * plane + half (float2 type) or half (float type)
          int pr = j0 + r;
          if (pr >= NR) pr -= NR;
          const bool right = q & 1;
          if (nv == TW) {
            if (wide) hpass_half_row<float2, true, HBLOCK>((q & 2 ? Pzw : Pxy) + pr * ES + woff, right, nv);
            else hpass_half_row<float, true, HBLOCK>(Pe + pr * ES + woff, right, nv);
          } else {
            if (wide) hpass_half_row<float2, false, HBLOCK>((q & 2 ? Pzw : Pxy) + pr * ES + woff, right, nv);
            else hpass_half_row<float, false, HBLOCK>(Pe + pr * ES + woff, right, nv);
          }
        }
      }
      mbar_arrive(s_bar + (FBW_STAGES + s % FBW_STAGES) * 8);   // full_b[stage]
      j0 += nrows;
      if (j0 >= NR) j0 -= NR;
    }
  } else {
    // =========================== C warps: vertical running sums + solve ===========================
#ifdef FBW_REGS_A
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FBW_REGS_C));
#endif
    const int ct = t - NA - NB;
    const bool act = ct < TW && x0 + ct < w;
    const int col = woff + (ct < HL ? ct : ct + 2 * M); // where step B left this column's sums
#if FBW_PAIR
    const unsigned r_empty = mapa_u32(s_bar + 2 * FBW_STAGES * 8, (unsigned)(rank ^ 1));   // the neighbour's empty_c[]
#endif
    float2* __restrict__ fo = a.flow_out + (size_t)pair * a.flow_out_pair_stride;
    float2 vxy = make_float2(0.f, 0.f), vzw = vxy;
    float ve = 0.f;
    // HBLOCK instantiation (the two coarsest levels, or every level on request): the vertical running sums are carried
    // in double -- a float running sum keeps the absolute rounding error of the rows that passed through the window,
    // and at near-singular pixels below bright texture that error decides the result (numpy restatement on real
    // footage: with direct horizontal AND vertical sums at the two coarsest levels the path disagrees with cv2 on as
    // few pixels as cv2's plain build does).  Sums of fifteen floats are exact in double, so each output is the
    // correctly rounded sum of its own inputs whatever the restart period.
    double dv0 = 0., dv1 = 0., dv2 = 0., dv3 = 0., dv4 = 0.;
    const float eps = 1e-3f / (a.inv_area * a.inv_area);
    // per-pair flow statistics folded into the last iteration (STATS): every C thread sums its own column, the CTA
    // adds one set of fixed-point partials (same layout as flow_stats_accum, pathfinder.cu)
    constexpr bool do_stats = STATS;
    float st_m = 0.f, st_x = 0.f, st_y = 0.f, st_mx = 0.f;
    // The C warps also feed the copy engine: L2 prefetch of the rows step A will need FBW_PF_BLOCKS blocks from now.
    // One lane per C warp issues the bulk prefetches of a row straight after each other, with uniform addresses (no
    // per-lane serialisation loop around the copy-engine instruction: round 1 had five lanes issue one prefetch each,
    // 45 instructions per prefetch, and the C warps -- the one stage that never waits -- spent half their time there).
    const char* pf_r0 = (const char*)(a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride);
    const char* pf_fl = MODE == 1 ? (const char*)(a.flow_in + (size_t)pair * a.flow_in_pair_stride) : nullptr;
    const size_t pf_r1 = (size_t)4 * a.r_frame_stride, pf_e = (size_t)16 * a.plane_stride;
    const int pf_xs = max(xa, 0);
    const int pf_cols = min(EW, pitch - pf_xs);
    const int cwu = __shfl_sync(0xffffffffu, ct >> 5, 0);
    auto prefetch_block = [&](int sb) {                 // rows of M that block sb adds: image rows [y0, y0 + nrows)
      if (sb >= nblk || (ct & 31) != 0 || pf_cols <= 0) return;
      const int nrows = sb == 0 ? 2 * M + RB : RB;
      const int y0 = sb == 0 ? ys - M : ys + sb * RB + M;
      for (int r = cwu; r < nrows; r += FBW_C_WARPS) {
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
          if (HBLOCK) { dv0 = 0.; dv1 = 0.; dv2 = 0.; dv3 = 0.; dv4 = 0.; }
#pragma unroll
          for (int k = 0; k < 2 * M; ++k) {
            const int e = pn * ES + col;
            if (HBLOCK) {
              const float2 pxy = Pxy[e], pzw = Pzw[e];
              dv0 += (double)pxy.x; dv1 += (double)pxy.y; dv2 += (double)pzw.x; dv3 += (double)pzw.y;
              dv4 += (double)Pe[e];
            } else {
              vxy = add2(vxy, Pxy[e]);
              vzw = add2(vzw, Pzw[e]);
              ve += Pe[e];
            }
            if (++pn == NR) pn = 0;
          }
        } else {
          pn += 2 * M;
          if (pn >= NR) pn -= NR;
        }
        int pold = po;
        const int nr = min(RB, ye - yb);
        float2* orow = fo + (size_t)yb * a.out_pitch + (x0 + ct);
        // This serial walk is one warp per scheduler, in-order issue, so it is written for instruction-level
        // parallelism: four rows per step, all 24 shared-memory reads first, then
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
          if (HBLOCK) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              dv0 += (double)nxy[k].x; dv1 += (double)nxy[k].y; dv2 += (double)nzw[k].x; dv3 += (double)nzw[k].y;
              dv4 += (double)ne[k];
              wxy[k] = make_float2((float)dv0, (float)dv1); wzw[k] = make_float2((float)dv2, (float)dv3);
              we[k] = (float)dv4;
              dv0 -= (double)oxy[k].x; dv1 -= (double)oxy[k].y; dv2 -= (double)ozw[k].x; dv3 -= (double)ozw[k].y;
              dv4 -= (double)oe[k];
            }
          } else {
            wxy[0] = add2(vxy, nxy[0]); wzw[0] = add2(vzw, nzw[0]); we[0] = ve + ne[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) {
              wxy[k] = add2(wxy[k - 1], sub2(nxy[k], oxy[k - 1]));
              wzw[k] = add2(wzw[k - 1], sub2(nzw[k], ozw[k - 1]));
              we[k] = we[k - 1] + (ne[k] - oe[k - 1]);
            }
            vxy = sub2(wxy[3], oxy[3]); vzw = sub2(wzw[3], ozw[3]); ve = we[3] - oe[3];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) solve_store(wxy[k], wzw[k], we[k]);
        }
        for (; r < nr; ++r) {                           // last rows of the image
          const int en = pn * ES + col, eo = pold * ES + col;
          if (HBLOCK) {
            const float2 pxy = Pxy[en], pzw = Pzw[en], qxy = Pxy[eo], qzw = Pzw[eo];
            dv0 += (double)pxy.x; dv1 += (double)pxy.y; dv2 += (double)pzw.x; dv3 += (double)pzw.y; dv4 += (double)Pe[en];
            solve_store(make_float2((float)dv0, (float)dv1), make_float2((float)dv2, (float)dv3), (float)dv4);
            dv0 -= (double)qxy.x; dv1 -= (double)qxy.y; dv2 -= (double)qzw.x; dv3 -= (double)qzw.y; dv4 -= (double)Pe[eo];
          } else {
            vxy = add2(vxy, Pxy[en]); vzw = add2(vzw, Pzw[en]); ve += Pe[en];
            solve_store(vxy, vzw, ve);
            vxy = sub2(vxy, Pxy[eo]); vzw = sub2(vzw, Pzw[eo]); ve -= Pe[eo];
          }
          if (++pn == NR) pn = 0;
          if (++pold == NR) pold = 0;
        }
      }
      mbar_arrive(s_bar + (2 * FBW_STAGES + s % FBW_STAGES) * 8);   // empty_c[stage]: its oldest rows may be reused
#if FBW_PAIR
      // ... and the neighbour's step A may overwrite the halo columns it sent for this block (step B read them before
      // it released full_b): one arrival per C warp on the neighbour's empty_c[stage]
      __syncwarp();
#ifndef FBW_PAIR_NOREMOTE
      if ((ct & 31) == 0) mbar_arrive_remote(r_empty + (s % FBW_STAGES) * 8);
#endif
#endif
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
      // the ring is dead by now (every block has been consumed): its first bytes hold the C warps' partials
      unsigned long long* s_acc = (unsigned long long*)smem;              // [3][FBW_C_WARPS]
      unsigned int* s_mx = (unsigned int*)(s_acc + 3 * FBW_C_WARPS);      // [FBW_C_WARPS]
      asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");               // every C warp is past its last ring read
      const double Q = 1048576.0;
      if ((ct & 31) == 0) {
        const int wdx = ct >> 5;
        s_acc[0 * FBW_C_WARPS + wdx] = (unsigned long long)llrint((double)st_m * Q);
        s_acc[1 * FBW_C_WARPS + wdx] = (unsigned long long)llrint((double)st_x * Q);
        s_acc[2 * FBW_C_WARPS + wdx] = (unsigned long long)llrint((double)st_y * Q);
        s_mx[wdx] = __float_as_uint(st_mx);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
      if (ct < 4) {
        unsigned long long* acc = a.stats_acc + (size_t)pair * 4;
        if (ct < 3) {
          unsigned long long v = 0;
#pragma unroll
          for (int k = 0; k < FBW_C_WARPS; ++k) v += s_acc[ct * FBW_C_WARPS + k];
          atomicAdd(acc + ct, v);
        } else {
          unsigned int v = 0;
#pragma unroll
          for (int k = 0; k < FBW_C_WARPS; ++k) v = max(v, s_mx[k]);
          atomicMax((unsigned int*)(acc + 3), v);
        }
      }
    }
  }
#if FBW_PAIR
  cluster_sync_all();                                   // neither CTA leaves while the other may still send to it
#endif
}
