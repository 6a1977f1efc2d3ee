// K12: the reference's own downstream logic on the device, one CTA per frame.
//   vector filter      pathfinder_viewer.py:159-178 (get_flow_lk after the LK call)
//   danger intensity   pathfinder_viewer.py:210-217 (draw_sparse_lamps before drawing)
// numpy semantics kept: float32 throughout, np.median (mean of the two middle values), np.percentile(.,99)
// (linear interpolation), np.int32() truncation toward zero, float64 sqrt for the integer flow modulus.
// Plus flow_stats: per-pair statistics of a dense flow field (what draw_flow / draw_hsv consume).
#include <math.h>

#include "common.cuh"

namespace b2of {

constexpr int PF_THREADS = 1024;
constexpr int PF_MAX_PTS = 32768;  // 128 KB of float keys in shared memory

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < PF_THREADS / 32; ++i) s += red[i];  // fixed order: deterministic
  return s;
}

// exact k-th smallest (k = 0 .. n-1) of n non-negative floats in shared memory: most-significant-digit radix
// selection, 8 bits per pass; lanes with the same digit are counted with one shared-memory atomic per warp.
// Called by all threads of the block; returns the same value in every thread.
__device__ float select_kth(const float* s_key, int n, int k, unsigned int* s_hist, unsigned int* s_sel) {
  const int t = threadIdx.x;
  unsigned int prefix = 0, mask = 0;
  int kk = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (t < 256) s_hist[t] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += PF_THREADS) {
      const int i = i0 + t;
      const bool in = i < n;
      const unsigned int u = in ? __float_as_uint(s_key[i]) : 0u;
      const bool take = in && (u & mask) == prefix;
      const unsigned int d = (u >> shift) & 255u;
      const unsigned int active = __ballot_sync(0xffffffffu, take);
      if (take) {
        const unsigned int peers = __match_any_sync(active, d);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[d], (unsigned int)__popc(peers));
      }
    }
    __syncthreads();
    if (t == 0) {
      unsigned int cum = 0, d = 0;
      for (; d < 255; ++d) {
        const unsigned int c = s_hist[d];
        if (cum + c > (unsigned int)kk) break;
        cum += c;
      }
      *s_sel = (d << 16) | 0;               // digit
      s_hist[0] = cum;                        // elements below the chosen digit
    }
    __syncthreads();
    const unsigned int d = *s_sel >> 16;
    kk -= (int)s_hist[0];
    prefix |= d << shift;
    mask |= 255u << shift;
    __syncthreads();
  }
  return __uint_as_float(prefix);
}

__global__ void __launch_bounds__(PF_THREADS) pathfinder_filter(const float* __restrict__ pts, size_t pts_bstride,
                                                                 const float* __restrict__ next_pts, int n,
                                                                 int width, int height, int32_t* __restrict__ kept_pts,
                                                                 int32_t* __restrict__ kept_flow,
                                                                 uint8_t* __restrict__ danger_v,
                                                                 uint8_t* __restrict__ mask_out,
                                                                 int32_t* __restrict__ n_kept,
                                                                 float* __restrict__ stats, int np2, int mode,
                                                                 int32_t* __restrict__ all_pts,
                                                                 int32_t* __restrict__ all_next) {
  extern __shared__ float s_key[];  // the n moduli
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int s_sel;
  __shared__ float red[PF_THREADS / 32];
  __shared__ int warp_off[PF_THREADS / 32 + 1];
  __shared__ int s_base;
  const int b = blockIdx.x, t = threadIdx.x;
  const float* P = pts + (size_t)b * pts_bstride * 2;
  const float* Q = next_pts + (size_t)b * n * 2;
  const float hw = (float)(width / 2), hh = (float)(height / 2);
  float sum_mag = 0.f, max_mag = 0.f, sum_dx = 0.f, sum_dy = 0.f;
  for (int i = t; i < n; i += PF_THREADS) {
    float key = 0.f;
    {
      float x = P[2 * i], y = P[2 * i + 1];
      float fx = Q[2 * i] - x, fy = Q[2 * i + 1] - y;
      float mod = sqrtf(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
      float ddx = hw - x, ddy = hh - y;
      float mm = sqrtf(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
      key = __fmul_rn(__fdiv_rn(mod, 5.f + sqrtf(mm)), 30.f);
      sum_mag += mod; max_mag = fmaxf(max_mag, mod); sum_dx += fx; sum_dy += fy;
    }
    s_key[i] = key;
  }
  __syncthreads();
  // np.median / np.percentile(., 99) in float32 need four order statistics, not the sorted array: exact radix
  // selection on the bit patterns (the moduli are non-negative, so bit order is value order; NaN-free input assumed)
  float med, p99;
  {
    const float k_mid = select_kth(s_key, n, n / 2, s_hist, &s_sel);
    if (n & 1) med = k_mid;
    else med = __fdiv_rn(__fadd_rn(select_kth(s_key, n, n / 2 - 1, s_hist, &s_sel), k_mid), 2.f);
    double vi = 0.99 * (double)(n - 1);
    int lo = (int)floor(vi);
    int hi = lo + 1 < n ? lo + 1 : n - 1;
    float tt = (float)(vi - (double)lo);
    float a = select_kth(s_key, n, lo, s_hist, &s_sel);
    float c = hi == lo ? a : select_kth(s_key, n, hi, s_hist, &s_sel);
    float diff = c - a;
    p99 = tt >= 0.5f ? c - diff * (1.f - tt) : a + diff * tt;
  }
  // second pass: mask + ordered compaction
  int kept_total = 0;
  float sum_v = 0.f;
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += PF_THREADS) {
    int i = i0 + t;
    bool keep = false;
    int px = 0, py = 0, qx = 0, qy = 0;
    if (i < n) {
      float x = P[2 * i], y = P[2 * i + 1];
      float fx = Q[2 * i] - x, fy = Q[2 * i + 1] - y;
      float ang = atan2f(fy, fx);
      float mod = sqrtf(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
      float ddx = hw - x, ddy = hh - y;
      float mm = sqrtf(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
      float m2 = __fmul_rn(__fdiv_rn(mod, 5.f + sqrtf(mm)), 30.f);
      float gx = __fmul_rn(m2, cosf(ang)), gy = __fmul_rn(m2, sinf(ang));
      qx = (int)(__fadd_rn(__fadd_rn(x, gx), 0.5f));
      qy = (int)(__fadd_rn(__fadd_rn(y, gy), 0.5f));
      px = (int)(x + 0.5f);
      py = (int)(y + 0.5f);
      // mode 0: the viewer's band  median < m < p99 (pathfinder_viewer.py:171);
      // mode 1: the development script's  m > 1.2 * median  (DenseOF.py:228; float32 product, as numpy 2 computes it)
      keep = mode == 1 ? m2 > __fmul_rn(med, 1.2f) : (med < m2) && (m2 < p99);
      mask_out[(size_t)b * n + i] = keep ? 1 : 0;
      if (all_pts) {                      // every point as the reference rounds it (:169-170): the overlay's input
        const size_t oa = ((size_t)b * n + i) * 2;
        all_pts[oa] = px; all_pts[oa + 1] = py;
        all_next[oa] = qx; all_next[oa + 1] = qy;
      }
    }
    unsigned int ballot = __ballot_sync(0xffffffffu, keep);
    int lane = t & 31, wid = t >> 5;
    if (lane == 0) warp_off[wid] = __popc(ballot);
    __syncthreads();
    if (t == 0) {
      int acc = s_base;
      for (int wi = 0; wi < PF_THREADS / 32; ++wi) { int c = warp_off[wi]; warp_off[wi] = acc; acc += c; }
      warp_off[PF_THREADS / 32] = acc;
    }
    __syncthreads();
    if (keep) {
      int slot = warp_off[wid] + __popc(ballot & ((1u << lane) - 1));
      size_t o = ((size_t)b * n + slot) * 2;
      int fxi = qx - px, fyi = qy - py;
      kept_pts[o] = px; kept_pts[o + 1] = py;
      kept_flow[o] = fxi; kept_flow[o + 1] = fyi;
      double m = sqrt((double)(fxi * fxi + fyi * fyi));
      double v = fmin(50.0 + m * 2.0, 255.0);
      uint8_t vb = (uint8_t)v;
      danger_v[(size_t)b * n + slot] = vb;
      sum_v += (float)vb;
    }
    __syncthreads();
    if (t == 0) s_base = warp_off[PF_THREADS / 32];
    __syncthreads();
  }
  kept_total = s_base;
  float tot_mag = block_sum(sum_mag, red);
  float tot_dx = block_sum(sum_dx, red);
  float tot_dy = block_sum(sum_dy, red);
  float tot_v = block_sum(sum_v, red);
  // block max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_mag = fmaxf(max_mag, __shfl_xor_sync(0xffffffffu, max_mag, o));
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = max_mag;
  __syncthreads();
  if (t == 0) {
    float mx = 0.f;
    for (int i = 0; i < PF_THREADS / 32; ++i) mx = fmaxf(mx, red[i]);
    n_kept[b] = kept_total;
    float* s = stats + (size_t)b * B2OF_STATS_WIDTH;
    s[0] = tot_mag / (float)n; s[1] = mx; s[2] = tot_dx / (float)n; s[3] = tot_dy / (float)n;
    s[4] = med; s[5] = p99; s[6] = (float)kept_total; s[7] = tot_v;
  }
}

int pathfinder_filter_dev(const float* pts, size_t pts_bstride, const float* next_pts, int n_pts, int batch, int width,
                          int height, int mode, int32_t* kept_pts, int32_t* kept_flow, uint8_t* danger_v, uint8_t* mask,
                          int32_t* n_kept, float* stats, int32_t* all_pts, int32_t* all_next, cudaStream_t st) {
  const char* fn = "pathfinder_filter";
  B2OF_ASSERT(n_pts >= 1 && batch >= 0 && width > 0 && height > 0, fn);
  B2OF_ASSERT(mode == B2OF_FILTER_VIEWER || mode == B2OF_FILTER_DENSEOF, fn);
  B2OF_ASSERT((all_pts == nullptr) == (all_next == nullptr), fn);
  B2OF_ASSERT(pts && next_pts && kept_pts && kept_flow && danger_v && mask && n_kept && stats, fn);
  if (n_pts > PF_MAX_PTS) return fail(B2OF_E_UNSUPPORTED, "more than %d points per frame", PF_MAX_PTS);
  if (batch == 0) return B2OF_OK;
  int np2 = 1;
  while (np2 < n_pts) np2 <<= 1;
  size_t smem = (size_t)np2 * sizeof(float);
  static PerDeviceMax max_set;
  if (smem > 48 * 1024 && max_set.raise(smem))
    B2OF_CUDA(cudaFuncSetAttribute(pathfinder_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pathfinder_filter<<<batch, PF_THREADS, smem, st>>>(pts, pts_bstride, next_pts, n_pts, width, height, kept_pts,
                                                     kept_flow, danger_v, mask, n_kept, stats, np2, mode,
                                                     all_pts, all_next);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

// ---- dense flow sampled on the measurement grid (what draw_flow does on its 14-px grid,
// DenseOF.py:44-50): next_pts = pts + flow[int(y), int(x)], ready for pathfinder_filter in place of the LK result ----
__global__ void __launch_bounds__(256) flow_sample_grid(const float2* __restrict__ flow, int rows, int cols,
                                                         const float* __restrict__ pts, size_t pts_bstride, int n,
                                                         float* __restrict__ next_pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  const float* P = pts + (size_t)b * pts_bstride * 2;
  float x = P[2 * i], y = P[2 * i + 1];
  int xi = clampi((int)x, 0, cols - 1), yi = clampi((int)y, 0, rows - 1);
  float2 f = flow[(size_t)b * rows * cols + (size_t)yi * cols + xi];
  next_pts[((size_t)b * n + i) * 2] = x + f.x;
  next_pts[((size_t)b * n + i) * 2 + 1] = y + f.y;
}

int flow_sample_dev(const float* flow, int n_pairs, int rows, int cols, const float* pts, size_t pts_bstride, int n_pts,
                    float* next_pts, cudaStream_t st) {
  const char* fn = "flow_sample";
  B2OF_ASSERT(n_pairs >= 0 && rows > 0 && cols > 0 && n_pts >= 0, fn);
  if (n_pairs == 0 || n_pts == 0) return B2OF_OK;
  B2OF_ASSERT(flow != nullptr && pts != nullptr && next_pts != nullptr, fn);
  flow_sample_grid<<<dim3(cdiv(n_pts, 256), n_pairs), 256, 0, st>>>((const float2*)flow, rows, cols, pts, pts_bstride,
                                                                   n_pts, next_pts);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

// ---- dense flow statistics: deterministic fixed-point accumulation, then finalize in place ----
// scratch layout inside the 8-float stats row: [0:2) u64 sum|f| (Q20), [2:4) i64 sum dx, [4:6) i64 sum dy, [6] max bits
__global__ void __launch_bounds__(256) flow_stats_accum(const float2* __restrict__ flow, size_t n_px,
                                                         float* __restrict__ stats) {
  const int b = blockIdx.y;
  const float2* f = flow + (size_t)b * n_px;
  float sm = 0.f, sx = 0.f, sy = 0.f, mx = 0.f;
  // two pixels per 16-byte load (n_px * 8 bytes per pair keeps every pair 16-byte aligned when n_px is even)
  const size_t n2 = ((n_px & 1) || ((size_t)f & 15)) ? 0 : n_px / 2;
  const float4* f4 = (const float4*)f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n2; i += (size_t)gridDim.x * 256) {
    float4 v = __ldg(f4 + i);
    float m0 = sqrtf(v.x * v.x + v.y * v.y), m1 = sqrtf(v.z * v.z + v.w * v.w);
    sm += m0 + m1; sx += v.x + v.z; sy += v.y + v.w; mx = fmaxf(mx, fmaxf(m0, m1));
  }
  for (size_t i = 2 * n2 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n_px; i += (size_t)gridDim.x * 256) {
    float2 v = f[i];
    float m = sqrtf(v.x * v.x + v.y * v.y);
    sm += m; sx += v.x; sy += v.y; mx = fmaxf(mx, m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  // fixed-point (Q20) per-warp partials are integers: summing the block's eight warps in shared memory first and
  // issuing one set of atomics per block gives the same bits as eight sets, with an eighth of the contention
  __shared__ unsigned long long s_acc[3][8];
  __shared__ unsigned int s_mx[8];
  const double Q = 1048576.0;
  if ((threadIdx.x & 31) == 0) {
    const int wdx = threadIdx.x >> 5;
    s_acc[0][wdx] = (unsigned long long)llrint((double)sm * Q);
    s_acc[1][wdx] = (unsigned long long)llrint((double)sx * Q);   // two's complement wrap = signed add
    s_acc[2][wdx] = (unsigned long long)llrint((double)sy * Q);
    s_mx[wdx] = __float_as_uint(mx);                              // mx >= 0: bit order == value order
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long* acc = (unsigned long long*)(stats + (size_t)b * B2OF_STATS_WIDTH);
    if (threadIdx.x < 3) {
      unsigned long long v = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) v += s_acc[threadIdx.x][k];
      atomicAdd(acc + threadIdx.x, v);
    } else {
      unsigned int v = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) v = max(v, s_mx[k]);
      atomicMax((unsigned int*)(acc + 3), v);
    }
  }
}

__global__ void flow_stats_final(float* __restrict__ stats, size_t n_px, int n_pairs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_pairs) return;
  float* s = stats + (size_t)b * B2OF_STATS_WIDTH;
  unsigned long long* acc = (unsigned long long*)s;
  const double Q = 1048576.0;
  double sm = (double)acc[0] / Q, sx = (double)(long long)acc[1] / Q, sy = (double)(long long)acc[2] / Q;
  float mx = __uint_as_float(*(unsigned int*)(acc + 3));
  s[0] = (float)(sm / (double)n_px); s[1] = mx; s[2] = (float)(sx / (double)n_px); s[3] = (float)(sy / (double)n_px);
  s[4] = s[5] = s[6] = s[7] = 0.f;
}

// ---- dense flow as a picture: draw_hsv, pathfinder_viewer.py:124-141 ----
//   ang = arctan2(fy, fx) + pi;  v = sqrt(fx^2 + fy^2)                          (float32, as numpy computes them)
//   hsv = uint8(ang * (180/pi/2)), 255, uint8(min(4 v, 255));  bgr = cv2.cvtColor(hsv, COLOR_HSV2BGR)
// The cvtColor call is third-party (opencv color_hsv.simd.hpp): float32 h * (6/180), v * (1/255), the table
// {v, v(1-s), v(1-s f), v(1-s(1-f))} with s = 1, times 255, truncated (oracle/pathfinder.py::hsv2bgr_u8, pinned
// against cv2 for every (h, v) at the saturation the reference writes).  arctan2 is evaluated in double and rounded
// once.  numpy 2's float32 arctan2 is a SIMD routine that is NOT correctly rounded (it differs from the rounded double
// value on 38 % of random inputs, by an ulp), so where ang * 28.65 lands within an ulp of an integer the truncated hue
// differs by one: 1 to 4 pixels per million on random fields (scripts/gpu_stress_overlay.py), and which pixels
// depends on the host's numpy build.
__global__ void __launch_bounds__(256) flow_hsv_bgr(const float2* __restrict__ flow, size_t n_px,
                                                     uint8_t* __restrict__ bgr) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_px) return;
  const float2 f = flow[i];
  const float ang = __fadd_rn((float)atan2((double)f.y, (double)f.x), 3.14159274101257324f);   // + float32(pi)
  const float v = __fsqrt_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)));
  const int hq = (int)__fmul_rn(ang, 28.6478897565411604f);                // uint8(ang * (180/pi/2)): truncation
  const int vq = (int)fminf(__fmul_rn(v, 4.f), 255.f);
  // HSV -> BGR, s = 255
  float h = __fmul_rn((float)(hq & 255), (float)(6.0 / 180.0));
  const float s = __fmul_rn(255.f, (float)(1.0 / 255.0));
  const float vf = __fmul_rn((float)vq, (float)(1.0 / 255.0));
  if (h >= 6.f) h = __fsub_rn(h, 6.f);
  const float sec = floorf(h);
  const float fr = __fsub_rn(h, sec);
  int sector = (int)sec;
  if (sector >= 6) sector -= 6;
  float tab[4];
  tab[0] = vf;
  tab[1] = __fmul_rn(vf, __fsub_rn(1.f, s));
  tab[2] = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(s, fr)));
  tab[3] = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, fr))));
  // sector -> (b, g, r) table indices, 2 bits each: {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}
  const unsigned code[6] = {1u | 3u << 2 | 0u << 4, 1u | 0u << 2 | 2u << 4, 3u | 0u << 2 | 1u << 4,
                            0u | 2u << 2 | 1u << 4, 0u | 1u << 2 | 3u << 4, 2u | 1u << 2 | 0u << 4};
  const unsigned c = code[sector];
  auto q = [&](unsigned k) {
    const float x = __fmul_rn(tab[k], 255.f);
    return (uint8_t)min(max((int)x, 0), 255);
  };
  uint8_t* o = bgr + 3 * i;
  o[0] = q(c & 3u); o[1] = q((c >> 2) & 3u); o[2] = q((c >> 4) & 3u);
}

int flow_hsv_dev(const float* flow, int n_pairs, int rows, int cols, uint8_t* bgr, cudaStream_t st) {
  const char* fn = "draw_hsv";
  B2OF_ASSERT(n_pairs >= 0 && rows > 0 && cols > 0, fn);
  if (n_pairs == 0) return B2OF_OK;
  B2OF_ASSERT(flow != nullptr && bgr != nullptr, fn);
  const size_t n_px = (size_t)n_pairs * rows * cols;
  flow_hsv_bgr<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>((const float2*)flow, n_px, bgr);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

// second half of flow_stats for accumulators filled elsewhere (the last fb_iter_ws launch of a pass)
int flow_stats_finalize_dev(float* stats, size_t n_px, int n_pairs, cudaStream_t st) {
  if (n_pairs <= 0) return B2OF_OK;
  flow_stats_final<<<cdiv(n_pairs, 128), 128, 0, st>>>(stats, n_px, n_pairs);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

int flow_stats_dev(const float* flow, int n_pairs, int rows, int cols, float* stats, cudaStream_t st) {
  const char* fn = "flow_stats";
  B2OF_ASSERT(n_pairs >= 0 && rows > 0 && cols > 0, fn);
  if (n_pairs == 0) return B2OF_OK;
  B2OF_ASSERT(flow != nullptr && stats != nullptr, fn);
  B2OF_ASSERT(((uintptr_t)stats & 7) == 0, fn);
  size_t n_px = (size_t)rows * cols;
  B2OF_CUDA(cudaMemsetAsync(stats, 0, (size_t)n_pairs * B2OF_STATS_WIDTH * sizeof(float), st));
  int bx = (int)((n_px + 256 * 16 - 1) / (256 * 16));
  if (bx > 148) bx = 148;
  flow_stats_accum<<<dim3(bx, n_pairs), 256, 0, st>>>((const float2*)flow, n_px, stats);
  B2OF_LAUNCH_CHECK();
  flow_stats_final<<<cdiv(n_pairs, 128), 128, 0, st>>>(stats, n_px, n_pairs);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

}  // namespace b2of
