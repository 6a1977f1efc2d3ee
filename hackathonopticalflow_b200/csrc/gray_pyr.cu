// K1 bgr2gray_u8 and K2 pyrdown_u8: bit-exact integer front-end.
//   K1 replaces cv2.cvtColor(BGR2GRAY)  (pathfinder_viewer.py:244, :280; DenseOF.py:481, :510; SparseOF.py:28)
//   K2 replaces the pyrDown chain inside cv2.calcOpticalFlowPyrLK (pathfinder_viewer.py:156-158; SparseOF.py:35-36)
// Both are HBM-bound byte kernels: 16-byte vector loads/stores, shared-memory halo tile for the stencil.
#include "common.cuh"

namespace b2of {

// ----------------------------------------------------------------------------------------------
// K1: gray = (3735 B + 19235 G + 9798 R + 16384) >> 15   
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t luma(uint32_t b, uint32_t g, uint32_t r) {
  return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// flat path: the whole batch is one contiguous run of pixels, 16 pixels (48 B in, 16 B out) per thread
__global__ void __launch_bounds__(256) bgr2gray_flat16(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                        size_t n_groups) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n_groups; i += stride) {
    uint4 a = __ldg(src + 3 * i), b = __ldg(src + 3 * i + 1), c = __ldg(src + 3 * i + 2);
    uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t acc = 0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        int byte = 3 * (4 * q + p);
        uint32_t bb = (w[byte >> 2] >> (8 * (byte & 3))) & 0xffu;
        uint32_t gg = (w[(byte + 1) >> 2] >> (8 * ((byte + 1) & 3))) & 0xffu;
        uint32_t rr = (w[(byte + 2) >> 2] >> (8 * ((byte + 2) & 3))) & 0xffu;
        acc |= luma(bb, gg, rr) << (8 * p);
      }
      o[q] = acc;
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// general path: arbitrary steps / alignment, one pixel per thread
__global__ void __launch_bounds__(256) bgr2gray_strided(const uint8_t* __restrict__ src, size_t src_step,
                                                         size_t src_bstride, uint8_t* __restrict__ dst,
                                                         size_t dst_step, size_t dst_bstride, int rows, int cols) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= cols) return;
  const uint8_t* s = src + blockIdx.z * src_bstride + (size_t)y * src_step + 3 * (size_t)x;
  dst[blockIdx.z * dst_bstride + (size_t)y * dst_step + x] = (uint8_t)luma(s[0], s[1], s[2]);
}

int bgr2gray_dev(const uint8_t* bgr, int rows, int cols, size_t src_step, size_t src_bstride, uint8_t* gray,
                 size_t dst_step, size_t dst_bstride, int batch, cudaStream_t st) {
  if (rows == 0 || cols == 0 || batch == 0) return B2OF_OK;
  bool contiguous = src_step == (size_t)cols * 3 && dst_step == (size_t)cols &&
                    (batch == 1 || (src_bstride == src_step * rows && dst_bstride == dst_step * rows));
  size_t n = (size_t)rows * cols * batch;
  bool aligned = ((uintptr_t)bgr % 16 == 0) && ((uintptr_t)gray % 16 == 0);
  if (contiguous && aligned && n >= 16) {
    size_t groups = n / 16;
    int blocks = (int)((groups + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    bgr2gray_flat16<<<blocks, 256, 0, st>>>((const uint4*)bgr, (uint4*)gray, groups);
    B2OF_LAUNCH_CHECK();
    size_t done = groups * 16;
    if (done < n) {  // < 16 trailing pixels
      bgr2gray_strided<<<dim3(1, 1, 1), 32, 0, st>>>(bgr + 3 * done, 0, 0, gray + done, 0, 0, 1, (int)(n - done));
      B2OF_LAUNCH_CHECK();
    }
    return B2OF_OK;
  }
  dim3 grid(cdiv(cols, 256), rows, batch);
  bgr2gray_strided<<<grid, 256, 0, st>>>(bgr, src_step, src_bstride, gray, dst_step, dst_bstride, rows, cols);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

// ----------------------------------------------------------------------------------------------
// K2: pyrDown u8. dst(y,x) = (sum_{i,j} k_i k_j src(refl(2y+i-2), refl(2x+j-2)) + 128) >> 8, k = [1 4 6 4 1]
// One CTA makes a 64x16 dst tile from a 35-row x 160-byte source tile held in shared memory.
// ----------------------------------------------------------------------------------------------
constexpr int PD_TW = 64, PD_TH = 16;
constexpr int PD_SROWS = 2 * PD_TH + 3;  // 35
constexpr int PD_SCOLS = 160;            // [2*tx0-16, 2*tx0+144): 16-byte aligned window around the 131 needed columns

__global__ void __launch_bounds__(256) pyrdown_u8_kernel(const uint8_t* __restrict__ src, size_t src_step,
                                                          size_t src_bstride, int rows, int cols,
                                                          uint8_t* __restrict__ dst, size_t dst_step,
                                                          size_t dst_bstride, int drows, int dcols, int vec_ok) {
  __shared__ __align__(16) uint8_t s_src[PD_SROWS][PD_SCOLS];
  __shared__ uint16_t s_h[PD_SROWS][PD_TW];
  const int tx0 = blockIdx.x * PD_TW, ty0 = blockIdx.y * PD_TH;
  const uint8_t* sb = src + blockIdx.z * src_bstride;
  uint8_t* db = dst + blockIdx.z * dst_bstride;
  const int gx0 = 2 * tx0 - 16;  // global x of s_src column 0
  const int gy0 = 2 * ty0 - 2;   // global y of s_src row 0
  const int t = threadIdx.x;
  bool interior_x = vec_ok && gx0 >= 0 && gx0 + PD_SCOLS <= cols;
  if (interior_x) {
    // 35 rows x 10 uint4
    for (int i = t; i < PD_SROWS * (PD_SCOLS / 16); i += 256) {
      int r = i / (PD_SCOLS / 16), c = i % (PD_SCOLS / 16);
      int gy = reflect101(gy0 + r, rows);
      uint4 v = __ldg((const uint4*)(sb + (size_t)gy * src_step + gx0) + c);
      *((uint4*)&s_src[r][c * 16]) = v;
    }
  } else {
    // only columns 14..146 are consumed (global 2*tx0-2 .. 2*tx0+130)
    for (int i = t; i < PD_SROWS * 136; i += 256) {
      int r = i / 136, c = 12 + i % 136;
      int gy = reflect101(gy0 + r, rows);
      int gx = reflect101(gx0 + c, cols);
      s_src[r][c] = sb[(size_t)gy * src_step + gx];
    }
  }
  __syncthreads();
  for (int i = t; i < PD_SROWS * PD_TW; i += 256) {
    int r = i / PD_TW, x = i % PD_TW;
    const uint8_t* p = &s_src[r][16 + 2 * x - 2];
    s_h[r][x] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
  }
  __syncthreads();
  // 16 rows x 16 groups of 4 pixels
  int y = t / 16, x4 = (t % 16) * 4;
  int oy = ty0 + y;
  if (oy >= drows) return;
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int x = x4 + k;
    uint32_t v = s_h[2 * y][x] + 4u * s_h[2 * y + 1][x] + 6u * s_h[2 * y + 2][x] + 4u * s_h[2 * y + 3][x] +
                 s_h[2 * y + 4][x];
    o[k] = (v + 128u) >> 8;
  }
  uint8_t* drow = db + (size_t)oy * dst_step;
  int ox = tx0 + x4;
  if (ox + 3 < dcols && (((uintptr_t)(drow + ox)) & 3) == 0) {
    *((uint32_t*)(drow + ox)) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ox + k < dcols) drow[ox + k] = (uint8_t)o[k];
  }
}

// K2 as a register-resident stream (the shipped form for 16-byte-aligned sources; the tile kernel above takes the rest).
// thread = eight output pixels of PS_TH consecutive output rows: a source row is ONE 16-byte load (its sixteen own
// pixels) + the aligned words left and right of it (columns -2, -1 and +16); the 5-tap row sums are IDP.4A dot products
// of byte windows with (1, 4, 6, 4) plus the fifth tap -- two instructions per output -- and stay packed two to a
// register: a row sum is <= 4080 and the vertical combination <= 65280, so (A + E) + 4 (B + D) + 6 C, the + 128 and
// the >> 8 run on both halves of a register at once.  Walking down, three of an output row's five row sums are the
// previous output row's; the two new source rows are requested one output row ahead.  No shared memory, no barrier.
// The reflected columns at the frame edges are the edge threads' own pixels; only a ragged right edge (cols not a
// multiple of 16) takes a byte-wise path.  1080p x 64 frames: 131 us (tile kernel) -> 44.5 us = 3.7 TB/s of algorithmic bytes.
#ifndef PS_TH
#define PS_TH 16
#endif
#ifndef PS_MB
#define PS_MB 8
#endif
struct PdRow { uint32_t wl, w0, w1, w2, w3, wr; };

__device__ __noinline__ PdRow pd_load_edge(const uint8_t* __restrict__ row, int sx, int cols) {
  PdRow r;
  auto b = [&](int c) { return (uint32_t)row[reflect101(c, cols)]; };
  auto w = [&](int c) { return b(c) | (b(c + 1) << 8) | (b(c + 2) << 16) | (b(c + 3) << 24); };
  r.wl = (b(sx - 2) << 16) | (b(sx - 1) << 24);
  r.w0 = w(sx); r.w1 = w(sx + 4); r.w2 = w(sx + 8); r.w3 = w(sx + 12);
  r.wr = b(sx + 16);
  return r;
}

// `ragged`: the thread's sixteen own columns cross the right frame edge (cols not a multiple of 16).  Otherwise the
// only reflected columns are -2, -1 of the first thread of a row (= its own columns 2, 1) and column cols of the last
// (= its own column 14): fixed up from the thread's own vector, no extra path.
__device__ __forceinline__ PdRow pd_load(const uint8_t* __restrict__ row, int sx, int cols, bool ragged) {
  if (ragged) return pd_load_edge(row, sx, cols);
  PdRow r;
  const uint4 v = __ldg((const uint4*)(row + sx));
  r.w0 = v.x; r.w1 = v.y; r.w2 = v.z; r.w3 = v.w;
  r.wl = sx >= 4 ? __ldg((const uint32_t*)(row + sx - 4)) : ((v.x & 0x00ff0000u) | ((v.x & 0x0000ff00u) << 16));
  r.wr = sx + 16 < cols ? __ldg((const uint32_t*)(row + sx + 16)) : ((v.w >> 16) & 0xffu);
  return r;
}

// row sums of the eight outputs, packed: (h0 | h1 << 16, h2 | h3 << 16, h4 | h5 << 16, h6 | h7 << 16)
__device__ __forceinline__ uint4 pd_hsum(const PdRow& r) {
  const uint32_t K = 0x04060401u, B0 = 0x00000001u, B2 = 0x00010000u;     // taps (1, 4, 6, 4); the fifth tap's byte
  const uint32_t m0 = __byte_perm(r.wl, r.w0, 0x5432), m1 = __byte_perm(r.w0, r.w1, 0x5432),
                 m2 = __byte_perm(r.w1, r.w2, 0x5432), m3 = __byte_perm(r.w2, r.w3, 0x5432);
  const uint32_t h0 = __dp4a(r.w0, B2, __dp4a(m0, K, 0u)), h1 = __dp4a(r.w1, B0, __dp4a(r.w0, K, 0u));
  const uint32_t h2 = __dp4a(r.w1, B2, __dp4a(m1, K, 0u)), h3 = __dp4a(r.w2, B0, __dp4a(r.w1, K, 0u));
  const uint32_t h4 = __dp4a(r.w2, B2, __dp4a(m2, K, 0u)), h5 = __dp4a(r.w3, B0, __dp4a(r.w2, K, 0u));
  const uint32_t h6 = __dp4a(r.w3, B2, __dp4a(m3, K, 0u)), h7 = __dp4a(r.wr, B0, __dp4a(r.w3, K, 0u));
  return make_uint4(h0 | (h1 << 16), h2 | (h3 << 16), h4 | (h5 << 16), h6 | (h7 << 16));
}

template <bool RAGGED>     // RAGGED = false: cols is a multiple of 16, no byte-wise path in the kernel at all
__global__ void __launch_bounds__(128, PS_MB) pyrdown_u8_stream(const uint8_t* __restrict__ src, size_t src_step,
                                                          size_t src_bstride, int rows, int cols,
                                                          uint8_t* __restrict__ dst, size_t dst_step,
                                                          size_t dst_bstride, int drows, int dcols) {
  const int x = (blockIdx.x * 128 + threadIdx.x) * 8;           // first of this thread's eight output columns
  const int y0 = blockIdx.y * PS_TH;
  if (x >= dcols) return;
  const int sx = 2 * x;
  const bool fast = RAGGED && sx + 16 > cols;      // (the `ragged` flag of pd_load)
  const uint8_t* sb = src + blockIdx.z * src_bstride;
  uint8_t* db = dst + blockIdx.z * dst_bstride;
  auto srow = [&](int sy) { return sb + (size_t)reflect101(sy, rows) * src_step; };
  uint4 hA = pd_hsum(pd_load(srow(2 * y0 - 2), sx, cols, fast));
  uint4 hB = pd_hsum(pd_load(srow(2 * y0 - 1), sx, cols, fast));
  uint4 hC = pd_hsum(pd_load(srow(2 * y0), sx, cols, fast));
  PdRow n0 = pd_load(srow(2 * y0 + 1), sx, cols, fast), n1 = pd_load(srow(2 * y0 + 2), sx, cols, fast);
  const int y1 = min(y0 + PS_TH, drows);
  const bool wide = x + 8 <= dcols && ((((uintptr_t)db + x) | dst_step) & 7) == 0;
#pragma unroll 1
  for (int y = y0; y < y1; ++y) {
    const PdRow c0 = n0, c1 = n1;
    if (y + 1 < y1) { n0 = pd_load(srow(2 * y + 3), sx, cols, fast); n1 = pd_load(srow(2 * y + 4), sx, cols, fast); }
    const uint4 hD = pd_hsum(c0), hE = pd_hsum(c1);
    auto vert = [](uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
      return (((a + e) + 4u * (b + d) + 6u * c + 0x00800080u) >> 8) & 0x00ff00ffu;
    };
    const uint32_t v01 = vert(hA.x, hB.x, hC.x, hD.x, hE.x), v23 = vert(hA.y, hB.y, hC.y, hD.y, hE.y);
    const uint32_t v45 = vert(hA.z, hB.z, hC.z, hD.z, hE.z), v67 = vert(hA.w, hB.w, hC.w, hD.w, hE.w);
    const uint32_t o0 = __byte_perm(v01, v23, 0x6420), o1 = __byte_perm(v45, v67, 0x6420);
    uint8_t* o = db + (size_t)y * dst_step + x;
    if (wide) {
      *(uint2*)o = make_uint2(o0, o1);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (x + k < dcols) o[k] = (uint8_t)((k < 4 ? o0 : o1) >> (8 * (k & 3)));
    }
    hA = hC; hB = hD; hC = hE;
  }
}

int pyrdown_dev(const uint8_t* src, int rows, int cols, size_t src_step, size_t src_bstride, uint8_t* dst,
                size_t dst_step, size_t dst_bstride, int batch, cudaStream_t st) {
  if (rows == 0 || cols == 0 || batch == 0) return B2OF_OK;
  int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
  int vec_ok = ((uintptr_t)src % 16 == 0) && (src_step % 16 == 0) && (src_bstride % 16 == 0);
  if (vec_ok && cols >= 24 && rows >= 2) {
    dim3 grid(cdiv(cdiv(dcols, 8), 128), cdiv(drows, PS_TH), batch);
    if (cols % 16 == 0)
      pyrdown_u8_stream<false><<<grid, 128, 0, st>>>(src, src_step, src_bstride, rows, cols, dst, dst_step, dst_bstride,
                                                     drows, dcols);
    else
      pyrdown_u8_stream<true><<<grid, 128, 0, st>>>(src, src_step, src_bstride, rows, cols, dst, dst_step, dst_bstride,
                                                    drows, dcols);
    B2OF_LAUNCH_CHECK();
    return B2OF_OK;
  }
  dim3 grid(cdiv(dcols, PD_TW), cdiv(drows, PD_TH), batch);
  pyrdown_u8_kernel<<<grid, 256, 0, st>>>(src, src_step, src_bstride, rows, cols, dst, dst_step, dst_bstride, drows,
                                          dcols, vec_ok);
  B2OF_LAUNCH_CHECK();
  return B2OF_OK;
}

}  // namespace b2of
