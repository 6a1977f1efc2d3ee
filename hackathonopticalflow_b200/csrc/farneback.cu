// K3-K6: dense Farneback flow, replaces cv2.calcOpticalFlowFarneback as called at DenseOF.py:147-156
// (wrapper DenseOF.py:127-157, call site :520).  Arithmetic spec: SURVEY.md App. A.3 (opencv optflowgf.cpp,
// GaussianBlur, resize -- third-party, restated in oracle/farneback.py).
//
// Data layout in HBM (all float32, per pyramid level k of size w_k x h_k, pitch_k = w_k rounded up to 32):
//   I_k   [frame][h_k][pitch_k]            blurred + resampled level image            (K3)
//   R_k   [frame][5][h_k][pitch_k]         polynomial expansion, PLANAR channels      (K4)
//   F_k   [pair][h_k][pitch_k] float2      flow ping/pong buffers                     (K5)
// M (the 5-channel matrix field) never leaves the SM: UpdateMatrices, the (2m+1)^2 window sum and the 2x2
// solve are one kernel per (level, iteration); the x(1/pyr_scale) bilinear flow upsample is folded into
// the first iteration of each level.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b2of {

constexpr int FB_MIN_SIZE = 32;
constexpr int FB_MAX_POLY_N = 16;

struct PolyConst {
  float g[FB_MAX_POLY_N + 1], xg[FB_MAX_POLY_N + 1], xxg[FB_MAX_POLY_N + 1];
  float ig11, ig03, ig33, ig55;
};

struct FbLevel {
  int k, w, h, pitch, ksz;
  double scale, sigma;
  // device tables (owned by the plan)
  float* taps;            // ksz blur taps
  int *sx0, *sx1, *sy0, *sy1;  // resize source indices
  float *fx, *fy;         // resize fractions
  // flow upsample tables from the next-coarser level (valid when k < top level)
  int *ux0, *ux1, *uy0, *uy1;
  float *ufx, *ufy;
  // fused level-image kernel: output tile and the largest source region any tile needs (0 = use two-pass path)
  int f_tw, f_th, f_rw, f_rh, f_rw_pad;
  size_t f_smem;
  // regular (exact 1/S) form: K combined taps, or r_K == 0
  int r_K, r_S, r_c0;
  float r_c[24];
};

struct FbPlan {
  int rows, cols;
  b2of_farneback_params p;
  std::vector<FbLevel> lv;  // coarsest first
  PolyConst pc;
  float* gauss_taps;  // winsize/2+1 taps for the FARNEBACK_GAUSSIAN window (device)
  // OPTFLOW_USE_INITIAL_FLOW: INTER_AREA decimation tables full-res -> coarsest level (device, CSR form)
  int *ax_start, *ax_src, *ay_start, *ay_src;
  float *ax_alpha, *ay_alpha;
  int area_fast, area_sx, area_sy;   // exact integer factors: plain block average
  void* table_block;
  size_t S;  // sum of h*pitch over levels (floats per plane per frame)
};

// ----------------------------------------------------------------------------------------------
// host-side plan: level sizes, Gaussian taps, resize tables -- computed exactly as cv2 does
// ----------------------------------------------------------------------------------------------
static void gaussian_kernel(int n, double sigma, std::vector<float>& out) {
  out.resize(n);
  static const float small3[] = {0.25f, 0.5f, 0.25f};
  static const float small5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  static const float small7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
  if (sigma <= 0 && (n == 1 || n == 3 || n == 5 || n == 7)) {
    const float* t = n == 3 ? small3 : n == 5 ? small5 : small7;
    if (n == 1) { out[0] = 1.f; return; }
    for (int i = 0; i < n; ++i) out[i] = t[i];
    return;
  }
  double s = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
  std::vector<double> k(n);
  double sum = 0;
  for (int i = 0; i < n; ++i) {
    double x = i - (n - 1) * 0.5;
    k[i] = exp(-(x * x) / (2.0 * s * s));
    sum += k[i];
  }
  for (int i = 0; i < n; ++i) out[i] = (float)(k[i] / sum);
}

static void linear_tables(int dst_n, int src_n, std::vector<int>& s0, std::vector<int>& s1, std::vector<float>& f) {
  s0.resize(dst_n); s1.resize(dst_n); f.resize(dst_n);
  double scale = (double)src_n / dst_n;
  for (int i = 0; i < dst_n; ++i) {
    float ff = (float)((i + 0.5) * scale - 0.5);
    int s = (int)floorf(ff);
    ff -= s;
    if (s < 0) { s = 0; ff = 0; }
    if (s >= src_n - 1) { s = src_n - 1; ff = 0; }
    s0[i] = s;
    s1[i] = s + 1 < src_n ? s + 1 : src_n - 1;
    f[i] = ff;
  }
}

static void invert6(double G[6][6], double inv[6][6]) {
  double a[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { a[i][j] = G[i][j]; a[i][j + 6] = i == j; }
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    for (int j = 0; j < 12; ++j) { double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
    double d = a[c][c];
    for (int j = 0; j < 12; ++j) a[c][j] /= d;
    for (int r = 0; r < 6; ++r) if (r != c) {
      double m = a[r][c];
      if (m != 0) for (int j = 0; j < 12; ++j) a[r][j] -= m * a[c][j];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) inv[i][j] = a[i][j + 6];
}

static void poly_constants(int n, double sigma, PolyConst& pc) {
  if (sigma < 1.1920929e-07) sigma = n * 0.3;
  std::vector<float> g(2 * n + 1);
  double s = 0;
  for (int x = -n; x <= n; ++x) { g[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
  s = 1. / s;
  for (int x = -n; x <= n; ++x) g[x + n] = (float)(g[x + n] * s);
  double G[6][6] = {{0}};
  for (int y = -n; y <= n; ++y)
    for (int x = -n; x <= n; ++x) {
      double gg = (double)g[y + n] * g[x + n];
      G[0][0] += gg;
      G[1][1] += gg * x * x;
      G[3][3] += gg * x * x * x * x;
      G[5][5] += gg * x * x * y * y;
    }
  G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
  G[4][4] = G[3][3];
  G[3][4] = G[4][3] = G[5][5];
  double inv[6][6];
  invert6(G, inv);
  memset(&pc, 0, sizeof pc);
  for (int k = 0; k <= n; ++k) {
    pc.g[k] = g[n + k];
    pc.xg[k] = (float)(k * g[n + k]);
    pc.xxg[k] = (float)(k * k * g[n + k]);
  }
  pc.ig11 = (float)inv[1][1];
  pc.ig03 = (float)inv[0][3];
  pc.ig33 = (float)inv[3][3];
  pc.ig55 = (float)inv[5][5];
}

// (K, S) pairs with a compiled regular level-image kernel: the reference's four levels
static bool level_regular_supported(int K, int S) {
  return (K == 3 && S == 1) || (K == 4 && S == 2) || (K == 10 && S == 4) || (K == 20 && S == 8);
}

struct PlanKey {
  int dev, rows, cols, levels, winsize, poly_n, flags;
  double pyr_scale, poly_sigma;
  bool operator<(const PlanKey& o) const {
    return memcmp(this, &o, sizeof(PlanKey)) < 0;
  }
};
static std::mutex g_plan_mu;
static std::map<PlanKey, FbPlan*> g_plans;

static int build_plan(int rows, int cols, const b2of_farneback_params& p, FbPlan** out) {
  FbPlan* pl = new FbPlan();
  pl->rows = rows; pl->cols = cols; pl->p = p;
  int k = 0;
  double scale = 1;
  for (; k < p.levels; ++k) {
    scale *= p.pyr_scale;
    if (cols * scale < FB_MIN_SIZE || rows * scale < FB_MIN_SIZE) break;
  }
  int top = k;
  std::vector<std::vector<float>> fl;
  std::vector<std::vector<int>> il;
  for (int lvl = top; lvl >= 0; --lvl) {
    FbLevel L{};
    L.k = lvl;
    double sc = 1;
    for (int i = 0; i < lvl; ++i) sc *= p.pyr_scale;
    L.scale = sc;
    L.sigma = (1. / sc - 1) * 0.5;
    int ks = cv_round(L.sigma * 5) | 1;
    L.ksz = ks > 3 ? ks : 3;
    L.w = cv_round(cols * sc);
    L.h = cv_round(rows * sc);
    L.pitch = (int)align_up(L.w, 32);
    pl->lv.push_back(L);
  }
  // pack all tables into one device block
  std::vector<char> host;
  auto put = [&](const void* src, size_t bytes) {
    size_t off = align_up(host.size(), 16);
    host.resize(off + bytes);
    memcpy(host.data() + off, src, bytes);
    return off;
  };
  struct Offs { size_t taps, sx0, sx1, sy0, sy1, fx, fy, ux0, ux1, uy0, uy1, ufx, ufy; };
  std::vector<Offs> offs(pl->lv.size());
  pl->S = 0;
  for (size_t i = 0; i < pl->lv.size(); ++i) {
    FbLevel& L = pl->lv[i];
    pl->S += (size_t)L.h * L.pitch;
    std::vector<float> taps;
    gaussian_kernel(L.ksz, L.sigma, taps);
    offs[i].taps = put(taps.data(), taps.size() * 4);
    std::vector<int> s0, s1;
    std::vector<float> f;
    linear_tables(L.w, cols, s0, s1, f);
    offs[i].sx0 = put(s0.data(), s0.size() * 4); offs[i].sx1 = put(s1.data(), s1.size() * 4); offs[i].fx = put(f.data(), f.size() * 4);
    std::vector<int> t0, t1;
    std::vector<float> tf;
    linear_tables(L.h, rows, t0, t1, tf);
    offs[i].sy0 = put(t0.data(), t0.size() * 4); offs[i].sy1 = put(t1.data(), t1.size() * 4); offs[i].fy = put(tf.data(), tf.size() * 4);
    {
      // regular form: sx0 = S*x + off, sx1 = sx0 + (f != 0), f constant -- in both axes with the same S, off, f
      L.r_K = 0;
      std::vector<float> taps_h;
      gaussian_kernel(L.ksz, L.sigma, taps_h);
      auto regular = [&](const std::vector<int>& a0, const std::vector<int>& a1, const std::vector<float>& ff, int n,
                         int src_n, int& S, int& off, float& fr) {
        if (n < 2) return false;
        S = a0[1] - a0[0]; off = a0[0]; fr = ff[0];
        if (S < 1) return false;
        for (int q = 0; q < n; ++q) {
          if (a0[q] != S * q + off || ff[q] != fr) return false;
          if (fr != 0.f && (a1[q] != a0[q] + 1 || a1[q] > src_n - 1)) return false;
        }
        return true;
      };
      int Sx, Sy, ox, oy; float fxr, fyr;
      if (regular(s0, s1, f, L.w, cols, Sx, ox, fxr) && regular(t0, t1, tf, L.h, rows, Sy, oy, fyr) && Sx == Sy &&
          ox == oy && fxr == fyr) {
        int K = L.ksz + (fxr != 0.f ? 1 : 0);
        // source overshoot at the borders must stay below one reflection (reflect_once)
        if (level_regular_supported(K, Sx) && K < std::min(rows, cols)) {
          for (int j = 0; j < K; ++j) {
            float a = j < L.ksz ? taps_h[j] * (1.f - fxr) : 0.f;
            float b = (j >= 1 && fxr != 0.f) ? taps_h[j - 1] * fxr : 0.f;
            L.r_c[j] = a + b;
          }
          L.r_K = K; L.r_S = Sx; L.r_c0 = ox - L.ksz / 2;
        }
      }
    }
    {
      // pick the largest output tile whose source region fits in ~100 KB of shared memory (2 CTAs/SM)
      const int cand[4][2] = {{64, 16}, {32, 16}, {32, 8}, {16, 8}};
      L.f_tw = 0;
      int r = L.ksz / 2;
      for (int c = 0; c < 4 && !L.f_tw; ++c) {
        int tw = cand[c][0], th = cand[c][1], rw = 0, rh = 0;
        for (int x0 = 0; x0 < L.w; x0 += tw) {
          int x1 = std::min(x0 + tw, L.w) - 1;
          rw = std::max(rw, s1[x1] + r - (s0[x0] - r) + 1);
        }
        for (int y0 = 0; y0 < L.h; y0 += th) {
          int y1 = std::min(y0 + th, L.h) - 1;
          rh = std::max(rh, t1[y1] + r - (t0[y0] - r) + 1);
        }
        int rw_pad = (int)align_up(rw, 16);
        size_t smem = (size_t)rh * rw_pad + (size_t)rh * tw * sizeof(float);
        if (smem <= 100 * 1024) {
          L.f_tw = tw; L.f_th = th; L.f_rw = rw; L.f_rh = rh; L.f_rw_pad = rw_pad; L.f_smem = smem;
        }
      }
    }
    if (i > 0) {
      const FbLevel& C = pl->lv[i - 1];
      linear_tables(L.w, C.w, s0, s1, f);
      offs[i].ux0 = put(s0.data(), s0.size() * 4); offs[i].ux1 = put(s1.data(), s1.size() * 4); offs[i].ufx = put(f.data(), f.size() * 4);
      linear_tables(L.h, C.h, s0, s1, f);
      offs[i].uy0 = put(s0.data(), s0.size() * 4); offs[i].uy1 = put(s1.data(), s1.size() * 4); offs[i].ufy = put(f.data(), f.size() * 4);
    }
  }
  // FARNEBACK_GAUSSIAN window taps (float32, as cv2)
  int m = p.winsize / 2;
  std::vector<float> gk(m + 1);
  {
    double sigma = m * 0.3;
    float s = 1.f;
    gk[0] = 1.f;
    for (int i = 1; i <= m; ++i) {
      float t = (float)exp(-i * i / (2 * sigma * sigma));
      gk[i] = t;
      s += t * 2;
    }
    s = 1.f / s;
    for (int i = 0; i <= m; ++i) gk[i] = gk[i] * s;
  }
  size_t gk_off = put(gk.data(), gk.size() * 4);
  // INTER_AREA tables (cv::computeResizeAreaTab) from the full-resolution flow to the coarsest level
  auto area_tab = [](int ssize, int dsize, std::vector<int>& start, std::vector<int>& src, std::vector<float>& alpha) {
    double scale = (double)ssize / dsize;
    start.assign(1, 0); src.clear(); alpha.clear();
    for (int dx = 0; dx < dsize; ++dx) {
      double fsx1 = dx * scale, fsx2 = fsx1 + scale;
      double cell = std::min(scale, ssize - fsx1);
      int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
      sx2 = std::min(sx2, ssize - 1);
      sx1 = std::min(sx1, sx2);
      if (sx1 - fsx1 > 1e-3) { src.push_back(sx1 - 1); alpha.push_back((float)((sx1 - fsx1) / cell)); }
      for (int sx = sx1; sx < sx2; ++sx) { src.push_back(sx); alpha.push_back((float)(1.0 / cell)); }
      if (fsx2 - sx2 > 1e-3) { src.push_back(sx2); alpha.push_back((float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)); }
      start.push_back((int)src.size());
    }
  };
  std::vector<int> axs, axi, ays, ayi;
  std::vector<float> axa, aya;
  const FbLevel& C0 = pl->lv[0];
  area_tab(cols, C0.w, axs, axi, axa);
  area_tab(rows, C0.h, ays, ayi, aya);
  {
    double scx = (double)cols / C0.w, scy = (double)rows / C0.h;
    int ix = (int)nearbyint(scx), iy = (int)nearbyint(scy);
    pl->area_fast = fabs(scx - ix) < 2.220446049250313e-16 && fabs(scy - iy) < 2.220446049250313e-16;
    pl->area_sx = ix; pl->area_sy = iy;
  }
  size_t o_axs = put(axs.data(), axs.size() * 4), o_axi = put(axi.data(), axi.size() * 4 + 4), o_axa = put(axa.data(), axa.size() * 4 + 4);
  size_t o_ays = put(ays.data(), ays.size() * 4), o_ayi = put(ayi.data(), ayi.size() * 4 + 4), o_aya = put(aya.data(), aya.size() * 4 + 4);
  cudaError_t e = cudaMalloc(&pl->table_block, host.size());
  if (e != cudaSuccess) { delete pl; return fail(B2OF_E_NOMEM, "cudaMalloc(plan tables) failed: %s", cudaGetErrorString(e)); }
  e = cudaMemcpy(pl->table_block, host.data(), host.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(pl->table_block); delete pl; return fail(B2OF_E_CUDA, "plan table upload failed: %s", cudaGetErrorString(e)); }
  char* base = (char*)pl->table_block;
  for (size_t i = 0; i < pl->lv.size(); ++i) {
    FbLevel& L = pl->lv[i];
    L.taps = (float*)(base + offs[i].taps);
    L.sx0 = (int*)(base + offs[i].sx0); L.sx1 = (int*)(base + offs[i].sx1); L.fx = (float*)(base + offs[i].fx);
    L.sy0 = (int*)(base + offs[i].sy0); L.sy1 = (int*)(base + offs[i].sy1); L.fy = (float*)(base + offs[i].fy);
    if (i > 0) {
      L.ux0 = (int*)(base + offs[i].ux0); L.ux1 = (int*)(base + offs[i].ux1); L.ufx = (float*)(base + offs[i].ufx);
      L.uy0 = (int*)(base + offs[i].uy0); L.uy1 = (int*)(base + offs[i].uy1); L.ufy = (float*)(base + offs[i].ufy);
    }
  }
  pl->gauss_taps = (float*)(base + gk_off);
  pl->ax_start = (int*)(base + o_axs); pl->ax_src = (int*)(base + o_axi); pl->ax_alpha = (float*)(base + o_axa);
  pl->ay_start = (int*)(base + o_ays); pl->ay_src = (int*)(base + o_ayi); pl->ay_alpha = (float*)(base + o_aya);
  poly_constants(p.poly_n, p.poly_sigma, pl->pc);
  *out = pl;
  return B2OF_OK;
}

// b2of_release(): drop every cached plan (their device tables) -- the caller guarantees no call is in flight
void farneback_release() {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_plans) {
    cudaSetDevice(kv.first.dev);
    cudaFree(kv.second->table_block);
    delete kv.second;
  }
  g_plans.clear();
  cudaSetDevice(cur);
}

static int get_plan(int rows, int cols, const b2of_farneback_params& p, FbPlan** out) {
  PlanKey key;
  memset(&key, 0, sizeof key);
  cudaGetDevice(&key.dev);
  key.rows = rows; key.cols = cols; key.levels = p.levels; key.winsize = p.winsize; key.poly_n = p.poly_n;
  key.flags = p.flags & B2OF_OPTFLOW_FARNEBACK_GAUSSIAN;
  key.pyr_scale = p.pyr_scale; key.poly_sigma = p.poly_sigma;
  std::lock_guard<std::mutex> lock(g_plan_mu);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) { *out = it->second; return B2OF_OK; }
  FbPlan* pl = nullptr;
  int rc = build_plan(rows, cols, p, &pl);
  if (rc) return rc;
  g_plans[key] = pl;
  *out = pl;
  return B2OF_OK;
}

// ----------------------------------------------------------------------------------------------
// K3: level image = resize_linear(GaussianBlur(float(frame)))  -- two separable passes, each fusing the
// 1-D blur (REFLECT_101) with the 1-D linear resample.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fb_level_hpass(const uint8_t* __restrict__ frames, size_t step,
                                                       size_t frame_stride, int W, int H, float* __restrict__ T,
                                                       int wk, int pitch, size_t t_frame_stride,
                                                       const float* __restrict__ taps, int ksz,
                                                       const int* __restrict__ sx0, const int* __restrict__ sx1,
                                                       const float* __restrict__ fxs) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= wk) return;
  const uint8_t* row = frames + blockIdx.z * frame_stride + (size_t)y * step;
  int r = ksz >> 1;
  int a = sx0[x], b = sx1[x];
  float f = fxs[x];
  float va = 0.f, vb = 0.f;
  bool interior = a - r >= 0 && b + r < W;
  if (interior) {
    for (int t = 0; t < ksz; ++t) {
      float kt = __ldg(taps + t);
      va = fmaf(kt, (float)row[a + t - r], va);
      vb = fmaf(kt, (float)row[b + t - r], vb);
    }
  } else {
    for (int t = 0; t < ksz; ++t) {
      float kt = __ldg(taps + t);
      va = fmaf(kt, (float)row[reflect101(a + t - r, W)], va);
      vb = fmaf(kt, (float)row[reflect101(b + t - r, W)], vb);
    }
  }
  T[blockIdx.z * t_frame_stride + (size_t)y * pitch + x] = va * (1.f - f) + vb * f;
}

__global__ void __launch_bounds__(128) fb_level_vpass(const float* __restrict__ T, int H, int pitch,
                                                       size_t t_frame_stride, float* __restrict__ I, int wk, int hk,
                                                       size_t i_frame_stride, const float* __restrict__ taps, int ksz,
                                                       const int* __restrict__ sy0, const int* __restrict__ sy1,
                                                       const float* __restrict__ fys) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= wk) return;
  const float* tb = T + blockIdx.z * t_frame_stride + x;
  int r = ksz >> 1;
  int a = sy0[y], b = sy1[y];
  float f = fys[y];
  float va = 0.f, vb = 0.f;
  for (int t = 0; t < ksz; ++t) {
    float kt = __ldg(taps + t);
    va = fmaf(kt, tb[(size_t)reflect101(a + t - r, H) * pitch], va);
    vb = fmaf(kt, tb[(size_t)reflect101(b + t - r, H) * pitch], vb);
  }
  I[blockIdx.z * i_frame_stride + (size_t)y * pitch + x] = va * (1.f - f) + vb * f;
}

// ----------------------------------------------------------------------------------------------
// K3 (fused form): one kernel per level straight from the u8 frame.  The CTA stages the source region of its
// output tile in shared memory (REFLECT_101 applied while loading), runs the horizontal blur+resample into a
// second shared buffer and the vertical blur+resample out of it.  Used whenever the region fits in shared
// memory; otherwise the two-pass kernels above (global scratch T) take over.
// ----------------------------------------------------------------------------------------------
struct LevelFusedArgs {
  const uint8_t* frames; size_t step, frame_stride; int W, H;
  float* I; int wk, hk, pitch; size_t i_frame_stride;
  const float* taps; int ksz;
  const int *sx0, *sx1, *sy0, *sy1; const float *fx, *fy;
  int tw, th;          // output tile
  int rw, rh;          // max source region (cols, rows) over all tiles
  int rw_pad;          // rw rounded up to 16
};

__global__ void __launch_bounds__(256) fb_level_fused(LevelFusedArgs a) {
  extern __shared__ __align__(16) unsigned char lsm[];
  uint8_t* s_src = lsm;                                        // [rh][rw_pad]
  float* s_h = (float*)(lsm + (size_t)a.rh * a.rw_pad);        // [rh][tw]
  const int x0 = blockIdx.x * a.tw, y0 = blockIdx.y * a.th;
  const int x1 = min(x0 + a.tw, a.wk) - 1, y1 = min(y0 + a.th, a.hk) - 1;
  const int r = a.ksz >> 1;
  const int c_lo = a.sx0[x0] - r, c_hi = a.sx1[x1] + r;
  const int r_lo = a.sy0[y0] - r, r_hi = a.sy1[y1] + r;
  const int nc = c_hi - c_lo + 1, nr = r_hi - r_lo + 1;
  const uint8_t* fb = a.frames + blockIdx.z * a.frame_stride;
  const int t = threadIdx.x;
  for (int i = t; i < nr * nc; i += 256) {
    int rr = i / nc, cc = i - rr * nc;
    int gy = reflect101(r_lo + rr, a.H), gx = reflect101(c_lo + cc, a.W);
    s_src[rr * a.rw_pad + cc] = fb[(size_t)gy * a.step + gx];
  }
  __syncthreads();
  const int tw = x1 - x0 + 1;
  for (int i = t; i < nr * tw; i += 256) {
    int rr = i / tw, x = i - rr * tw;
    int ia = a.sx0[x0 + x] - c_lo - r, ib = a.sx1[x0 + x] - c_lo - r;
    float f = a.fx[x0 + x];
    const uint8_t* row = s_src + rr * a.rw_pad;
    float va = 0.f, vb = 0.f;
    for (int k = 0; k < a.ksz; ++k) {
      float kt = __ldg(a.taps + k);
      va = fmaf(kt, (float)row[ia + k], va);
      vb = fmaf(kt, (float)row[ib + k], vb);
    }
    s_h[rr * a.tw + x] = va * (1.f - f) + vb * f;
  }
  __syncthreads();
  const int th = y1 - y0 + 1;
  float* ob = a.I + blockIdx.z * a.i_frame_stride;
  for (int i = t; i < th * tw; i += 256) {
    int y = i / tw, x = i - y * tw;
    int ia = a.sy0[y0 + y] - r_lo - r, ib = a.sy1[y0 + y] - r_lo - r;
    float f = a.fy[y0 + y];
    float va = 0.f, vb = 0.f;
    for (int k = 0; k < a.ksz; ++k) {
      float kt = __ldg(a.taps + k);
      va = fmaf(kt, s_h[(ia + k) * a.tw + x], va);
      vb = fmaf(kt, s_h[(ib + k) * a.tw + x], vb);
    }
    ob[(size_t)(y0 + y) * a.pitch + x0 + x] = va * (1.f - f) + vb * f;
  }
}


// ----------------------------------------------------------------------------------------------
// K3 (regular form): exact 1/S down-scales (the reference's pyr_scale = 0.5 gives S = 1, 2, 4, 8).  There the
// resample taps are the same for every output pixel (sx0 = S*x + off, sx1 = sx0 + 1, f constant), so blur and
// linear resample collapse into ONE K-tap filter per axis (K = ksz, or ksz + 1 when f != 0) whose taps sit in
// the constant bank.  Source region as u8 in shared memory (word loads), horizontal pass with lanes along rows
// into a transposed float buffer, vertical pass with lanes along x: bank-conflict-free for every S.
// ----------------------------------------------------------------------------------------------
struct LevelRegArgs {
  const uint8_t* frames; size_t step, frame_stride; int W, H;
  float* I; int wk, hk, pitch; size_t i_frame_stride;
  int c0;              // first source column/row of output 0 (= sx0[0] - ksz/2); output x starts at S*x + c0
  float c[24];         // combined taps
};

constexpr int LR_TW = 64;
// output rows per CTA: taller tiles for the fine levels (the 64 x 16 tiles were CTA-launch bound: 132,600 CTAs for
// 65 frames at 1080p)
__host__ __device__ constexpr int lr_th(int S) { return S == 1 ? 64 : S == 2 ? 32 : S == 4 ? 16 : 8; }
__host__ __device__ constexpr int lr_rwp(int K, int S) { return ((S * (LR_TW - 1) + K + 3 + 3) / 4 + 1) * 4; }
__host__ __device__ constexpr int lr_rh(int K, int S) { return S * (lr_th(S) - 1) + K; }
__host__ __device__ constexpr size_t lr_smem(int K, int S) {
  return (size_t)lr_rh(K, S) * lr_rwp(K, S) + (size_t)lr_rh(K, S) * LR_TW * sizeof(float);
}

__device__ __forceinline__ int reflect_once(int i, int n) {  // REFLECT_101 for |overshoot| < n
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// threads per CTA: 512 for the coarsest level (its 59 KB source region allows three CTAs per SM: more warps per CTA
// instead), 256 otherwise
__host__ __device__ constexpr int lr_nt(int S) { return S >= 8 ? 512 : 256; }

template <int K, int S>
__global__ void __launch_bounds__(lr_nt(S)) fb_level_regular(LevelRegArgs a) {
  constexpr int NT = lr_nt(S), NQ = NT / LR_TW;
  constexpr int TH = lr_th(S), RWP = lr_rwp(K, S), RH = lr_rh(K, S), RW = S * (LR_TW - 1) + K;
  extern __shared__ __align__(16) unsigned char lsm[];
  uint8_t* s_src = lsm;                               // [RH][RWP] u8
  float* s_h = (float*)(lsm + (size_t)RH * RWP);      // [RH][LR_TW]  (RH*RWP is a multiple of 4)
  const int x0 = blockIdx.x * LR_TW, y0 = blockIdx.y * TH;
  const int gx0 = S * x0 + a.c0, gy0 = S * y0 + a.c0;
  const int lpad = gx0 & 3;                           // region starts at the word boundary below gx0
  const int gxa = gx0 - lpad;
  constexpr int NW = (3 + RW + 3) / 4;                // words per region row (upper bound)
  const uint8_t* fb = a.frames + blockIdx.z * a.frame_stride;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool cols_in = gxa >= 0 && gxa + 4 * NW <= a.W && ((a.step | (size_t)fb) & 3) == 0;
  // only what valid outputs consume is loaded (keeps the border overshoot below one reflection)
  const int rows_needed = S * (min(TH, a.hk - y0) - 1) + K;
  const int cols_needed = lpad + S * (min(LR_TW, a.wk - x0) - 1) + K;
  for (int rr = warp; rr < rows_needed; rr += NT / 32) {
    const uint8_t* row = fb + (size_t)reflect_once(gy0 + rr, a.H) * a.step;
    if (cols_in) {
      for (int wc = lane; wc < NW; wc += 32)
        *((uint32_t*)(s_src + rr * RWP) + wc) = __ldg((const uint32_t*)(row + gxa) + wc);
    } else {
      for (int cc = lane; cc < cols_needed; cc += 32) s_src[rr * RWP + cc] = row[reflect_once(gxa + cc, a.W)];
    }
  }
  __syncthreads();
  const int x = t & (LR_TW - 1), q = t >> 6;          // 64 columns x NQ row phases
  {
    const uint8_t* p = s_src + lpad + S * x;
    for (int rr = q; rr < rows_needed; rr += NQ) {
      const uint8_t* pr = p + rr * RWP;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) v = fmaf(a.c[j], (float)pr[j], v);
      s_h[rr * LR_TW + x] = v;
    }
  }
  __syncthreads();
  if (x0 + x < a.wk) {
    float* ob = a.I + blockIdx.z * a.i_frame_stride + x0 + x;
#pragma unroll
    for (int y = q; y < TH; y += NQ) {
      if (y0 + y >= a.hk) break;
      const float* p = s_h + (S * y) * LR_TW + x;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) v = fmaf(a.c[j], p[j * LR_TW], v);
      ob[(size_t)(y0 + y) * a.pitch] = v;
    }
  }
}

template <int K, int S>
static void launch_level_regular(const LevelRegArgs& ra, int frames, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    cudaFuncSetAttribute(fb_level_regular<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lr_smem(K, S));
  dim3 g(cdiv(ra.wk, LR_TW), cdiv(ra.hk, lr_th(S)), frames);
  fb_level_regular<K, S><<<g, lr_nt(S), lr_smem(K, S), st>>>(ra);
}


// ----------------------------------------------------------------------------------------------
// K3 (full-resolution level, streaming form): the regular (K, S) = (3, 1) level -- a 3-tap blur, no resampling -- is a
// pure byte-in / float-out stream (N + 4 N bytes per frame), and the tile kernel above spends it on two shared-memory
// passes and two barriers.  Here a thread owns four adjacent columns (one aligned 32-bit load per row, one 16-byte
// store per row) and walks down L0_TH rows: the horizontal pass runs on the thread's own four bytes plus one byte
// from each neighbour lane (shuffle; the first / last lane of a warp fetches the neighbouring word itself, at the
// image border the reflected pixel is one of the thread's own bytes), the vertical pass on the three most recent
// filtered rows in registers.  Eight row loads are in flight per thread.  Same taps and the same fused multiply-add
// order as fb_level_regular<3, 1>: bit-identical.
// ----------------------------------------------------------------------------------------------
constexpr int L0_TH = 30;                 // output rows per thread (32 input rows = four groups of eight loads)
constexpr int L0_WARPS = 5;               // warps per CTA, side by side: 640 columns (1920 = 3 x 640)
struct Level0Args {
  const uint8_t* frames; size_t step, frame_stride; int W, H;
  float* I; int pitch; size_t i_frame_stride;
  float c0, c1, c2;
};

__global__ void __launch_bounds__(32 * L0_WARPS) fb_level0_stream(Level0Args a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gx = (blockIdx.x * L0_WARPS + warp) * 128 + 4 * lane;    // first of this thread's four columns
  if (gx - 4 * lane >= a.W) return;                                    // whole warp beyond the row
  const int y0 = blockIdx.y * L0_TH;
  const uint8_t* fb = a.frames + blockIdx.z * a.frame_stride;
  float* ob = a.I + blockIdx.z * a.i_frame_stride + gx;
  const bool live = gx < a.W;                                          // W % 4 == 0: a word is inside or outside
  const int gxc = live ? gx : a.W - 4;                                 // idle lanes re-read the last word
  const bool left_edge = gxc == 0, right_edge = gxc + 4 >= a.W;
  const float c0 = a.c0, c1 = a.c1, c2 = a.c2;
  float4 hA = make_float4(0.f, 0.f, 0.f, 0.f), hB = hA;
  uint32_t wv[8], we[8], wn[8], wen[8];
  auto load_group = [&](int g, uint32_t* dv, uint32_t* de) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint8_t* row = fb + (size_t)reflect101(y0 - 1 + 8 * g + k, a.H) * a.step;
      dv[k] = __ldg((const uint32_t*)(row + gxc));
      de[k] = 0u;
      if (lane == 0 && !left_edge) de[k] = __ldg((const uint32_t*)(row + gxc - 4));
      if (lane == 31 && !right_edge) de[k] = __ldg((const uint32_t*)(row + gxc + 4));
    }
  };
  load_group(0, wv, we);
  constexpr int NG = (L0_TH + 2) / 8;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g + 1 < NG) load_group(g + 1, wn, wen);          // the next eight rows are in flight while these are filtered
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t w = wv[k];
      uint32_t wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
      if (lane == 0) wl = we[k];
      if (lane == 31) wr = we[k];
      if (left_edge) wl = w << 16;                                     // reflect101: column -1 is column 1
      if (right_edge) wr = w >> 16;                                    // column W is column W - 2
      const float pm = (float)(wl >> 24), p0 = (float)(w & 255u), p1 = (float)((w >> 8) & 255u),
                  p2 = (float)((w >> 16) & 255u), p3 = (float)(w >> 24), p4 = (float)(wr & 255u);
      float4 hC;
      hC.x = fmaf(c2, p1, fmaf(c1, p0, c0 * pm));
      hC.y = fmaf(c2, p2, fmaf(c1, p1, c0 * p0));
      hC.z = fmaf(c2, p3, fmaf(c1, p2, c0 * p1));
      hC.w = fmaf(c2, p4, fmaf(c1, p3, c0 * p2));
      const int i = 8 * g + k;                                         // input row y0 - 1 + i; output row y0 + i - 2
      const int yo = y0 + i - 2;
      if (i >= 2 && yo < a.H && live) {
        float4 o;
        o.x = fmaf(c2, hC.x, fmaf(c1, hB.x, c0 * hA.x));
        o.y = fmaf(c2, hC.y, fmaf(c1, hB.y, c0 * hA.y));
        o.z = fmaf(c2, hC.z, fmaf(c1, hB.z, c0 * hA.z));
        o.w = fmaf(c2, hC.w, fmaf(c1, hB.w, c0 * hA.w));
        *(float4*)(ob + (size_t)yo * a.pitch) = o;
      }
      hA = hB; hB = hC;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { wv[k] = wn[k]; we[k] = wen[k]; }
  }
}

// ----------------------------------------------------------------------------------------------
// K3 (coarse levels in one pass): the reference's pyramid (pyr_scale 0.5, three coarse levels) filters the SAME
// full-resolution frame three times with the regular (K, S) forms (4, 2), (10, 4), (20, 8).  One CTA stages the
// source region of an 8 x 8 tile of the coarsest level once -- as floats, one conversion per byte instead of one per
// tap -- and produces that tile together with the 16 x 16 and 32 x 32 tiles of the two finer levels that the same
// 64 x 64 source pixels feed: the frame is read once instead of three times and 65 x 3 launches of tiny tiles become
// one.  Per level the arithmetic is fb_level_regular's (same taps, same fused multiply-add order): bit-identical.
// ----------------------------------------------------------------------------------------------
constexpr int LC_REG = 76;                // source region edge: rows [64 by - 6, 64 by + 70), same for columns
constexpr int LC_LPAD = 2;                // columns of left padding: the staged region starts at 64 bx - 8 (word aligned)
constexpr int LC_RP = 84;                 // region pitch in floats: 16-byte rows, and 84 * 4 B = 80 (mod 128) puts the
                                          // eight rows of a quarter warp on eight different 16-byte bank groups
constexpr int LC_NT = 288;                // 9 warps: three per level in the horizontal pass
struct LevelCoarseArgs {
  const uint8_t* frames; size_t step, frame_stride; int W, H;
  float* I[3];                             // level images, S = 2, 4, 8
  int wk[3], hk[3], pitch[3];
  size_t i_frame_stride[3];
  float c2[4], c4[10], c8[20];             // combined taps
};

// horizontal pass of one region row for outputs [X0, X1) of one level: the row segment is read once, as aligned
// 16-byte loads, into registers; every tap of every output is then a register operand (the first version read one
// shared-memory word per tap and was bound by the number of LDS instructions)
template <int K, int S, int COFF, int P, int X0, int X1>
__device__ __forceinline__ void lc_hrow(const float* __restrict__ row, float* __restrict__ sh_row,
                                        const float* __restrict__ c) {
  constexpr int A0 = (COFF + S * X0) & ~3, A1 = (COFF + S * (X1 - 1) + K + 3) & ~3, NV = (A1 - A0) / 4;
  float v[4 * NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const float4 f = *(const float4*)(row + A0 + 4 * q);
    v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
  }
#pragma unroll
  for (int x = X0; x < X1; ++x) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = fmaf(c[j], v[COFF + S * x + j - A0], acc);
    sh_row[x] = acc;
  }
}

template <int K, int S, int T, int P>
__device__ __forceinline__ void lc_vpass(const float* __restrict__ sh, const float* __restrict__ c, int idx,
                                         float* __restrict__ out, int x0, int y0, int wk, int hk, int pitch) {
  const int y = idx / T, x = idx - y * T;            // lanes run along x
  if (x0 + x >= wk || y0 + y >= hk) return;
  const float* p = sh + (S * y) * P + x;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) v = fmaf(c[j], p[j * P], v);
  out[(size_t)(y0 + y) * pitch + x0 + x] = v;
}

template <int K, int S, int NY, int P>
__device__ __forceinline__ void lc_vwalk(const float* __restrict__ sh, const float* __restrict__ c, int x, int yf,
                                         float* __restrict__ out, int x0, int y0, int wk, int hk, int pitch) {
  if (x0 + x >= wk || y0 + yf >= hk) return;
  constexpr int NR = S * (NY - 1) + K;
  float v[NR];
  const float* p = sh + (S * yf) * P + x;
#pragma unroll
  for (int r = 0; r < NR; ++r) v[r] = p[r * P];
  float* o = out + (size_t)(y0 + yf) * pitch + x0 + x;
#pragma unroll
  for (int y = 0; y < NY; ++y) {
    if (y0 + yf + y >= hk) break;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = fmaf(c[j], v[S * y + j], acc);
    o[(size_t)y * pitch] = acc;
  }
}

__global__ void __launch_bounds__(LC_NT) fb_levels_coarse(LevelCoarseArgs a) {
  __shared__ __align__(16) float s_src[LC_REG * LC_RP];   // source region as floats
  __shared__ float s_h1[66 * 33], s_h2[70 * 17], s_h3[76 * 9];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int gx0 = 64 * blockIdx.x - 8, gy0 = 64 * blockIdx.y - 6;
  constexpr int NW = (LC_LPAD + LC_REG + 3) / 4;      // 20 words per region row
  const uint8_t* fb = a.frames + blockIdx.z * a.frame_stride;
  // aligned 4-byte loads, all of them in flight before the first conversion.  Rows are reflected per word (in range:
  // one compare); a word with a column outside the frame (the first two / last few words of a region row in the
  // CTAs of the first / last column of the grid) is assembled from four reflected byte loads.  (The first version
  // sent every CTA that touches the frame border -- 18 % of them at 1080p -- down a byte-per-thread path with a full
  // reflect per byte, which cost more than all the interior CTAs' staging together.)
  {
    const bool aligned = ((a.step | (size_t)fb) & 3) == 0;
    constexpr int PER = (LC_REG * NW + LC_NT - 1) / LC_NT;
    uint32_t wv[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = t + LC_NT * k;
      const int rr = i / NW, wc = i - rr * NW;
      wv[k] = 0u;
      if (i < LC_REG * NW) {
        const uint8_t* row = fb + (size_t)reflect101(gy0 + rr, a.H) * a.step;
        const int gx = gx0 + 4 * wc;
        if (aligned && gx >= 0 && gx + 4 <= a.W) {
          wv[k] = __ldg((const uint32_t*)(row + gx));
        } else {
          wv[k] = (uint32_t)row[reflect101(gx, a.W)] | ((uint32_t)row[reflect101(gx + 1, a.W)] << 8) |
                  ((uint32_t)row[reflect101(gx + 2, a.W)] << 16) | ((uint32_t)row[reflect101(gx + 3, a.W)] << 24);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = t + LC_NT * k;
      const int rr = i / NW, wc = i - rr * NW;
      if (i < LC_REG * NW)
        *(float4*)(s_src + rr * LC_RP + 4 * wc) = make_float4((float)(wv[k] & 255u), (float)((wv[k] >> 8) & 255u),
                                                              (float)((wv[k] >> 16) & 255u), (float)(wv[k] >> 24));
    }
  }
  __syncthreads();
  // horizontal passes, thread = (level, region row): warps 0-2 the S = 8 level (76 rows), 3-5 S = 4 (70 rows from
  // region row 3), 6-8 S = 2 (66 rows from region row 5); column offsets = c0 - g0 + LPAD = 2, 5, 7
  {
    const int lv = warp / 3, r = t - lv * 96;
    if (lv == 0) {
      if (r < 76) {
        lc_hrow<20, 8, 2, 9, 0, 4>(s_src + r * LC_RP, s_h3 + r * 9, a.c8);
        lc_hrow<20, 8, 2, 9, 4, 8>(s_src + r * LC_RP, s_h3 + r * 9, a.c8);
      }
    } else if (lv == 1) {
      if (r < 70) {
        lc_hrow<10, 4, 5, 17, 0, 8>(s_src + (r + 3) * LC_RP, s_h2 + r * 17, a.c4);
        lc_hrow<10, 4, 5, 17, 8, 16>(s_src + (r + 3) * LC_RP, s_h2 + r * 17, a.c4);
      }
    } else if (r < 66) {
      lc_hrow<4, 2, 7, 33, 0, 16>(s_src + (r + 5) * LC_RP, s_h1 + r * 33, a.c2);
      lc_hrow<4, 2, 7, 33, 16, 32>(s_src + (r + 5) * LC_RP, s_h1 + r * 33, a.c2);
    }
  }
  __syncthreads();
  // vertical passes: a thread owns one output column of one level and a run of NY consecutive rows; the filtered
  // rows the run needs are read once into registers (consecutive outputs share K - S of their K rows) and the output
  // index is the thread's own -- the first version looped over a flat output index with a level switch, a division
  // and K shared-memory reads per output, a third of the kernel's instructions.  Warps 0-3: the 32 x 32 tile of the
  // S = 2 level (8 rows each), warps 4-5: the 16 x 16 tile of S = 4 (4 rows per half warp), warp 6: the 8 x 8 tile
  // of S = 8 (2 rows per quarter warp).
  if (warp < 4)
    lc_vwalk<4, 2, 8, 33>(s_h1, a.c2, lane, 8 * warp, a.I[0] + blockIdx.z * a.i_frame_stride[0], 32 * blockIdx.x,
                          32 * blockIdx.y, a.wk[0], a.hk[0], a.pitch[0]);
  else if (warp < 6)
    lc_vwalk<10, 4, 4, 17>(s_h2, a.c4, lane & 15, 4 * (2 * (warp - 4) + (lane >> 4)),
                           a.I[1] + blockIdx.z * a.i_frame_stride[1], 16 * blockIdx.x, 16 * blockIdx.y, a.wk[1], a.hk[1],
                           a.pitch[1]);
  else if (warp == 6)
    lc_vwalk<20, 8, 2, 9>(s_h3, a.c8, lane & 7, 2 * (lane >> 3), a.I[2] + blockIdx.z * a.i_frame_stride[2],
                          8 * blockIdx.x, 8 * blockIdx.y, a.wk[2], a.hk[2], a.pitch[2]);
}

// ----------------------------------------------------------------------------------------------
// K4: polynomial expansion. 64x32 output tile per CTA; the (32+2n)x(64+2n) input tile and the three
// vertically filtered rows live in shared memory; REPLICATE borders.
// Output layout per frame and level: Ra = float4 plane (ch0..3) followed by Rb = float plane (ch4), so the
// bilinear gather of the warped frame in K5 is 2 loads per corner instead of 5.
// ----------------------------------------------------------------------------------------------
constexpr int PE_TW = 64, PE_TH = 32;

__global__ void __launch_bounds__(256) fb_polyexp(const float* __restrict__ I, int w, int h, int pitch,
                                                   size_t i_frame_stride, float* __restrict__ R,
                                                   size_t plane_stride, size_t r_frame_stride, PolyConst pc, int n) {
  extern __shared__ float smem[];
  const int sw = PE_TW + 2 * n;  // tile width with halo
  const int sh = PE_TH + 2 * n;
  float* s_in = smem;                    // [sh][sw]
  float* s_r0 = s_in + sh * sw;          // [PE_TH][sw]
  float* s_r1 = s_r0 + PE_TH * sw;
  float* s_r2 = s_r1 + PE_TH * sw;
  const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
  const float* ib = I + blockIdx.z * i_frame_stride;
  const int t = threadIdx.x;
  for (int i = t; i < sh * sw; i += 256) {
    int r = i / sw, c = i % sw;
    int gy = clampi(y0 + r - n, 0, h - 1), gx = clampi(x0 + c - n, 0, w - 1);
    s_in[i] = ib[(size_t)gy * pitch + gx];
  }
  __syncthreads();
  // vertical pass
  for (int i = t; i < PE_TH * sw; i += 256) {
    int r = i / sw, c = i % sw;
    const float* col = s_in + (r + n) * sw + c;
    float c0 = col[0];
    float r0 = c0 * pc.g[0], r1 = 0.f, r2 = 0.f;
    for (int k = 1; k <= n; ++k) {
      float a = col[-k * sw], b = col[k * sw];
      float sum = a + b;
      r0 = fmaf(pc.g[k], sum, r0);
      r1 = fmaf(pc.xg[k], b - a, r1);
      r2 = fmaf(pc.xxg[k], sum, r2);
    }
    s_r0[i] = r0; s_r1[i] = r1; s_r2[i] = r2;
  }
  __syncthreads();
  // horizontal pass: 64x32 outputs, 8 per thread (same column, rows ty, ty+4, ...)
  const int tx = t & 63, ty = t >> 6;
  float* rb = R + blockIdx.z * r_frame_stride;
  float4* ra4 = (float4*)rb;
  float* rb1 = rb + 4 * plane_stride;
  for (int r = ty; r < PE_TH; r += 4) {
    int gx = x0 + tx, gy = y0 + r;
    if (gx >= w || gy >= h) continue;
    const float* p0 = s_r0 + r * sw + tx + n;
    const float* p1 = s_r1 + r * sw + tx + n;
    const float* p2 = s_r2 + r * sw + tx + n;
    float b1 = p0[0] * pc.g[0], b3 = p1[0] * pc.g[0], b5 = p2[0] * pc.g[0];
    float b2 = 0.f, b4 = 0.f, b6 = 0.f;
    for (int k = 1; k <= n; ++k) {
      float tg = p0[k] + p0[-k];
      b1 = fmaf(tg, pc.g[k], b1);
      b4 = fmaf(tg, pc.xxg[k], b4);
      b2 = fmaf(p0[k] - p0[-k], pc.xg[k], b2);
      b3 = fmaf(p1[k] + p1[-k], pc.g[k], b3);
      b6 = fmaf(p1[k] - p1[-k], pc.xg[k], b6);
      b5 = fmaf(p2[k] + p2[-k], pc.g[k], b5);
    }
    size_t o = (size_t)gy * pitch + gx;
    ra4[o] = make_float4(b3 * pc.ig11, b2 * pc.ig11, b1 * pc.ig03 + b5 * pc.ig33, b1 * pc.ig03 + b4 * pc.ig33);
    rb1[o] = b6 * pc.ig55;
  }
}

// ----------------------------------------------------------------------------------------------
// K4 (specialised, poly_n known at compile time): 64x32 output tile, 320 threads.
//   phase 1: thread = (halo column, 8-row segment); 8+2N rows go global -> registers (all loads in flight), the
//            three vertical filters come out of a register window, results to shared memory
//   phase 2: thread = two adjacent x (8-byte LDS of the filtered rows), warps over rows; 32 B + 8 B stores per thread
// ----------------------------------------------------------------------------------------------
// FUSE0: the level is the full-resolution one (the reference's level 0: 3-tap blur, no resampling) and its image is
// computed here, from the uint8 frame, instead of being read back from HBM: the CTA stages the source pixels of its
// halo tile, runs fb_level_regular<3, 1>'s two passes on them in shared memory (same taps, same fused multiply-add
// order: same bits) and phase 1 takes its 18 rows from there.  Saves the level-image kernel and an 8 B/px round trip.
struct Poly0Src {
  const uint8_t* frames; size_t step, frame_stride;
  float c[3];                               // combined taps of the regular (3, 1) level form
};

template <int N, bool FUSE0>
__global__ void __launch_bounds__(320) fb_polyexp_n(const float* __restrict__ I, int w, int h, int pitch,
                                                     size_t i_frame_stride, float* __restrict__ R,
                                                     size_t plane_stride, size_t r_frame_stride, PolyConst pc,
                                                     Poly0Src src) {
  constexpr int CW = PE_TW + 2 * N;        // halo columns (74 for N = 5)
  constexpr int CP = (CW + 7) / 8 * 8;     // phase-1 thread columns (80), also the smem pitch (even)
  constexpr int SEG = 8, NSEG = PE_TH / SEG;
  static_assert(CP * NSEG <= 320, "phase 1 does not fit the block");
  __shared__ __align__(16) float s_r[3][PE_TH][CP];
  constexpr int IH = PE_TH + 2 * N;        // rows of the level-image halo tile (42)
  __shared__ float s_i[FUSE0 ? IH * CP : 1];
  const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
  const float* ib = I + blockIdx.z * i_frame_stride;
  const int t = threadIdx.x;
  if constexpr (FUSE0) {
    // source tile: image rows y0 - N - 1 .. y0 + PE_TH + N, columns from the word boundary below x0 - N - 1; both
    // scratch arrays live in s_r, which phase 1 only writes after the barrier that ends this block
    constexpr int SH = IH + 2, LP = 2;     // x0 - N - 1 = x0 - 6: two columns above a multiple of four (x0 % 64 == 0)
    constexpr int SWP = ((LP + CW + 2 + 3) / 4) * 4;      // 80 staged columns
    static_assert(N == 5, "the staging offsets are written for poly_n = 5");
    uint8_t* s_u = (uint8_t*)&s_r[0][0][0];               // [SH][SWP] uint8
    float* s_h = (float*)(s_u + ((SH * SWP + 15) & ~15)); // [SH][CP] horizontal pass
    static_assert(((SH * SWP + 15) & ~15) + SH * CP * 4 <= sizeof(s_r), "scratch does not fit s_r");
    const uint8_t* fb = src.frames + blockIdx.z * src.frame_stride;
    const int gx0 = x0 - N - 1 - LP, gy0 = y0 - N - 1;
    const bool inside = gx0 >= 0 && gy0 >= 0 && gx0 + SWP <= w && gy0 + SH <= h && ((src.step | (size_t)fb) & 3) == 0;
    if (inside) {
      for (int i = t; i < SH * (SWP / 4); i += 320) {
        const int rr = i / (SWP / 4), wc = i - rr * (SWP / 4);
        ((uint32_t*)s_u)[rr * (SWP / 4) + wc] = __ldg((const uint32_t*)(fb + (size_t)(gy0 + rr) * src.step + gx0) + wc);
      }
    } else {
      for (int i = t; i < SH * SWP; i += 320) {
        const int rr = i / SWP, cc = i - rr * SWP;
        s_u[i] = fb[(size_t)reflect101(gy0 + rr, h) * src.step + reflect101(gx0 + cc, w)];
      }
    }
    __syncthreads();
    // horizontal pass at the CLAMPED halo column (the expansion replicates the level image at its border, the level
    // image itself reflects the frame: the staged tile holds reflect101 of every coordinate it covers)
    for (int i = t; i < SH * CW; i += 320) {
      const int rr = i / CW, c = i - rr * CW;
      const int tc = clampi(x0 + c - N, 0, w - 1) - gx0;  // staged column of the halo column's own pixel
      const uint8_t* pr = s_u + rr * SWP + tc - 1;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) v = fmaf(src.c[j], (float)pr[j], v);
      s_h[rr * CP + c] = v;
    }
    __syncthreads();
    for (int i = t; i < IH * CW; i += 320) {
      const int r = i / CW, c = i - r * CW;
      const int tr = clampi(y0 + r - N, 0, h - 1) - gy0;
      const float* ph = s_h + (tr - 1) * CP + c;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) v = fmaf(src.c[j], ph[j * CP], v);
      s_i[r * CP + c] = v;
    }
    __syncthreads();
  }
  {
    const int seg = t / CP, c = t - seg * CP;
    if (seg < NSEG && c < CW) {
      const int gx = clampi(x0 + c - N, 0, w - 1);
      const int yb = y0 + seg * SEG - N;
      float v[SEG + 2 * N];
#pragma unroll
      for (int j = 0; j < SEG + 2 * N; ++j)
        v[j] = FUSE0 ? s_i[(seg * SEG + j) * CP + c] : ib[(size_t)clampi(yb + j, 0, h - 1) * pitch + gx];
#pragma unroll
      for (int j = 0; j < SEG; ++j) {
        float r0 = v[j + N] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
          float a = v[j + N - k], b = v[j + N + k];
          float sum = a + b;
          r0 = fmaf(pc.g[k], sum, r0);
          r1 = fmaf(pc.xg[k], b - a, r1);
          r2 = fmaf(pc.xxg[k], sum, r2);
        }
        const int r = seg * SEG + j;
        s_r[0][r][c] = r0; s_r[1][r][c] = r1; s_r[2][r][c] = r2;
      }
    }
  }
  __syncthreads();
  {
    const int warp = t >> 5, lane = t & 31;
    const int xl = 2 * lane;                 // local x of the first of two outputs
    const int gx = x0 + xl;
    float* rb = R + blockIdx.z * r_frame_stride;
    float4* ra4 = (float4*)rb;
    float* rb1 = rb + 4 * plane_stride;
    for (int r = warp; r < PE_TH; r += 10) {
      const int gy = y0 + r;
      if (gy >= h || gx >= w) continue;
      float p0[2 * N + 2], p1[2 * N + 2], p2[2 * N + 2];
#pragma unroll
      for (int j = 0; j < N + 1; ++j) {
        float2 a = *(const float2*)&s_r[0][r][xl + 2 * j];
        float2 b = *(const float2*)&s_r[1][r][xl + 2 * j];
        float2 c = *(const float2*)&s_r[2][r][xl + 2 * j];
        p0[2 * j] = a.x; p0[2 * j + 1] = a.y;
        p1[2 * j] = b.x; p1[2 * j + 1] = b.y;
        p2[2 * j] = c.x; p2[2 * j + 1] = c.y;
      }
      float4 oa[2];
      float ob[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ctr = e + N;
        float b1 = p0[ctr] * pc.g[0], b3 = p1[ctr] * pc.g[0], b5 = p2[ctr] * pc.g[0];
        float b2 = 0.f, b4 = 0.f, b6 = 0.f;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
          float tg = p0[ctr + k] + p0[ctr - k];
          b1 = fmaf(tg, pc.g[k], b1);
          b4 = fmaf(tg, pc.xxg[k], b4);
          b2 = fmaf(p0[ctr + k] - p0[ctr - k], pc.xg[k], b2);
          b3 = fmaf(p1[ctr + k] + p1[ctr - k], pc.g[k], b3);
          b6 = fmaf(p1[ctr + k] - p1[ctr - k], pc.xg[k], b6);
          b5 = fmaf(p2[ctr + k] + p2[ctr - k], pc.g[k], b5);
        }
        oa[e] = make_float4(b3 * pc.ig11, b2 * pc.ig11, b1 * pc.ig03 + b5 * pc.ig33, b1 * pc.ig03 + b4 * pc.ig33);
        ob[e] = b6 * pc.ig55;
      }
      size_t o = (size_t)gy * pitch + gx;
      ra4[o] = oa[0];
      if (gx + 1 < w) {
        ra4[o + 1] = oa[1];
        *(float2*)(rb1 + o) = make_float2(ob[0], ob[1]);   // o is even: pitch % 32 == 0, gx even
      } else {
        rb1[o] = ob[0];
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// K4 (packed form, the one that runs for poly_n = 5 / 7): fb_polyexp_n is issue-bound (77 % of the issue slots, more
// than half of them FFMA / FADD), so this form does the same arithmetic two rows at a time in packed f32x2
// instructions (FFMA2 / FADD2 / FMUL2: one issue slot, two IEEE operations -- same operations, same order, same bits).
// The pairing is across ROWS r and r + 16 of the 64 x 32 tile, because a pair of rows slides through both filters
// together (pairing adjacent columns or adjacent rows would need the window at two alignments):
//   phase 1: thread = (halo column, 8-row segment of the upper AND of the lower half tile); the 2 x 18 input values
//            are loaded as the two halves of 18 register pairs, the three vertical filters run packed, results go
//            to shared memory as (row r, row r + 16) pairs: layout [channel][r & 15][column][r >> 4]
//   phase 2: thread = two adjacent columns x (row r, row r + 16): six 16-byte LDS per channel bring the 12 window
//            positions as ready-made register pairs; horizontal filters packed; the final scaling and the stores are
//            scalar (a float4 of one row takes one half of four different pairs).
// ----------------------------------------------------------------------------------------------
#ifndef FB_POLY_MINB
#define FB_POLY_MINB 4     // CTAs per SM the register allocation aims at (64 registers)
#endif
__device__ __forceinline__ float2 f2_bc(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 f2_neg(float2 v) { return make_float2(-v.x, -v.y); }

template <int N>
__global__ void __launch_bounds__(256, FB_POLY_MINB) fb_polyexp_p(const float* __restrict__ I, int w, int h, int pitch,
                                                     size_t i_frame_stride, float* __restrict__ R,
                                                     size_t plane_stride, size_t r_frame_stride, PolyConst pc) {
  constexpr int CW = PE_TW + 2 * N;        // halo columns; column c is image column x0 - N + c
  constexpr int HR = PE_TH / 2;            // rows per half tile: row r pairs with row r + HR
  constexpr int SEG = 8, NSEG = HR / SEG;
  static_assert(PE_TH == 32 && CW * NSEG <= 256 && (CW % 2) == 0, "tile geometry");
  __shared__ __align__(16) float2 s_r[3][HR][CW];
  const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
  const float* ib = I + blockIdx.z * i_frame_stride;
  const int t = threadIdx.x;
  if (t < CW * NSEG) {
    const int seg = t / CW, c = t - seg * CW;
    const float* col = ib + clampi(x0 + c - N, 0, w - 1);
    const int yb = y0 + seg * SEG - N;
    float2 v[SEG + 2 * N];
    if (y0 >= N && y0 + PE_TH + N <= h) {    // no row of the halo tile is clamped (uniform over the CTA)
      const float* plo = col + (size_t)yb * pitch;
      const float* phi = plo + (size_t)HR * pitch;
#pragma unroll
      for (int j = 0; j < SEG + 2 * N; ++j) {
        v[j].x = plo[j * pitch];
        v[j].y = phi[j * pitch];
      }
    } else {
#pragma unroll
      for (int j = 0; j < SEG + 2 * N; ++j) {
        v[j].x = col[(size_t)clampi(yb + j, 0, h - 1) * pitch];
        v[j].y = col[(size_t)clampi(yb + j + HR, 0, h - 1) * pitch];
      }
    }
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      float2 r0 = __fmul2_rn(v[j + N], f2_bc(pc.g[0]));
      float2 r1 = make_float2(0.f, 0.f), r2 = r1;
#pragma unroll
      for (int k = 1; k <= N; ++k) {
        const float2 a = v[j + N - k], b = v[j + N + k];
        const float2 sum = __fadd2_rn(a, b);
        r0 = __ffma2_rn(f2_bc(pc.g[k]), sum, r0);
        r1 = __ffma2_rn(f2_bc(pc.xg[k]), __fadd2_rn(b, f2_neg(a)), r1);
        r2 = __ffma2_rn(f2_bc(pc.xxg[k]), sum, r2);
      }
      const int r = seg * SEG + j;
      s_r[0][r][c] = r0; s_r[1][r][c] = r1; s_r[2][r][c] = r2;
    }
  }
  __syncthreads();
  {
    const int warp = t >> 5, lane = t & 31;
    const int xl = 2 * lane;                 // local x of the first of two outputs
    const int gx = x0 + xl;
    float* rb = R + blockIdx.z * r_frame_stride;
    float4* ra4 = (float4*)rb;
    float* rb1 = rb + 4 * plane_stride;
    if (gx < w) {
#pragma unroll 1
      for (int r = warp; r < HR; r += 8) {
        if (y0 + r >= h) break;
        float2 b1[2], b2[2], b3[2], b4[2], b5[2], b6[2];
        {
          float2 p0[2 * N + 2];
#pragma unroll
          for (int q = 0; q < N + 1; ++q) {
            const float4 f = *(const float4*)&s_r[0][r][xl + 2 * q];
            p0[2 * q] = make_float2(f.x, f.y); p0[2 * q + 1] = make_float2(f.z, f.w);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ctr = e + N;
            b1[e] = __fmul2_rn(p0[ctr], f2_bc(pc.g[0]));
            b2[e] = make_float2(0.f, 0.f); b4[e] = b2[e];
#pragma unroll
            for (int k = 1; k <= N; ++k) {
              const float2 tg = __fadd2_rn(p0[ctr + k], p0[ctr - k]);
              b1[e] = __ffma2_rn(tg, f2_bc(pc.g[k]), b1[e]);
              b4[e] = __ffma2_rn(tg, f2_bc(pc.xxg[k]), b4[e]);
              b2[e] = __ffma2_rn(__fadd2_rn(p0[ctr + k], f2_neg(p0[ctr - k])), f2_bc(pc.xg[k]), b2[e]);
            }
          }
        }
        {
          float2 p1[2 * N + 2];
#pragma unroll
          for (int q = 0; q < N + 1; ++q) {
            const float4 f = *(const float4*)&s_r[1][r][xl + 2 * q];
            p1[2 * q] = make_float2(f.x, f.y); p1[2 * q + 1] = make_float2(f.z, f.w);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ctr = e + N;
            b3[e] = __fmul2_rn(p1[ctr], f2_bc(pc.g[0]));
            b6[e] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 1; k <= N; ++k) {
              b3[e] = __ffma2_rn(__fadd2_rn(p1[ctr + k], p1[ctr - k]), f2_bc(pc.g[k]), b3[e]);
              b6[e] = __ffma2_rn(__fadd2_rn(p1[ctr + k], f2_neg(p1[ctr - k])), f2_bc(pc.xg[k]), b6[e]);
            }
          }
        }
        {
          float2 p2[2 * N + 2];
#pragma unroll
          for (int q = 0; q < N + 1; ++q) {
            const float4 f = *(const float4*)&s_r[2][r][xl + 2 * q];
            p2[2 * q] = make_float2(f.x, f.y); p2[2 * q + 1] = make_float2(f.z, f.w);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ctr = e + N;
            b5[e] = __fmul2_rn(p2[ctr], f2_bc(pc.g[0]));
#pragma unroll
            for (int k = 1; k <= N; ++k)
              b5[e] = __ffma2_rn(__fadd2_rn(p2[ctr + k], p2[ctr - k]), f2_bc(pc.g[k]), b5[e]);
          }
        }
        // final scaling and stores, one row of the pair at a time (scalar: same expressions as fb_polyexp_n)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int gy = y0 + r + hh * HR;
          if (gy >= h) break;
          float4 oa[2];
          float ob[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float c1 = hh ? b1[e].y : b1[e].x, c2 = hh ? b2[e].y : b2[e].x, c3 = hh ? b3[e].y : b3[e].x;
            const float c4 = hh ? b4[e].y : b4[e].x, c5 = hh ? b5[e].y : b5[e].x, c6 = hh ? b6[e].y : b6[e].x;
            oa[e] = make_float4(c3 * pc.ig11, c2 * pc.ig11, c1 * pc.ig03 + c5 * pc.ig33, c1 * pc.ig03 + c4 * pc.ig33);
            ob[e] = c6 * pc.ig55;
          }
          const size_t o = (size_t)gy * pitch + gx;
          if (gx + 1 < w) {
            // both records in ONE 32-byte store (o is even: pitch % 32 == 0, gx even): two 16-byte stores per lane
            // leave every 32-byte sector half written by each instruction and nearly double the L1 -> L2 write traffic
            // (ncu: 151.6 M sectors for 84 M sectors of output)
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ra4 + o), "f"(oa[0].x), "f"(oa[0].y),
                         "f"(oa[0].z), "f"(oa[0].w), "f"(oa[1].x), "f"(oa[1].y), "f"(oa[1].z), "f"(oa[1].w)
                         : "memory");
            *(float2*)(rb1 + o) = make_float2(ob[0], ob[1]);
          } else {
            ra4[o] = oa[0];
            rb1[o] = ob[0];
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// K5/K6: one fused kernel per (level, iteration):
//   flow_in (zero | previous iteration | bilinear x(1/pyr_scale) upsample of the coarser level)
//   -> UpdateMatrices on a (T+2m)^2 halo tile (bilinear warp of R1, border attenuation)   [smem M: float4 + float]
//   -> separable (2m+1)^2 window sum, BORDER_REPLICATE (box, or Gaussian taps with flag 256):
//        horizontal pass IN PLACE (one thread owns a whole row: output x only needs inputs >= x),
//        vertical pass straight into the solve
//   -> per-pixel 2x2 solve -> flow_out
// Packed f32x2 adds (FADD2) carry 4 of the 5 channels two at a time.
// This is the GENERAL form (any window radius at run time, box or Gaussian window: template <NT, 0, 0, GAUSS, MODE>;
// the compile-time tile / radius parameters CT / CM are kept for experiments).  The reference's own window
// (winsize 15, box) runs on the warp-specialised strip kernel fb_iter_ws (fb_ws.cuh).
// ----------------------------------------------------------------------------------------------
constexpr int IT_THREADS = 512;

struct IterArgs {
  const float* R;          // level base, [frame]{float4 plane ch0..3, float plane ch4}
  size_t r_frame_stride, plane_stride;
  int pitch, w, h;
  int pair_frame_step;     // 1: sequence (pair p = frames p, p+1); 2: independent pairs (2p, 2p+1)
  int mode;                // 0 zero flow, 1 flow_in at this level, 2 upsample coarse flow
  const float2* flow_in;   // mode 1: [pair][h][pitch]; mode 2: coarse [pair][ch][cpitch]
  size_t flow_in_pair_stride;
  int in_pitch;            // in float2
  int cw, ch;
  const int *ux0, *ux1, *uy0, *uy1;
  const float *ufx, *ufy;
  float up_mult;
  int up_exact2;           // mode 2: this level is exactly twice the coarser one in both axes
  float2* flow_out;
  size_t flow_out_pair_stride;
  int out_pitch;           // in float2
  int m;                   // winsize / 2
  int tile;                // output tile edge (generic kernel)
  int nb;                  // vertically adjacent tiles walked by one CTA
  float inv_area;          // 1 / winsize^2 (box)
  const float* gtaps;      // Gaussian window taps or nullptr
  unsigned long long* stats_acc;  // fb_iter_ws, last iteration only: per-pair fixed-point flow statistics, or nullptr
};

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 sub4(float4 a, float4 b) {  // a - b (exact: fma with -1)
  const float2 m1 = make_float2(-1.f, -1.f);
  float2 lo = __ffma2_rn(make_float2(b.x, b.y), m1, make_float2(a.x, a.y));
  float2 hi = __ffma2_rn(make_float2(b.z, b.w), m1, make_float2(a.z, a.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 fma4s(float4 v, float s, float4 acc) {  // acc + v*s
  const float2 ss = make_float2(s, s);
  float2 lo = __ffma2_rn(make_float2(v.x, v.y), ss, make_float2(acc.x, acc.y));
  float2 hi = __ffma2_rn(make_float2(v.z, v.w), ss, make_float2(acc.z, acc.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float border_w(int i, int n) {
  // product of the 5-px attenuation table from both ends (UpdateMatrices)
  const float tab[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
  float s = 1.f;
  if (i < 5) s *= tab[i];
  if (i >= n - 5) s *= tab[n - 1 - i];
  return s;
}

__device__ __forceinline__ float2 solve2x2(float g11, float g12, float g22, float h1, float h2) {
  float idet = 1.f / (g11 * g22 - g12 * g12 + 1e-3f);
  return make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct FbCorner { float4 a0, a1; float e0, e1; };   // two horizontally adjacent R1 records (ch0..3, ch4)


__device__ __forceinline__ void pin(int& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void pin(unsigned& v) { asm volatile("" : "+r"(v)); }
template <typename T>
__device__ __forceinline__ void pin(T*& v) { asm volatile("" : "+l"(v)); }

// explicit state-space loads / stores: pinned pointers lose their address space, and 32-bit shared addresses
// save the generic -> shared conversion per access
template <int OFF>
__device__ __forceinline__ float4 ldg_f4(const float4* p) {
  float4 v;
  asm("ld.global.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ float ldg_f1(const float* p) {
  float v;
  asm("ld.global.f32 %0, [%1+%2];" : "=f"(v) : "l"(p), "n"(OFF));
  return v;
}
__device__ __forceinline__ float2 ldg_f2(const float2* p) {
  float2 v;
  asm("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
// streamed operands of fb_iter_ws (R0 record, flow vector: each byte used once by the CTA) and its gathered ones (R1
// corners: re-used by the next row / the next run), with cache hints behind FBW_STREAM_HINT / FBW_GATHER_HINT
#ifndef FBW_STREAM_HINT
#define FBW_STREAM_HINT 0
#endif
#ifndef FBW_GATHER_HINT
#define FBW_GATHER_HINT 0
#endif
#if FBW_STREAM_HINT == 1
#define FBW_LDS_Q ".L1::no_allocate"
#elif FBW_STREAM_HINT == 2
#define FBW_LDS_Q ".L1::evict_first"
#else
#define FBW_LDS_Q ""
#endif
#if FBW_GATHER_HINT == 1
#define FBW_LDG_Q ".L1::evict_last"
#else
#define FBW_LDG_Q ""
#endif
template <int OFF>
__device__ __forceinline__ float4 lds_f4(const float4* p) {
  float4 v;
  asm("ld.global" FBW_LDS_Q ".v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ float lds_f1(const float* p) {
  float v;
  asm("ld.global" FBW_LDS_Q ".f32 %0, [%1+%2];" : "=f"(v) : "l"(p), "n"(OFF));
  return v;
}
__device__ __forceinline__ float2 lds_f2(const float2* p) {
  float2 v;
  asm("ld.global" FBW_LDS_Q ".v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
template <int OFF>
__device__ __forceinline__ float4 ldgat_f4(const float4* p) {
  float4 v;
  asm("ld.global" FBW_LDG_Q ".v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ float ldgat_f1(const float* p) {
  float v;
  asm("ld.global" FBW_LDG_Q ".f32 %0, [%1+%2];" : "=f"(v) : "l"(p), "n"(OFF));
  return v;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_f4(unsigned addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts_f1(unsigned addr, float x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}

// flow vector of pixel (x, y) of this level: o = y * pitch + x (MODE 1: the flow buffers share the level's pitch)
// a * wa + b * wb with both products rounded (no FMA contraction): the arithmetic of cv2's linear resize
__device__ __forceinline__ float lerp_nc(float a, float wa, float b, float wb) {
  return __fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb));
}

template <int MODE>
__device__ __forceinline__ float2 fetch_flow_m(const IterArgs& a, const float2* fin, int o, int y, int xa, int xb,
                                               float ufx) {
  if (MODE == 0) return make_float2(0.f, 0.f);
  if (MODE == 1) return lds_f2(fin + o);
  int ya, yb;
  float fy;
  if (a.up_exact2) {
    // exact x2 upsample (the reference's pyr_scale 0.5 on even sizes): cv2's linear-resize row table in closed form
    // (source position y / 2 - 0.25, clamped at both ends) instead of three table loads per row
    if (y & 1) { ya = y >> 1; yb = min(ya + 1, a.ch - 1); fy = ya >= a.ch - 1 ? 0.f : 0.25f; }
    else if (y == 0) { ya = 0; yb = min(1, a.ch - 1); fy = 0.f; }
    else { ya = (y >> 1) - 1; yb = ya + 1; fy = 0.75f; }
  } else {
    ya = a.uy0[y]; yb = a.uy1[y];
    fy = a.ufy[y];
  }
  float2 p00 = fin[ya * a.in_pitch + xa], p01 = fin[ya * a.in_pitch + xb];
  float2 p10 = fin[yb * a.in_pitch + xa], p11 = fin[yb * a.in_pitch + xb];
  // lerp_nc: cv2's resize rounds both products before the add; a contracted FMA (either way round) is other bits
  const float gx = 1.f - ufx, gy = 1.f - fy;
  float tx0 = lerp_nc(p00.x, gx, p01.x, ufx), ty0 = lerp_nc(p00.y, gx, p01.y, ufx);
  float tx1 = lerp_nc(p10.x, gx, p11.x, ufx), ty1 = lerp_nc(p10.y, gx, p11.y, ufx);
  return make_float2(__fmul_rn(lerp_nc(tx0, gy, tx1, fy), a.up_mult), __fmul_rn(lerp_nc(ty0, gy, ty1, fy), a.up_mult));
}

template <int NT, int CT, int CM, bool GAUSS, int MODE>
__global__ void __launch_bounds__(NT, CT ? 2 : 1) fb_iter(IterArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int T = CT ? CT : a.tile;
  const int m = CT ? CM : a.m;
  const int E = T + 2 * m;          // halo tile edge
  const int ES = E | 1;             // odd row stride (in elements): conflict-free row walks
  const int HL = T >> 1;            // outputs [0,HL) are summed left->right, [HL,T) right->left (see step B)
  float4* sM4 = (float4*)smem;      // [E][ES]  (M0..M3)   rows form a ring (see `off`)
  float* sM1 = smem + 4 * E * ES;   // [E][ES]  (M4)
  // launch order: pair index fastest, so the CTAs of pairs p and p + 1 on the same tile run together and the
  // expansion of frame p + 1 (R1 of pair p, R0 of pair p + 1) comes from HBM once and from L2 the second time
  const int pair = blockIdx.x;
  const int bx = blockIdx.y, by = blockIdx.z;
  const int x0 = bx * T;
  const int w = a.w, h = a.h, pitch = a.pitch;
  const float* base0 = a.R + (size_t)pair * a.pair_frame_step * a.r_frame_stride;
  const float4* __restrict__ R0a = (const float4*)base0;
  const float* __restrict__ R0b = base0 + 4 * a.plane_stride;
  const float4* __restrict__ R1a = (const float4*)(base0 + a.r_frame_stride);
  const float* __restrict__ R1b = base0 + a.r_frame_stride + 4 * a.plane_stride;
  const float2* __restrict__ fin = MODE ? a.flow_in + (size_t)pair * a.flow_in_pair_stride : nullptr;
  float2* __restrict__ fo = a.flow_out + (size_t)pair * a.flow_out_pair_stride;
  const int t = threadIdx.x;

  // per-thread constants of step A: one halo column per thread.  The column -> lane map is padded on the left so
  // that lane 0 of every warp sits on an 8-pixel boundary of the image (x0 is a multiple of 8 for T % 8 == 0):
  // the streamed float4 / float2 / float row loads then fall on whole 128-byte lines (4 wavefronts instead of ~7.5).
  const int PADL = (8 - (m & 7)) & 7;         // 1 for m = 7
  const int EW = (E + PADL + 7) & ~7;         // padded columns per row run (72 for E = 70)
  const int RP = NT / EW;                     // row runs per tile (7)
  const int ty0 = t / EW, ix = t - ty0 * EW - PADL;
  const int ty = (ix >= 0 && ix < E) ? ty0 : RP;   // padding lanes sit step A out
  const int x = clampi(x0 - m + ix, 0, w - 1);
  const float xf = (float)x;
  // cv2's own test, unsigned wrap-around included (it misfires when a dimension is below 10 px; kept for parity)
  const bool xb_border = (unsigned)(x - 5) >= (unsigned)(w - 10);
  const float bwx = border_w(x, w);
  int uxa = 0, uxb = 0;
  float ufx = 0.f;
  if (MODE == 2) { uxa = a.ux0[x]; uxb = a.ux1[x]; ufx = a.ufx[x]; }

  // The CTA walks `nb` vertically adjacent tiles.  The last 2m horizontally-summed rows of one tile are the
  // first 2m rows of the next one, so they stay in shared memory (ring of E rows, logical row l of the current
  // tile lives at physical row (l + off) mod E) and only T new rows of M are computed: halo recompute drops
  // from E*E/(T*T) to E/T per tile after the first.
  int off = 0;
  for (int c = 0; c < a.nb; ++c) {
    const int y0 = (by * a.nb + c) * T;
    if (y0 >= h) break;
    const int lstart = c == 0 ? 0 : 2 * m;    // first logical row that is new in this tile
    const int nrows = E - lstart;

    // ---- step A: M on the new halo rows (positions clamped to the image = BORDER_REPLICATE of M) ----
    // thread = one halo column and a run of RS consecutive rows.  Walking down a column, the bottom corners of
    // one pixel's bilinear gather are the top corners of the next one whenever the integer part of the warp
    // advanced by exactly one row (almost always: the flow is smooth), so they are carried in registers and only
    // the two new corners are loaded.  The loop is unrolled by two with the carried / new corner sets swapping
    // roles (no register moves); the flow vector of the next row is fetched one row ahead.
    if (ty < RP) {
      const int RS = (nrows + RP - 1) / RP;
      int l = lstart + ty * RS;
      const int l_end = min(l + RS, E);
      // first touches come from HBM (flow, R0, most of the R1 neighbourhood): pull the run into L2 now
      for (int r = l; r < l_end; ++r) {
        const int yy = clampi(y0 - m + r, 0, h - 1);
        const int o = yy * pitch + x;
        if (MODE == 1) prefetch_l2(fin + o);
        prefetch_l2(R0a + o);
        prefetch_l2(R0b + o);
        prefetch_l2(R1a + o);
        prefetch_l2(R1b + o);
      }
      if (l < l_end) {
        int pr = l + off;
        if (pr >= E) pr -= E;
        // shared-memory byte addresses of this thread's M record (float4 plane, float plane), kept incrementally
        unsigned sa4 = smem_u32(sM4 + pr * ES + ix), sa1 = smem_u32(sM1 + pr * ES + ix);
        const unsigned sa4_end = smem_u32(sM4 + E * ES + ix);
        int yu = y0 - m + l;                                       // unclamped image row of logical row l
        // loop invariants pinned in registers (ptxas otherwise re-derives them from the constant bank every row)
        const float4* r0a = R0a; const float* r0b = R0b; const float4* r1a = R1a; const float* r1b = R1b;
        const float2* fi = fin;
        int wm1 = w - 1, hm1 = h - 1, pit = pitch;
        int pitb = h > 1 ? pitch : 0;                              // keeps the unused bottom-corner loads in bounds
        int thr = xb_border ? 0 : h - 10;                          // (unsigned)(y - 5) >= thr  <=>  border pixel
        pin(r0a); pin(r0b); pin(r1a); pin(r1b); pin(fi); pin(wm1); pin(hm1); pin(pit); pin(pitb); pin(thr);
        FbCorner cA, cB;
        cA.a0 = cA.a1 = make_float4(0.f, 0.f, 0.f, 0.f);
        cA.e0 = cA.e1 = 0.f;
        cB = cA;
        int o_carry = -1 << 30;                                    // R1 offset the carried corners came from
        int yA = min(max(yu, 0), hm1), yB = yA;
        int oA = yA * pit + x, oB = oA;
        float2 dA = fetch_flow_m<MODE>(a, fi, oA, yA, uxa, uxb, ufx), dB = dA;
        auto row = [&](const float2 d, float2& dn, const int y, int& yn, const int o, int& on, bool has_next,
                       FbCorner& top, FbCorner& bot) {
          const float4 q = ldg_f4<0>(r0a + o);
          const float q4 = ldg_f1<0>(r0b + o);
          ++yu;
          if (has_next) {
            yn = min(max(yu, 0), hm1);
            on = yn * pit + x;
            dn = fetch_flow_m<MODE>(a, fi, on, yn, uxa, uxb, ufx);
          }
          float fx = xf + d.x, fy = (float)y + d.y;
          const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
          fx -= (float)x1; fy -= (float)y1;
          const bool inside = (unsigned)x1 < (unsigned)wm1 && (unsigned)y1 < (unsigned)hm1;
          const int ot = inside ? y1 * pit + x1 : 0;
          if (ot != o_carry) {                  // top corners are not the carried bottom corners: load them
            top.a0 = ldg_f4<0>(r1a + ot); top.a1 = ldg_f4<16>(r1a + ot);
            top.e0 = ldg_f1<0>(r1b + ot); top.e1 = ldg_f1<4>(r1b + ot);
          }
          const int ob = ot + pitb;
          bot.a0 = ldg_f4<0>(r1a + ob); bot.a1 = ldg_f4<16>(r1a + ob);
          bot.e0 = ldg_f1<0>(r1b + ob); bot.e1 = ldg_f1<4>(r1b + ob);
          o_carry = ob;
          const float gx = 1.f - fx, gy = 1.f - fy;
          const float a00 = gx * gy, a01 = fx * gy, a10 = gx * fy, a11 = fx * fy;
          float r2 = a00 * top.a0.x + a01 * top.a1.x + a10 * bot.a0.x + a11 * bot.a1.x;
          float r3 = a00 * top.a0.y + a01 * top.a1.y + a10 * bot.a0.y + a11 * bot.a1.y;
          float r4 = a00 * top.a0.z + a01 * top.a1.z + a10 * bot.a0.z + a11 * bot.a1.z;
          float r5 = a00 * top.a0.w + a01 * top.a1.w + a10 * bot.a0.w + a11 * bot.a1.w;
          float r6 = a00 * top.e0 + a01 * top.e1 + a10 * bot.e0 + a11 * bot.e1;
          r2 = inside ? r2 : 0.f;
          r3 = inside ? r3 : 0.f;
          r4 = inside ? r4 : q.z;               // (q + q) * 0.5 = q, (q4 + q4) * 0.25 = q4 * 0.5: exact
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
            const float s = bwx * border_w(y, h);
            r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
          }
          sts_f4(sa4, r4 * r4 + r6 * r6, (r4 + r5) * r6, r5 * r5 + r6 * r6, r4 * r2 + r6 * r3);
          sts_f1(sa1, r6 * r2 + r5 * r3);
          sa4 += ES * 16; sa1 += ES * 4;
          if (sa4 == sa4_end) { sa4 -= E * ES * 16; sa1 -= E * ES * 4; }   // ring wrap
        };
        for (; l + 1 < l_end; l += 2) {
          row(dA, dB, yA, yB, oA, oB, true, cA, cB);
          row(dB, dA, yB, yA, oB, oA, l + 2 < l_end, cB, cA);
        }
        if (l < l_end) row(dA, dB, yA, yB, oA, oB, false, cA, cB);
      }
    }
    __syncthreads();

    // ---- step B: horizontal window sums IN PLACE on the new rows.  One thread owns half a row of one plane:
    //   left half : outputs x in [0,HL)  computed left->right,  stored at position x        (reads positions >= x)
    //   right half: outputs x in [HL,T)  computed right->left,  stored at position x + 2m   (reads positions <= x+2m)
    // positions [HL, HL+2m) are written by neither, so the two halves never race.
    {
      const int gs = (E + 31) & ~31;            // group stride: each group starts on a warp boundary
      const int g = t / gs, r = t - g * gs;
      if (g < 4 && r < nrows) {
        int pr = lstart + r + off;
        if (pr >= E) pr -= E;
        const bool right = g & 1;
        if (g < 2) {
          float4* row = sM4 + pr * ES;
          if (GAUSS) {
            if (!right) {
              for (int xx = 0; xx < HL; ++xx) {
                float4 s = fma4s(row[xx + m], a.gtaps[0], make_float4(0.f, 0.f, 0.f, 0.f));
                for (int k = 1; k <= m; ++k) s = fma4s(add4(row[xx + m - k], row[xx + m + k]), a.gtaps[k], s);
                row[xx] = s;
              }
            } else {
              for (int xx = T - 1; xx >= HL; --xx) {
                float4 s = fma4s(row[xx + m], a.gtaps[0], make_float4(0.f, 0.f, 0.f, 0.f));
                for (int k = 1; k <= m; ++k) s = fma4s(add4(row[xx + m - k], row[xx + m + k]), a.gtaps[k], s);
                row[xx + 2 * m] = s;
              }
            }
          } else if (!right) {
            float4 s = row[0];
            for (int k = 1; k < 2 * m; ++k) s = add4(s, row[k]);
            for (int xx = 0; xx < HL; ++xx) {
              s = add4(s, row[xx + 2 * m]);
              float4 old = row[xx];
              row[xx] = s;
              s = sub4(s, old);
            }
          } else {
            float4 s = row[T];
            for (int k = 1; k < 2 * m; ++k) s = add4(s, row[T + k]);
            for (int xx = T - 1; xx >= HL; --xx) {
              s = add4(s, row[xx]);
              float4 old = row[xx + 2 * m];
              row[xx + 2 * m] = s;
              s = sub4(s, old);
            }
          }
        } else {
          float* row = sM1 + pr * ES;
          if (GAUSS) {
            if (!right) {
              for (int xx = 0; xx < HL; ++xx) {
                float s = row[xx + m] * a.gtaps[0];
                for (int k = 1; k <= m; ++k) s = fmaf(row[xx + m - k] + row[xx + m + k], a.gtaps[k], s);
                row[xx] = s;
              }
            } else {
              for (int xx = T - 1; xx >= HL; --xx) {
                float s = row[xx + m] * a.gtaps[0];
                for (int k = 1; k <= m; ++k) s = fmaf(row[xx + m - k] + row[xx + m + k], a.gtaps[k], s);
                row[xx + 2 * m] = s;
              }
            }
          } else if (!right) {
            float s = row[0];
            for (int k = 1; k < 2 * m; ++k) s += row[k];
            for (int xx = 0; xx < HL; ++xx) {
              s += row[xx + 2 * m];
              float old = row[xx];
              row[xx] = s;
              s -= old;
            }
          } else {
            float s = row[T];
            for (int k = 1; k < 2 * m; ++k) s += row[T + k];
            for (int xx = T - 1; xx >= HL; --xx) {
              s += row[xx];
              float old = row[xx + 2 * m];
              row[xx + 2 * m] = s;
              s -= old;
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- step C: vertical window sums + 2x2 solve; thread = (column, row segment) ----
    {
      const int nseg = NT / T;
      const int segr = (T + nseg - 1) / nseg;
      const int seg = t / T, xo = t - seg * T;
      const int gx = x0 + xo;
      const int xs = xo < HL ? xo : xo + 2 * m;   // where step B left this column's sums
      if (seg < nseg && gx < w) {
        const int r0 = seg * segr;
        const int r1 = min(r0 + segr, T);
        if (GAUSS) {
          for (int y = r0; y < r1; ++y) {
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float s1 = 0.f;
            for (int k = 0; k <= 2 * m; ++k) {
              int pr = y + k + off;
              if (pr >= E) pr -= E;
              float wk = a.gtaps[k < m ? m - k : k - m];
              s4 = fma4s(sM4[pr * ES + xs], wk, s4);
              s1 = fmaf(sM1[pr * ES + xs], wk, s1);
            }
            int gy = y0 + y;
            if (gy < h) fo[(size_t)gy * a.out_pitch + gx] = solve2x2(s4.x, s4.y, s4.z, s4.w, s1);
          }
        } else if (r0 < r1) {
          int pt = r0 + off;                    // tail (oldest row of the window)
          if (pt >= E) pt -= E;
          int ph = pt;                          // head
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float s1 = 0.f;
          for (int k = 0; k < 2 * m; ++k) {
            s4 = add4(s4, sM4[ph * ES + xs]);
            s1 += sM1[ph * ES + xs];
            if (++ph == E) ph = 0;
          }
          for (int y = r0; y < r1; ++y) {
            s4 = add4(s4, sM4[ph * ES + xs]);
            s1 += sM1[ph * ES + xs];
            if (++ph == E) ph = 0;
            int gy = y0 + y;
            if (gy < h) {
              float sc = a.inv_area;
              fo[(size_t)gy * a.out_pitch + gx] = solve2x2(s4.x * sc, s4.y * sc, s4.z * sc, s4.w * sc, s1 * sc);
            }
            s4 = sub4(s4, sM4[pt * ES + xs]);
            s1 -= sM1[pt * ES + xs];
            if (++pt == E) pt = 0;
          }
        }
      }
    }
    off += T;
    if (off >= E) off -= E;
    if (c + 1 < a.nb) __syncthreads();          // step C still reads the rows the next tile's step A overwrites
  }
}

#include "fb_ws.cuh"

// ----------------------------------------------------------------------------------------------
// OPTFLOW_USE_INITIAL_FLOW: flow_coarsest = resize(flow0, INTER_AREA) * scale   (optflowgf.cpp, first level)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fb_initflow_area(const float2* __restrict__ flow0, int W, size_t pair_stride,
                                                         float2* __restrict__ out, int w, int h, int pitch,
                                                         size_t out_pair_stride, const int* __restrict__ xs,
                                                         const int* __restrict__ xi, const float* __restrict__ xa,
                                                         const int* __restrict__ ys, const int* __restrict__ yi,
                                                         const float* __restrict__ ya, int fast, int sx, int sy,
                                                         float mult) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const float2* f = flow0 + blockIdx.z * pair_stride;
  float ax = 0.f, ay = 0.f;
  if (fast) {   // ResizeAreaFast: plain sum of the sx*sy block in row-major order, then * 1/area
    for (int j = 0; j < sy; ++j)
      for (int i = 0; i < sx; ++i) {
        float2 v = f[(size_t)(y * sy + j) * W + x * sx + i];
        ax += v.x; ay += v.y;
      }
    float sc = 1.f / (float)(sx * sy);
    ax *= sc; ay *= sc;
  } else {      // ResizeArea: rows weighted by beta, columns by alpha, float accumulation
    for (int j = ys[y]; j < ys[y + 1]; ++j) {
      const float2* row = f + (size_t)yi[j] * W;
      float rx = 0.f, ry = 0.f;
      for (int i = xs[x]; i < xs[x + 1]; ++i) {
        float2 v = row[xi[i]];
        rx = fmaf(v.x, xa[i], rx); ry = fmaf(v.y, xa[i], ry);
      }
      ax = fmaf(ya[j], rx, ax); ay = fmaf(ya[j], ry, ay);
    }
  }
  out[blockIdx.z * out_pair_stride + (size_t)y * pitch + x] = make_float2(ax * mult, ay * mult);
}

// ----------------------------------------------------------------------------------------------
// iterations == 0: cv2 still walks the pyramid, so the result is the initial flow (zero, or the caller's field
// area-resized to the coarsest level) carried up by the x(1/pyr_scale) bilinear resize of every level.
// ----------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) fb_flow_only(IterArgs a) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, pair = blockIdx.z;
  if (x >= a.w) return;
  const float2* fin = MODE ? a.flow_in + (size_t)pair * a.flow_in_pair_stride : nullptr;
  int xa = 0, xb = 0;
  float fx = 0.f;
  if (MODE == 2) { xa = a.ux0[x]; xb = a.ux1[x]; fx = a.ufx[x]; }
  a.flow_out[(size_t)pair * a.flow_out_pair_stride + (size_t)y * a.out_pitch + x] =
      fetch_flow_m<MODE>(a, fin, y * a.pitch + x, y, xa, xb, fx);
}

// Exact x2 upsample of the coarser level's flow (the reference's pyr_scale 0.5 on even sizes) as its own pass:
// thread (i, j) owns the coarse cell between columns i, i + 1 and rows j, j + 1 -- four loads -- and writes the four
// fine pixels that interpolate inside it (x = 2i + 1, 2i + 2; y = 2j + 1, 2j + 2; fractions 0.25 / 0.75, 0 at the
// clamped ends).  Same expression as fetch_flow_m<2>, same bits; a quarter of its loads per pixel and no tables.
__global__ void __launch_bounds__(256) fb_upsample2x(IterArgs a) {
  // thread = the fine pixel pair (2k, 2k + 1) of four fine rows 4m .. 4m + 3: its two pixels are ONE aligned 16-byte
  // store per row (a thread per coarse cell owns pixels 2i + 1, 2i + 2 -- two 8-byte stores 16 bytes apart across the
  // lanes, every 32-byte sector half written per instruction).  cv2's linear-resize table in closed form:
  //   x = 0 -> (0, 1, 0);  x = 2q + 1 -> (q, min(q + 1, n - 1), q >= n - 1 ? 0 : 0.25);  x = 2q + 2 -> (q, q + 1, 0.75)
  const int k = blockIdx.x * 64 + (threadIdx.x & 63), m = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int pair = blockIdx.z;
  if (2 * k >= a.w || 4 * m >= a.h) return;
  const float2* fin = a.flow_in + (size_t)pair * a.flow_in_pair_stride;
  const int cA = max(k - 1, 0), cB = k, cC = min(k + 1, a.cw - 1);
  // coarse rows 2m - 1 .. 2m + 2 (clamped), three columns each
  float2 p[4][3];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float2* row = fin + (size_t)min(max(2 * m - 1 + r, 0), a.ch - 1) * a.in_pitch;
    p[r][0] = row[cA]; p[r][1] = row[cB]; p[r][2] = row[cC];
  }
  const float fx0 = k == 0 ? 0.f : 0.75f, fx1 = k >= a.cw - 1 ? 0.f : 0.25f;
  const bool first = k == 0;                            // pixel 0 reads columns (0, min(1, n - 1)) = (cA, cC)
  float2* out = a.flow_out + (size_t)pair * a.flow_out_pair_stride + 2 * k;
#pragma unroll
  for (int dy = 0; dy < 4; ++dy) {
    const int y = 4 * m + dy;
    if (y >= a.h) break;
    // rows (ra, rb) as indices into p[] (coarse row 2m - 1 + index) and the weight of rb
    int ra, rb;
    float fy;
    if (dy == 0) { ra = y == 0 ? 1 : 0; rb = y == 0 ? 2 : 1; fy = y == 0 ? 0.f : 0.75f; }   // y = 2 (2m - 1) + 2
    else if (dy == 1) { ra = 1; rb = 2; fy = 2 * m >= a.ch - 1 ? 0.f : 0.25f; }             // y = 2 (2m) + 1
    else if (dy == 2) { ra = 1; rb = 2; fy = 0.75f; }                                       // y = 2 (2m) + 2
    else { ra = 2; rb = 3; fy = 2 * m + 1 >= a.ch - 1 ? 0.f : 0.25f; }                      // y = 2 (2m + 1) + 1
    float2 qa[3], qb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {                       // (dy == 0 picks between two register rows: selects, no indexing)
      qa[c] = dy == 0 ? (y == 0 ? p[1][c] : p[0][c]) : p[ra][c];
      qb[c] = dy == 0 ? (y == 0 ? p[2][c] : p[1][c]) : p[rb][c];
    }
    float2 o[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float fx = e ? fx1 : fx0;
      const float2 p00 = e ? qa[1] : qa[0], p01 = e ? qa[2] : (first ? qa[2] : qa[1]);
      const float2 p10 = e ? qb[1] : qb[0], p11 = e ? qb[2] : (first ? qb[2] : qb[1]);
      const float gx = 1.f - fx, gy = 1.f - fy;
      const float tx0 = lerp_nc(p00.x, gx, p01.x, fx), ty0 = lerp_nc(p00.y, gx, p01.y, fx);
      const float tx1 = lerp_nc(p10.x, gx, p11.x, fx), ty1 = lerp_nc(p10.y, gx, p11.y, fx);
      o[e] = make_float2(__fmul_rn(lerp_nc(tx0, gy, tx1, fy), a.up_mult), __fmul_rn(lerp_nc(ty0, gy, ty1, fy), a.up_mult));
    }
    *(float4*)(out + (size_t)y * a.out_pitch) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
  }
}

// ----------------------------------------------------------------------------------------------
// host driver
// ----------------------------------------------------------------------------------------------
struct FbWorkspace {
  float* T;       // [frames][rows][pitch0]
  float* I;       // [frames][S]
  float* R;       // per level: [frames][5][h][pitch]; level bases via offsets
  float2* FA;     // per level ping: [pairs][h][pitch]
  float2* FB;     // per level pong
  size_t bytes;
};

static size_t fb_layout(const FbPlan* pl, int pairs, int frames, void* base, size_t cap, FbWorkspace* ws) {
  Arena ar(base, cap);
  size_t pitch0 = align_up(pl->cols, 32);
  ws->T = ar.take<float>((size_t)frames * pl->rows * pitch0);
  ws->I = ar.take<float>((size_t)frames * pl->S);
  ws->R = ar.take<float>((size_t)frames * 5 * pl->S);
  ws->FA = ar.take<float2>((size_t)pairs * pl->S);
  ws->FB = ar.take<float2>((size_t)pairs * pl->S);
  ws->bytes = align_up(ar.off, 256);
  return ws->bytes;
}

static int check_params(int rows, int cols, const b2of_farneback_params* p) {
  const char* fn = "calcOpticalFlowFarneback";
  B2OF_ASSERT(p != nullptr, fn);
  B2OF_ASSERT(rows > 0 && cols > 0, fn);
  B2OF_ASSERT(p->pyr_scale > 0 && p->pyr_scale < 1, fn);
  B2OF_ASSERT(p->levels >= 0, fn);
  B2OF_ASSERT(p->winsize >= 1, fn);
  B2OF_ASSERT(p->iterations >= 0, fn);
  B2OF_ASSERT(p->poly_n >= 1, fn);
  if (p->poly_n > FB_MAX_POLY_N) return fail(B2OF_E_UNSUPPORTED, "poly_n > %d is not supported", FB_MAX_POLY_N);
  if (p->winsize / 2 > 45) return fail(B2OF_E_UNSUPPORTED, "winsize > 91 is not supported");
  return B2OF_OK;
}

size_t farneback_workspace_bytes(int rows, int cols, const b2of_farneback_params* p, int chunk_pairs, int shared) {
  if (check_params(rows, cols, p)) return 0;
  FbPlan* pl;
  if (get_plan(rows, cols, *p, &pl)) return 0;
  FbWorkspace ws;
  int frames = shared ? chunk_pairs + 1 : 2 * chunk_pairs;
  return fb_layout(pl, chunk_pairs, frames, nullptr, 0, &ws);
}

// per-frame work: level images + polynomial expansion for `frames` frames; frame i lands in workspace
// slot slot0 + i*slot_step (independent pairs interleave prev/next frames: slot_step 2)
static int fb_frames(const FbPlan* pl, const FbWorkspace& ws, const uint8_t* frames_dev, size_t step,
                     size_t frame_stride, int frames, int slot0, int slot_step, int total_slots, cudaStream_t st) {
  const int W = pl->cols, H = pl->rows;
  size_t pitch0 = align_up(W, 32);
  size_t lvl_off = 0;
  // the reference's three coarse levels ((K, S) = (4, 2), (10, 4), (20, 8)) come out of one pass over the frame
  bool fused_level[8] = {};
  {
    int idx[3] = {-1, -1, -1};
    const int wantK[3] = {4, 10, 20}, wantS[3] = {2, 4, 8};
    for (size_t li = 0; li < pl->lv.size() && li < 8; ++li)
      for (int q = 0; q < 3; ++q)
        if (pl->lv[li].r_K == wantK[q] && pl->lv[li].r_S == wantS[q]) idx[q] = (int)li;
    if (idx[0] >= 0 && idx[1] >= 0 && idx[2] >= 0) {
      LevelCoarseArgs ca{};
      ca.frames = frames_dev; ca.step = step; ca.frame_stride = frame_stride; ca.W = W; ca.H = H;
      size_t off = 0;
      std::vector<size_t> lvoff(pl->lv.size());
      for (size_t li = 0; li < pl->lv.size(); ++li) { lvoff[li] = off; off += (size_t)pl->lv[li].h * pl->lv[li].pitch; }
      int gx = 1, gy = 1;
      const int wantC0[3] = {-1, -3, -6};          // the kernel's region offsets are compiled in for these
      bool ok = true;
      for (int q = 0; q < 3; ++q) {
        const FbLevel& L = pl->lv[idx[q]];
        const size_t plane = (size_t)L.h * L.pitch;
        ca.I[q] = ws.I + (size_t)total_slots * lvoff[idx[q]] + (size_t)slot0 * plane;
        ca.wk[q] = L.w; ca.hk[q] = L.h; ca.pitch[q] = L.pitch; ca.i_frame_stride[q] = plane * slot_step;
        ok = ok && L.r_c0 == wantC0[q];
        const int T = 64 / wantS[q];
        gx = std::max(gx, cdiv(L.w, T)); gy = std::max(gy, cdiv(L.h, T));
      }
      memcpy(ca.c2, pl->lv[idx[0]].r_c, sizeof ca.c2);
      memcpy(ca.c4, pl->lv[idx[1]].r_c, sizeof ca.c4);
      memcpy(ca.c8, pl->lv[idx[2]].r_c, sizeof ca.c8);
      if (ok) {
        double bytes = (double)W * H;
        for (int q = 0; q < 3; ++q) bytes += 4.0 * ca.wk[q] * ca.hk[q];
        {
          ProfScope ps(PT_FB_LEVEL_H, st, (double)frames * bytes);
          fb_levels_coarse<<<dim3(gx, gy, frames), LC_NT, 0, st>>>(ca);
        }
        B2OF_LAUNCH_CHECK();
        for (int q = 0; q < 3; ++q) fused_level[idx[q]] = true;
      }
    }
  }
  for (size_t li = 0; li < pl->lv.size(); ++li) {
    const FbLevel& L = pl->lv[li];
    size_t plane = (size_t)L.h * L.pitch;
    float* Tb = ws.T + (size_t)slot0 * H * pitch0;
    float* Ib = ws.I + (size_t)total_slots * lvl_off + (size_t)slot0 * plane;
    float* Rb = ws.R + (size_t)total_slots * 5 * lvl_off + (size_t)slot0 * 5 * plane;
    size_t t_stride = (size_t)H * pitch0 * slot_step, i_stride = plane * slot_step, r_stride = 5 * plane * slot_step;
#ifndef FB_FUSE_LEVEL0
#define FB_FUSE_LEVEL0 0   // measured: the level-image kernel goes (-0.27 ms per 65 frames) but the expansion, already
#endif                     // issue-bound at 77 %, pays +0.46 ms for the two extra passes: off (profiles/README.md)
    // the full-resolution level in its regular (3, 1) form with poly_n = 5: its image is computed inside the expansion
    const bool fuse0 = FB_FUSE_LEVEL0 && L.r_K == 3 && L.r_S == 1 && L.r_c0 == -1 && pl->p.poly_n == 5 &&
                       L.w == W && L.h == H && W >= 4 && H >= 4;
    if (fuse0) {
      // no level image in HBM at all
    } else if (li < 8 && fused_level[li]) {
      // level image already written by fb_levels_coarse
    } else if (L.r_K) {
      LevelRegArgs ra{};
      ra.frames = frames_dev; ra.step = step; ra.frame_stride = frame_stride; ra.W = W; ra.H = H;
      ra.I = Ib; ra.wk = L.w; ra.hk = L.h; ra.pitch = L.pitch; ra.i_frame_stride = i_stride;
      ra.c0 = L.r_c0;
      memcpy(ra.c, L.r_c, sizeof ra.c);
      {
        ProfScope ps(PT_FB_LEVEL_H, st, (double)frames * ((double)W * H + 4.0 * L.h * L.w));
        if (L.r_K == 3 && L.r_S == 1 && L.r_c0 == -1 && L.w == W && L.h == H && W % 4 == 0 && W >= 8 && H >= 2 &&
            ((step | frame_stride | (size_t)frames_dev) & 3) == 0) {
          Level0Args la{frames_dev, step, frame_stride, W, H, Ib, L.pitch, i_stride, L.r_c[0], L.r_c[1], L.r_c[2]};
          fb_level0_stream<<<dim3(cdiv(W, 128 * L0_WARPS), cdiv(H, L0_TH), frames), 32 * L0_WARPS, 0, st>>>(la);
        } else if (L.r_K == 3) launch_level_regular<3, 1>(ra, frames, st);
        else if (L.r_K == 4) launch_level_regular<4, 2>(ra, frames, st);
        else if (L.r_K == 10) launch_level_regular<10, 4>(ra, frames, st);
        else launch_level_regular<20, 8>(ra, frames, st);
      }
      B2OF_LAUNCH_CHECK();
    } else if (L.f_tw) {
      LevelFusedArgs fa{};
      fa.frames = frames_dev; fa.step = step; fa.frame_stride = frame_stride; fa.W = W; fa.H = H;
      fa.I = Ib; fa.wk = L.w; fa.hk = L.h; fa.pitch = L.pitch; fa.i_frame_stride = i_stride;
      fa.taps = L.taps; fa.ksz = L.ksz;
      fa.sx0 = L.sx0; fa.sx1 = L.sx1; fa.sy0 = L.sy0; fa.sy1 = L.sy1; fa.fx = L.fx; fa.fy = L.fy;
      fa.tw = L.f_tw; fa.th = L.f_th; fa.rw = L.f_rw; fa.rh = L.f_rh; fa.rw_pad = L.f_rw_pad;
      dim3 gf(cdiv(L.w, L.f_tw), cdiv(L.h, L.f_th), frames);
      {
        // algorithmic bytes: the u8 frame once + the f32 level image
        ProfScope ps(PT_FB_LEVEL_H, st, (double)frames * ((double)W * H + 4.0 * L.h * L.w));
        fb_level_fused<<<gf, 256, L.f_smem, st>>>(fa);
      }
      B2OF_LAUNCH_CHECK();
    } else {
      dim3 g1(cdiv(L.w, 128), H, frames);
      {
        ProfScope ps(PT_FB_LEVEL_H, st, (double)frames * ((double)W * H + 4.0 * H * L.w));
        fb_level_hpass<<<g1, 128, 0, st>>>(frames_dev, step, frame_stride, W, H, Tb, L.w, L.pitch, t_stride, L.taps,
                                           L.ksz, L.sx0, L.sx1, L.fx);
      }
      B2OF_LAUNCH_CHECK();
      dim3 g2(cdiv(L.w, 128), L.h, frames);
      {
        ProfScope ps(PT_FB_LEVEL_V, st, (double)frames * (4.0 * H * L.w + 4.0 * L.h * L.w));
        fb_level_vpass<<<g2, 128, 0, st>>>(Tb, H, L.pitch, t_stride, Ib, L.w, L.h, i_stride, L.taps, L.ksz, L.sy0,
                                           L.sy1, L.fy);
      }
      B2OF_LAUNCH_CHECK();
    }
    int n = pl->p.poly_n;
    size_t smem = ((size_t)(PE_TH + 2 * n) * (PE_TW + 2 * n) + 3 * (size_t)PE_TH * (PE_TW + 2 * n)) * sizeof(float);
    dim3 g3(cdiv(L.w, PE_TW), cdiv(L.h, PE_TH), frames);
    {
      ProfScope ps(PT_FB_POLYEXP, st, (double)frames * 24.0 * L.w * L.h);
      Poly0Src ps0{};
      if (fuse0) {
        ps0.frames = frames_dev; ps0.step = step; ps0.frame_stride = frame_stride;
        ps0.c[0] = L.r_c[0]; ps0.c[1] = L.r_c[1]; ps0.c[2] = L.r_c[2];
        fb_polyexp_n<5, true><<<g3, 320, 0, st>>>(Ib, L.w, L.h, L.pitch, i_stride, Rb, plane, r_stride, pl->pc, ps0);
#ifndef FB_POLY_PACKED
#define FB_POLY_PACKED 1
#endif
      } else if (FB_POLY_PACKED && n == 5) fb_polyexp_p<5><<<g3, 256, 0, st>>>(Ib, L.w, L.h, L.pitch, i_stride, Rb, plane, r_stride, pl->pc);
      else if (FB_POLY_PACKED && n == 7) fb_polyexp_p<7><<<g3, 256, 0, st>>>(Ib, L.w, L.h, L.pitch, i_stride, Rb, plane, r_stride, pl->pc);
      else if (n == 5) fb_polyexp_n<5, false><<<g3, 320, 0, st>>>(Ib, L.w, L.h, L.pitch, i_stride, Rb, plane, r_stride, pl->pc, ps0);
      else if (n == 7) fb_polyexp_n<7, false><<<g3, 320, 0, st>>>(Ib, L.w, L.h, L.pitch, i_stride, Rb, plane, r_stride, pl->pc, ps0);
      else fb_polyexp<<<g3, 256, smem, st>>>(Ib, L.w, L.h, L.pitch, i_stride, Rb, plane, r_stride, pl->pc, n);
    }
    B2OF_LAUNCH_CHECK();
    lvl_off += plane;
  }
  return B2OF_OK;
}

static PerDeviceOnce g_attr_once;
static void set_func_attrs() {
  const int big = 227 * 1024;
#define B2OF_ATTR(NT_, CT_, CM_, G_)                                                                        \
  cudaFuncSetAttribute(fb_iter<NT_, CT_, CM_, G_, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);    \
  cudaFuncSetAttribute(fb_iter<NT_, CT_, CM_, G_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);    \
  cudaFuncSetAttribute(fb_iter<NT_, CT_, CM_, G_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  B2OF_ATTR(512, 0, 0, false)
  B2OF_ATTR(512, 0, 0, true)
#undef B2OF_ATTR
#define B2OF_WS_ATTR(MODE, STATS, HB) \
  cudaFuncSetAttribute(fb_iter_ws<MODE, STATS, HB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FBW_SMEM)
  B2OF_WS_ATTR(0, false, false); B2OF_WS_ATTR(1, false, false); B2OF_WS_ATTR(2, false, false);
  B2OF_WS_ATTR(0, true, false); B2OF_WS_ATTR(1, true, false); B2OF_WS_ATTR(2, true, false);
  B2OF_WS_ATTR(0, false, true); B2OF_WS_ATTR(1, false, true); B2OF_WS_ATTR(2, false, true);
  B2OF_WS_ATTR(0, true, true); B2OF_WS_ATTR(1, true, true); B2OF_WS_ATTR(2, true, true);
#undef B2OF_WS_ATTR
  cudaFuncSetAttribute(fb_polyexp, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(fb_level_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
}

// iterations for `pairs` pairs whose frames sit in workspace slots (pair p -> slots p*fstep, p*fstep+1)
int flow_stats_dev(const float*, int, int, int, float*, cudaStream_t);          // pathfinder.cu
int flow_stats_finalize_dev(float*, size_t, int, cudaStream_t);

// Pairs [pair0, pair0 + pairs) of a chunk of pairs_total pairs whose per-frame work (fb_frames) is in the workspace.
// `concurrent`: another range of the same chunk runs on a second stream at the same time (fb_pairs).
static int fb_pairs_range(const FbPlan* pl, const FbWorkspace& ws, int pairs_total, int pair0, int pairs, int fstep,
                          int total_slots, float* flow_out, float* stats, const b2of_farneback_params& p,
                          bool concurrent, cudaStream_t st) {
  // `p` is THIS call's parameter block: the cached plan only fixes what its key holds (sizes, windows, taps)
  const int call_flags = p.flags;
  flow_out += (size_t)pair0 * pl->rows * pl->cols * 2;
  if (stats) stats += (size_t)pair0 * B2OF_STATS_WIDTH;
  const int m = p.winsize / 2;
  const bool gauss = (p.flags & B2OF_OPTFLOW_FARNEBACK_GAUSSIAN) != 0;
  // the reference's window runs on the strip kernel; anything else on the general kernel with the largest tile whose
  // halo fits one SM
  const bool fast = !gauss && m == FBS_M;
  int tile = 64;
  while (tile > 16 && (size_t)5 * (tile + 2 * m) * ((tile + 2 * m) | 1) * sizeof(float) > 200 * 1024) tile -= 8;
  const int E = tile + 2 * m;
  size_t smem = (size_t)5 * E * (E | 1) * sizeof(float);
  if (!fast && (smem > 227 * 1024 || E > IT_THREADS / 4))
    return fail(B2OF_E_UNSUPPORTED, "winsize %d needs %zu B of shared memory", p.winsize, smem);
  static std::atomic<int> n_sm_dev[B2OF_MAX_DEVICES];
  int n_sm = n_sm_dev[current_device()].load();
  if (n_sm == 0) {
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, current_device()) != cudaSuccess || n_sm <= 0)
      n_sm = 148;
    n_sm_dev[current_device()].store(n_sm);
  }
  size_t lvl_off = 0;
  const float2* coarse = nullptr;
  int cpitch = 0, cw = 0, ch = 0;
  size_t cstride = 0;
  for (size_t li = 0; li < pl->lv.size(); ++li) {
    const FbLevel& L = pl->lv[li];
    const bool last_level = li + 1 == pl->lv.size();
    size_t plane = (size_t)L.h * L.pitch;
    float2* A = ws.FA + (size_t)pairs_total * lvl_off + (size_t)pair0 * plane;
    float2* B = ws.FB + (size_t)pairs_total * lvl_off + (size_t)pair0 * plane;
    IterArgs a{};
    a.R = ws.R + (size_t)total_slots * 5 * lvl_off + (size_t)pair0 * fstep * 5 * plane;
    a.r_frame_stride = 5 * plane;
    a.plane_stride = plane;
    a.pitch = L.pitch; a.w = L.w; a.h = L.h;
    a.pair_frame_step = fstep;
    a.m = m;
    a.tile = tile;
    a.inv_area = (float)(1.0 / ((double)p.winsize * p.winsize));
    a.gtaps = gauss ? pl->gauss_taps : nullptr;
    a.ux0 = L.ux0; a.ux1 = L.ux1; a.uy0 = L.uy0; a.uy1 = L.uy1; a.ufx = L.ufx; a.ufy = L.ufy;
    a.up_mult = (float)(1.0 / p.pyr_scale);
    const float2* cur = nullptr;  // flow at this level after the previous iteration
    int cur_pitch = 0;
    size_t cur_stride = 0;
    // vertical streaming: as many tiles per CTA as still leaves >= ~4 waves of CTAs
    const int cx = cdiv(L.w, tile), cy = cdiv(L.h, tile);
    int nb = (int)(((long long)cx * cy * pairs) / (4 * 296));
    const int nb_cap = 8;
    nb = nb < 1 ? 1 : (nb > nb_cap ? nb_cap : nb);
    nb = cdiv(cy, cdiv(cy, nb));                 // balance the row groups
    a.nb = nb;
    dim3 grid(pairs, cx, cdiv(cy, nb));
    int iters = p.iterations;
    const bool flow_only = iters == 0;              // cv2: the level's initial flow is its result
    if (flow_only) iters = 1;
    for (int it = 0; it < iters; ++it) {
      if (it == 0) {
        if (!coarse && (call_flags & B2OF_OPTFLOW_USE_INITIAL_FLOW)) {   // per call: the plan is shared
          // initial flow of every pair sits in the caller's output buffer: area-resize it into the pong buffer
          dim3 gi(cdiv(L.w, 256), L.h, pairs);
          fb_initflow_area<<<gi, 256, 0, st>>>((const float2*)flow_out, pl->cols, (size_t)pl->cols * pl->rows, B, L.w,
                                               L.h, L.pitch, plane, pl->ax_start, pl->ax_src, pl->ax_alpha,
                                               pl->ay_start, pl->ay_src, pl->ay_alpha,
                                               (L.w == pl->cols && L.h == pl->rows) ? 1 : pl->area_fast,
                                               (L.w == pl->cols && L.h == pl->rows) ? 1 : pl->area_sx,
                                               (L.w == pl->cols && L.h == pl->rows) ? 1 : pl->area_sy, (float)L.scale);
          B2OF_LAUNCH_CHECK();
          a.mode = 1; a.flow_in = B; a.flow_in_pair_stride = plane; a.in_pitch = L.pitch;
        } else if (coarse) {
          a.mode = 2; a.flow_in = coarse; a.flow_in_pair_stride = cstride; a.in_pitch = cpitch; a.cw = cw; a.ch = ch;
          a.up_exact2 = L.h == 2 * ch && L.w == 2 * cw && ch >= 2;
        } else {
          a.mode = 0; a.flow_in = nullptr;
        }
      } else {
        a.mode = 1; a.flow_in = cur; a.flow_in_pair_stride = cur_stride; a.in_pitch = cur_pitch;
      }
      bool final_out = last_level && it == iters - 1;
      float2* dst;
      int dpitch;
      size_t dstride;
      if (final_out) {
        dst = (float2*)flow_out; dpitch = L.w; dstride = (size_t)L.w * L.h;
      } else {
        dst = (it & 1) ? B : A; dpitch = L.pitch; dstride = plane;
      }
      a.flow_out = dst; a.out_pitch = dpitch; a.flow_out_pair_stride = dstride;
      a.stats_acc = nullptr;
#ifndef FB_SEPARATE_UPSAMPLE
#define FB_SEPARATE_UPSAMPLE 1
#endif
      if (FB_SEPARATE_UPSAMPLE && fast && !flow_only && a.mode == 2) {
        // The x(1/pyr_scale) upsample of the coarser level's flow as its own small pass into the level's second
        // flow buffer (free during the first iteration), so that the first iteration reads its flow the way the
        // others do.  Folded into the fused kernel (four gathers and a second dependent load per pixel in step A)
        // the first launch of a level took 27 % longer than the later ones: 2.29 vs 1.81 ms at 1080p x 64 pairs,
        // against 0.2 ms for this pass.  Same arithmetic (fetch_flow_m<2>), same bits.
        IterArgs u = a;
        u.flow_out = B; u.out_pitch = L.pitch; u.flow_out_pair_stride = plane;
        dim3 gf(cdiv(L.w, 256), L.h, pairs);
        {
          ProfScope ps(PT_FB_UPSAMPLE, st, pairs * (8.0 * cw * ch + 8.0 * L.w * L.h));
          if (u.up_exact2) fb_upsample2x<<<dim3(cdiv(cw, 64), cdiv(L.h, 16), pairs), 256, 0, st>>>(u);
          else fb_flow_only<2><<<gf, 256, 0, st>>>(u);
        }
        B2OF_LAUNCH_CHECK();
        a.mode = 1; a.flow_in = B; a.flow_in_pair_stride = plane; a.in_pitch = L.pitch;
      }
      if (flow_only) {
        dim3 gf(cdiv(L.w, 256), L.h, pairs);
        if (a.mode == 0) fb_flow_only<0><<<gf, 256, 0, st>>>(a);
        else if (a.mode == 1) fb_flow_only<1><<<gf, 256, 0, st>>>(a);
        else fb_flow_only<2><<<gf, 256, 0, st>>>(a);
        B2OF_LAUNCH_CHECK();
        cur = dst; cur_pitch = dpitch; cur_stride = dstride;
        continue;
      }
      if (final_out && stats && fast) {
        // the last launch also reduces its flow field to the per-pair statistics (fixed-point partial sums)
        B2OF_CUDA(cudaMemsetAsync(stats, 0, (size_t)pairs * B2OF_STATS_WIDTH * sizeof(float), st));
        a.stats_acc = (unsigned long long*)stats;
      }
      {
        // algorithmic bytes of this launch: R0 + R1 (20 B/px each), flow in (8 B/px, or 8 B per coarse px, or none),
        // flow out (8 B/px)
        double px = (double)L.w * L.h;
        double bytes = pairs * (48.0 * px + (a.mode == 1 ? 8.0 * px : a.mode == 2 ? 8.0 * cw * ch : 0.0));
        ProfScope ps(last_level ? PT_FB_ITER_FINEST : PT_FB_ITER_COARSE, st, bytes);
#define B2OF_ITER_LAUNCH(NT_, CT_, CM_, G_)                                                 \
  do {                                                                                      \
    if (a.mode == 0) fb_iter<NT_, CT_, CM_, G_, 0><<<grid, NT_, smem, st>>>(a);             \
    else if (a.mode == 1) fb_iter<NT_, CT_, CM_, G_, 1><<<grid, NT_, smem, st>>>(a);        \
    else fb_iter<NT_, CT_, CM_, G_, 2><<<grid, NT_, smem, st>>>(a);                         \
  } while (0)
        if (fast) {
          // warp-specialised strip kernel: one CTA per SM, 112-column strips, nseg row segments of nb 16-row blocks
          // (at least two waves of CTAs when the batch is small)
          IterArgs b = a;
#if FBW_PAIR
          const int nstrips = 2 * cdiv(L.w, 2 * FBS_TW), blocks = cdiv(L.h, FBW_RB);   // CTA pairs (cluster 1 x 2 x 1)
#elif defined(FBW_CLUSTER_ONLY)
          const int nstrips = 2 * cdiv(cdiv(L.w, FBS_TW), 2), blocks = cdiv(L.h, FBW_RB);
#else
          const int nstrips = cdiv(L.w, FBS_TW), blocks = cdiv(L.h, FBW_RB);
#endif
          // row segments: the split that minimises (waves of CTAs) x (rows per CTA + the 30 rows of pipeline fill
          // and halo); nb is a multiple of the refresh period, so results do not depend on the split or the batch
          int best_nb = cdiv(blocks, FBW_REFRESH) * FBW_REFRESH;
          long long best_cost = -1;
          for (int nbc = FBW_REFRESH; nbc <= cdiv(blocks, FBW_REFRESH) * FBW_REFRESH; nbc += FBW_REFRESH) {
            const long long ctas = (long long)pairs * nstrips * cdiv(blocks, nbc);
            const long long rows = (nbc < blocks ? nbc : blocks) * FBW_RB + 30;
            // with a second range of the chunk on another stream the SMs a partial last wave leaves idle run that
            // range's CTAs: what counts is the total work, as long as this launch alone can fill the machine
            const long long cost = (concurrent && ctas >= n_sm) ? ctas * rows : ((ctas + n_sm - 1) / n_sm) * rows * n_sm;
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_nb = nbc; }
          }
          b.nb = best_nb;
          if (b.mode == 1 && b.in_pitch != L.pitch)     // the kernel indexes flow_in with the level's own pitch
            return fail(B2OF_E_BADARG, "internal: flow_in pitch %d != level pitch %d", b.in_pitch, L.pitch);
          dim3 gs(pairs, nstrips, cdiv(blocks, b.nb));
          // horizontal window sums: per block of 15 (fb_ws.cuh) at the two coarsest levels -- that is where a sliding
          // sum's carried rounding error decides near-singular pixels (numpy restatement on real footage: blocked sums
          // at these two levels alone give the whole gain, at the finest level alone none), and they are 8 % of the
          // iteration time -- and sliding at the finer ones, unless the call asks for blocked sums everywhere
          // (B2OF_FARNEBACK_BLOCKED_SUMS)
#ifndef FBW_HBLOCK_COARSE_LEVELS
#define FBW_HBLOCK_COARSE_LEVELS 2
#endif
          const bool hblock = (call_flags & B2OF_FARNEBACK_BLOCKED_SUMS) != 0 ||
                              ((int)li < FBW_HBLOCK_COARSE_LEVELS && !last_level);
#define B2OF_WS_LAUNCH(MODE, STATS, HB) fb_iter_ws<MODE, STATS, HB><<<gs, FBW_NT, FBW_SMEM, st>>>(b)
#define B2OF_WS_MODE(STATS, HB) \
  do { if (b.mode == 0) B2OF_WS_LAUNCH(0, STATS, HB); else if (b.mode == 1) B2OF_WS_LAUNCH(1, STATS, HB); \
       else B2OF_WS_LAUNCH(2, STATS, HB); } while (0)
          if (b.stats_acc) { if (hblock) B2OF_WS_MODE(true, true); else B2OF_WS_MODE(true, false); }
          else { if (hblock) B2OF_WS_MODE(false, true); else B2OF_WS_MODE(false, false); }
#undef B2OF_WS_MODE
#undef B2OF_WS_LAUNCH
        } else if (gauss) B2OF_ITER_LAUNCH(512, 0, 0, true);
        else B2OF_ITER_LAUNCH(512, 0, 0, false);
#undef B2OF_ITER_LAUNCH
      }
      B2OF_LAUNCH_CHECK();
      cur = dst; cur_pitch = dpitch; cur_stride = dstride;
    }
    coarse = cur; cpitch = cur_pitch; cstride = cur_stride; cw = L.w; ch = L.h;
    lvl_off += plane;
  }
  if (stats) {
    if (fast && p.iterations > 0) return flow_stats_finalize_dev(stats, (size_t)pl->rows * pl->cols, pairs, st);
    return flow_stats_dev(flow_out, pairs, pl->rows, pl->cols, stats, st);
  }
  return B2OF_OK;
}

// Side streams of a device (fb_pairs), created on first use, destroyed by b2of_release().
constexpr int FB_MAX_RANGES = 8;
struct FbFork {
  cudaStream_t s[FB_MAX_RANGES - 1] = {};
  cudaEvent_t fork = nullptr, join[FB_MAX_RANGES - 1] = {};
  bool ok = false;
  std::mutex mu;
};
static FbFork g_fork[B2OF_MAX_DEVICES];
static std::mutex g_fork_mu;

static FbFork* get_fork() {
  std::lock_guard<std::mutex> lock(g_fork_mu);
  FbFork& f = g_fork[current_device()];
  if (!f.ok) {
    bool good = cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < FB_MAX_RANGES - 1 && good; ++i)
      good = cudaStreamCreateWithFlags(&f.s[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&f.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!good) {
      cudaGetLastError();
      return nullptr;
    }
    f.ok = true;
  }
  return &f;
}

void farneback_release_streams() {
  std::lock_guard<std::mutex> lock(g_fork_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < B2OF_MAX_DEVICES; ++d) {
    FbFork& f = g_fork[d];
    if (!f.ok) continue;
    cudaSetDevice(d);
    cudaEventDestroy(f.fork);
    for (int i = 0; i < FB_MAX_RANGES - 1; ++i) {
      cudaStreamDestroy(f.s[i]); cudaEventDestroy(f.join[i]);
      f.s[i] = nullptr; f.join[i] = nullptr;
    }
    f.fork = nullptr; f.ok = false;
  }
  cudaSetDevice(cur);
}

#ifndef FB_RANGES
#define FB_RANGES 8
#endif
// The pairs of a chunk are independent once the per-frame work is done.  A chunk of sixteen pairs or more is walked as
// two, and one of 32 or more as FB_RANGES = 4, contiguous ranges on as many streams (measured at 64 pairs, 1080p: one
// stream 7207 pairs/s, two 7500, three 7470, four 7630, six 7450, eight 7680): every launch of the level / iteration sequence then has launches of
// the other ranges queued next to it, and the SMs that a partial last wave of CTAs would leave idle (one 160 KB CTA
// per SM: 7.8 waves at the finest level of 64 pairs, 2.2 and 1.3 at the two coarsest) pick up their CTAs.  Same
// kernels, same arguments per pair, same bits.  With per-kernel profiling on (b2of_profile_enable) the chunk runs as
// one range on the caller's stream, so that a launch's events time that launch alone.
static int fb_pairs(const FbPlan* pl, const FbWorkspace& ws, int pairs, int fstep, int total_slots, float* flow_out,
                    float* stats, const b2of_farneback_params& p, cudaStream_t st) {
  static const int n_env = getenv("B2OF_STREAMS") ? atoi(getenv("B2OF_STREAMS")) : FB_RANGES;   // developer A/B knob
  // a power of two; a range keeps at least two pairs and enough CTAs at the finest level to fill min_fill of the SMs
  static const int fill_env = getenv("B2OF_RANGE_FILL") ? atoi(getenv("B2OF_RANGE_FILL")) : 90;   // percent of the SMs
  const long long strips = cdiv(pl->cols, FBS_TW);
  int nr = 1;
  while (2 * nr <= n_env && 2 * nr <= FB_MAX_RANGES && pairs / (2 * nr) >= 2 &&
         (pairs / (2 * nr)) * strips * 100 >= (long long)fill_env * 148)
    nr *= 2;
  const bool fast = !(p.flags & B2OF_OPTFLOW_FARNEBACK_GAUSSIAN) && p.winsize / 2 == FBS_M;
  FbFork* f = nullptr;
  if (nr > 1 && fast && p.iterations > 0 && g_prof_on.load(std::memory_order_relaxed) == 0)
    f = get_fork();
  if (!f) return fb_pairs_range(pl, ws, pairs, 0, pairs, fstep, total_slots, flow_out, stats, p, false, st);
  // one host thread at a time enqueues on a device's side streams: a wait captures the event's record at the time of
  // the call, so the events can be re-recorded by the next caller as soon as this one has enqueued its joins
  std::lock_guard<std::mutex> lock(f->mu);
  B2OF_CUDA(cudaEventRecord(f->fork, st));
  int rc = 0;
  for (int r = 0; r < nr; ++r) {
    const int p0 = (int)((long long)pairs * r / nr), p1 = (int)((long long)pairs * (r + 1) / nr);
    cudaStream_t sr = r == 0 ? st : f->s[r - 1];
    if (r > 0) cudaStreamWaitEvent(sr, f->fork, 0);
    const int rcr = fb_pairs_range(pl, ws, pairs, p0, p1 - p0, fstep, total_slots, flow_out, stats, p, true, sr);
    if (rcr && !rc) rc = rcr;
    // always join: the caller's stream must not run ahead of a side stream, error or not
    if (r > 0) { cudaEventRecord(f->join[r - 1], sr); cudaStreamWaitEvent(st, f->join[r - 1], 0); }
  }
  return rc;
}

int farneback_dev(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride, int n_pairs, int shared,
                  int rows, int cols, const b2of_farneback_params* p, float* flow, float* stats, void* workspace,
                  size_t workspace_bytes, cudaStream_t st) {
  int rc = check_params(rows, cols, p);
  if (rc) return rc;
  const char* fn = "calcOpticalFlowFarneback";
  B2OF_ASSERT(prev != nullptr && next != nullptr && flow != nullptr, fn);
  B2OF_ASSERT(step >= (size_t)cols, fn);
  B2OF_ASSERT(stats == nullptr || ((uintptr_t)stats & 7) == 0, fn);
  if (n_pairs <= 0) return B2OF_OK;
  FbPlan* pl;
  rc = get_plan(rows, cols, *p, &pl);
  if (rc) return rc;
  if (g_attr_once.first()) set_func_attrs();
  // largest chunk the workspace can hold
  FbWorkspace ws;
  int chunk = n_pairs;
  while (chunk >= 1) {
    int frames = shared ? chunk + 1 : 2 * chunk;
    if (fb_layout(pl, chunk, frames, nullptr, 0, &ws) <= workspace_bytes) break;
    chunk = chunk > 1 ? (chunk + 1) / 2 : 0;
  }
  if (chunk < 1)
    return fail(B2OF_E_NOMEM, "farneback workspace too small: %zu B given, %zu B needed for one pair", workspace_bytes,
                fb_layout(pl, 1, 2, nullptr, 0, &ws));
  size_t flow_pair = (size_t)rows * cols * 2;
  for (int p0 = 0; p0 < n_pairs; p0 += chunk) {
    int np = n_pairs - p0 < chunk ? n_pairs - p0 : chunk;
    int frames = shared ? np + 1 : 2 * np;
    fb_layout(pl, np, frames, workspace, workspace_bytes, &ws);
    if (shared) {
      rc = fb_frames(pl, ws, prev + (size_t)p0 * frame_stride, step, frame_stride, frames, 0, 1, frames, st);
      if (rc) return rc;
    } else {
      // independent pairs: prev[i] -> slot 2i, next[i] -> slot 2i+1
      rc = fb_frames(pl, ws, prev + (size_t)p0 * frame_stride, step, frame_stride, np, 0, 2, frames, st);
      if (rc) return rc;
      rc = fb_frames(pl, ws, next + (size_t)p0 * frame_stride, step, frame_stride, np, 1, 2, frames, st);
      if (rc) return rc;
    }
    rc = fb_pairs(pl, ws, np, shared ? 1 : 2, frames, flow + (size_t)p0 * flow_pair,
                  stats ? stats + (size_t)p0 * B2OF_STATS_WIDTH : nullptr, *p, st);
    if (rc) return rc;
  }
  return B2OF_OK;
}

}  // namespace b2of
