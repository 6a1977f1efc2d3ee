// extern "C" surface of libb2of.so (see include/b2of.h) plus the host-buffer entry points that give the
// cv2-call contract (host arrays in, host arrays out).
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"

namespace b2of {

static thread_local char t_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
}

// ---- per-kernel timing ----
std::atomic<int> g_prof_on{0};
namespace {
struct ProfRec { int tag; cudaEvent_t e0, e1; double bytes; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
const char* const kProfNames[PT_COUNT] = {"bgr2gray", "pyrdown", "fb_level_hpass", "fb_level_vpass", "fb_polyexp",
                                          "fb_iter_finest", "fb_iter_coarse", "lk_scharr", "lk_track", "gftt_mineig",
                                          "gftt_nms_compact", "gftt_select", "pathfinder_filter", "flow_stats",
                                          "fb_upsample"};
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
void prof_mark(int tag, cudaStream_t st, bool end, double bytes) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!end) {
    ProfRec r{tag, prof_event(), prof_event(), bytes};
    cudaEventRecord(r.e0, st);
    g_prof_recs.push_back(r);
  } else {
    for (size_t i = g_prof_recs.size(); i-- > 0;)
      if (g_prof_recs[i].tag == tag) { cudaEventRecord(g_prof_recs[i].e1, st); break; }
  }
}

// implemented in the per-subsystem translation units
int bgr2gray_dev(const uint8_t*, int, int, size_t, size_t, uint8_t*, size_t, size_t, int, cudaStream_t);
int pyrdown_dev(const uint8_t*, int, int, size_t, size_t, uint8_t*, size_t, size_t, int, cudaStream_t);
size_t farneback_workspace_bytes(int, int, const b2of_farneback_params*, int, int);
int farneback_dev(const uint8_t*, const uint8_t*, size_t, size_t, int, int, int, int, const b2of_farneback_params*,
                  float*, float*, void*, size_t, cudaStream_t);
void farneback_release();
void farneback_release_streams();
size_t pyrlk_workspace_bytes(int, int, const b2of_lk_params*, int);
int pyrlk_dev(const uint8_t*, const uint8_t*, size_t, size_t, int, int, int, const float*, size_t, int, float*,
              uint8_t*, float*, const b2of_lk_params*, void*, size_t, cudaStream_t);
size_t gftt_workspace_bytes(int, int, const b2of_gftt_params*, int);
int gftt_dev(const uint8_t*, const uint8_t*, size_t, size_t, int, int, int, const b2of_gftt_params*, float*, int, int*,
             void*, size_t, cudaStream_t);
int pathfinder_filter_dev(const float*, size_t, const float*, int, int, int, int, int, int32_t*, int32_t*, uint8_t*,
                          uint8_t*, int32_t*, float*, int32_t*, int32_t*, cudaStream_t);
int overlay_vectors_dev(const int32_t*, const int32_t*, const uint8_t*, int, int, int, int, int, uint8_t*, cudaStream_t);
int overlay_lamps_dev(const int32_t*, const uint8_t*, const int32_t*, int, int, int, int, uint8_t*, cudaStream_t);
int flow_stats_dev(const float*, int, int, int, float*, cudaStream_t);
int flow_sample_dev(const float*, int, int, int, const float*, size_t, int, float*, cudaStream_t);
int flow_hsv_dev(const float*, int, int, int, uint8_t*, cudaStream_t);

// ----------------------------------------------------------------------------------------------
// host-call context: one per device, grow-only device buffers + non-blocking streams.
// Host calls serialise on the context mutex (cv2 callers are single-threaded scripts; several Python
// threads may call concurrently and simply queue).
// ----------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return B2OF_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = align_up(n + n / 8, 1 << 20);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(B2OF_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return B2OF_OK;
  }
};

struct HostCtx {
  std::mutex mu;
  cudaStream_t st[3] = {nullptr, nullptr, nullptr};  // compute, copy-in, copy-out
  cudaEvent_t ev[8] = {};
  DevBuf in[2], out[2], ws, aux[4];
  bool ready = false;
  int init() {
    if (ready) return B2OF_OK;
    for (auto& s : st) B2OF_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto& e : ev) B2OF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ready = true;
    return B2OF_OK;
  }
};

static std::mutex g_ctx_mu;
static HostCtx* g_ctx[64] = {};

static int get_ctx(HostCtx** out) {
  int dev = 0;
  B2OF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(B2OF_E_BADARG, "bad device %d", dev);
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  if (!g_ctx[dev]) g_ctx[dev] = new HostCtx();
  *out = g_ctx[dev];
  return B2OF_OK;
}

// copy `rows` rows of `width_bytes` between host and device honouring a host step
static int copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, int rows,
                  cudaMemcpyKind kind, cudaStream_t st) {
  if (dpitch == width_bytes && spitch == width_bytes) {
    B2OF_CUDA(cudaMemcpyAsync(dst, src, width_bytes * rows, kind, st));
  } else {
    B2OF_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, kind, st));
  }
  return B2OF_OK;
}

}  // namespace b2of

using namespace b2of;

extern "C" {

int b2of_version(void) { return B2OF_VERSION; }

// Frees what the library caches between calls: the host-call contexts (streams, events, grow-only device buffers)
// of every device and the Farneback plans (device tables).  The caller guarantees that no call is in flight; the
// next call rebuilds what it needs.
int b2of_release(void) {
  int cur = 0;
  cudaGetDevice(&cur);
  {
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    for (int d = 0; d < 64; ++d) {
      HostCtx* c = g_ctx[d];
      if (!c) continue;
      cudaSetDevice(d);
      {
        std::lock_guard<std::mutex> l2(c->mu);
        if (c->ready) {
          for (auto& s : c->st) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
          for (auto& e : c->ev) cudaEventDestroy(e);
        }
        for (DevBuf* b : {&c->in[0], &c->in[1], &c->out[0], &c->out[1], &c->ws, &c->aux[0], &c->aux[1], &c->aux[2],
                          &c->aux[3]})
          if (b->p) cudaFree(b->p);
      }
      delete c;
      g_ctx[d] = nullptr;
    }
  }
  cudaSetDevice(cur);
  farneback_release();
  farneback_release_streams();
  b2of_profile_reset();
  return B2OF_OK;
}
const char* b2of_last_error(void) { return t_err; }
unsigned long long b2of_launch_count(void) { return g_launches.load(); }

// ---- per-kernel timing (CUDA events on the launching stream) ----
void b2of_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
int b2of_profile_tag_count(void) { return PT_COUNT; }
const char* b2of_profile_tag_name(int tag) { return tag >= 0 && tag < PT_COUNT ? kProfNames[tag] : ""; }
void b2of_profile_reset(void) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
  g_prof_recs.clear();
}
int b2of_profile_read(int tag, double* ms_total, unsigned long long* launches, double* bytes_total) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  double ms = 0, bytes = 0;
  unsigned long long n = 0;
  for (auto& r : g_prof_recs) {
    if (r.tag != tag) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return fail(B2OF_E_CUDA, "profile event sync failed");
    float t = 0;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) return fail(B2OF_E_CUDA, "profile elapsed failed");
    ms += t; bytes += r.bytes; ++n;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = n;
  if (bytes_total) *bytes_total = bytes;
  return B2OF_OK;
}

// ---- K1 ----
int b2of_bgr2gray_u8_dev(const uint8_t* bgr, int rows, int cols, size_t src_step, size_t src_bstride, uint8_t* gray,
                         size_t dst_step, size_t dst_bstride, int batch, void* stream) {
  const char* fn = "cvtColor";
  B2OF_ASSERT(rows >= 0 && cols >= 0 && batch >= 0, fn);
  B2OF_ASSERT((bgr != nullptr && gray != nullptr) || (size_t)rows * cols * batch == 0, fn);
  B2OF_ASSERT(src_step >= (size_t)cols * 3 && dst_step >= (size_t)cols, fn);
  return bgr2gray_dev(bgr, rows, cols, src_step, src_bstride, gray, dst_step, dst_bstride, batch,
                      (cudaStream_t)stream);
}

int b2of_bgr2gray_u8_host(const uint8_t* bgr, int rows, int cols, size_t src_step, uint8_t* gray, size_t dst_step) {
  const char* fn = "cvtColor";
  B2OF_ASSERT(rows >= 0 && cols >= 0, fn);
  if ((size_t)rows * cols == 0) return B2OF_OK;
  B2OF_ASSERT(bgr != nullptr && gray != nullptr, fn);
  HostCtx* c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c->mu);
  if ((rc = c->init())) return rc;
  size_t n = (size_t)rows * cols;
  if ((rc = c->in[0].ensure(n * 3))) return rc;
  if ((rc = c->out[0].ensure(n))) return rc;
  cudaStream_t st = c->st[0];
  if ((rc = copy2d(c->in[0].p, (size_t)cols * 3, bgr, src_step, (size_t)cols * 3, rows, cudaMemcpyHostToDevice, st)))
    return rc;
  if ((rc = bgr2gray_dev((const uint8_t*)c->in[0].p, rows, cols, (size_t)cols * 3, 0, (uint8_t*)c->out[0].p, cols, 0, 1,
                         st)))
    return rc;
  if ((rc = copy2d(gray, dst_step, c->out[0].p, cols, cols, rows, cudaMemcpyDeviceToHost, st))) return rc;
  B2OF_CUDA(cudaStreamSynchronize(st));
  return B2OF_OK;
}

// ---- K2 ----
int b2of_pyrdown_u8_dev(const uint8_t* src, int rows, int cols, size_t src_step, size_t src_bstride, uint8_t* dst,
                        size_t dst_step, size_t dst_bstride, int batch, void* stream) {
  const char* fn = "pyrDown";
  B2OF_ASSERT(rows >= 0 && cols >= 0 && batch >= 0, fn);
  B2OF_ASSERT((src != nullptr && dst != nullptr) || (size_t)rows * cols * batch == 0, fn);
  B2OF_ASSERT(src_step >= (size_t)cols && dst_step >= (size_t)((cols + 1) / 2), fn);
  return pyrdown_dev(src, rows, cols, src_step, src_bstride, dst, dst_step, dst_bstride, batch, (cudaStream_t)stream);
}

int b2of_pyrdown_u8_host(const uint8_t* src, int rows, int cols, size_t src_step, uint8_t* dst, size_t dst_step) {
  const char* fn = "pyrDown";
  B2OF_ASSERT(rows > 0 && cols > 0, fn);
  B2OF_ASSERT(src != nullptr && dst != nullptr, fn);
  HostCtx* c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c->mu);
  if ((rc = c->init())) return rc;
  int dr = (rows + 1) / 2, dc = (cols + 1) / 2;
  size_t pitch = align_up(cols, 16), dpitch = align_up(dc, 16);
  if ((rc = c->in[0].ensure(pitch * rows))) return rc;
  if ((rc = c->out[0].ensure(dpitch * dr))) return rc;
  cudaStream_t st = c->st[0];
  B2OF_CUDA(cudaMemcpy2DAsync(c->in[0].p, pitch, src, src_step, cols, rows, cudaMemcpyHostToDevice, st));
  if ((rc = pyrdown_dev((const uint8_t*)c->in[0].p, rows, cols, pitch, 0, (uint8_t*)c->out[0].p, dpitch, 0, 1, st)))
    return rc;
  B2OF_CUDA(cudaMemcpy2DAsync(dst, dst_step, c->out[0].p, dpitch, dc, dr, cudaMemcpyDeviceToHost, st));
  B2OF_CUDA(cudaStreamSynchronize(st));
  return B2OF_OK;
}

// ---- K3-K6 ----
size_t b2of_farneback_workspace_bytes(int rows, int cols, const b2of_farneback_params* p, int chunk_pairs,
                                      int shared_frames) {
  if (chunk_pairs < 1) chunk_pairs = 1;
  return farneback_workspace_bytes(rows, cols, p, chunk_pairs, shared_frames);
}

int b2of_farneback_pairs_dev(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride, int n_pairs,
                             int rows, int cols, const b2of_farneback_params* p, float* flow, void* ws, size_t ws_bytes,
                             void* stream) {
  return farneback_dev(prev, next, step, frame_stride, n_pairs, 0, rows, cols, p, flow, nullptr, ws, ws_bytes,
                       (cudaStream_t)stream);
}

int b2of_farneback_sequence_dev(const uint8_t* frames, size_t step, size_t frame_stride, int n_frames, int rows,
                                int cols, const b2of_farneback_params* p, float* flow, void* ws, size_t ws_bytes,
                                void* stream) {
  if (n_frames < 2) return B2OF_OK;
  return farneback_dev(frames, frames + frame_stride, step, frame_stride, n_frames - 1, 1, rows, cols, p, flow, nullptr,
                       ws, ws_bytes, (cudaStream_t)stream);
}

int b2of_farneback_sequence_stats_dev(const uint8_t* frames, size_t step, size_t frame_stride, int n_frames, int rows,
                                      int cols, const b2of_farneback_params* p, float* flow, float* stats, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (n_frames < 2) return B2OF_OK;
  return farneback_dev(frames, frames + frame_stride, step, frame_stride, n_frames - 1, 1, rows, cols, p, flow, stats,
                       ws, ws_bytes, (cudaStream_t)stream);
}

// Pipelined host form shared by the pairs and the sequence entry points: H2D (st[1]) -> compute (st[0]) -> D2H
// (st[2]), `chunk` pairs per stage, double-buffered device input/output.
//   shared == 0: pair i = (prev + i*frame_stride, next + i*frame_stride)
//   shared == 1: frames at prev + i*frame_stride, i in [0, n_pairs]; pair i = (frame i, frame i+1)
static int farneback_host_pipeline(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride,
                                   int n_pairs, int shared, int rows, int cols, const b2of_farneback_params* p,
                                   float* flow) {
  const char* fn = "calcOpticalFlowFarneback";
  B2OF_ASSERT(prev != nullptr && (shared || next != nullptr) && flow != nullptr, fn);
  B2OF_ASSERT(rows > 0 && cols > 0 && step >= (size_t)cols, fn);
  if (n_pairs <= 0) return B2OF_OK;
  size_t one = b2of_farneback_workspace_bytes(rows, cols, p, 1, shared);
  if (one == 0) return B2OF_E_BADARG;  // message already set
  HostCtx* c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c->mu);
  if ((rc = c->init())) return rc;
  const size_t frame = (size_t)rows * cols, flow_pair = frame * 2 * sizeof(float);
  static const int host_chunk = [] { const char* e = getenv("B2OF_HOST_CHUNK"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : v; }();
  int chunk = n_pairs < host_chunk ? n_pairs : host_chunk;
  const int fpc = shared ? chunk + 1 : 2 * chunk;   // device frames per chunk
  size_t ws_bytes = b2of_farneback_workspace_bytes(rows, cols, p, chunk, shared);
  if ((rc = c->ws.ensure(ws_bytes))) return rc;
  for (int b = 0; b < 2; ++b) {
    if ((rc = c->in[b].ensure(frame * fpc))) return rc;
    if ((rc = c->out[b].ensure(flow_pair * chunk))) return rc;
  }
  cudaStream_t s_c = c->st[0], s_in = c->st[1], s_out = c->st[2];
  // events: ev[0+b] input b ready, ev[2+b] compute b done, ev[4+b] output b drained
  int n_chunks = (n_pairs + chunk - 1) / chunk;
  for (int ci = 0; ci < n_chunks; ++ci) {
    int b = ci & 1;
    int p0 = ci * chunk, np = n_pairs - p0 < chunk ? n_pairs - p0 : chunk;
    uint8_t* din = (uint8_t*)c->in[b].p;
    float* dout = (float*)c->out[b].p;
    if (ci >= 2) B2OF_CUDA(cudaStreamWaitEvent(s_in, c->ev[2 + b], 0));  // compute that read this input is done
    if (shared) {
      const uint8_t* src = prev + (size_t)p0 * frame_stride;
      if (step == (size_t)cols && frame_stride == frame) {
        B2OF_CUDA(cudaMemcpyAsync(din, src, frame * (np + 1), cudaMemcpyHostToDevice, s_in));
      } else {
        for (int i = 0; i <= np; ++i)
          if ((rc = copy2d(din + (size_t)i * frame, cols, src + (size_t)i * frame_stride, step, cols, rows,
                           cudaMemcpyHostToDevice, s_in)))
            return rc;
      }
    } else {
      for (int i = 0; i < np; ++i) {
        if ((rc = copy2d(din + (size_t)(2 * i) * frame, cols, prev + (size_t)(p0 + i) * frame_stride, step, cols,
                         rows, cudaMemcpyHostToDevice, s_in)))
          return rc;
        if ((rc = copy2d(din + (size_t)(2 * i + 1) * frame, cols, next + (size_t)(p0 + i) * frame_stride, step,
                         cols, rows, cudaMemcpyHostToDevice, s_in)))
          return rc;
      }
    }
    if (p->flags & B2OF_OPTFLOW_USE_INITIAL_FLOW) {
      // the caller's flow array is the initial estimate: it must be on the device before the pass, and the
      // previous D2H out of this slot must have drained
      if (ci >= 2) B2OF_CUDA(cudaStreamWaitEvent(s_in, c->ev[4 + b], 0));
      B2OF_CUDA(cudaMemcpyAsync(dout, flow + (size_t)p0 * frame * 2, flow_pair * np, cudaMemcpyHostToDevice, s_in));
    }
    B2OF_CUDA(cudaEventRecord(c->ev[b], s_in));
    B2OF_CUDA(cudaStreamWaitEvent(s_c, c->ev[b], 0));
    if (ci >= 2) B2OF_CUDA(cudaStreamWaitEvent(s_c, c->ev[4 + b], 0));  // previous output in this slot drained
    if (shared)
      rc = farneback_dev(din, din + frame, cols, frame, np, 1, rows, cols, p, dout, nullptr, c->ws.p, c->ws.cap, s_c);
    else
      rc = farneback_dev(din, din + frame, cols, 2 * frame, np, 0, rows, cols, p, dout, nullptr, c->ws.p, c->ws.cap,
                         s_c);
    if (rc) return rc;
    B2OF_CUDA(cudaEventRecord(c->ev[2 + b], s_c));
    B2OF_CUDA(cudaStreamWaitEvent(s_out, c->ev[2 + b], 0));
    B2OF_CUDA(cudaMemcpyAsync(flow + (size_t)p0 * frame * 2, dout, flow_pair * np, cudaMemcpyDeviceToHost, s_out));
    B2OF_CUDA(cudaEventRecord(c->ev[4 + b], s_out));
  }
  B2OF_CUDA(cudaStreamSynchronize(s_out));
  B2OF_CUDA(cudaStreamSynchronize(s_c));
  return B2OF_OK;
}

int b2of_farneback_pairs_host(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride, int n_pairs,
                              int rows, int cols, const b2of_farneback_params* p, float* flow) {
  return farneback_host_pipeline(prev, next, step, frame_stride, n_pairs, 0, rows, cols, p, flow);
}

int b2of_farneback_sequence_host(const uint8_t* frames, size_t step, size_t frame_stride, int n_frames, int rows,
                                 int cols, const b2of_farneback_params* p, float* flow) {
  if (n_frames < 2) return B2OF_OK;
  return farneback_host_pipeline(frames, nullptr, step, frame_stride, n_frames - 1, 1, rows, cols, p, flow);
}

int b2of_farneback_host(const uint8_t* prev, const uint8_t* next, size_t step, int rows, int cols,
                        const b2of_farneback_params* p, float* flow) {
  return b2of_farneback_pairs_host(prev, next, step, 0, 1, rows, cols, p, flow);
}

// ---- K10-K11 ----
size_t b2of_pyrlk_workspace_bytes(int rows, int cols, const b2of_lk_params* p, int batch) {
  return pyrlk_workspace_bytes(rows, cols, p, batch);
}

int b2of_pyrlk_dev(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride, int batch, int rows,
                   int cols, const float* prev_pts, size_t pts_batch_stride, int n_pts, float* next_pts,
                   uint8_t* status, float* err, const b2of_lk_params* p, void* ws, size_t ws_bytes, void* stream) {
  return pyrlk_dev(prev, next, step, frame_stride, batch, rows, cols, prev_pts, pts_batch_stride, n_pts, next_pts,
                   status, err, p, ws, ws_bytes, (cudaStream_t)stream);
}

int b2of_pyrlk_host(const uint8_t* prev, const uint8_t* next, size_t step, int rows, int cols, const float* prev_pts,
                    int n_pts, float* next_pts, uint8_t* status, float* err, const b2of_lk_params* p) {
  const char* fn = "calcOpticalFlowPyrLK";
  B2OF_ASSERT(prev != nullptr && next != nullptr, fn);
  B2OF_ASSERT(rows > 0 && cols > 0 && step >= (size_t)cols, fn);
  B2OF_ASSERT(n_pts >= 0, fn);
  if (n_pts == 0) return B2OF_OK;
  B2OF_ASSERT(prev_pts != nullptr && next_pts != nullptr && status != nullptr && err != nullptr && p != nullptr, fn);
  size_t ws_bytes = pyrlk_workspace_bytes(rows, cols, p, 1);
  if (ws_bytes == 0) return B2OF_E_BADARG;
  HostCtx* c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c->mu);
  if ((rc = c->init())) return rc;
  size_t pitch = align_up(cols, 16), frame = pitch * rows;
  if ((rc = c->in[0].ensure(2 * frame))) return rc;
  if ((rc = c->ws.ensure(ws_bytes))) return rc;
  size_t pts_bytes = (size_t)n_pts * 2 * sizeof(float);
  if ((rc = c->aux[0].ensure(pts_bytes))) return rc;       // prev pts
  if ((rc = c->aux[1].ensure(pts_bytes))) return rc;       // next pts
  if ((rc = c->aux[2].ensure((size_t)n_pts))) return rc;   // status
  if ((rc = c->aux[3].ensure((size_t)n_pts * 4))) return rc;  // err
  cudaStream_t st = c->st[0];
  uint8_t* din = (uint8_t*)c->in[0].p;
  B2OF_CUDA(cudaMemcpy2DAsync(din, pitch, prev, step, cols, rows, cudaMemcpyHostToDevice, st));
  B2OF_CUDA(cudaMemcpy2DAsync(din + frame, pitch, next, step, cols, rows, cudaMemcpyHostToDevice, st));
  B2OF_CUDA(cudaMemcpyAsync(c->aux[0].p, prev_pts, pts_bytes, cudaMemcpyHostToDevice, st));
  if (p->flags & B2OF_OPTFLOW_USE_INITIAL_FLOW)
    B2OF_CUDA(cudaMemcpyAsync(c->aux[1].p, next_pts, pts_bytes, cudaMemcpyHostToDevice, st));
  if ((rc = pyrlk_dev(din, din + frame, pitch, 0, 1, rows, cols, (const float*)c->aux[0].p, 0, n_pts,
                      (float*)c->aux[1].p, (uint8_t*)c->aux[2].p, (float*)c->aux[3].p, p, c->ws.p, c->ws.cap, st)))
    return rc;
  B2OF_CUDA(cudaMemcpyAsync(next_pts, c->aux[1].p, pts_bytes, cudaMemcpyDeviceToHost, st));
  B2OF_CUDA(cudaMemcpyAsync(status, c->aux[2].p, (size_t)n_pts, cudaMemcpyDeviceToHost, st));
  B2OF_CUDA(cudaMemcpyAsync(err, c->aux[3].p, (size_t)n_pts * 4, cudaMemcpyDeviceToHost, st));
  B2OF_CUDA(cudaStreamSynchronize(st));
  return B2OF_OK;
}

// ---- K7-K9 ----
size_t b2of_gftt_workspace_bytes(int rows, int cols, const b2of_gftt_params* p, int batch) {
  return gftt_workspace_bytes(rows, cols, p, batch);
}

int b2of_gftt_dev(const uint8_t* img, const uint8_t* mask, size_t step, size_t frame_stride, int batch, int rows,
                  int cols, const b2of_gftt_params* p, float* corners, int cap, int* n_corners, void* ws,
                  size_t ws_bytes, void* stream) {
  return gftt_dev(img, mask, step, frame_stride, batch, rows, cols, p, corners, cap, n_corners, ws, ws_bytes,
                  (cudaStream_t)stream);
}

int b2of_gftt_host(const uint8_t* img, const uint8_t* mask, size_t step, size_t mask_step, int rows, int cols,
                   const b2of_gftt_params* p, float* corners, int cap, int* n_corners) {
  const char* fn = "goodFeaturesToTrack";
  B2OF_ASSERT(img != nullptr && p != nullptr && n_corners != nullptr, fn);
  B2OF_ASSERT(rows > 0 && cols > 0 && step >= (size_t)cols, fn);
  B2OF_ASSERT(cap >= 0 && (corners != nullptr || cap == 0), fn);
  size_t ws_bytes = gftt_workspace_bytes(rows, cols, p, 1);
  if (ws_bytes == 0) return B2OF_E_BADARG;
  HostCtx* c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(c->mu);
  if ((rc = c->init())) return rc;
  size_t pitch = align_up(cols, 16), frame = pitch * rows;
  if ((rc = c->in[0].ensure(2 * frame))) return rc;
  if ((rc = c->ws.ensure(ws_bytes))) return rc;
  if ((rc = c->aux[0].ensure((size_t)cap * 8 + 16))) return rc;
  if ((rc = c->aux[1].ensure(16))) return rc;
  cudaStream_t st = c->st[0];
  uint8_t* din = (uint8_t*)c->in[0].p;
  B2OF_CUDA(cudaMemcpy2DAsync(din, pitch, img, step, cols, rows, cudaMemcpyHostToDevice, st));
  if (mask) B2OF_CUDA(cudaMemcpy2DAsync(din + frame, pitch, mask, mask_step, cols, rows, cudaMemcpyHostToDevice, st));
  if ((rc = gftt_dev(din, mask ? din + frame : nullptr, pitch, 0, 1, rows, cols, p, (float*)c->aux[0].p, cap,
                     (int*)c->aux[1].p, c->ws.p, c->ws.cap, st)))
    return rc;
  B2OF_CUDA(cudaMemcpyAsync(n_corners, c->aux[1].p, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2OF_CUDA(cudaStreamSynchronize(st));
  int n = *n_corners;
  if (n > cap) n = cap;
  if (n > 0) B2OF_CUDA(cudaMemcpy(corners, c->aux[0].p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return B2OF_OK;
}

// ---- K12 ----
int b2of_pathfinder_filter_dev(const float* pts, size_t pts_batch_stride, const float* next_pts, int n_pts, int batch,
                               int width, int height, int mode, int32_t* kept_pts, int32_t* kept_flow,
                               uint8_t* danger_v, uint8_t* mask, int32_t* n_kept, float* stats, int32_t* all_pts,
                               int32_t* all_next, void* stream) {
  return pathfinder_filter_dev(pts, pts_batch_stride, next_pts, n_pts, batch, width, height, mode, kept_pts, kept_flow,
                               danger_v, mask, n_kept, stats, all_pts, all_next, (cudaStream_t)stream);
}

int b2of_overlay_vectors_dev(const int32_t* all_pts, const int32_t* all_next, const uint8_t* mask, int n_pts, int batch,
                             int rows, int cols, int draw_bad, uint8_t* layer_bgr, void* stream) {
  return overlay_vectors_dev(all_pts, all_next, mask, n_pts, batch, rows, cols, draw_bad, layer_bgr,
                             (cudaStream_t)stream);
}

int b2of_overlay_lamps_dev(const int32_t* kept_pts, const uint8_t* danger_v, const int32_t* n_kept, int n_pts, int batch,
                           int rows, int cols, uint8_t* bgr, void* stream) {
  return overlay_lamps_dev(kept_pts, danger_v, n_kept, n_pts, batch, rows, cols, bgr, (cudaStream_t)stream);
}

int b2of_flow_sample_dev(const float* flow, int n_pairs, int rows, int cols, const float* pts, size_t pts_batch_stride,
                         int n_pts, float* next_pts, void* stream) {
  return flow_sample_dev(flow, n_pairs, rows, cols, pts, pts_batch_stride, n_pts, next_pts, (cudaStream_t)stream);
}

int b2of_flow_hsv_dev(const float* flow, int n_pairs, int rows, int cols, uint8_t* bgr, void* stream) {
  return flow_hsv_dev(flow, n_pairs, rows, cols, bgr, (cudaStream_t)stream);
}

int b2of_flow_stats_dev(const float* flow, int n_pairs, int rows, int cols, float* stats, void* stream) {
  return flow_stats_dev(flow, n_pairs, rows, cols, stats, (cudaStream_t)stream);
}

}  // extern "C"
