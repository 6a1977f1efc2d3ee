/*
 * b2of.h -- C-ABI of libb2of.so, the B200 (sm_100a) optical-flow engine.
 *
 * This is the drop-in boundary for the per-frame-pair hot path of
 * spirinis/HackathonOpticalFlow.  The reference defines no plugin API of its own:
 * the path sits behind four `cv2` Python calls, so every entry point below names
 * the reference call site (file:line) whose arithmetic it replaces.  Shorthands:
 *   viewer.py   = /root/reference/pathfinder_viewer.py
 *   DenseOF.py  = /root/reference/<dev dir>/DenseOF.py
 *   SparseOF.py = /root/reference/<dev dir>/SparseOF.py
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy / cv types.
 *   - `*_dev` pointers are CUDA device pointers on the current device; `stream`
 *     is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - `*_host` entry points take ordinary host memory, do their own H2D/D2H on an
 *     internal stream and return after the result is in the caller's buffer --
 *     exactly the contract of the cv2 call they replace.
 *   - return value: 0 on success, negative B2OF_E_* on failure; the message is
 *     available from b2of_last_error() (thread-local).  Argument failures mirror
 *     the cv2 assertion the reference would have hit (cv2.error -215).
 *   - images are 8-bit single channel, row-major, `step` bytes between rows.
 *   - dense flow is float32 (rows, cols, 2) interleaved (dx, dy), C-contiguous.
 *   - there is NO CPU fallback anywhere behind this header.
 */
#ifndef B2OF_H
#define B2OF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2OF_VERSION 102

#define B2OF_OK 0
#define B2OF_E_BADARG (-215) /* cv2's StsAssert code: same meaning */
#define B2OF_E_CUDA (-1)
#define B2OF_E_NOMEM (-4)
#define B2OF_E_UNSUPPORTED (-213) /* cv2's StsNotImplemented */

/* cv2 flag values (cv2.OPTFLOW_*), kept numerically identical */
#define B2OF_OPTFLOW_USE_INITIAL_FLOW 4
#define B2OF_OPTFLOW_LK_GET_MIN_EIGENVALS 8
#define B2OF_OPTFLOW_FARNEBACK_GAUSSIAN 256
/* library extension (no cv2 counterpart; a bit cv2 does not use): the 15x15 box sums of the fused iteration kernel take
   their horizontal pass per block of 15 instead of as a sliding sum and carry their vertical running sums in double at
   EVERY pyramid level -- each output is then the rounded sum of its own inputs only, so near-singular pixels next to
   bright texture land closer to cv2.  Without the flag the two coarsest levels already do (that is where it decides the
   result: on real footage the pixels beyond 0.5 px drop from 0.15 / 0.23 % to 0.01 / 0.15 % on the two hard clips, where
   cv2 differs from itself on 0 / 0.04 %; 2 % fewer pairs/s) and the finer levels use float sliding sums; with the flag,
   14 % fewer pairs/s and no measurable further change -- a diagnostic, not a recommendation.  Ignored by the general
   kernels (other window sizes, Gaussian window), whose sums are direct. */
#define B2OF_FARNEBACK_BLOCKED_SUMS 0x10000
#define B2OF_TERM_COUNT 1
#define B2OF_TERM_EPS 2

int b2of_version(void);
const char* b2of_last_error(void);
/* frees everything the library caches between calls (host-call streams / events / staging buffers of every device,
 * Farneback plans and their device tables).  No call may be in flight; later calls rebuild what they need. */
int b2of_release(void);
/* number of kernel launches issued by this library in this process (all threads) */
unsigned long long b2of_launch_count(void);

/* Optional per-kernel timing: when enabled, every kernel launch of the tagged families is bracketed by CUDA events
 * on the stream it is launched on.  b2of_profile_read synchronises those events and returns the summed device
 * time, launch count and algorithmic bytes (DESIGN.md per-kernel figures) recorded for one tag since the last reset. */
void b2of_profile_enable(int on);
void b2of_profile_reset(void);
int b2of_profile_tag_count(void);
const char* b2of_profile_tag_name(int tag);
int b2of_profile_read(int tag, double* ms_total, unsigned long long* launches, double* bytes_total);

/* ---- K1: cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) --------------------------------
 * replaces viewer.py:244, :280; DenseOF.py:481, :510; SparseOF.py:28.
 * `batch` images, `src_batch_stride`/`dst_batch_stride` bytes apart. Bit-exact. */
int b2of_bgr2gray_u8_dev(const uint8_t* bgr_dev, int rows, int cols, size_t src_step, size_t src_batch_stride,
                         uint8_t* gray_dev, size_t dst_step, size_t dst_batch_stride, int batch, void* stream);
int b2of_bgr2gray_u8_host(const uint8_t* bgr, int rows, int cols, size_t src_step, uint8_t* gray, size_t dst_step);

/* ---- K2: cv2.pyrDown on 8-bit gray (the chain buildOpticalFlowPyramid runs inside
 * cv2.calcOpticalFlowPyrLK; viewer.py:156-158, SparseOF.py:35-36). Bit-exact.
 * dst is ((rows+1)/2, (cols+1)/2). */
int b2of_pyrdown_u8_dev(const uint8_t* src_dev, int rows, int cols, size_t src_step, size_t src_batch_stride,
                        uint8_t* dst_dev, size_t dst_step, size_t dst_batch_stride, int batch, void* stream);
int b2of_pyrdown_u8_host(const uint8_t* src, int rows, int cols, size_t src_step, uint8_t* dst, size_t dst_step);

/* ---- K3-K6: cv2.calcOpticalFlowFarneback ---------------------------------------
 * replaces DenseOF.py:147-156 (wrapper calculate_optical_flow DenseOF.py:127-157,
 * call site :520; reference parameters 0.5, 3, 15, 3, 5, 1.2, 0). */
typedef struct b2of_farneback_params {
  double pyr_scale;
  int levels;
  int winsize;
  int iterations;
  int poly_n;
  double poly_sigma;
  int flags;
} b2of_farneback_params;

/* bytes of device scratch needed to process `chunk_pairs` pairs at once.
 * shared_frames != 0: the pairs are consecutive frames of one sequence, per-frame
 * work (level images, polynomial expansion) is done once per frame. */
size_t b2of_farneback_workspace_bytes(int rows, int cols, const b2of_farneback_params* p, int chunk_pairs,
                                      int shared_frames);

/* n_pairs independent pairs: pair i is (prev_dev + i*frame_stride, next_dev + i*frame_stride);
 * flow_dev + i*rows*cols*2 receives its flow.  The workspace decides how many pairs
 * run per pass (>= 1 pair's worth is required). */
int b2of_farneback_pairs_dev(const uint8_t* prev_dev, const uint8_t* next_dev, size_t step, size_t frame_stride,
                             int n_pairs, int rows, int cols, const b2of_farneback_params* p, float* flow_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream);

/* n_frames consecutive frames -> n_frames-1 flows (frame i -> frame i+1). */
int b2of_farneback_sequence_dev(const uint8_t* frames_dev, size_t step, size_t frame_stride, int n_frames, int rows,
                                int cols, const b2of_farneback_params* p, float* flow_dev, void* workspace_dev,
                                size_t workspace_bytes, void* stream);

/* b2of_farneback_sequence_dev plus the per-pair flow statistics of b2of_flow_stats_dev (float32 (n_frames-1, 8),
 * 8-byte aligned), reduced inside the last iteration kernel instead of by a second pass over the flow fields.
 * stats_dev may be NULL. */
int b2of_farneback_sequence_stats_dev(const uint8_t* frames_dev, size_t step, size_t frame_stride, int n_frames,
                                      int rows, int cols, const b2of_farneback_params* p, float* flow_dev,
                                      float* stats_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* the cv2 call itself: host prev/next in, host flow out (flow must not be NULL; the
 * Python shim allocates it when the caller passed None, as cv2 does). */
int b2of_farneback_host(const uint8_t* prev, const uint8_t* next, size_t step, int rows, int cols,
                        const b2of_farneback_params* p, float* flow);
/* pipelined host form: n_pairs pairs from host memory, copies overlapped with compute */
int b2of_farneback_pairs_host(const uint8_t* prev, const uint8_t* next, size_t step, size_t frame_stride, int n_pairs,
                              int rows, int cols, const b2of_farneback_params* p, float* flow);

/* pipelined host form for a video: n_frames consecutive host frames -> n_frames-1 host flows */
int b2of_farneback_sequence_host(const uint8_t* frames, size_t step, size_t frame_stride, int n_frames, int rows,
                                 int cols, const b2of_farneback_params* p, float* flow);

/* ---- K10-K11: cv2.calcOpticalFlowPyrLK ------------------------------------------
 * replaces viewer.py:156-158 / DenseOF.py:183-185 (45x45 grid form, prev = current
 * frame) and SparseOF.py:35-36 (15x15 track form, forward + backward). */
typedef struct b2of_lk_params {
  int win_w, win_h;
  int max_level;
  int crit_type; /* B2OF_TERM_* bits */
  int crit_max_count;
  double crit_eps;
  int flags;
  double min_eig_threshold;
} b2of_lk_params;

size_t b2of_pyrlk_workspace_bytes(int rows, int cols, const b2of_lk_params* p, int batch);
/* batch image pairs, each tracking the same number of points.
 * prev_pts/next_pts: float32 (batch, n_pts, 2); status uint8 (batch, n_pts); err float32 (batch, n_pts).
 * next_pts is read only when flags has USE_INITIAL_FLOW. pts_batch_stride (in points) may be 0
 * to track one shared point set (the viewer's grid) in every pair. */
int b2of_pyrlk_dev(const uint8_t* prev_dev, const uint8_t* next_dev, size_t step, size_t frame_stride, int batch,
                   int rows, int cols, const float* prev_pts_dev, size_t pts_batch_stride, int n_pts,
                   float* next_pts_dev, uint8_t* status_dev, float* err_dev, const b2of_lk_params* p,
                   void* workspace_dev, size_t workspace_bytes, void* stream);
int b2of_pyrlk_host(const uint8_t* prev, const uint8_t* next, size_t step, int rows, int cols, const float* prev_pts,
                    int n_pts, float* next_pts, uint8_t* status, float* err, const b2of_lk_params* p);

/* ---- K7-K9: cv2.goodFeaturesToTrack ---------------------------------------------
 * replaces SparseOF.py:69 (feature_params SparseOF.py:10-13, mask SparseOF.py:61-66). */
typedef struct b2of_gftt_params {
  int max_corners;
  double quality_level;
  double min_distance;
  int block_size;
  int gradient_size; /* 3 only */
  int use_harris;
  double k;
} b2of_gftt_params;

size_t b2of_gftt_workspace_bytes(int rows, int cols, const b2of_gftt_params* p, int batch);
/* corners_dev: float32 (batch, corners_cap, 2) (x,y); n_corners_dev: int32 (batch). mask may be NULL. */
int b2of_gftt_dev(const uint8_t* img_dev, const uint8_t* mask_dev, size_t step, size_t frame_stride, int batch,
                  int rows, int cols, const b2of_gftt_params* p, float* corners_dev, int corners_cap,
                  int* n_corners_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
int b2of_gftt_host(const uint8_t* img, const uint8_t* mask, size_t step, size_t mask_step, int rows, int cols,
                   const b2of_gftt_params* p, float* corners, int corners_cap, int* n_corners);

/* ---- K12: the reference's own downstream logic, per frame, on the device ----------
 * vector filter viewer.py:159-178 and danger intensity viewer.py:210-217.
 * For each of `batch` frames: inputs prev points (shared grid if pts_batch_stride==0)
 * and LK next points; outputs, compacted in input order:
 *   kept_pts int32 (batch, n_pts, 2), kept_flow int32 (batch, n_pts, 2),
 *   danger_v uint8 (batch, n_pts), mask uint8 (batch, n_pts), n_kept int32 (batch),
 *   stats float32 (batch, 8): mean|flow|, max|flow|, mean dx, mean dy,
 *                             median modulus, p99 modulus, n_kept, sum danger V. */
#define B2OF_STATS_WIDTH 8
/* mode: which of the reference's two mask rules is applied to the normalised moduli m
 *   B2OF_FILTER_VIEWER   median(m) < m < percentile(m, 99)      viewer.py:171
 *   B2OF_FILTER_DENSEOF  m > 1.2 * median(m)                    DenseOF.py:228 */
#define B2OF_FILTER_VIEWER 0
#define B2OF_FILTER_DENSEOF 1
/* all_pts_dev / all_next_dev (both NULL or both given): int32 (batch, n_pts, 2), EVERY point and its normalised end
 * point as the reference rounds them (viewer.py:169-170) -- what the overlay (b2of_overlay_vectors_dev) draws. */
int b2of_pathfinder_filter_dev(const float* pts_dev, size_t pts_batch_stride, const float* next_pts_dev, int n_pts,
                               int batch, int width, int height, int mode, int32_t* kept_pts_dev,
                               int32_t* kept_flow_dev, uint8_t* danger_v_dev, uint8_t* mask_dev, int32_t* n_kept_dev,
                               float* stats_dev, int32_t* all_pts_dev, int32_t* all_next_dev, void* stream);

/* ---- overlay layers, composited on the device (the reference's drawing calls) ------------------------
 * vector layer, viewer.py:179-191: cv2.polylines of the kept vectors (mask == 1) in BGR (0,0,255), cv2.circle of
 * radius 1 in (255,0,255) at their start points, then -- draw_bad != 0, the reference's draw_bad_flow -- the rejected
 * vectors and their start points in (255,255,0).  layer_bgr_dev: uint8 (batch, rows, cols, 3), cleared first.
 * Pixel for pixel what cv2 draws (clipLine + 8-connected LineIterator, midpoint circle). */
int b2of_overlay_vectors_dev(const int32_t* all_pts_dev, const int32_t* all_next_dev, const uint8_t* mask_dev, int n_pts,
                             int batch, int rows, int cols, int draw_bad, uint8_t* layer_bgr_dev, void* stream);
/* lamp layer, viewer.py:210-222 (draw_sparse_lamps): HSV (0,255,V) at every kept point converted as
 * cv2.cvtColor(HSV2BGR) does, then a filled cv2.circle of radius 6 in the point's own colour.
 * kept_pts_dev / danger_v_dev / n_kept_dev are b2of_pathfinder_filter_dev's outputs. */
int b2of_overlay_lamps_dev(const int32_t* kept_pts_dev, const uint8_t* danger_v_dev, const int32_t* n_kept_dev,
                           int n_pts, int batch, int rows, int cols, uint8_t* bgr_dev, void* stream);

/* dense flow sampled on a point set (the grid): next_pts[b][i] = pts[i] + flow[b][int(y_i)][int(x_i)], float32
 * (batch, n_pts, 2) -- feeds b2of_pathfinder_filter_dev with the dense field instead of LK (what draw_flow samples,
 * DenseOF.py:44-50).  pts_batch_stride (in points) may be 0 for one shared grid. */
int b2of_flow_sample_dev(const float* flow_dev, int n_pairs, int rows, int cols, const float* pts_dev,
                         size_t pts_batch_stride, int n_pts, float* next_pts_dev, void* stream);

/* dense flow as a picture -- draw_hsv, pathfinder_viewer.py:124-141 (DenseOF.py:113-121): hue = direction,
 * value = 4 x length, converted as cv2.cvtColor(hsv, COLOR_HSV2BGR) does; uint8 (n_pairs, rows, cols, 3) BGR */
int b2of_flow_hsv_dev(const float* flow_dev, int n_pairs, int rows, int cols, uint8_t* bgr_dev, void* stream);

/* dense-flow statistics (what draw_flow / draw_hsv consume, DenseOF.py:44-50, :113-121):
 * per pair float32[8]: mean|flow|, max|flow|, mean dx, mean dy, 0, 0, 0, 0 */
int b2of_flow_stats_dev(const float* flow_dev, int n_pairs, int rows, int cols, float* stats_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2OF_H */
