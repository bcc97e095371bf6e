/* uvip_orb.h — C ABI of the B200-native ORB front-end (libuvip_orb.so).
 *
 * This is the drop-in boundary for the one hot path of chintha/U-VIP-SLAM that this repository
 * replaces: USLAM::ORBextractor and the descriptor path of USLAM::ORBmatcher.  The reference has no
 * FFI of its own (it is one C++ process); the entry points below are what its two classes would bind
 * if their bodies were moved behind a C boundary, and u-vip-slam_b200/host/ORBextractor.h / ORBmatcher.h
 * are those classes re-written as thin forwarders (INTEGRATION.md shows the patch a maintainer applies).
 * Each entry point cites the reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns an int status
 * (UVIP_OK = 0), never throws, never falls back to a CPU implementation.  "host" pointers are ordinary
 * process memory; "device" pointers (suffix _device) are CUDA device memory on the handle's GPU and the
 * call is asynchronous on the given cudaStream_t (passed as void*).  NULL selects the handle's OWN non-blocking stream,
 * which is not ordered against any stream of the caller: order later work behind it with uvip_extractor_stream() /
 * uvip_matcher_stream() (an event on that stream) or wait with uvip_extractor_status() / uvip_matcher_sync().  NULL is
 * NOT the legacy default stream; pass cudaStreamLegacy ((void*)0x1) for that — what frameworks whose "default stream"
 * handle is 0 (torch) must do.
 * A handle is not re-entrant (the reference extractor is stateful too, include/ORBextractor.h:90-91);
 * use one handle per calling thread — Tracking, LocalMapping and LoopClosing each own one matcher.
 */
#ifndef UVIP_ORB_H
#define UVIP_ORB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVIP_ABI_VERSION 1

enum {
    UVIP_OK = 0,
    UVIP_ERR_ARG = -1,          /* bad pointer / size / parameter */
    UVIP_ERR_CAPACITY = -2,     /* caller buffer or internal candidate list too small (nothing is truncated silently) */
    UVIP_ERR_CUDA = -3,         /* CUDA runtime error; uvip_last_error() has the text */
    UVIP_ERR_UNSUPPORTED = -4,  /* shape outside the supported envelope (see uvip_extractor_create) */
    UVIP_ERR_NO_DEVICE = -5     /* no usable CUDA device: there is no CPU fallback */
};

/* cv::KeyPoint layout, 28 bytes (OpenCV core/types.hpp; used at include/ORBextractor.h:56-58) */
typedef struct uvip_keypoint {
    float   x, y;       /* pt, level-0 pixels */
    float   size;       /* (float)(int)(31 * scale[octave])     src/ORBextractor.cc:820,829 */
    float   angle;      /* IC_Angle, degrees [0,360)             src/ORBextractor.cc:125-152 */
    float   response;   /* FAST score                            src/ORBextractor.cc:792-799 */
    int32_t octave;
    int32_t class_id;   /* -1 */
} uvip_keypoint;

/* ---- extractor: replaces USLAM::ORBextractor (include/ORBextractor.h:45-94) ------------------------------- */
typedef struct uvip_extractor uvip_extractor;

typedef struct uvip_extractor_params {
    /* the five constructor arguments, include/ORBextractor.h:51 */
    int32_t nfeatures;
    float   scale_factor;
    int32_t nlevels;
    int32_t score_type;     /* 0 HARRIS_SCORE, 1 FAST_SCORE — accepted and ignored like the reference's live path */
    int32_t fast_th;
    /* compile-time constants of the reference, exposed so tests can vary them (0 = reference value) */
    int32_t retry_th;       /* 7   src/ORBextractor.cc:798 */
    int32_t cell;           /* 30  src/ORBextractor.cc:752 */
    /* capacity of the device working set */
    int32_t device;         /* CUDA device ordinal */
    int32_t max_width, max_height;   /* largest frame; levels smaller than 64x64 are unsupported */
    int32_t max_batch;      /* frames resident per launch group (>=1) */
} uvip_extractor_params;

/* ORBextractor::ORBextractor (src/ORBextractor.cc:458-512): builds the scale / quota / umax tables, uploads
 * the 512-point pattern, allocates the pyramid working set for max_batch frames. */
int   uvip_extractor_create(const uvip_extractor_params* params, uvip_extractor** out);
int   uvip_extractor_destroy(uvip_extractor* ex);
/* ORBextractor::GetLevels / GetScaleFactor (include/ORBextractor.h:60-64) */
int   uvip_extractor_levels(const uvip_extractor* ex);
float uvip_extractor_scale_factor(const uvip_extractor* ex);
/* constructor tables for inspection: scale[nlevels], inv_scale[nlevels], quota[nlevels], umax[16] (any may be NULL) */
int   uvip_extractor_tables(const uvip_extractor* ex, float* scale, float* inv_scale, int32_t* quota, int32_t* umax);

/* ORBextractor::operator() (include/ORBextractor.h:56-58, src/ORBextractor.cc:849-961), host buffers.
 *   image        CV_8UC1 rows of `stride` bytes; NULL or w/h <= 0 = empty image: returns UVIP_OK, outputs untouched
 *   (mask)       the reference ignores it; it has no parameter here
 *   kps,n_inout  in: *n_inout incoming level-0 keypoints (kept only when full_detect == 0); out: cleared and replaced
 *   desc         cap x 32 bytes, row i belongs to kps[i]
 *   grid         Eigen::MatrixXi storage (column-major int32, grid_rows x grid_cols); read and incremented when
 *                full_detect == 0, untouched otherwise; may be NULL when full_detect != 0
 *   min_px_dist, full_detect, num_needed   as in the reference signature */
int   uvip_extract(uvip_extractor* ex, const uint8_t* image, int w, int h, int stride,
                   uvip_keypoint* kps, int* n_inout, int cap, uint8_t* desc,
                   int32_t* grid, int grid_rows, int grid_cols, int min_px_dist,
                   int full_detect, int num_needed);

/* The same operator() over a batch of equally sized frames with FullDetect=true semantics per frame.
 * frames: nframes images, frame f starts at frames + f*frame_pitch.  Outputs per frame f:
 * n_out[f] keypoints at kps + f*cap and desc + f*cap*32.  Host buffers (copies are inside the call). */
int   uvip_extract_batch(uvip_extractor* ex, const uint8_t* frames, int nframes, int w, int h, int stride,
                         size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc);
/* The two halves of uvip_extract_batch for callers that stream a sequence: submit enqueues the copies and kernels of one
 * batch and returns a ticket (0 or 1) without waiting; wait blocks until that batch's outputs are in the host buffers
 * and returns its status.  At most two batches are in flight per handle, so the H2D copy of batch i+1 overlaps the
 * kernels of batch i and the D2H copy of batch i-1.  The host buffers must stay valid (and should be pinned) until the
 * wait returns; uvip_extract on the same handle is refused while a ticket is outstanding. */
int   uvip_extract_batch_submit(uvip_extractor* ex, const uint8_t* frames, int nframes, int w, int h, int stride,
                                size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc, int* ticket);
int   uvip_extract_batch_wait(uvip_extractor* ex, int ticket);
/* uvip_extract_batch_submit followed, ON THE DEVICE, by the brute-force kNN2 of consecutive frames (frame f = queries, frame f+1 =
 * train; include/utils.h:81-111 semantics, see uvip_knn2): descriptors are matched where the extractor wrote them and never make
 * the host round trip that uvip_extract_batch_wait + uvip_knn2_batch would force.  knn_idx / knn_dist: (nframes - 1) x cap x 2
 * int32 host buffers, pair f at + f*cap*2, query i of the pair at + 2*i; rows >= n_out[f] are unspecified.  `m` supplies the kNN
 * scratch and must not be used by another thread until the ticket has been waited for (uvip_extract_batch_wait). */
struct uvip_matcher;
int   uvip_extract_match_batch_submit(uvip_extractor* ex, struct uvip_matcher* m, const uint8_t* frames, int nframes, int w, int h,
                                      int stride, size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc,
                                      int32_t* knn_idx, int32_t* knn_dist, int* ticket);
/* Device-resident variant: all pointers are device memory; asynchronous on `stream`.
 * uvip_extractor_status() after synchronising reports capacity overflow of the launch group. */
int   uvip_extract_batch_device(uvip_extractor* ex, const uint8_t* d_frames, int nframes, int w, int h, int stride,
                                size_t frame_pitch, uvip_keypoint* d_kps, int32_t* d_n_out, int cap, uint8_t* d_desc,
                                void* stream);
/* frames > 0: uvip_extract_batch_device enqueues a launch group as sub-batches of that many frames, each run through ALL stages before
 * the next starts, so that a sub-batch's pyramid (1.33 MB per 752x480 frame and plane set) is still in the 126 MB L2 when FAST, the
 * blur and the descriptor stage read it back: less DRAM traffic, more (smaller) launches.  0 = one launch per stage for the whole group
 * (default).  Results are identical.  The debug taps address the frames of the LAST sub-batch. */
int   uvip_extractor_set_subbatch(uvip_extractor* ex, int frames);
int   uvip_extractor_status(uvip_extractor* ex);   /* synchronises the handle; UVIP_OK or the sticky error of the last group */
void* uvip_extractor_stream(uvip_extractor* ex);   /* the handle's own cudaStream_t (what a NULL stream argument selects) */
/* how often uvip_extract had to capture + instantiate its CUDA graph (once per call SHAPE: frame geometry, cap, FullDetect,
 * grid geometry — not per num_featsneeded / number of incoming keypoints, which change every frame at src/Tracking.cc:946) */
long long uvip_extractor_graph_captures(const uvip_extractor* ex);

/* debug taps for parity tests (valid after an extract call, frame < nframes of that call) */
int   uvip_get_pyramid_level(uvip_extractor* ex, int frame, int level, int blurred,
                             uint8_t* dst, int dstride, int* w, int* h);      /* interior, src/ORBextractor.cc:963-1004 / :942 */
int   uvip_get_raw_corners(uvip_extractor* ex, int frame, int level,
                           int32_t* xs, int32_t* ys, int32_t* scores, int cap, int* n); /* per-cell FAST output, window coords,
                                                                                reference order (src/ORBextractor.cc:772-812) */
int   uvip_get_level_keypoints(uvip_extractor* ex, int frame, int level,
                               int32_t* xs, int32_t* ys, int32_t* scores, int cap, int* n); /* quadtree winners, level coords,
                                                                                list order (src/ORBextractor.cc:817-831) */
/* HarrisResponses (src/ORBextractor.cc:80-121: blockSize 7, k = 0.04 at its only call site :663) at n points (level
 * coordinates) of the UNBLURRED pyramid level of a frame of the last extract call.  This is the scoring half of the reference's
 * dead detector path ComputeKeyPoints (:536-746); the quota-cell distribution around it is not built. */
int   uvip_harris_responses(uvip_extractor* ex, int frame, int level, const float* xs, const float* ys, int n, int block_size,
                            float harris_k, float* out);
/* OPTIONAL mode: the reference's dead detector path ORBextractor::ComputeKeyPoints (src/ORBextractor.cc:536-746, SURVEY 8a row
 * E8; no caller reaches it in the reference) on the pyramid of a frame of the last extract call: quota cells (:547-561),
 * cv::FAST per cell + retry at threshold 5 when <= 3 corners (:645-652), HarrisResponses when the extractor was created with
 * HARRIS_SCORE (:655-659), quota redistribution (:683-709), KeyPointsFilter::retainBest per cell and per level (:718-741),
 * orientation (:744-745).  kps = [nlevels][cap_per_level] in level coordinates (= allKeypoints[level]).  Which keypoints of
 * EQUAL response survive a cut is decided by std::nth_element here as in the reference (OpenCV's retainBest calls it). */
int   uvip_compute_keypoints_quota(uvip_extractor* ex, int frame, uvip_keypoint* kps, int32_t* n_per_level, int cap_per_level);
/* how many of this library's kernels the handle has launched so far (bench.py's gpu_launches) */
long long uvip_extractor_launch_count(const uvip_extractor* ex);
/* per-stage device time (CUDA events recorded between the stage kernels on the launching stream), summed over
 * the launch groups since profiling was enabled (at most the last 256).  Stages: 0 pyramid (import + 7 resizes),
 * 1 FAST, 2 quadtree, 3 blur, 4 select, 5 orientation + descriptors.  Measurement utility for bench.py. */
#define UVIP_NUM_STAGES 6
int   uvip_extractor_profile(uvip_extractor* ex, int enable);
int   uvip_extractor_stage_ms(uvip_extractor* ex, float* ms /* [UVIP_NUM_STAGES] */, int* ngroups);

/* ---- next row N3: the step immediately before the path on every frame when `Enhance` is set ------------ */
/* cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply(src, dst) as called at src/Tracking.cc:425-431 (clip 4, 12 x 12
 * tiles), 8-bit, bit-exact against OpenCV.  dst may alias src.  Host buffers / device-resident batch. */
int   uvip_clahe(uvip_extractor* ex, const uint8_t* src, int w, int h, int stride, double clip_limit, int tiles_x, int tiles_y,
                 uint8_t* dst, int dst_stride);
int   uvip_clahe_batch_device(uvip_extractor* ex, const uint8_t* d_src, int nframes, int w, int h, int stride, size_t frame_pitch,
                              double clip_limit, int tiles_x, int tiles_y, uint8_t* d_dst, int dst_stride, size_t dst_pitch, void* stream);

/* ---- matcher: replaces the descriptor path of USLAM::ORBmatcher (include/ORBmatcher.h:41-94) ------------- */
typedef struct uvip_matcher uvip_matcher;

enum { UVIP_TH_HIGH = 100, UVIP_TH_LOW = 50, UVIP_HISTO_LENGTH = 30 };   /* src/ORBmatcher.cc:40-42 */

int   uvip_matcher_create(int device, uvip_matcher** out);
int   uvip_matcher_destroy(uvip_matcher* m);
long long uvip_matcher_launch_count(const uvip_matcher* m);
void* uvip_matcher_stream(uvip_matcher* m);        /* the handle's own cudaStream_t (what a NULL stream argument selects) */
int   uvip_matcher_sync(uvip_matcher* m);          /* waits for everything queued on the handle's own stream */

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1794-1810) for n row pairs: out[i] = popcount(a_i ^ b_i) */
int   uvip_descriptor_distance(uvip_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* out);

/* Brute-force k=2 nearest neighbours in Hamming space: the best/second-best scan every ORBmatcher search shares
 * (e.g. src/ORBmatcher.cc:201-226) run over the whole train set == BFMatcher(NORM_HAMMING).knnMatch(k=2) of
 * haloc::Utils::ratioMatching (include/utils.h:81-111).  Order = (distance, train index) ascending.
 * idx2/dist2 are nq x 2; a missing neighbour has idx -1, dist 257.  Host buffers. */
int   uvip_knn2(uvip_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2);
/* device-resident; train rows carry global indices idx_base + row so shards of a database can be merged */
int   uvip_knn2_device(uvip_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int idx_base,
                       int32_t* d_idx2, int32_t* d_dist2, void* stream);
/* npairs independent (query set, train set) problems in one launch: pair p matches d_q + p*q_pitch (nq[p] rows)
 * against d_t + p*t_pitch (nt[p] rows); results at d_idx2/d_dist2 + p*2*res_pitch.  d_nq/d_nt are device int32. */
int   uvip_knn2_batch_device(uvip_matcher* m, const uint8_t* d_q, const int32_t* d_nq, size_t q_pitch,
                             const uint8_t* d_t, const int32_t* d_nt, size_t t_pitch, int npairs, int max_nq,
                             int32_t* d_idx2, int32_t* d_dist2, size_t res_pitch, void* stream);
/* host-buffer form of the batch call (copies are inside the call): frame-to-frame matching of a whole sequence */
int   uvip_knn2_batch(uvip_matcher* m, const uint8_t* q, const int32_t* nq, size_t q_pitch,
                      const uint8_t* t, const int32_t* nt, size_t t_pitch, int npairs, int max_nq,
                      int32_t* idx2, int32_t* dist2, size_t res_pitch);
/* merge `nparts` partial top-2 lists (each nq x 2, parts are part_stride int32 apart) by (distance, global index);
 * shard-count invariant.  Used after the NCCL all-gather of the database-sharded kNN. */
int   uvip_knn2_merge_device(uvip_matcher* m, const int32_t* d_idx_parts, const int32_t* d_dist_parts, int nparts,
                             size_t part_stride, int nq, int32_t* d_idx2, int32_t* d_dist2, void* stream);
/* ratio test of include/utils.h:104-108: match[i] = idx2[2i] if dist0 <= dist1 * ratio (float x double), else -1 */
int   uvip_ratio_filter(uvip_matcher* m, const int32_t* idx2, const int32_t* dist2, int nq, double ratio,
                        int32_t* match, int* nmatches);
/* rotation-consistency histogram (bin code src/ORBmatcher.cc:232-241, ComputeThreeMaxima :1748-1789, rollback
 * :263-281): match[i] (train index or -1) is cleared unless bin(angle_a[i] - angle_b[match[i]]) is one of the
 * three fullest bins.  Host buffers. */
int   uvip_rot_hist_filter(uvip_matcher* m, int32_t* match, int n, const float* angle_a, const float* angle_b,
                           int* nkept);

/* Frame keypoint grid (src/FrameKTL.cc:250-264, PosInGrid :426-436): CSR over cols x rows cells,
 * cell id = ix*rows + iy, items ascending keypoint index.  cell_start has cols*rows+1 entries. */
int   uvip_grid_build(uvip_matcher* m, const float* kx, const float* ky, int n,
                      float min_x, float min_y, float inv_w, float inv_h, int cols, int rows,
                      int32_t* cell_start, int32_t* cell_items);

/* Grid-windowed search with claims: ORBmatcher::SearchByProjection(FrameKTL&, vector<MapPoint*>&, th)
 * (src/ORBmatcher.cc:49-125) when mode == 0, and the search loop of SearchByProjection(FrameKTL&, KeyFrame*, ...)
 * (src/ORBmatcher.cc:1683-1715) when mode == 1 — which is also the loop of SearchByProjection(KeyFrame*, Scw, ...)
 * (:372-399, taken = vpMatched) — and the loop of Fuse (:1075-1100) when mode == 4; candidates come from
 * FrameKTL::GetFeaturesInArea (src/FrameKTL.cc:359-424) = KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:952-992, same
 * cells and the same inclusive |dx| <= r test) followed by the explicit level test [l-1, l] of the keyframe-side loops.  The caller (shim) projects the map points and supplies per query
 * (u, v, radius, minLevel, maxLevel, descriptor).  taken[] (nk): -1 = free, anything else = keypoint already has a
 * map point; on return claimed keypoints hold the claiming query index.  match[q] = keypoint index or -1.
 * The sequential claim order of the reference is reproduced exactly. */
typedef struct uvip_search_params {
    int32_t mode;        /* 0: top-2 with same-level ratio rule; 1: best only; 4: best only without claims (Fuse,
                          * src/ORBmatcher.cc:1075-1100: taken[] is neither read nor written, several queries may return
                          * the same keypoint and the caller replays Replace / AddObservation in query order);
                          * 6: top-2 without levels, accept (float)best <= (float)best2 * ratio and best <= th_dist, claims
                          * (WindowSearch :409-516 and SearchByProjection(F1, F2, windowSize, ...) :519-596 — uncalled in this fork) */
    int32_t th_dist;     /* UVIP_TH_HIGH for mode 0, ORBdist for mode 1 */
    float   ratio;       /* mfNNratio */
    float   min_x, min_y, inv_w, inv_h;   /* FrameKTL::mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv */
    int32_t cols, rows;  /* FRAME_GRID_COLS 64, FRAME_GRID_ROWS 48 (include/FrameKTL.h:45-46) */
} uvip_search_params;
int   uvip_search_window(uvip_matcher* m, const uvip_search_params* sp,
                         const float* qu, const float* qv, const float* qr, const int32_t* qmin_level,
                         const int32_t* qmax_level, const uint8_t* qdesc, int nq,
                         const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, int nk,
                         const int32_t* cell_start, const int32_t* cell_items,
                         int32_t* taken, int32_t* match, int* nmatches);
/* The matching core of SearchForTriangulation (src/ORBmatcher.cc:893-952 with CheckDistEpipolarLine :136-153): per query
 * (a keypoint of keyframe 1 without a map point) the candidates of the same vocabulary node (CSR lists, caller-filtered to
 * keypoints of keyframe 2 without a map point) with distance <= th_dist are ordered by (distance, index); among those within
 * round(2 * best distance) the first whose squared distance to the epipolar line is below kthr[idx] = 3.84 * sigma2(octave)
 * is matched and claimed (taken[] as in uvip_search_window).  qline[q] = (a, b, c, a*a + b*b) of x1' F12 in float, as :139-146
 * computes them; kx, ky = keypoint coordinates of keyframe 2.  Queries in array order = the reference's node-major order. */
int   uvip_search_lists_epipolar(uvip_matcher* m, int th_dist, const uint8_t* qdesc, const float* qline, int nq,
                                 const int32_t* cand_start, const int32_t* cand_idx, const uint8_t* kdesc, const float* kx, const float* ky,
                                 const double* kthr, int nk, int32_t* taken, int32_t* match, int* nmatches);
/* uvip_search_window for one frame WITHOUT a caller-side grid: one upload, frame grid + search on the device, one download,
 * one synchronisation.  This is what the shim's per-frame searches call (e.g. SearchByProjection(F, local map points, th),
 * src/Tracking.cc:2228); uvip_grid_build + uvip_search_window remain for callers that keep a frame's grid. */
int   uvip_search_frame(uvip_matcher* m, const uvip_search_params* sp,
                        const float* qu, const float* qv, const float* qr, const int32_t* qmin_level,
                        const int32_t* qmax_level, const uint8_t* qdesc, int nq,
                        const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, int nk,
                        int32_t* taken, int32_t* match, int* nmatches);
/* The same search for a batch of device-resident frames (BASELINE config 3 as a throughput workload; frames shard with no
 * exchange because claims never cross frames).  Frame f owns queries [f*q_stride, f*q_stride + d_nq[f]) and keypoints
 * [f*k_stride, f*k_stride + d_nk[f]) of every array; the frame grids (src/FrameKTL.cc:250-264) are built on the device by the
 * same call.  d_taken (nframes x k_stride) in/out and d_match (nframes x q_stride) as above; d_counts[2f] = matches of
 * frame f, d_counts[2f+1] = claim rounds it took.  Asynchronous on `stream` (NULL = the handle's stream). */
int   uvip_search_window_batch_device(uvip_matcher* m, const uvip_search_params* sp, int nframes,
                                      const float* d_qu, const float* d_qv, const float* d_qr, const int32_t* d_qmin_level,
                                      const int32_t* d_qmax_level, const uint8_t* d_qdesc, const int32_t* d_nq, int q_stride,
                                      const float* d_kx, const float* d_ky, const int32_t* d_octave, const uint8_t* d_kdesc,
                                      const int32_t* d_nk, int k_stride,
                                      int32_t* d_taken, int32_t* d_match, int32_t* d_counts, void* stream);
/* Node-restricted search with claims over explicit candidate lists (CSR: cand_start[nq+1], cand_idx[]): the inner
 * loops of SearchByBoW(KeyFrame*, FrameKTL&, ...) (src/ORBmatcher.cc:186-245; mode 2: best <= th and
 * (float)best < ratio*(float)best2) and SearchByBoW(KeyFrame*, KeyFrame*, ...) (:751-811; mode 3: best < th, same
 * ratio rule); mode 1 = best only, best <= th (Fuse / SearchByProjection(KF,Scw) with caller-built candidate lists,
 * :1075-1100).  Queries are processed in array order (list them node-major like the reference's merge-join);
 * taken/match as in uvip_search_window.  The DBoW2 feature vectors that define the lists stay on the caller's side. */
int   uvip_search_lists(uvip_matcher* m, int mode, int th_dist, float ratio, const uint8_t* qdesc, int nq,
                        const int32_t* cand_start, const int32_t* cand_idx, const uint8_t* kdesc, int nk,
                        int32_t* taken, int32_t* match, int* nmatches);
/* next row N4, the descriptor half: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:197-270) for a
 * batch of map points.  The observed descriptors of point p are rows start[p] .. start[p+1]-1 of desc (start[0] == 0).
 * best_idx[p] = index inside the point's own list of the descriptor with the least median Hamming distance to the
 * others (sorted row element floor((N-1)/2), diagonal included; first index on ties, :250-264), best_median[p] that
 * median (may be NULL); both -1 for a point without descriptors.  Host buffers. */
int   uvip_distinctive_descriptors(uvip_matcher* m, const uint8_t* desc, const int32_t* start, int npoints,
                                   int32_t* best_idx, int32_t* best_median);
/* ORBmatcher::RadiusByViewingCos (src/ORBmatcher.cc:127-133) */
/* haloc hash (next row N4, second half): haloc::Hash::getHash (src/hash.cpp:57-85; called at src/KeyFrame.cc:322,328 and
 * src/LoopClosing.cc:136) for nsets descriptor sets, rows start[s]..start[s+1]-1 of desc (CSR).  proj = the caller's num_proj
 * projection vectors of proj_len floats (the reference's r_, drawn from rand() seeded with time(NULL), :95-147 — which is why
 * they are an input here); hash = nsets x (num_proj * 32) floats, products and running sums in float in row order, exactly
 * as :70-79.  uvip_haloc_match = haloc::Hash::match (:190-206) of one query against n stored hashes
 * (KeyFrameDatabase::DetectLoopCandidatesHaloc, src/KeyFrameDatabase.cc:98-118). */
int   uvip_haloc_hash(uvip_matcher* m, const uint8_t* desc, const int32_t* start, int nsets, const float* proj, int num_proj,
                      int proj_len, float* hash);
int   uvip_haloc_match(uvip_matcher* m, const float* query, const float* table, int n, int len, float* score);
float uvip_radius_by_viewing_cos(float view_cos);

/* ---- next row N2: DBoW2 vocabulary-tree descent ------------------------------------------------------ */
/* ORBVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) (src/FrameKTL.cc:439-446, src/KeyFrame.cc:203-210;
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1119-1195) = per descriptor a Hamming tree descent (:1218-1259) + host-side
 * map bookkeeping.  The descent is the kernel; the caller (shim / Python mirror) loads the vocabulary text file, flattens
 * the tree and accumulates the BowVector / FeatureVector from the per-feature results.
 * Flat tree: children of node i = child_ids[child_start[i] .. child_start[i+1]) in file order; node 0 is the root; a node
 * without children is a word (node_word = word id, else -1).  L = depth of the vocabulary (header of the text file). */
typedef struct uvip_vocabulary uvip_vocabulary;
int   uvip_vocabulary_create(int device, int nnodes, const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc,
                             const double* node_weight, const int32_t* node_word, int L, uvip_vocabulary** out);
int   uvip_vocabulary_destroy(uvip_vocabulary* v);
/* per feature: word id, weight of the word, and the node on the path at level L - levelsup (0 = root) */
int   uvip_bow_transform(uvip_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, int32_t* node_id, double* weight);
int   uvip_bow_transform_device(uvip_vocabulary* v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word_id, int32_t* d_node_id,
                                double* d_weight, void* stream);

/* ---- next row N1: KLT front end ------------------------------------------------------------------------ */
/* cv::buildOpticalFlowPyramid(im, pyr, Size(win,win), max_level) with derivatives (src/FrameKTL.cc:76) and
 * cv::calcOpticalFlowPyrLK(pyr0, pyr1, pts0, pts1, status, err, Size(win,win), max_level, criteria, flags)
 * (src/Tracking.cc:1044-1047).  A handle owns `nslots` device pyramids (a frame's pyramid is built once and serves as
 * "next" for one call and "prev" for the following one).  pyrDown / Scharr / window interpolation are bit-exact against
 * OpenCV; tracked positions agree within ~1e-3 px (float reduction order), status flags are identical.
 * flags: 4 = OPTFLOW_USE_INITIAL_FLOW, 8 = OPTFLOW_LK_GET_MIN_EIGENVALS (the reference passes both). */
typedef struct uvip_klt uvip_klt;
int   uvip_klt_create(int device, int max_width, int max_height, int win, int max_level, int nslots, uvip_klt** out);
int   uvip_klt_destroy(uvip_klt* k);
int   uvip_klt_build_pyramid(uvip_klt* k, int slot, const uint8_t* image, int w, int h, int stride, int* nlevels);
int   uvip_klt_get_level(uvip_klt* k, int slot, int level, uint8_t* img, int16_t* der_xy, int* w, int* h);   /* debug tap */
int   uvip_klt_track(uvip_klt* k, int slot_prev, int slot_next, const float* prev_pts, float* next_pts_inout, int n, int max_level,
                     int max_iter, double epsilon, int flags, double min_eig_threshold, uint8_t* status, float* err);
/* buildOpticalFlowPyramid + calcOpticalFlowPyrLK for a SEQUENCE resident in device memory (throughput form of the two calls above):
 * pyramids of nframes <= nslots frames into slots 0 .. nframes-1, then frame f -> f+1 tracked for every f in one launch.
 * d_prev_pts / d_next_pts / d_status / d_err: (nframes - 1) x npts (x 2 floats); d_next_pts holds the initial guesses when
 * OPTFLOW_USE_INITIAL_FLOW (4) is set.  Asynchronous on `stream`. */
int   uvip_klt_track_sequence_device(uvip_klt* k, const uint8_t* d_frames, int nframes, int w, int h, int stride, size_t frame_pitch,
                                     const float* d_prev_pts, float* d_next_pts, int npts, int max_level, int max_iter, double epsilon,
                                     int flags, double min_eig_threshold, uint8_t* d_status, float* d_err, void* stream);
/* cv::findFundamentalMat(pts0, pts1, FM_RANSAC, threshold, ...) of src/Tracking.cc:1062, of which the reference keeps the inlier mask:
 * 7-point hypotheses scored by OpenCV's residual (max of the two squared point-to-epipolar-line distances, double, cast to float,
 * compared with (float)(threshold^2)); the mask of the best hypothesis is returned together with its count and, optionally, its F
 * (row-major, Frobenius norm 1; pts1^T F pts0 = 0).  pts: n x 2 float, n >= 15 (below that OpenCV runs LMedS instead:
 * UVIP_ERR_UNSUPPORTED).  cv::RNG's sample sequence is not reproducible, so the samples come from a counter-based generator and nhyp
 * hypotheses are always evaluated (OpenCV: at most 1000 with early stop; 2048 is a good default): the inlier SET is what is pinned
 * against cv2 (tests/golden/cv2_ransac.npz), the result itself is deterministic and identical to the CPU oracle's. */
int   uvip_klt_ransac_fundamental(uvip_klt* k, const float* pts0, const float* pts1, int n, double threshold, int nhyp, uint8_t* mask,
                                  double* F, int* ninliers);
long long uvip_klt_launch_count(const uvip_klt* k);

/* ---- misc ---------------------------------------------------------------------------------------------- */
int         uvip_abi_version(void);
const char* uvip_last_error(void);          /* thread-local text of the last failure */
int         uvip_device_count(void);
/* measurement utility (no reference counterpart): sustained __popc throughput of `device` in popc/s — the
 * denominator of the Hamming-kNN roofline */
int         uvip_popc_peak(int device, int iters, double* popc_per_s);

/* ---- bench / test utility (not part of the hot path) ----------------------------------------------------------------------------
 * SURVEY.md Appendix B `synth_frame(seed, W, H, dx, dy, noise_seed)` on the device, byte-identical to the numpy generator of
 * u-vip-slam_b200/synth.py: frame f is written at d_out + f*frame_pitch (w*h bytes, row-major).  d_seeds / d_noise_seeds: int64 per frame,
 * d_dxy: (dx, dy) int32 pairs, |dx|, |dy| <= 32.  BASELINE config 5 (8 x 4096 frames of 1280x1024) is generated with it on the GPU box.
 * There is no handle: `stream` is used as given (NULL = the legacy default stream). */
int   uvip_synth_frames_device(const long long* d_seeds, const int* d_dxy, const long long* d_noise_seeds, int nframes, int w, int h,
                               uint8_t* d_out, size_t frame_pitch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UVIP_ORB_H */
