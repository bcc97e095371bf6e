// ref_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own USLAM::ORBextractor, compiled from
// /root/reference/src/ORBextractor.cc where it lies (see oracle/Makefile target `ref`) against the stand-in headers of
// this directory.  Used by tests/ to pin the oracle (and through it the CUDA path) against the reference's real code
// for everything that is not an OpenCV primitive.
//
// DistributeOctTree sorts (size, node POINTER) pairs (src/ORBextractor.cc:1151), so equal-size nodes are ordered by
// their heap addresses — implementation-defined in the reference.  While a call runs, operator new below hands out
// monotonically increasing addresses from an arena (no reuse), which realises the behaviour the oracle and the CUDA
// path pin: among equal sizes the later-created node comes first.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "ORBextractor.h"

namespace {
struct Arena { char* base; size_t cap, used; bool on; };
thread_local Arena g_arena = {0, 0, 0, false};
inline bool in_arena(void* p) { return g_arena.base && (char*)p >= g_arena.base && (char*)p < g_arena.base + g_arena.cap; }
void* arena_alloc(size_t n)
{
    if (g_arena.on) {
        const size_t a = (g_arena.used + 15) & ~(size_t)15;
        if (a + n <= g_arena.cap) { g_arena.used = a + n; return g_arena.base + a; }
        abort();                       // arena exhausted: never silently fall back to reusable heap addresses
    }
    void* p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
}
void* operator new(size_t n) { return arena_alloc(n); }
void* operator new[](size_t n) { return arena_alloc(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete(void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }

extern "C" {

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } ref_keypoint;

void* ref_create(int nfeatures, float scale_factor, int nlevels, int score_type, int fast_th)
{
    return new USLAM::ORBextractor(nfeatures, scale_factor, nlevels, score_type, fast_th);
}
void ref_destroy(void* h) { delete (USLAM::ORBextractor*)h; }
int ref_levels(void* h) { return ((USLAM::ORBextractor*)h)->GetLevels(); }
float ref_scale_factor(void* h) { return ((USLAM::ORBextractor*)h)->GetScaleFactor(); }

// ORBextractor::operator()(image, mask, keypoints, descriptors, grid_2d, min_px_dist, FullDetect, num_featsneeded).
// kps/n_inout carry the incoming keypoints in and the result out; grid is column-major rows x cols (Eigen::MatrixXi),
// read and updated when full_detect == 0.  Returns 0, -2 if cap is too small.  arena_mb bounds the call's allocations
// (0 = 1024 MB, < 0 = no arena).
int ref_extract(void* h, const uint8_t* img, int w, int h_, int stride, ref_keypoint* kps, int* n_inout, int cap, uint8_t* desc,
                int32_t* grid, int grid_rows, int grid_cols, int min_px_dist, int full_detect, int num_needed, int arena_mb)
{
    USLAM::ORBextractor* ex = (USLAM::ORBextractor*)h;
    Arena& A = g_arena;
    const size_t want = (size_t)(arena_mb > 0 ? arena_mb : 1024) << 20;
    if (A.cap < want) { free(A.base); A.base = (char*)malloc(want); A.cap = A.base ? want : 0; }
    if (!A.base) return -1;
    int rc = 0, nout = 0;
    A.used = 0; A.on = arena_mb >= 0;          // arena_mb < 0: plain malloc addresses (to observe the reference's nondeterminism)
    {
        cv::Mat image(h_, w, CV_8UC1, (void*)img, (size_t)stride), descriptors;
        if (w == 0 || h_ == 0) image = cv::Mat();
        std::vector<cv::KeyPoint> keypoints;
        keypoints.reserve((size_t)*n_inout);
        for (int i = 0; i < *n_inout; i++) {
            cv::KeyPoint k(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
            keypoints.push_back(k);
        }
        Eigen::MatrixXi g(grid ? grid_rows : 1, grid ? grid_cols : 1);
        if (grid) memcpy(g.data(), grid, sizeof(int) * (size_t)grid_rows * grid_cols);
        int mpd = min_px_dist;
        (*ex)(image, cv::Mat(), keypoints, descriptors, g, mpd, full_detect != 0, num_needed);
        nout = (int)keypoints.size();
        if (nout > cap) rc = -2;
        else {
            for (int i = 0; i < nout; i++) {
                const cv::KeyPoint& k = keypoints[i];
                ref_keypoint o = {k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
                kps[i] = o;
            }
            if (!image.empty() && !descriptors.empty())
                for (int i = 0; i < nout && i < descriptors.rows; i++) memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
            if (grid) memcpy(grid, g.data(), sizeof(int) * (size_t)grid_rows * grid_cols);
        }
    }
    A.on = false;
    // everything the call allocated through operator new is dead here (locals of operator(), `keypoints`, the list nodes);
    // the extractor's pyramid lives in malloc'ed stand-in Mat buffers, so the arena can be rewound by the next call
    *n_inout = nout;
    return rc;
}

// Frame-parallel batch for CPU-baseline timing: every OpenMP thread owns one reference extractor (the class is stateful,
// include/ORBextractor.h:90-91) and its own arena.  frames are contiguous w*h each; FullDetect, no incoming keypoints.
// arena_mb as in ref_extract.
int ref_extract_batch(int nfeatures, float scale_factor, int nlevels, int score_type, int fast_th, const uint8_t* frames, int nframes,
                      int w, int h_, ref_keypoint* kps, int* n_out, int cap, uint8_t* desc, int threads, int arena_mb)
{
    int bad = 0;
#pragma omp parallel num_threads(threads > 0 ? threads : 1) reduction(| : bad)
    {
        void* ex = ref_create(nfeatures, scale_factor, nlevels, score_type, fast_th);
#pragma omp for schedule(dynamic, 1)
        for (int f = 0; f < nframes; f++) {
            int n = 0;
            const int rc = ref_extract(ex, frames + (size_t)f * w * h_, w, h_, w, kps + (size_t)f * cap, &n, cap, desc + (size_t)f * cap * 32,
                                       0, 0, 0, 1, 1, 0, arena_mb);
            n_out[f] = rc ? 0 : n;
            bad |= rc != 0;
        }
        ref_destroy(ex);
    }
    return bad ? -2 : 0;
}

// The reference's dead detector path (src/ORBextractor.cc:536-746), reached through a subclass because ComputeKeyPoints and
// ComputePyramid are protected and no caller uses them: returns the keypoints ComputeKeyPoints leaves on `level` (level
// coordinates, response = HarrisResponses when the extractor was created with HARRIS_SCORE).  The SELECTION among equal
// responses depends on std::nth_element (KeyPointsFilter::retainBest), i.e. on the C++ library; the responses do not.
struct DeadPathProbe : public USLAM::ORBextractor {
    static int run(USLAM::ORBextractor* ex, const cv::Mat& image, int level, ref_keypoint* out, int cap)
    {
        DeadPathProbe* p = static_cast<DeadPathProbe*>(ex);
        p->ComputePyramid(image);
        std::vector<std::vector<cv::KeyPoint> > all;
        p->ComputeKeyPoints(all);
        if (level < 0 || level >= (int)all.size()) return -1;
        const int n = (int)all[(size_t)level].size();
        for (int i = 0; i < n && i < cap; i++) {
            const cv::KeyPoint& k = all[(size_t)level][(size_t)i];
            ref_keypoint o = {k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
            out[i] = o;
        }
        return n;
    }
};
int ref_dead_path_keypoints(void* h, const uint8_t* img, int w, int h_, int stride, int level, ref_keypoint* out, int cap)
{
    cv::Mat image(h_, w, CV_8UC1, (void*)img, (size_t)stride);
    return DeadPathProbe::run((USLAM::ORBextractor*)h, image, level, out, cap);
}

}  // extern "C"
