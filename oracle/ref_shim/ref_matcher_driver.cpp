// ref_matcher_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own USLAM::ORBmatcher, compiled
// from /root/reference/src/ORBmatcher.cc where it lies (oracle/Makefile target `ref`) against the stand-in types of
// slam_standin.h.  Each entry point builds the reference's pointer-rich scene (FrameKTL / KeyFrame / MapPoint stand-ins)
// from flat arrays, calls the reference's member function, and flattens the result.  Used by tests/ to pin the oracle's
// matcher restatement (and through it the CUDA kernels and the C++ shim) against the reference's compiled code.
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <memory>
#include <set>
#include <vector>
#include "ORBmatcher.h"
#include "m8_scene.h"

using namespace USLAM;

namespace {

// wall time of the last reference member call alone (scene construction excluded), for the CPU-baseline legs of bench.py
thread_local double g_call_seconds = 0.0;
struct CallTimer {
    timespec t0;
    CallTimer() { clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~CallTimer() { timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); g_call_seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec); }
};

cv::Mat desc_mat(const uint8_t* d, int n) { cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1); if (n > 0) memcpy(m.data, d, (size_t)n * 32); return m; }
cv::Mat vec3(const float* p) { cv::Mat m(3, 1, CV_32F); for (int i = 0; i < 3; i++) m.at<float>(i) = p[i]; return m; }

void fill_keys(std::vector<cv::KeyPoint>& keys, int n, const float* kx, const float* ky, const int32_t* octave, const float* angle)
{
    keys.resize(n);
    for (int i = 0; i < n; i++) keys[i] = cv::KeyPoint(kx[i], ky[i], 31.f, angle ? angle[i] : -1.f, 0.f, octave ? octave[i] : 0, -1);
}

void fill_featvec(DBoW2::FeatureVector& fv, int nnodes, const int32_t* node_id, const int32_t* start, const int32_t* idx)
{
    for (int i = 0; i < nnodes; i++) {
        std::vector<unsigned int>& v = fv[(DBoW2::NodeId)node_id[i]];
        for (int j = start[i]; j < start[i + 1]; j++) v.push_back((unsigned)idx[j]);
    }
}

struct FrameScene {
    FrameKTL F;
    std::vector<MapPoint> pre;         // map points already attached to keypoints before the call
    void init(int nk, const float* kx, const float* ky, const int32_t* octave, const float* angle, const uint8_t* desc, const float* bounds,
              int nlevels, const float* sf, const int32_t* taken)
    {
        fill_keys(F.mvKeysUn, nk, kx, ky, octave, angle);
        F.mvKeys = F.mvKeysUn;
        F.mDescriptors = desc_mat(desc, nk);
        F.mnMinX = bounds[0]; F.mnMaxX = bounds[1]; F.mnMinY = bounds[2]; F.mnMaxY = bounds[3];
        F.mvScaleFactors.assign(sf, sf + nlevels); F.mnScaleLevels = nlevels;
        F.grid.build(F.mvKeysUn, F.mnMinX, F.mnMaxX, F.mnMinY, F.mnMaxY);
        F.mvpMapPoints.assign(nk, (MapPoint*)0);
        F.mvbOutlier.assign(nk, false);
        pre.resize(nk);
        for (int i = 0; i < nk; i++) if (taken && taken[i] != -1) F.mvpMapPoints[i] = &pre[i];
    }
};

}  // namespace

extern "C" {

double refm_last_call_seconds(void) { return g_call_seconds; }

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1794-1810)
int refm_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int32_t wa[8], wb[8];
    memcpy(wa, a, 32); memcpy(wb, b, 32);
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8UC1, wa), cv::Mat(1, 32, CV_8UC1, wb));
}

// ORBmatcher::SearchByProjection(FrameKTL&, const vector<MapPoint*>&, th)  (src/ORBmatcher.cc:49-125)
// taken[k] != -1: keypoint k already has a map point.  owner[k] = index of the map point that claimed keypoint k, -1 none,
// -2 for pre-taken keypoints.  Returns the reference's return value (nmatches).
int refm_search_by_projection_mps(int nk, const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, const float* bounds,
                                  int nlevels, const float* sf, const int32_t* taken,
                                  int nq, const float* u, const float* v, const int32_t* level, const float* view_cos, const uint8_t* in_view,
                                  const uint8_t* bad, const uint8_t* qdesc, float th, float nnratio, int32_t* owner)
{
    FrameScene S;
    S.init(nk, kx, ky, octave, 0, kdesc, bounds, nlevels, sf, taken);
    std::vector<MapPoint> mps(nq);
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; i++) {
        MapPoint& m = mps[i];
        m.mbTrackInView = in_view ? in_view[i] != 0 : true; m.bad = bad ? bad[i] != 0 : false;
        m.mnTrackScaleLevel = level[i]; m.mTrackProjX = u[i]; m.mTrackProjY = v[i]; m.mTrackViewCos = view_cos[i];
        m.desc = desc_mat(qdesc + (size_t)i * 32, 1);
        vp[i] = &m;
    }
    ORBmatcher matcher(nnratio, true);
    int n;
    { CallTimer timer; n = matcher.SearchByProjection(S.F, vp, th); }
    for (int k = 0; k < nk; k++) {
        MapPoint* p = S.F.mvpMapPoints[k];
        owner[k] = !p ? -1 : (p >= &mps[0] && p < &mps[0] + nq ? (int32_t)(p - &mps[0]) : -2);
    }
    return n;
}

// ORBmatcher::SearchByProjection(FrameKTL& CurrentFrame, KeyFrame*, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:1622-1746).  Keyframe slot i: has_mp / bad / found flags, world position pos[3i..], descriptor,
// min distance invariance, keypoint angle.  Tcw is the current pose, 4x4 row-major float.
int refm_search_by_projection_kf(int nk, const float* kx, const float* ky, const int32_t* octave, const float* kangle, const uint8_t* kdesc,
                                 const float* bounds, int nlevels, const float* sf, const int32_t* taken, const float* Tcw, const float* intr,
                                 int np, const uint8_t* has_mp, const uint8_t* bad, const uint8_t* found, const float* pos, const float* min_dist,
                                 const uint8_t* pdesc, const float* pangle, float th, int orb_dist, float nnratio, int check_ori, int32_t* owner)
{
    FrameScene S;
    S.init(nk, kx, ky, octave, kangle, kdesc, bounds, nlevels, sf, taken);
    S.F.mTcw = cv::Mat(4, 4, CV_32F);
    memcpy(S.F.mTcw.data, Tcw, 64);
    S.F.fx = intr[0]; S.F.fy = intr[1]; S.F.cx = intr[2]; S.F.cy = intr[3];
    KeyFrame KF;
    std::vector<MapPoint> mps(np);
    KF.mapPoints.assign(np, (MapPoint*)0);
    KF.keysUn.resize(np);
    std::set<MapPoint*> already;
    for (int i = 0; i < np; i++) {
        KF.keysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, pangle[i], 0.f, 0, -1);
        if (!has_mp[i]) continue;
        MapPoint& m = mps[i];
        m.bad = bad[i] != 0; m.pos = vec3(pos + 3 * (size_t)i); m.minDist = min_dist[i]; m.desc = desc_mat(pdesc + (size_t)i * 32, 1);
        KF.mapPoints[i] = &m;
        if (found[i]) already.insert(&m);
    }
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByProjection(S.F, &KF, already, th, orb_dist);
    for (int k = 0; k < nk; k++) {
        MapPoint* p = S.F.mvpMapPoints[k];
        owner[k] = !p ? -1 : (p >= &mps[0] && p < &mps[0] + np ? (int32_t)(p - &mps[0]) : -2);
    }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, FrameKTL&, vpMapPointMatches)  (src/ORBmatcher.cc:155-284)
// match_of_frame_kp[k] = keyframe slot whose map point was assigned to frame keypoint k, else -1.
int refm_search_by_bow_kf_frame(int nkf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* has_mp, const uint8_t* bad,
                                int kf_nodes, const int32_t* kf_node_id, const int32_t* kf_start, const int32_t* kf_idx,
                                int nk, const uint8_t* f_desc, const float* f_angle,
                                int f_nodes, const int32_t* f_node_id, const int32_t* f_start, const int32_t* f_idx,
                                float nnratio, int check_ori, int32_t* match_of_frame_kp)
{
    KeyFrame KF;
    std::vector<MapPoint> mps(nkf);
    KF.mapPoints.assign(nkf, (MapPoint*)0);
    KF.keysUn.resize(nkf);
    KF.descriptors = desc_mat(kf_desc, nkf);
    for (int i = 0; i < nkf; i++) {
        KF.keysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, kf_angle[i], 0.f, 0, -1);
        if (has_mp[i]) { mps[i].bad = bad[i] != 0; KF.mapPoints[i] = &mps[i]; }
    }
    fill_featvec(KF.featVec, kf_nodes, kf_node_id, kf_start, kf_idx);
    FrameKTL F;
    F.mvKeys.resize(nk);
    for (int i = 0; i < nk; i++) F.mvKeys[i] = cv::KeyPoint(0.f, 0.f, 31.f, f_angle[i], 0.f, 0, -1);
    F.mvKeysUn = F.mvKeys;
    F.mDescriptors = desc_mat(f_desc, nk);
    F.mvpMapPoints.assign(nk, (MapPoint*)0);
    fill_featvec(F.mFeatVec, f_nodes, f_node_id, f_start, f_idx);
    std::vector<MapPoint*> matches;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&KF, F, matches);
    for (int k = 0; k < nk; k++) match_of_frame_kp[k] = matches[k] ? (int32_t)(matches[k] - &mps[0]) : -1;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)  (src/ORBmatcher.cc:715-850)
// match12[i] = slot in keyframe 2 matched to slot i of keyframe 1, else -1.
int refm_search_by_bow_kf_kf(int n1, const uint8_t* desc1, const float* angle1, const uint8_t* has1, const uint8_t* bad1,
                             int nodes1, const int32_t* node_id1, const int32_t* start1, const int32_t* idx1,
                             int n2, const uint8_t* desc2, const float* angle2, const uint8_t* has2, const uint8_t* bad2,
                             int nodes2, const int32_t* node_id2, const int32_t* start2, const int32_t* idx2,
                             float nnratio, int check_ori, int32_t* match12)
{
    KeyFrame K1, K2;
    std::vector<MapPoint> m1(n1), m2(n2);
    struct { KeyFrame* K; std::vector<MapPoint>* m; int n; const uint8_t* d; const float* a; const uint8_t* has; const uint8_t* bad; } side[2] = {
        {&K1, &m1, n1, desc1, angle1, has1, bad1}, {&K2, &m2, n2, desc2, angle2, has2, bad2}};
    for (auto& s : side) {
        s.K->mapPoints.assign(s.n, (MapPoint*)0);
        s.K->keysUn.resize(s.n);
        s.K->descriptors = desc_mat(s.d, s.n);
        for (int i = 0; i < s.n; i++) {
            s.K->keysUn[i] = cv::KeyPoint(0.f, 0.f, 31.f, s.a[i], 0.f, 0, -1);
            if (s.has[i]) { (*s.m)[i].bad = s.bad[i] != 0; s.K->mapPoints[i] = &(*s.m)[i]; }
        }
    }
    fill_featvec(K1.featVec, nodes1, node_id1, start1, idx1);
    fill_featvec(K2.featVec, nodes2, node_id2, start2, idx2);
    std::vector<MapPoint*> matches;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&K1, &K2, matches);
    for (int i = 0; i < n1; i++) match12[i] = matches[i] ? (int32_t)(matches[i] - &m2[0]) : -1;
    return n;
}

// ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>&, th)  (src/ORBmatcher.cc:1016-1134).  Keyframe: keypoints, descriptors, pose
// (Rcw 3x3 row-major, tcw, Ow), intrinsics, image bounds, scale factors, kf_has_mp/kf_bad per slot.  Map points: world
// position, normal, min/max distance, descriptor, null/bad/in-keyframe flags.
// action[i]: 0 = nothing, 1 = replaced by the keyframe's map point at slot target[i], 2 = added to the keyframe at slot
// target[i], 3 = hit a slot whose map point is bad (counted, nothing done).  Returns nFused.
int refm_fuse(int nk, const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, const float* bounds, int nlevels,
              const float* sf, const float* Rcw, const float* tcw, const float* Ow, const float* intr, const uint8_t* kf_has_mp,
              const uint8_t* kf_bad, int np, const uint8_t* is_null, const uint8_t* bad, const uint8_t* in_kf, const float* pos,
              const float* normal, const float* min_dist, const float* max_dist, const uint8_t* pdesc, float th, int32_t* action,
              int32_t* target)
{
    KeyFrame KF;
    fill_keys(KF.keysUn, nk, kx, ky, octave, 0);
    KF.descriptors = desc_mat(kdesc, nk);
    KF.set_bounds((int)bounds[0], (int)bounds[1], (int)bounds[2], (int)bounds[3]);
    KF.scaleFactors.assign(sf, sf + nlevels);
    KF.Rcw = cv::Mat(3, 3, CV_32F); memcpy(KF.Rcw.data, Rcw, 36);
    KF.tcw = vec3(tcw); KF.Ow = vec3(Ow);
    KF.fx = intr[0]; KF.fy = intr[1]; KF.cx = intr[2]; KF.cy = intr[3];
    std::vector<MapPoint> kfmp(nk), mps(np);
    KF.mapPoints.assign(nk, (MapPoint*)0);
    for (int i = 0; i < nk; i++) if (kf_has_mp[i]) { kfmp[i].bad = kf_bad[i] != 0; KF.mapPoints[i] = &kfmp[i]; }
    std::vector<MapPoint*> vp(np, (MapPoint*)0);
    for (int i = 0; i < np; i++) {
        if (is_null[i]) continue;
        MapPoint& m = mps[i];
        m.bad = bad[i] != 0; m.pos = vec3(pos + 3 * (size_t)i); m.normal = vec3(normal + 3 * (size_t)i);
        m.minDist = min_dist[i]; m.maxDist = max_dist[i]; m.desc = desc_mat(pdesc + (size_t)i * 32, 1);
        if (in_kf[i]) m.obs[&KF] = 0;
        vp[i] = &m;
    }
    std::vector<MapPoint*> before(KF.mapPoints);
    ORBmatcher matcher(0.6f, true);
    const int n = matcher.Fuse(&KF, vp, th);
    for (int i = 0; i < np; i++) {
        action[i] = 0; target[i] = -1;
        if (is_null[i]) continue;
        MapPoint& m = mps[i];
        if (m.replaced) { action[i] = 1; target[i] = (int32_t)(m.replaced - &kfmp[0]); if (m.replaced >= &mps[0] && m.replaced < &mps[0] + np) target[i] = -2 - (int32_t)(m.replaced - &mps[0]); }
        else if (!in_kf[i] && m.obs.count(&KF)) { action[i] = 2; target[i] = (int32_t)m.obs[&KF]; }
    }
    return n;
}

// the keyframe-side searches of row M8 on a scene bundle (m8_scene.h): 0 Fuse(KF,MPs)  1 Fuse(KF,Scw)  2 SearchByProjection(KF,Scw)
// 3 SearchBySim3.  Returns the reference's return value, or -1000 on I/O failure.
int refm_m8_run(const char* scene_path, const char* out_path, int which)
{
    m8::Bundle in, out;
    if (!in.load(scene_path) || (int)in.a.size() < m8::A_COUNT) return -1000;
    const int r = m8::run<ORBmatcher>(which, in, out);
    return out.save(out_path) ? r : -1000;
}

}  // extern "C"
