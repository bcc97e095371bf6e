// m8_scene.h — TEST INFRASTRUCTURE.  One scene description and one driver routine for the keyframe-side searches of
// SURVEY 8a row M8, compiled TWICE with the same stand-in FrameKTL / KeyFrame / MapPoint types (slam_standin.h):
//   * against the reference's own USLAM::ORBmatcher (src/ORBmatcher.cc, oracle/_ref/libref_orbmatcher.so, entry refm_m8_run)
//   * against the drop-in shim's USLAM::ORBmatcher (u-vip-slam_b200/host/ORBmatcher.h -> C-ABI -> CUDA; tests/cpp/test_shim_m8.cpp)
// Both read the same array bundle written by tests/test_reference_pin.py and dump the same result bundle, so the test is a
// byte comparison of what the two matchers did to identical scenes through identical call signatures.
#ifndef UVIP_M8_SCENE_H
#define UVIP_M8_SCENE_H
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace m8 {

// bundle file: int32 count, then per array: int32 nbytes + payload
struct Bundle {
    std::vector<std::vector<char> > a;
    bool load(const char* path)
    {
        FILE* f = fopen(path, "rb"); if (!f) return false;
        int32_t n = 0; if (fread(&n, 4, 1, f) != 1) { fclose(f); return false; }
        a.resize((size_t)n);
        for (int i = 0; i < n; i++) {
            int32_t nb = 0; if (fread(&nb, 4, 1, f) != 1) { fclose(f); return false; }
            a[(size_t)i].resize((size_t)nb);
            if (nb && fread(a[(size_t)i].data(), 1, (size_t)nb, f) != (size_t)nb) { fclose(f); return false; }
        }
        fclose(f); return true;
    }
    bool save(const char* path) const
    {
        FILE* f = fopen(path, "wb"); if (!f) return false;
        int32_t n = (int32_t)a.size(); fwrite(&n, 4, 1, f);
        for (size_t i = 0; i < a.size(); i++) { int32_t nb = (int32_t)a[i].size(); fwrite(&nb, 4, 1, f); if (nb) fwrite(a[i].data(), 1, (size_t)nb, f); }
        fclose(f); return true;
    }
    template <class T> const T* get(int i) const { return (const T*)a[(size_t)i].data(); }
    template <class T> int count(int i) const { return (int)(a[(size_t)i].size() / sizeof(T)); }
    template <class T> void put(const std::vector<T>& v) { a.push_back(std::vector<char>((const char*)v.data(), (const char*)v.data() + v.size() * sizeof(T))); }
};

inline cv::Mat mat_f32(const float* p, int r, int c) { cv::Mat m(r, c, CV_32F); memcpy(m.data, p, sizeof(float) * (size_t)r * c); return m; }
inline cv::Mat mat_desc(const uint8_t* d, int n) { cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1); if (n > 0) memcpy(m.data, d, (size_t)n * 32); return m; }

// array order of the scene bundle (written by tests/test_reference_pin.py::m8_bundle)
enum { A_PARAMS = 0,      // float: th, s12, nnratio
       A_SF,              // float[nlevels] scale factors (both keyframes)
       A_INTR,            // float[4] fx fy cx cy
       A_BOUNDS,          // int32[4] minX maxX minY maxY
       A_SCW,             // float[16]
       A_R12, A_T12,      // float[9], float[3]
       K1_XYOA,           // float[nk1*4]: x, y, octave, angle
       K1_DESC,           // uint8[nk1*32]
       K1_POSE,           // float[15]: Rcw(9) tcw(3) Ow(3)
       K1_MP,             // int32[nk1*3]: has map point, bad, pre-matched slot in K2 for SearchBySim3 (-1 none)
       K1_MPGEO,          // float[nk1*8]: pos(3) normal(3) minDist maxDist of the slot's map point
       K2_XYOA, K2_DESC, K2_POSE, K2_MP, K2_MPGEO,
       P_FLAGS,           // int32[np*4]: null, bad, already observed in K1 (Fuse) / already matched (Scw searches), -
       P_GEO,             // float[np*8]
       P_DESC,            // uint8[np*32]
       K1_MATCHED,        // int32[nk1]: vpMatched for SearchByProjection(KF,Scw): index into the point list or -1
       A_F12,             // float[9] fundamental matrix for SearchForTriangulation
       K1_NODE, K2_NODE,  // int32[nk]: vocabulary node of every keypoint (FeatureVector), -1 = none
       A_COUNT };

struct Scene {
    USLAM::KeyFrame K[2];
    std::vector<USLAM::MapPoint> kmp[2], pts;
    std::vector<USLAM::MapPoint*> vp;
    cv::Mat Scw, R12, t12, F12; float th, s12, nnratio;
    int nk[2], np;

    void fill_point(USLAM::MapPoint& m, const float* geo, const uint8_t* desc, bool bad)
    {
        m.bad = bad; m.pos = mat_f32(geo, 3, 1); m.normal = mat_f32(geo + 3, 3, 1); m.minDist = geo[6]; m.maxDist = geo[7];
        m.desc = mat_desc(desc, 1);
    }
    void build(const Bundle& B)
    {
        const float* prm = B.get<float>(A_PARAMS); th = prm[0]; s12 = prm[1]; nnratio = prm[2];
        const int nl = B.count<float>(A_SF);
        const float* intr = B.get<float>(A_INTR); const int32_t* bd = B.get<int32_t>(A_BOUNDS);
        Scw = mat_f32(B.get<float>(A_SCW), 4, 4); R12 = mat_f32(B.get<float>(A_R12), 3, 3); t12 = mat_f32(B.get<float>(A_T12), 3, 1);
        F12 = mat_f32(B.get<float>(A_F12), 3, 3);
        for (int k = 0; k < 2; k++) {
            const int o = k ? K2_XYOA : K1_XYOA;
            USLAM::KeyFrame& KF = K[k];
            nk[k] = B.count<float>(o) / 4;
            const float* xyoa = B.get<float>(o);
            KF.keysUn.resize((size_t)nk[k]);
            for (int i = 0; i < nk[k]; i++) KF.keysUn[(size_t)i] = cv::KeyPoint(xyoa[4 * i], xyoa[4 * i + 1], 31.f, xyoa[4 * i + 3], 0.f, (int)xyoa[4 * i + 2], -1);
            KF.descriptors = mat_desc(B.get<uint8_t>(o + 1), nk[k]);
            const float* pose = B.get<float>(o + 2);
            KF.Rcw = mat_f32(pose, 3, 3); KF.tcw = mat_f32(pose + 9, 3, 1); KF.Ow = mat_f32(pose + 12, 3, 1);
            KF.fx = intr[0]; KF.fy = intr[1]; KF.cx = intr[2]; KF.cy = intr[3];
            KF.scaleFactors.assign(B.get<float>(A_SF), B.get<float>(A_SF) + nl);
            KF.levelSigma2.resize((size_t)nl);
            for (int l = 0; l < nl; l++) KF.levelSigma2[(size_t)l] = KF.scaleFactors[(size_t)l] * KF.scaleFactors[(size_t)l];     // src/FrameKTL.cc: mvLevelSigma2
            const int32_t* node = B.get<int32_t>(k ? K2_NODE : K1_NODE);
            for (int i = 0; i < nk[k]; i++) if (node[i] >= 0) KF.featVec[(DBoW2::NodeId)node[i]].push_back((unsigned)i);
            KF.mnId = (unsigned long)k;
            KF.set_bounds(bd[0], bd[1], bd[2], bd[3]);
            const int32_t* mp = B.get<int32_t>(o + 3); const float* geo = B.get<float>(o + 4);
            kmp[k].resize((size_t)nk[k]);
            KF.mapPoints.assign((size_t)nk[k], (USLAM::MapPoint*)0);
            for (int i = 0; i < nk[k]; i++)
                if (mp[3 * i]) { fill_point(kmp[k][(size_t)i], geo + 8 * i, B.get<uint8_t>(o + 1) + 32 * (size_t)i, mp[3 * i + 1] != 0); kmp[k][(size_t)i].obs[&KF] = (size_t)i; KF.mapPoints[(size_t)i] = &kmp[k][(size_t)i]; }
        }
        np = B.count<int32_t>(P_FLAGS) / 4;
        pts.resize((size_t)np); vp.assign((size_t)np, (USLAM::MapPoint*)0);
        const int32_t* pf = B.get<int32_t>(P_FLAGS); const float* pg = B.get<float>(P_GEO);
        for (int i = 0; i < np; i++) {
            if (pf[4 * i]) continue;
            fill_point(pts[(size_t)i], pg + 8 * i, B.get<uint8_t>(P_DESC) + 32 * (size_t)i, pf[4 * i + 1] != 0);
            vp[(size_t)i] = &pts[(size_t)i];
        }
    }
    // identity of a map point: keyframe-1 slot k -> k, keyframe-2 slot k -> 100000 + k, list point i -> 200000 + i, none -> -1
    int id_of(const USLAM::MapPoint* p) const
    {
        if (!p) return -1;
        if (!kmp[0].empty() && p >= &kmp[0][0] && p < &kmp[0][0] + nk[0]) return (int)(p - &kmp[0][0]);
        if (!kmp[1].empty() && p >= &kmp[1][0] && p < &kmp[1][0] + nk[1]) return 100000 + (int)(p - &kmp[1][0]);
        if (!pts.empty() && p >= &pts[0] && p < &pts[0] + np) return 200000 + (int)(p - &pts[0]);
        return -2;
    }
};

// The two keyframes of the scene as FRAMES (FrameKTL stand-ins), for the frame-side overloads of include/ORBmatcher.h:49-72 that no
// caller of this fork reaches: same keypoints, descriptors, map points and poses
inline void frame_of(const Scene& S, int k, USLAM::FrameKTL& F, const int32_t* bd)
{
    const USLAM::KeyFrame& K = S.K[k];
    F.mvKeysUn = K.keysUn; F.mvKeys = K.keysUn;
    F.mDescriptors = K.descriptors.clone();
    F.mvpMapPoints = K.mapPoints;
    F.mvbOutlier.assign(K.keysUn.size(), false);
    for (size_t i = 0; i < F.mvbOutlier.size(); i++) F.mvbOutlier[i] = (i % 17) == 3;
    F.mvScaleFactors = K.scaleFactors; F.mnScaleLevels = (int)K.scaleFactors.size();
    F.fx = K.fx; F.fy = K.fy; F.cx = K.cx; F.cy = K.cy;
    F.mnMinX = (float)bd[0]; F.mnMaxX = (float)bd[1]; F.mnMinY = (float)bd[2]; F.mnMaxY = (float)bd[3];
    F.mfGridElementWidthInv = (float)FRAME_GRID_COLS / (F.mnMaxX - F.mnMinX); F.mfGridElementHeightInv = (float)FRAME_GRID_ROWS / (F.mnMaxY - F.mnMinY);
    F.grid.build(F.mvKeysUn, F.mnMinX, F.mnMaxX, F.mnMinY, F.mnMaxY);
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) F.mTcw.at<float>(r, c) = K.Rcw.at<float>(r, c); F.mTcw.at<float>(r, 3) = K.tcw.at<float>(r); }
    F.mnId = (unsigned long)k;
}

// which: 0 Fuse(KF, MPs, th)   1 Fuse(KF, Scw, MPs, th)   2 SearchByProjection(KF, Scw, MPs, vpMatched, th)   3 SearchBySim3
//        4 SearchForTriangulation(KF1, KF2, F12, keys1, keys2, pairs)
//        5 WindowSearch(F1, F2, (int)th, matches2)    6 SearchByProjection(F1, F2, (int)th, matches2)    7 SearchForInitialization(F1, F2,
//        prev, matches12, (int)th)    8 SearchByProjection(F2 = current, F1 = last, th)    9 WindowSearch with octave limits 1..3
template <class Matcher>
int run(int which, const Bundle& in, Bundle& out)
{
    Scene S; S.build(in);
    Matcher matcher(S.nnratio, true);
    if (which >= 5) {
        const int32_t* bd = in.get<int32_t>(A_BOUNDS);
        USLAM::FrameKTL F1, F2; frame_of(S, 0, F1, bd); frame_of(S, 1, F2, bd);
        std::vector<int32_t> ret(1, 0), slot_owner, replaced, obs_slot;
        if (which == 5 || which == 9 || which == 6) {
            std::vector<USLAM::MapPoint*> m2;
            if (which == 5) ret[0] = matcher.WindowSearch(F1, F2, (int)S.th, m2);
            else if (which == 9) ret[0] = matcher.WindowSearch(F1, F2, (int)S.th, m2, 1, 3);
            else ret[0] = matcher.SearchByProjection(F1, F2, (int)S.th, m2);
            for (size_t k = 0; k < m2.size(); k++) slot_owner.push_back(S.id_of(m2[k]));
        } else if (which == 7) {
            std::vector<cv::Point2f> prev(F1.mvKeysUn.size()); std::vector<int> m12;
            for (size_t i = 0; i < prev.size(); i++) { prev[i] = F1.mvKeysUn[i].pt; prev[i].x += (float)((int)(i % 5) - 2) * 0.5f; prev[i].y -= (float)((int)(i % 3) - 1) * 0.5f; }
            for (size_t i = 0; i < F1.mvKeysUn.size(); i++) if (i % 3 != 2) { F1.mvKeysUn[i].octave = 0; }      // the function only looks at level 0
            for (size_t i = 0; i < F2.mvKeysUn.size(); i++) if (i % 3 != 2) { F2.mvKeysUn[i].octave = 0; }
            F2.grid.build(F2.mvKeysUn, F2.mnMinX, F2.mnMaxX, F2.mnMinY, F2.mnMaxY);
            ret[0] = matcher.SearchForInitialization(F1, F2, prev, m12, (int)S.th);
            for (size_t i = 0; i < m12.size(); i++) slot_owner.push_back(m12[i]);
            for (size_t i = 0; i < prev.size(); i++) { replaced.push_back((int32_t)(prev[i].x * 64.f)); replaced.push_back((int32_t)(prev[i].y * 64.f)); }
        } else {
            ret[0] = matcher.SearchByProjection(F2, F1, S.th);
            for (size_t k = 0; k < F2.mvpMapPoints.size(); k++) slot_owner.push_back(S.id_of(F2.mvpMapPoints[k]));
        }
        out.put(ret); out.put(slot_owner); out.put(replaced); out.put(obs_slot);
        return ret[0];
    }
    const int32_t* pf = in.get<int32_t>(P_FLAGS);
    std::vector<int32_t> ret(1, 0), slot_owner, replaced, obs_slot;
    USLAM::KeyFrame* pKF = &S.K[0];
    if (which == 0 || which == 1) {
        std::vector<USLAM::MapPoint*> vp;
        for (int i = 0; i < S.np; i++) {
            if (which == 1 && !S.vp[(size_t)i]) continue;            // Fuse(KF,Scw) dereferences every entry: no NULLs in its list
            if (S.vp[(size_t)i] && pf[4 * i + 2] && which == 0) S.pts[(size_t)i].obs[pKF] = 0;     // IsInKeyFrame(pKF)
            vp.push_back(S.vp[(size_t)i]);
        }
        // th < 0 in the bundle: call with the DEFAULT argument (include/ORBmatcher.h:85,88 — what src/LocalMapping.cc:1236,1261
        // rely on), so that a drop-in header with a different default fails the comparison
        if (S.th < 0.f) ret[0] = which == 0 ? matcher.Fuse(pKF, vp) : matcher.Fuse(pKF, S.Scw, vp);
        else ret[0] = which == 0 ? matcher.Fuse(pKF, vp, S.th) : matcher.Fuse(pKF, S.Scw, vp, S.th);
        for (int k = 0; k < S.nk[0]; k++) slot_owner.push_back(S.id_of(pKF->mapPoints[(size_t)k]));
        for (int k = 0; k < S.nk[0]; k++) replaced.push_back(S.id_of(S.kmp[0][(size_t)k].replaced));
        for (int i = 0; i < S.np; i++) replaced.push_back(S.id_of(S.pts[(size_t)i].replaced));
        for (int i = 0; i < S.np; i++) obs_slot.push_back(S.pts[(size_t)i].obs.count(pKF) && !(which == 0 && pf[4 * i + 2]) ? (int32_t)S.pts[(size_t)i].obs[pKF] : -1);
    } else if (which == 2) {
        std::vector<USLAM::MapPoint*> vp, matched((size_t)S.nk[0], (USLAM::MapPoint*)0);
        for (int i = 0; i < S.np; i++) if (S.vp[(size_t)i]) vp.push_back(S.vp[(size_t)i]);
        const int32_t* km = in.get<int32_t>(K1_MATCHED);
        for (int k = 0; k < S.nk[0]; k++) if (km[k] >= 0 && S.vp[(size_t)km[k]]) matched[(size_t)k] = S.vp[(size_t)km[k]];
        ret[0] = matcher.SearchByProjection(pKF, S.Scw, vp, matched, (int)S.th);
        for (int k = 0; k < S.nk[0]; k++) slot_owner.push_back(S.id_of(matched[(size_t)k]));
    } else if (which == 4) {
        std::vector<cv::KeyPoint> mk1, mk2; std::vector<std::pair<size_t, size_t> > pairs;
        ret[0] = matcher.SearchForTriangulation(&S.K[0], &S.K[1], S.F12, mk1, mk2, pairs);
        slot_owner.assign((size_t)S.nk[0], -1);
        for (size_t i = 0; i < pairs.size(); i++) {
            slot_owner[pairs[i].first] = (int32_t)pairs[i].second;
            replaced.push_back((int32_t)pairs[i].first); replaced.push_back((int32_t)pairs[i].second);
            // the returned keypoints must be the ones the pairs name
            obs_slot.push_back(mk1[i].pt.x == S.K[0].keysUn[pairs[i].first].pt.x && mk2[i].pt.y == S.K[1].keysUn[pairs[i].second].pt.y ? 1 : 0);
        }
    } else {
        const int32_t* mp1 = in.get<int32_t>(K1_MP);
        std::vector<USLAM::MapPoint*> m12((size_t)S.nk[0], (USLAM::MapPoint*)0);
        for (int k = 0; k < S.nk[0]; k++) if (mp1[3 * k + 2] >= 0 && S.K[1].mapPoints[(size_t)mp1[3 * k + 2]]) m12[(size_t)k] = S.K[1].mapPoints[(size_t)mp1[3 * k + 2]];
        ret[0] = matcher.SearchBySim3(&S.K[0], &S.K[1], m12, S.s12, S.R12, S.t12, S.th);
        for (int k = 0; k < S.nk[0]; k++) slot_owner.push_back(S.id_of(m12[(size_t)k]));
    }
    out.put(ret); out.put(slot_owner); out.put(replaced); out.put(obs_slot);
    return ret[0];
}

}  // namespace m8
#endif
