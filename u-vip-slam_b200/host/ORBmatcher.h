// ORBmatcher.h — drop-in for the descriptor path of the reference's include/ORBmatcher.h (chintha/U-VIP-SLAM, :41-94).
// Same class name, constructor, constants and call signatures; the pointer-rich containers of the reference
// (FrameKTL, MapPoint) are GATHERED into flat arrays, sent through the C-ABI (include/uvip_orb.h), and the results
// SCATTERED back, so Tracking / LocalMapping / LoopClosing stay untouched.  The member templates accept
// the reference's own FrameKTL / MapPoint types (they only use the members cited below), or mock types in tests.
#pragma once
#include <climits>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#ifdef UVIP_WITH_OPENCV
#include <opencv2/core/core.hpp>
#else
#include "uvip_compat.h"
#endif
#include "../../include/uvip_orb.h"

namespace USLAM {

class ORBmatcher {
public:
    // The reference builds a stack-local ORBmatcher in front of every search (src/Tracking.cc:2222,2387,2426,3015,
    // src/LocalMapping.cc:1080,1230, src/LoopClosing.cc:373,454,699), from three host threads.  Construction and destruction
    // therefore cost nothing here: the device handle (stream, scratch buffers, pinned staging) belongs to the CALLING THREAD
    // (thread_local, created at the thread's first search, destroyed when the thread exits) and is looked up per call, so an
    // instance can also be handed to another thread.
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    ~ORBmatcher() {}
    // device the calling thread's handle is created on (default 0); call before the thread's first search
    static void SetDevice(int device) { tls().device = device; }

    // src/ORBmatcher.cc:1794-1810.  One pair of 32-byte rows: stays a host popcount, exactly as SURVEY section 2 row 4
    // prescribes for MapPoint::ComputeDistinctiveDescriptors (N is tiny); batches go through uvip_descriptor_distance.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
    {
        const unsigned char* pa = a.ptr<unsigned char>(); const unsigned char* pb = b.ptr<unsigned char>();
        int dist = 0;
        for (int i = 0; i < 8; i++) {
            uint32_t x, y; std::memcpy(&x, pa + 4 * i, 4); std::memcpy(&y, pb + 4 * i, 4);
            dist += __builtin_popcount(x ^ y);
        }
        return dist;
    }

    // SearchByProjection(FrameKTL&, const vector<MapPoint*>&, th)  (src/ORBmatcher.cc:49-125)
    // FrameT needs: mvKeysUn, mDescriptors, mvpMapPoints, mvScaleFactors, mnMinX, mnMinY, mfGridElementWidthInv,
    //               mfGridElementHeightInv (include/FrameKTL.h);  MapPointT needs: mbTrackInView, isBad(), mnTrackScaleLevel,
    //               mTrackViewCos, mTrackProjX, mTrackProjY, GetDescriptor() (include/MapPoint.h)
    template <class FrameT, class MapPointT>
    int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th = 3)
    {
        ensure();
        const bool bFactor = th != 1.0;
        std::vector<float> qu, qv, qr; std::vector<int32_t> qmin, qmax; std::vector<unsigned char> qd; std::vector<MapPointT*> who;
        for (size_t i = 0; i < vpMapPoints.size(); i++) {
            MapPointT* pMP = vpMapPoints[i];
            if (!pMP->mbTrackInView) continue;
            if (pMP->isBad()) continue;
            const int lvl = pMP->mnTrackScaleLevel;
            float r = uvip_radius_by_viewing_cos(pMP->mTrackViewCos);
            if (bFactor) r *= th;
            qu.push_back(pMP->mTrackProjX); qv.push_back(pMP->mTrackProjY); qr.push_back(r * F.mvScaleFactors[lvl]);
            qmin.push_back(lvl - 1); qmax.push_back(lvl);
            const cv::Mat d = pMP->GetDescriptor();
            qd.insert(qd.end(), d.ptr(0), d.ptr(0) + 32);
            who.push_back(pMP);
        }
        const int nq = (int)who.size(), nk = (int)F.mvKeysUn.size();
        if (nq == 0) return 0;
        std::vector<float> kx((size_t)nk), ky((size_t)nk); std::vector<int32_t> oct((size_t)nk), taken((size_t)nk);
        std::vector<unsigned char> kd((size_t)nk * 32);
        for (int i = 0; i < nk; i++) {
            kx[i] = F.mvKeysUn[i].pt.x; ky[i] = F.mvKeysUn[i].pt.y; oct[i] = F.mvKeysUn[i].octave;
            taken[i] = F.mvpMapPoints[i] ? -2 : -1;
            std::memcpy(&kd[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
        }
        uvip_search_params sp;
        sp.mode = 0; sp.th_dist = TH_HIGH; sp.ratio = mfNNratio;
        sp.min_x = (float)F.mnMinX; sp.min_y = (float)F.mnMinY; sp.inv_w = F.mfGridElementWidthInv; sp.inv_h = F.mfGridElementHeightInv;
        sp.cols = 64; sp.rows = 48;
        std::vector<int32_t> match((size_t)nq);
        int nmatches = 0;
        check(uvip_search_frame(handle_, &sp, qu.data(), qv.data(), qr.data(), qmin.data(), qmax.data(), qd.data(), nq,
                                kx.data(), ky.data(), oct.data(), kd.data(), nk, taken.data(), match.data(), &nmatches), "uvip_search_frame");
        for (int q = 0; q < nq; q++) if (match[q] >= 0) F.mvpMapPoints[match[q]] = who[q];     // :119
        return nmatches;
    }

    // SearchByProjection(FrameKTL& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
    // (src/ORBmatcher.cc:1622-1746; relocalisation and keyframe tracking, src/Tracking.cc:2480,2494,3020,3026).  The host
    // side projects the keyframe's map points with the current pose exactly as :1650-1678 does (Rcw*x+tcw through
    // cv::gemm's small-matrix float path; -Rcw.t()*tcw and cv::norm accumulate in double), the window search + claims run in
    // uvip_search_window mode 1 and the rotation histogram in uvip_rot_hist_filter.
    // FrameT needs: mTcw (4x4 CV_32F), fx, fy, cx, cy, mnMinX/MaxX/MinY/MaxY, mvScaleFactors, mnScaleLevels, mvKeysUn,
    //               mDescriptors, mvpMapPoints, mfGridElement{Width,Height}Inv;  KeyFrameT: GetMapPointMatches(),
    //               GetKeyPointUn(i);  MapPointT: isBad(), GetWorldPos() (3x1 CV_32F), GetMinDistanceInvariance(), GetDescriptor()
    template <class FrameT, class KeyFrameT, class MapPointT>
    int SearchByProjection(FrameT& CurrentFrame, KeyFrameT* pKF, const std::set<MapPointT*>& sAlreadyFound, const float th, const int ORBdist)
    {
        ensure();
        float R[3][3], t[3], Ow[3];
        for (int r = 0; r < 3; r++) {
            const float* row = CurrentFrame.mTcw.template ptr<float>(r);
            for (int c = 0; c < 3; c++) R[r][c] = row[c];
            t[r] = row[3];
        }
        for (int c = 0; c < 3; c++) {                                              // Ow = -Rcw.t()*tcw  (:1628)
            double s = 0; for (int r = 0; r < 3; r++) s += (double)R[r][c] * (double)t[r];
            Ow[c] = (float)(-1.0 * s);
        }
        const std::vector<MapPointT*> vpMPs = pKF->GetMapPointMatches();
        std::vector<float> qu, qv, qr, qa; std::vector<int32_t> qmin, qmax; std::vector<unsigned char> qd; std::vector<MapPointT*> who;
        for (size_t i = 0; i < vpMPs.size(); i++) {
            MapPointT* pMP = vpMPs[i];
            if (!pMP) continue;
            if (pMP->isBad() || sAlreadyFound.count(pMP)) continue;
            const cv::Mat x3Dw = pMP->GetWorldPos();
            float X[3], xc3[3];
            for (int r = 0; r < 3; r++) X[r] = x3Dw.template ptr<float>(r)[0];
            for (int r = 0; r < 3; r++) {                                          // x3Dc = Rcw*x3Dw+tcw  (:1651)
                // cv::gemm's 3x3 path (core/src/matmul.cpp): float products and sums, then (float)(t*alpha + c*beta) in double
                float s = R[r][0] * X[0];
                s = s + R[r][1] * X[1];
                s = s + R[r][2] * X[2];
                xc3[r] = (float)((double)s + (double)t[r]);
            }
            const float xc = xc3[0], yc = xc3[1];
            const float invzc = (float)(1.0 / xc3[2]);
            const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
            const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
            if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            const float minDistance = pMP->GetMinDistanceInvariance();              // predicted scale level (:1666-1672)
            double n2 = 0; for (int r = 0; r < 3; r++) { const float po = X[r] - Ow[r]; n2 += (double)po * (double)po; }
            const float dist3D = (float)std::sqrt(n2);
            const float ratio = dist3D / minDistance;
            const std::vector<float>& sf = CurrentFrame.mvScaleFactors;
            const int lvl = std::min((int)(std::lower_bound(sf.begin(), sf.end(), ratio) - sf.begin()), (int)CurrentFrame.mnScaleLevels - 1);
            qu.push_back(u); qv.push_back(v); qr.push_back(th * sf[(size_t)lvl]);
            qmin.push_back(lvl - 1); qmax.push_back(lvl + 1);
            const cv::Mat d = pMP->GetDescriptor();
            qd.insert(qd.end(), d.ptr(0), d.ptr(0) + 32);
            qa.push_back(pKF->GetKeyPointUn(i).angle);
            who.push_back(pMP);
        }
        const int nq = (int)who.size(), nk = (int)CurrentFrame.mvKeysUn.size();
        if (nq == 0 || nk == 0) return 0;
        std::vector<float> kx((size_t)nk), ky((size_t)nk), ka((size_t)nk); std::vector<int32_t> oct((size_t)nk), taken((size_t)nk);
        std::vector<unsigned char> kd((size_t)nk * 32);
        for (int i = 0; i < nk; i++) {
            kx[i] = CurrentFrame.mvKeysUn[i].pt.x; ky[i] = CurrentFrame.mvKeysUn[i].pt.y; oct[i] = CurrentFrame.mvKeysUn[i].octave;
            ka[i] = CurrentFrame.mvKeysUn[i].angle;
            taken[i] = CurrentFrame.mvpMapPoints[i] ? -2 : -1;
            std::memcpy(&kd[(size_t)i * 32], CurrentFrame.mDescriptors.ptr(i), 32);
        }
        uvip_search_params sp;
        sp.mode = 1; sp.th_dist = ORBdist; sp.ratio = mfNNratio;
        sp.min_x = (float)CurrentFrame.mnMinX; sp.min_y = (float)CurrentFrame.mnMinY;
        sp.inv_w = CurrentFrame.mfGridElementWidthInv; sp.inv_h = CurrentFrame.mfGridElementHeightInv;
        sp.cols = 64; sp.rows = 48;
        std::vector<int32_t> match((size_t)nq);
        int nmatches = 0;
        check(uvip_search_frame(handle_, &sp, qu.data(), qv.data(), qr.data(), qmin.data(), qmax.data(), qd.data(), nq,
                                kx.data(), ky.data(), oct.data(), kd.data(), nk, taken.data(), match.data(), &nmatches), "uvip_search_frame");
        for (int q = 0; q < nq; q++) if (match[q] >= 0) CurrentFrame.mvpMapPoints[match[q]] = who[q];      // :1698
        if (mbCheckOrientation) {                                                                          // :1701-1743
            std::vector<int32_t> kept(match);
            check(uvip_rot_hist_filter(handle_, kept.data(), nq, qa.data(), ka.data(), &nmatches), "uvip_rot_hist_filter");
            for (int q = 0; q < nq; q++) if (match[q] >= 0 && kept[q] < 0) CurrentFrame.mvpMapPoints[match[q]] = static_cast<MapPointT*>(NULL);
        }
        return nmatches;
    }

    // ------------------------------------------------------------------------------------------------------------------
    // The four public overloads of include/ORBmatcher.h:49-72 that no caller of this fork reaches (inherited from ORB-SLAM): kept so
    // that the header swap drops no member.  WindowSearch and SearchByProjection(F1, F2, ...) run in uvip_search_frame mode 6 (top-2
    // without levels), SearchByProjection(CurrentFrame, LastFrame, th) in mode 1; SearchForInitialization lets a later keypoint STEAL an
    // earlier match (:641-665), a dependency the claim table cannot express: its candidates come from the frame's own
    // GetFeaturesInArea, all candidate distances from one uvip_descriptor_distance call, and the sequential bookkeeping is replayed here.
    // ------------------------------------------------------------------------------------------------------------------
protected:
    // flattened keypoints of a frame for uvip_search_frame
    template <class FrameT>
    struct FlatFrame {
        std::vector<float> kx, ky, ka; std::vector<int32_t> oct; std::vector<unsigned char> kd;
        explicit FlatFrame(const FrameT& F)
        {
            const int nk = (int)F.mvKeysUn.size();
            kx.resize((size_t)nk); ky.resize((size_t)nk); ka.resize((size_t)nk); oct.resize((size_t)nk); kd.resize((size_t)nk * 32 + 32);
            for (int i = 0; i < nk; i++) {
                kx[(size_t)i] = F.mvKeysUn[(size_t)i].pt.x; ky[(size_t)i] = F.mvKeysUn[(size_t)i].pt.y; ka[(size_t)i] = F.mvKeysUn[(size_t)i].angle;
                oct[(size_t)i] = F.mvKeysUn[(size_t)i].octave;
                std::memcpy(&kd[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
            }
        }
    };
    template <class FrameT>
    uvip_search_params frame_params(const FrameT& F, int mode, int th_dist)
    {
        uvip_search_params sp;
        sp.mode = mode; sp.th_dist = th_dist; sp.ratio = mfNNratio;
        sp.min_x = (float)F.mnMinX; sp.min_y = (float)F.mnMinY; sp.inv_w = F.mfGridElementWidthInv; sp.inv_h = F.mfGridElementHeightInv;
        sp.cols = 64; sp.rows = 48;
        return sp;
    }

public:
    // WindowSearch(F1, F2, windowSize, vpMapPointMatches2, minOctave, maxOctave)  (src/ORBmatcher.cc:409-516)
    template <class FrameT, class MapPointT>
    int WindowSearch(FrameT& F1, FrameT& F2, int windowSize, std::vector<MapPointT*>& vpMapPointMatches2, int minScaleLevel = -1, int maxScaleLevel = INT_MAX)
    {
        ensure();
        vpMapPointMatches2 = std::vector<MapPointT*>(F2.mvpMapPoints.size(), static_cast<MapPointT*>(NULL));
        const bool bMinLevel = minScaleLevel > 0, bMaxLevel = maxScaleLevel < INT_MAX;
        std::vector<float> qu, qv, qr, qa; std::vector<int32_t> ql; std::vector<unsigned char> qd; std::vector<MapPointT*> who;
        for (size_t i1 = 0; i1 < F1.mvpMapPoints.size(); i1++) {
            MapPointT* pMP1 = F1.mvpMapPoints[i1];
            if (!pMP1) continue;
            if (pMP1->isBad()) continue;
            const cv::KeyPoint& kp1 = F1.mvKeysUn[i1];
            const int level1 = kp1.octave;
            if (bMinLevel && level1 < minScaleLevel) continue;
            if (bMaxLevel && level1 > maxScaleLevel) continue;
            qu.push_back(kp1.pt.x); qv.push_back(kp1.pt.y); qr.push_back((float)windowSize); ql.push_back(level1); qa.push_back(kp1.angle);
            qd.insert(qd.end(), F1.mDescriptors.ptr((int)i1), F1.mDescriptors.ptr((int)i1) + 32);
            who.push_back(pMP1);
        }
        const int nq = (int)who.size(), nk = (int)F2.mvKeysUn.size();
        if (nq == 0 || nk == 0) return 0;
        const FlatFrame<FrameT> K(F2);
        std::vector<int32_t> taken((size_t)nk, -1), match((size_t)nq);
        uvip_search_params sp = frame_params(F2, 6, TH_HIGH);
        int nmatches = 0;
        check(uvip_search_frame(handle_, &sp, qu.data(), qv.data(), qr.data(), ql.data(), ql.data(), qd.data(), nq, K.kx.data(), K.ky.data(), K.oct.data(),
                                K.kd.data(), nk, taken.data(), match.data(), &nmatches), "uvip_search_frame");
        if (mbCheckOrientation) check(uvip_rot_hist_filter(handle_, match.data(), nq, qa.data(), K.ka.data(), &nmatches), "uvip_rot_hist_filter");   // :487-513
        for (int q = 0; q < nq; q++) if (match[(size_t)q] >= 0) vpMapPointMatches2[(size_t)match[(size_t)q]] = who[(size_t)q];
        return nmatches;
    }

    // SearchByProjection(F1, F2, windowSize, vpMapPointMatches2)  (src/ORBmatcher.cc:519-596): F1's map points projected with F2's pose
    template <class FrameT, class MapPointT>
    int SearchByProjection(FrameT& F1, FrameT& F2, int windowSize, std::vector<MapPointT*>& vpMapPointMatches2)
    {
        ensure();
        vpMapPointMatches2 = F2.mvpMapPoints;
        const std::set<MapPointT*> spMapPointsAlreadyFound(vpMapPointMatches2.begin(), vpMapPointMatches2.end());
        float R[3][3], t[3];
        for (int r = 0; r < 3; r++) { const float* row = F2.mTcw.template ptr<float>(r); for (int c = 0; c < 3; c++) R[r][c] = row[c]; t[r] = row[3]; }
        std::vector<float> qu, qv, qr; std::vector<int32_t> ql; std::vector<unsigned char> qd; std::vector<MapPointT*> who;
        for (size_t i1 = 0; i1 < F1.mvpMapPoints.size(); i1++) {
            MapPointT* pMP1 = F1.mvpMapPoints[i1];
            if (!pMP1) continue;
            if (pMP1->isBad() || spMapPointsAlreadyFound.count(pMP1)) continue;
            const int level1 = F1.mvKeysUn[i1].octave;
            float X[3], c3[3]; read3(pMP1->GetWorldPos(), X); mul_add3(R, X, t, c3);
            const float invz = (float)(1.0 / c3[2]);
            const float u2 = F2.fx * c3[0] * invz + F2.cx, v2 = F2.fy * c3[1] * invz + F2.cy;
            qu.push_back(u2); qv.push_back(v2); qr.push_back((float)windowSize); ql.push_back(level1);
            qd.insert(qd.end(), F1.mDescriptors.ptr((int)i1), F1.mDescriptors.ptr((int)i1) + 32);
            who.push_back(pMP1);
        }
        const int nq = (int)who.size(), nk = (int)F2.mvKeysUn.size();
        if (nq == 0 || nk == 0) return 0;
        const FlatFrame<FrameT> K(F2);
        std::vector<int32_t> taken((size_t)nk), match((size_t)nq);
        for (int i = 0; i < nk; i++) taken[(size_t)i] = vpMapPointMatches2[(size_t)i] ? -2 : -1;
        uvip_search_params sp = frame_params(F2, 6, TH_HIGH);
        int nmatches = 0;
        check(uvip_search_frame(handle_, &sp, qu.data(), qv.data(), qr.data(), ql.data(), ql.data(), qd.data(), nq, K.kx.data(), K.ky.data(), K.oct.data(),
                                K.kd.data(), nk, taken.data(), match.data(), &nmatches), "uvip_search_frame");
        for (int q = 0; q < nq; q++) if (match[(size_t)q] >= 0) vpMapPointMatches2[(size_t)match[(size_t)q]] = who[(size_t)q];
        return nmatches;
    }

    // SearchByProjection(CurrentFrame, LastFrame, th)  (src/ORBmatcher.cc:1507-1620): the last frame's map points, best-only, rotation histogram
    template <class FrameT>
    int SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th)
    {
        ensure();
        typedef typename std::remove_pointer<typename std::remove_reference<decltype(CurrentFrame.mvpMapPoints[0])>::type>::type MapPointT;
        float R[3][3], t[3];
        for (int r = 0; r < 3; r++) { const float* row = CurrentFrame.mTcw.template ptr<float>(r); for (int c = 0; c < 3; c++) R[r][c] = row[c]; t[r] = row[3]; }
        std::vector<float> qu, qv, qr, qa; std::vector<int32_t> qmin, qmax; std::vector<unsigned char> qd; std::vector<MapPointT*> who;
        for (size_t i = 0; i < LastFrame.mvpMapPoints.size(); i++) {
            MapPointT* pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            float X[3], c3[3]; read3(pMP->GetWorldPos(), X); mul_add3(R, X, t, c3);
            const float invzc = (float)(1.0 / c3[2]);
            const float u = CurrentFrame.fx * c3[0] * invzc + CurrentFrame.cx, v = CurrentFrame.fy * c3[1] * invzc + CurrentFrame.cy;
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
            if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            const int nPredictedOctave = LastFrame.mvKeys[i].octave;
            qu.push_back(u); qv.push_back(v); qr.push_back(th * CurrentFrame.mvScaleFactors[(size_t)nPredictedOctave]);
            qmin.push_back(nPredictedOctave - 1); qmax.push_back(nPredictedOctave + 1);
            qd.insert(qd.end(), LastFrame.mDescriptors.ptr((int)i), LastFrame.mDescriptors.ptr((int)i) + 32);
            qa.push_back(LastFrame.mvKeysUn[i].angle);
            who.push_back(pMP);
        }
        const int nq = (int)who.size(), nk = (int)CurrentFrame.mvKeysUn.size();
        if (nq == 0 || nk == 0) return 0;
        const FlatFrame<FrameT> K(CurrentFrame);
        std::vector<int32_t> taken((size_t)nk), match((size_t)nq);
        for (int i = 0; i < nk; i++) taken[(size_t)i] = CurrentFrame.mvpMapPoints[(size_t)i] ? -2 : -1;
        uvip_search_params sp = frame_params(CurrentFrame, 1, TH_HIGH);
        int nmatches = 0;
        check(uvip_search_frame(handle_, &sp, qu.data(), qv.data(), qr.data(), qmin.data(), qmax.data(), qd.data(), nq, K.kx.data(), K.ky.data(), K.oct.data(),
                                K.kd.data(), nk, taken.data(), match.data(), &nmatches), "uvip_search_frame");
        for (int q = 0; q < nq; q++) if (match[(size_t)q] >= 0) CurrentFrame.mvpMapPoints[(size_t)match[(size_t)q]] = who[(size_t)q];
        if (mbCheckOrientation) {
            std::vector<int32_t> kept(match);
            check(uvip_rot_hist_filter(handle_, kept.data(), nq, qa.data(), K.ka.data(), &nmatches), "uvip_rot_hist_filter");
            for (int q = 0; q < nq; q++) if (match[(size_t)q] >= 0 && kept[(size_t)q] < 0) CurrentFrame.mvpMapPoints[(size_t)match[(size_t)q]] = static_cast<MapPointT*>(NULL);
        }
        return nmatches;
    }

    // SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  (src/ORBmatcher.cc:598-713)
    template <class FrameT>
    int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10)
    {
        ensure();
        const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
        vnMatches12 = std::vector<int>((size_t)n1, -1);
        // candidates of every level-0 keypoint, in the order the reference's loop meets them, and their distances in one device call
        std::vector<int32_t> cs(1, 0), ci; std::vector<int> qi;
        for (int i1 = 0; i1 < n1; i1++) {
            if (F1.mvKeysUn[(size_t)i1].octave > 0) continue;
            const std::vector<size_t> vIndices2 = F2.GetFeaturesInArea(vbPrevMatched[(size_t)i1].x, vbPrevMatched[(size_t)i1].y, (float)windowSize, 0, 0);
            if (vIndices2.empty()) continue;
            for (size_t k = 0; k < vIndices2.size(); k++) ci.push_back((int32_t)vIndices2[k]);
            cs.push_back((int32_t)ci.size()); qi.push_back(i1);
        }
        const int npairs = (int)ci.size();
        std::vector<int32_t> dist((size_t)npairs + 1);
        if (npairs) {
            std::vector<unsigned char> a((size_t)npairs * 32), b((size_t)npairs * 32);
            for (size_t q = 0; q < qi.size(); q++)
                for (int j = cs[q]; j < cs[q + 1]; j++) {
                    std::memcpy(&a[(size_t)j * 32], F1.mDescriptors.ptr(qi[q]), 32);
                    std::memcpy(&b[(size_t)j * 32], F2.mDescriptors.ptr(ci[(size_t)j]), 32);
                }
            check(uvip_descriptor_distance(handle_, a.data(), b.data(), npairs, dist.data()), "uvip_descriptor_distance");
        }
        // the sequential part, as written at :620-680 (a later keypoint with a strictly smaller distance steals the match)
        int nmatches = 0;
        std::vector<int> vMatchedDistance((size_t)n2, INT_MAX), vnMatches21((size_t)n2, -1), accepted;   // accepted: i1 in rotHist push order
        std::vector<float> a1, a2;                                                        // angle pair of every accepted match, at acceptance
        for (size_t q = 0; q < qi.size(); q++) {
            const int i1 = qi[q];
            int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
            for (int j = cs[q]; j < cs[q + 1]; j++) {
                const int i2 = ci[(size_t)j], d = dist[(size_t)j];
                if (vMatchedDistance[(size_t)i2] <= d) continue;
                if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
                else if (d < bestDist2) bestDist2 = d;
            }
            if (bestDist <= TH_LOW && bestDist < (float)bestDist2 * mfNNratio) {
                if (vnMatches21[(size_t)bestIdx2] >= 0) { vnMatches12[(size_t)vnMatches21[(size_t)bestIdx2]] = -1; nmatches--; }
                vnMatches12[(size_t)i1] = bestIdx2; vnMatches21[(size_t)bestIdx2] = i1; vMatchedDistance[(size_t)bestIdx2] = bestDist;
                nmatches++;
                if (mbCheckOrientation) { accepted.push_back(i1); a1.push_back(F1.mvKeysUn[(size_t)i1].angle); a2.push_back(F2.mvKeysUn[(size_t)bestIdx2].angle); }
            }
        }
        if (mbCheckOrientation && !accepted.empty()) {
            // the histogram counts every ACCEPTED pair, stolen ones included (their bin entry stays, :667-676); only matches that are
            // still valid are removed afterwards (:694-698)
            std::vector<int32_t> m(accepted.size());
            for (size_t k = 0; k < m.size(); k++) m[k] = (int32_t)k;
            int kept = 0;
            check(uvip_rot_hist_filter(handle_, m.data(), (int)m.size(), a1.data(), a2.data(), &kept), "uvip_rot_hist_filter");
            for (size_t j = 0; j < accepted.size(); j++)
                if (m[j] < 0 && vnMatches12[(size_t)accepted[j]] >= 0) { vnMatches12[(size_t)accepted[j]] = -1; nmatches--; }
        }
        for (size_t i1 = 0; i1 < vnMatches12.size(); i1++)                                // :708-710
            if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[(size_t)vnMatches12[i1]].pt;
        return nmatches;
    }

    // SearchByBoW(KeyFrame*, FrameKTL&, vector<MapPoint*>&)  (src/ORBmatcher.cc:155-284): brute force restricted to features
    // of the same vocabulary node, TH_LOW + ratio, claims in the reference's node-major order, rotation histogram.
    // KeyFrameT needs: GetMapPointMatches(), GetFeatureVector() (a std::map<node id, vector<unsigned>>), GetDescriptor(i),
    //                  GetKeyPointUn(i);  FrameT needs: mvpMapPoints, mFeatVec, mDescriptors, mvKeys
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches)
    {
        ensure();
        const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPointT*>(F.mvpMapPoints.size(), static_cast<MapPointT*>(NULL));
        const auto vFeatVecKF = pKF->GetFeatureVector();
        std::vector<unsigned char> qd; std::vector<int32_t> cs(1, 0), ci; std::vector<unsigned> qidx;
        auto KFit = vFeatVecKF.begin(); auto Fit = F.mFeatVec.begin();
        while (KFit != vFeatVecKF.end() && Fit != F.mFeatVec.end()) {          // merge-join by node id (:176-252)
            if (KFit->first == Fit->first) {
                for (size_t iKF = 0; iKF < KFit->second.size(); iKF++) {
                    const unsigned realIdxKF = KFit->second[iKF];
                    MapPointT* pMP = vpMapPointsKF[realIdxKF];
                    if (!pMP) continue;
                    if (pMP->isBad()) continue;
                    const cv::Mat dKF = pKF->GetDescriptor(realIdxKF);
                    qd.insert(qd.end(), dKF.ptr(0), dKF.ptr(0) + 32);
                    for (size_t iF = 0; iF < Fit->second.size(); iF++) ci.push_back((int32_t)Fit->second[iF]);
                    cs.push_back((int32_t)ci.size());
                    qidx.push_back(realIdxKF);
                }
                ++KFit; ++Fit;
            } else if (KFit->first < Fit->first) KFit = vFeatVecKF.lower_bound(Fit->first);
            else Fit = F.mFeatVec.lower_bound(KFit->first);
        }
        const int nq = (int)qidx.size(), nk = (int)F.mvpMapPoints.size();
        if (nq == 0 || nk == 0) return 0;
        std::vector<unsigned char> kd((size_t)nk * 32);
        for (int i = 0; i < nk; i++) std::memcpy(&kd[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
        std::vector<int32_t> taken((size_t)nk, -1), match((size_t)nq);
        int nmatches = 0;
        if (ci.empty()) ci.push_back(0);
        check(uvip_search_lists(handle_, 2, TH_LOW, mfNNratio, qd.data(), nq, cs.data(), ci.data(), kd.data(), nk, taken.data(), match.data(), &nmatches),
              "uvip_search_lists");
        if (mbCheckOrientation) {                                               // :227-241, :255-281
            std::vector<float> a1((size_t)nq), a2((size_t)nk);
            for (int q = 0; q < nq; q++) a1[q] = pKF->GetKeyPointUn(qidx[q]).angle;
            for (int i = 0; i < nk; i++) a2[i] = F.mvKeys[i].angle;
            check(uvip_rot_hist_filter(handle_, match.data(), nq, a1.data(), a2.data(), &nmatches), "uvip_rot_hist_filter");
        }
        for (int q = 0; q < nq; q++) if (match[q] >= 0) vpMapPointMatches[match[q]] = vpMapPointsKF[qidx[q]];
        return nmatches;
    }

    // SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)  (src/ORBmatcher.cc:715-850, loop closing
    // src/LoopClosing.cc:399): same node-restricted top-2 as above between two keyframes, strict best < TH_LOW, ratio, claims on
    // the second keyframe's features, rotation histogram.  Candidates without a (good) map point are dropped while the
    // lists are built (:766-772).  KeyFrameT needs: GetKeyPointsUn(), GetFeatureVector(), GetMapPointMatches(), GetDescriptors()
    template <class KeyFrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12)
    {
        ensure();
        const std::vector<cv::KeyPoint> vKeysUn1 = pKF1->GetKeyPointsUn(), vKeysUn2 = pKF2->GetKeyPointsUn();
        const auto vFeatVec1 = pKF1->GetFeatureVector(); const auto vFeatVec2 = pKF2->GetFeatureVector();
        const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const cv::Mat Descriptors1 = pKF1->GetDescriptors(), Descriptors2 = pKF2->GetDescriptors();
        vpMatches12 = std::vector<MapPointT*>(vpMapPoints1.size(), static_cast<MapPointT*>(NULL));
        std::vector<unsigned char> qd; std::vector<int32_t> cs(1, 0), ci; std::vector<unsigned> qidx;
        auto f1it = vFeatVec1.begin(); auto f2it = vFeatVec2.begin();
        while (f1it != vFeatVec1.end() && f2it != vFeatVec2.end()) {
            if (f1it->first == f2it->first) {
                for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                    const unsigned idx1 = f1it->second[i1];
                    MapPointT* pMP1 = vpMapPoints1[idx1];
                    if (!pMP1) continue;
                    if (pMP1->isBad()) continue;
                    qd.insert(qd.end(), Descriptors1.ptr((int)idx1), Descriptors1.ptr((int)idx1) + 32);
                    for (size_t i2 = 0; i2 < f2it->second.size(); i2++) {
                        const unsigned idx2 = f2it->second[i2];
                        MapPointT* pMP2 = vpMapPoints2[idx2];
                        if (!pMP2 || pMP2->isBad()) continue;
                        ci.push_back((int32_t)idx2);
                    }
                    cs.push_back((int32_t)ci.size());
                    qidx.push_back(idx1);
                }
                ++f1it; ++f2it;
            } else if (f1it->first < f2it->first) f1it = vFeatVec1.lower_bound(f2it->first);
            else f2it = vFeatVec2.lower_bound(f1it->first);
        }
        const int nq = (int)qidx.size(), nk = (int)vpMapPoints2.size();
        if (nq == 0 || nk == 0) return 0;
        std::vector<unsigned char> kd((size_t)nk * 32);
        for (int i = 0; i < nk; i++) std::memcpy(&kd[(size_t)i * 32], Descriptors2.ptr(i), 32);
        std::vector<int32_t> taken((size_t)nk, -1), match((size_t)nq);
        int nmatches = 0;
        if (ci.empty()) ci.push_back(0);
        check(uvip_search_lists(handle_, 3, TH_LOW, mfNNratio, qd.data(), nq, cs.data(), ci.data(), kd.data(), nk, taken.data(), match.data(), &nmatches),
              "uvip_search_lists");
        if (mbCheckOrientation) {
            std::vector<float> a1((size_t)nq), a2((size_t)nk);
            for (int q = 0; q < nq; q++) a1[q] = vKeysUn1[qidx[q]].angle;
            for (int i = 0; i < nk; i++) a2[i] = vKeysUn2[(size_t)i].angle;
            check(uvip_rot_hist_filter(handle_, match.data(), nq, a1.data(), a2.data(), &nmatches), "uvip_rot_hist_filter");
        }
        for (int q = 0; q < nq; q++) if (match[q] >= 0) vpMatches12[qidx[q]] = vpMapPoints2[(size_t)match[q]];
        return nmatches;
    }

    // ------------------------------------------------------------------------------------------------------------------
    // Keyframe-side searches (LocalMapping / LoopClosing threads).  The host side keeps the reference's
    // geometry — projection, depth / image / distance / viewing-angle gates, lower_bound level prediction — with cv::Mat
    // arithmetic as OpenCV 3.4 evaluates it (R*x+t: gemm's small-matrix float path; anything transposed and cv::norm / dot:
    // double accumulation; Mat*scalar: float product with the scalar rounded to float), and hands the window search to uvip_search_window:
    // KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:952-992) + levels [l-1, l] + best-only, with claims (mode 1) or without
    // (mode 4).  KeyFrameT needs: fx, fy, cx, cy, mnMinX, mnMinY (protected in the reference: add `friend class ORBmatcher;`
    // or two getters), mfGridElementWidthInv, mfGridElementHeightInv, GetKeyPointsUn(), GetDescriptors(),
    // GetScaleFactors(), GetScaleLevels(), IsInImage(u, v), GetMapPoint(i), AddMapPoint(p, i), GetMapPoints(),
    // GetMapPointMatches(), GetRotation(), GetTranslation(), GetCameraCenter();  MapPointT: isBad(), IsInKeyFrame(kf),
    // GetWorldPos(), GetNormal(), GetMin/MaxDistanceInvariance(), GetDescriptor(), Replace(p), AddObservation(kf, i),
    // GetIndexInKeyFrame(kf).
    // ------------------------------------------------------------------------------------------------------------------
protected:
    struct Pose3 { float R[3][3], t[3], Ow[3]; };
    static void read3x3(const cv::Mat& m, float R[3][3]) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = m.template ptr<float>(r)[c]; }
    static void read3(const cv::Mat& m, float v[3]) { for (int r = 0; r < 3; r++) v[r] = m.template ptr<float>(r)[0]; }
    // Scw -> Rcw, tcw, Ow  (src/ORBmatcher.cc:298-302, :1145-1149)
    static void decompose_sim3(const cv::Mat& Scw, Pose3& P)
    {
        float s[3][4];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) s[r][c] = Scw.template ptr<float>(r)[c];
        double d = 0; for (int c = 0; c < 3; c++) d += (double)s[0][c] * (double)s[0][c];       // sRcw.row(0).dot(sRcw.row(0))
        const float scw = (float)std::sqrt(d);
        const float inv = (float)(1.0 / scw);                        // Mat / scalar = convertTo(alpha = 1/scalar): float product for 32f (convert.cpp)
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) P.R[r][c] = s[r][c] * inv; P.t[r] = s[r][3] * inv; }
        camera_center(P);
    }
    static void camera_center(Pose3& P)                                                          // Ow = -Rcw.t()*tcw: generic gemm, double
    {
        for (int c = 0; c < 3; c++) { double a = 0; for (int r = 0; r < 3; r++) a += (double)P.R[r][c] * (double)P.t[r]; P.Ow[c] = (float)(a * -1.0); }
    }
    static void mul_add3(const float R[3][3], const float x[3], const float t[3], float out[3])  // R*x + t, gemm small-matrix path
    {
        for (int r = 0; r < 3; r++) { float a = R[r][0] * x[0]; a = a + R[r][1] * x[1]; a = a + R[r][2] * x[2]; out[r] = (float)((double)a + (double)t[r]); }
    }
    static float norm3(const float v[3]) { double a = 0; for (int r = 0; r < 3; r++) a += (double)v[r] * (double)v[r]; return (float)std::sqrt(a); }

    struct KfQueries {
        std::vector<float> u, v, r; std::vector<int32_t> lo, hi; std::vector<unsigned char> d; std::vector<int> who;
        void push(float u_, float v_, float r_, int lvl, const cv::Mat& desc, int index)
        { u.push_back(u_); v.push_back(v_); r.push_back(r_); lo.push_back(lvl - 1); hi.push_back(lvl); d.insert(d.end(), desc.ptr(0), desc.ptr(0) + 32); who.push_back(index); }
    };
    // the shared tail of :1036-1086 / :324-362 / :1173-1215: image, distance and viewing-angle gates, level, radius
    template <class KeyFrameT, class MapPointT>
    static bool gate_and_level(KeyFrameT* pKF, MapPointT* pMP, const float p3Dw[3], const float Ow[3], float u, float v, const std::vector<float>& sf,
                               bool normal_gate, float dist3D_given, bool use_given, int& level)
    {
        if (!pKF->IsInImage(u, v)) return false;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        const float dist3D = use_given ? dist3D_given : norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) return false;
        if (normal_gate) {
            float Pn[3]; read3(pMP->GetNormal(), Pn);
            double dot = 0; for (int r = 0; r < 3; r++) dot += (double)PO[r] * (double)Pn[r];
            if (dot < 0.5 * dist3D) return false;
        }
        const float ratio = dist3D / minDistance;
        level = std::min((int)(std::lower_bound(sf.begin(), sf.end(), ratio) - sf.begin()), (int)sf.size() - 1);
        return true;
    }
    // KeyFrame::GetFeaturesInArea window + levels [l-1, l] + best-only over a flattened keyframe; match[q] = keypoint or -1
    template <class KeyFrameT>
    int keyframe_window_search(KeyFrameT* pKF, int mode, int th_dist, const KfQueries& Q, std::vector<int32_t>& taken, std::vector<int32_t>& match)
    {
        const std::vector<cv::KeyPoint> keys = pKF->GetKeyPointsUn();
        const cv::Mat D = pKF->GetDescriptors();
        const int nk = (int)keys.size(), nq = (int)Q.who.size();
        match.assign((size_t)nq, -1);
        if (nq == 0 || nk == 0) return 0;
        std::vector<float> kx((size_t)nk), ky((size_t)nk); std::vector<int32_t> oct((size_t)nk); std::vector<unsigned char> kd((size_t)nk * 32);
        for (int i = 0; i < nk; i++) { kx[i] = keys[i].pt.x; ky[i] = keys[i].pt.y; oct[i] = keys[i].octave; std::memcpy(&kd[(size_t)i * 32], D.ptr(i), 32); }
        if (taken.size() != (size_t)nk) taken.assign((size_t)nk, -1);
        uvip_search_params sp;
        sp.mode = mode; sp.th_dist = th_dist; sp.ratio = mfNNratio;
        sp.min_x = (float)pKF->mnMinX; sp.min_y = (float)pKF->mnMinY; sp.inv_w = pKF->mfGridElementWidthInv; sp.inv_h = pKF->mfGridElementHeightInv;
        sp.cols = 64; sp.rows = 48;
        int n = 0;
        check(uvip_search_frame(handle_, &sp, Q.u.data(), Q.v.data(), Q.r.data(), Q.lo.data(), Q.hi.data(), Q.d.data(), nq, kx.data(), ky.data(), oct.data(),
                                kd.data(), nk, taken.data(), match.data(), &n), "uvip_search_frame");
        return n;
    }

public:
    // Fuse(KeyFrame* pKF, vector<MapPoint*>& vpMapPoints, float th)  (src/ORBmatcher.cc:1016-1134; src/LocalMapping.cc:1236,1261)
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, std::vector<MapPointT*>& vpMapPoints, const float th = 2.5f)   // include/ORBmatcher.h:85
    {
        ensure();
        Pose3 P; read3x3(pKF->GetRotation(), P.R); read3(pKF->GetTranslation(), P.t); read3(pKF->GetCameraCenter(), P.Ow);
        const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
        const std::vector<float> sf = pKF->GetScaleFactors();
        KfQueries Q;
        for (size_t i = 0; i < vpMapPoints.size(); i++) {
            MapPointT* pMP = vpMapPoints[i];
            if (!pMP) continue;
            if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
            float X[3], c3[3]; read3(pMP->GetWorldPos(), X); mul_add3(P.R, X, P.t, c3);
            if (c3[2] < 0.0f) continue;
            const float invz = 1 / c3[2], x = c3[0] * invz, y = c3[1] * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            int lvl;
            if (!gate_and_level(pKF, pMP, X, P.Ow, u, v, sf, true, 0.f, false, lvl)) continue;
            Q.push(u, v, th * sf[(size_t)lvl], lvl, pMP->GetDescriptor(), (int)i);
        }
        std::vector<int32_t> taken, match;
        keyframe_window_search(pKF, 4, TH_LOW, Q, taken, match);
        int nFused = 0;
        for (size_t q = 0; q < match.size(); q++) {                            // :1115-1129, in map-point order
            if (match[q] < 0) continue;
            MapPointT* pMP = vpMapPoints[(size_t)Q.who[q]];
            MapPointT* pMPinKF = pKF->GetMapPoint((size_t)match[q]);
            if (pMPinKF) { if (!pMPinKF->isBad()) pMP->Replace(pMPinKF); }
            else { pMP->AddObservation(pKF, (size_t)match[q]); pKF->AddMapPoint(pMP, (size_t)match[q]); }
            nFused++;
        }
        return nFused;
    }

    // Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th)  (src/ORBmatcher.cc:1136-1265; src/LoopClosing.cc:704)
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, float th = 2.5f)
    {
        ensure();
        Pose3 P; decompose_sim3(Scw, P);
        const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
        const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();
        const std::vector<float> sf = pKF->GetScaleFactors();
        KfQueries Q;
        for (size_t i = 0; i < vpPoints.size(); i++) {
            MapPointT* pMP = vpPoints[i];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            float X[3], c3[3]; read3(pMP->GetWorldPos(), X); mul_add3(P.R, X, P.t, c3);
            if (c3[2] < 0.0f) continue;
            const float invz = (float)(1.0 / c3[2]), x = c3[0] * invz, y = c3[1] * invz;          // :1180 `1.0/`
            const float u = fx * x + cx, v = fy * y + cy;
            int lvl;
            if (!gate_and_level(pKF, pMP, X, P.Ow, u, v, sf, true, 0.f, false, lvl)) continue;
            Q.push(u, v, th * sf[(size_t)lvl], lvl, pMP->GetDescriptor(), (int)i);
        }
        std::vector<int32_t> taken, match;
        keyframe_window_search(pKF, 4, TH_LOW, Q, taken, match);
        int nFused = 0;
        for (size_t q = 0; q < match.size(); q++) {                            // :1246-1260: here the KEYFRAME's point is replaced
            if (match[q] < 0) continue;
            MapPointT* pMP = vpPoints[(size_t)Q.who[q]];
            MapPointT* pMPinKF = pKF->GetMapPoint((size_t)match[q]);
            if (pMPinKF) { if (!pMPinKF->isBad()) pMPinKF->Replace(pMP); }
            else { pMP->AddObservation(pKF, (size_t)match[q]); pKF->AddMapPoint(pMP, (size_t)match[q]); }
            nFused++;
        }
        return nFused;
    }

    // SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)
    // (src/ORBmatcher.cc:286-407; src/LoopClosing.cc:512): claims on vpMatched, so later points skip taken keypoints (mode 1)
    template <class KeyFrameT, class MapPointT>
    int SearchByProjection(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th)
    {
        ensure();
        Pose3 P; decompose_sim3(Scw, P);
        const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
        std::set<MapPointT*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
        spAlreadyFound.erase(static_cast<MapPointT*>(NULL));
        const std::vector<float> sf = pKF->GetScaleFactors();
        KfQueries Q;
        for (size_t i = 0; i < vpPoints.size(); i++) {
            MapPointT* pMP = vpPoints[i];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            float X[3], c3[3]; read3(pMP->GetWorldPos(), X); mul_add3(P.R, X, P.t, c3);
            if (c3[2] < 0.0) continue;
            const float invz = 1 / c3[2], x = c3[0] * invz, y = c3[1] * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            int lvl;
            if (!gate_and_level(pKF, pMP, X, P.Ow, u, v, sf, true, 0.f, false, lvl)) continue;
            Q.push(u, v, th * sf[(size_t)lvl], lvl, pMP->GetDescriptor(), (int)i);
        }
        std::vector<int32_t> taken(vpMatched.size()), match;
        for (size_t k = 0; k < vpMatched.size(); k++) taken[k] = vpMatched[k] ? -2 : -1;
        const int nmatches = keyframe_window_search(pKF, 1, TH_LOW, Q, taken, match);
        for (size_t q = 0; q < match.size(); q++) if (match[q] >= 0) vpMatched[(size_t)match[q]] = vpPoints[(size_t)Q.who[q]];
        return nmatches;
    }

    // SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, s12, R12, t12, th)
    // (src/ORBmatcher.cc:1267-1505; src/LoopClosing.cc:458): two best-only projection searches (TH_HIGH, no claims) and
    // the mutual-agreement check
    template <class KeyFrameT, class MapPointT>
    int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12,
                     const float th)
    {
        ensure();
        const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
        float R1w[3][3], t1w[3], R2w[3][3], t2w[3], r12[3][3], T12[3], sR12[3][3], sR21[3][3], t21[3];
        read3x3(pKF1->GetRotation(), R1w); read3(pKF1->GetTranslation(), t1w);
        read3x3(pKF2->GetRotation(), R2w); read3(pKF2->GetTranslation(), t2w);
        read3x3(R12, r12); read3(t12, T12);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
            sR12[r][c] = r12[r][c] * s12;                                                        // s12*R12
            sR21[r][c] = r12[c][r] * (float)(1.0 / s12);                                         // (1.0/s12)*R12.t(): transpose, then float scale
        }
        {   // t21 = -sR21*t12: the negated matrix times t12 through the small-matrix float path
            const float zero[3] = {0.f, 0.f, 0.f}; float neg[3][3];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) neg[r][c] = -sR21[r][c];
            mul_add3(neg, T12, zero, t21);
        }
        const std::vector<float> sf1 = pKF1->GetScaleFactors(), sf2 = pKF2->GetScaleFactors();
        const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
        std::vector<bool> vbAlreadyMatched1((size_t)N1, false), vbAlreadyMatched2((size_t)N2, false);
        for (int i = 0; i < N1; i++) {
            MapPointT* pMP = vpMatches12[(size_t)i];
            if (pMP) {
                vbAlreadyMatched1[(size_t)i] = true;
                const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
                if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[(size_t)idx2] = true;
            }
        }
        const float origin[3] = {0.f, 0.f, 0.f};
        KfQueries Q12, Q21;
        for (int i1 = 0; i1 < N1; i1++) {                                      // KF1's points into KF2 (:1316-1395)
            MapPointT* pMP = vpMapPoints1[(size_t)i1];
            if (!pMP || vbAlreadyMatched1[(size_t)i1]) continue;
            if (pMP->isBad()) continue;
            float X[3], c1[3], c2[3]; read3(pMP->GetWorldPos(), X); mul_add3(R1w, X, t1w, c1); mul_add3(sR21, c1, t21, c2);
            if (c2[2] < 0.0) continue;
            const float invz = (float)(1.0 / c2[2]), x = c2[0] * invz, y = c2[1] * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            int lvl;
            if (!gate_and_level(pKF2, pMP, c2, origin, u, v, sf2, false, norm3(c2), true, lvl)) continue;
            Q12.push(u, v, th * sf2[(size_t)lvl], lvl, pMP->GetDescriptor(), i1);
        }
        for (int i2 = 0; i2 < N2; i2++) {                                      // KF2's points into KF1 (:1398-1476)
            MapPointT* pMP = vpMapPoints2[(size_t)i2];
            if (!pMP || vbAlreadyMatched2[(size_t)i2]) continue;
            if (pMP->isBad()) continue;
            float X[3], c2[3], c1[3]; read3(pMP->GetWorldPos(), X); mul_add3(R2w, X, t2w, c2); mul_add3(sR12, c2, T12, c1);
            if (c1[2] < 0.0) continue;
            const float invz = (float)(1.0 / c1[2]), x = c1[0] * invz, y = c1[1] * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            int lvl;
            if (!gate_and_level(pKF1, pMP, c1, origin, u, v, sf1, false, norm3(c1), true, lvl)) continue;
            Q21.push(u, v, th * sf1[(size_t)lvl], lvl, pMP->GetDescriptor(), i2);
        }
        std::vector<int32_t> taken, m12, m21;
        keyframe_window_search(pKF2, 4, TH_HIGH, Q12, taken, m12);
        taken.clear();
        keyframe_window_search(pKF1, 4, TH_HIGH, Q21, taken, m21);
        std::vector<int> vnMatch1((size_t)N1, -1), vnMatch2((size_t)N2, -1);
        for (size_t q = 0; q < m12.size(); q++) if (m12[q] >= 0) vnMatch1[(size_t)Q12.who[q]] = m12[q];
        for (size_t q = 0; q < m21.size(); q++) if (m21[q] >= 0) vnMatch2[(size_t)Q21.who[q]] = m21[q];
        int nFound = 0;
        for (int i1 = 0; i1 < N1; i1++) {                                      // agreement (:1479-1502)
            const int idx2 = vnMatch1[(size_t)i1];
            if (idx2 >= 0 && vnMatch2[(size_t)idx2] == i1) { vpMatches12[(size_t)i1] = vpMapPoints2[(size_t)idx2]; nFound++; }
        }
        return nFound;
    }

    // SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, keys1, keys2, pairs)  (src/ORBmatcher.cc:852-1014;
    // src/LocalMapping.cc:1080): keypoints without a map point, same vocabulary node, dist <= TH_LOW, sorted by (dist, index),
    // first epipolar-consistent candidate within round(2 * best) wins and is claimed; rotation histogram; outputs in ascending
    // order of the first keyframe's index.  The epipolar lines x1' F12 are formed here in float as :139-141 does; the per-pair
    // chi-square test runs inside uvip_search_lists_epipolar.  KeyFrameT additionally needs GetFeatureVector(), GetSigma2(l).
    template <class KeyFrameT>
    int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, cv::Mat F12, std::vector<cv::KeyPoint>& vMatchedKeys1,
                               std::vector<cv::KeyPoint>& vMatchedKeys2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs)
    {
        ensure();
        const auto vpMapPoints1 = pKF1->GetMapPointMatches(); const auto vpMapPoints2 = pKF2->GetMapPointMatches();
        const std::vector<cv::KeyPoint> vKeysUn1 = pKF1->GetKeyPointsUn(), vKeysUn2 = pKF2->GetKeyPointsUn();
        const cv::Mat Descriptors1 = pKF1->GetDescriptors(), Descriptors2 = pKF2->GetDescriptors();
        const auto vFeatVec1 = pKF1->GetFeatureVector(); const auto vFeatVec2 = pKF2->GetFeatureVector();
        float F[3][3]; read3x3(F12, F);
        std::vector<unsigned char> qd; std::vector<float> ql; std::vector<int32_t> cs(1, 0), ci; std::vector<unsigned> qidx;
        auto f1it = vFeatVec1.begin(); auto f2it = vFeatVec2.begin();
        while (f1it != vFeatVec1.end() && f2it != vFeatVec2.end()) {
            if (f1it->first == f2it->first) {
                for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                    const size_t idx1 = f1it->second[i1];
                    if (vpMapPoints1[idx1]) continue;                                           // already a map point (:897-899)
                    const cv::KeyPoint& kp1 = vKeysUn1[idx1];
                    const float a = kp1.pt.x * F[0][0] + kp1.pt.y * F[1][0] + F[2][0];
                    const float b = kp1.pt.x * F[0][1] + kp1.pt.y * F[1][1] + F[2][1];
                    const float c = kp1.pt.x * F[0][2] + kp1.pt.y * F[1][2] + F[2][2];
                    const float den = a * a + b * b;
                    qd.insert(qd.end(), Descriptors1.ptr((int)idx1), Descriptors1.ptr((int)idx1) + 32);
                    ql.push_back(a); ql.push_back(b); ql.push_back(c); ql.push_back(den);
                    for (size_t i2 = 0; i2 < f2it->second.size(); i2++) { const size_t idx2 = f2it->second[i2]; if (!vpMapPoints2[idx2]) ci.push_back((int32_t)idx2); }
                    cs.push_back((int32_t)ci.size());
                    qidx.push_back((unsigned)idx1);
                }
                ++f1it; ++f2it;
            } else if (f1it->first < f2it->first) f1it = vFeatVec1.lower_bound(f2it->first);
            else f2it = vFeatVec2.lower_bound(f1it->first);
        }
        vMatchedKeys1.clear(); vMatchedKeys2.clear(); vMatchedPairs.clear();
        const int nq = (int)qidx.size(), nk = (int)vKeysUn2.size();
        if (nq == 0 || nk == 0) return 0;
        std::vector<unsigned char> kd((size_t)nk * 32); std::vector<float> kx((size_t)nk), ky((size_t)nk), a2((size_t)nk); std::vector<double> kthr((size_t)nk);
        for (int i = 0; i < nk; i++) {
            std::memcpy(&kd[(size_t)i * 32], Descriptors2.ptr(i), 32);
            kx[i] = vKeysUn2[(size_t)i].pt.x; ky[i] = vKeysUn2[(size_t)i].pt.y; a2[i] = vKeysUn2[(size_t)i].angle;
            kthr[i] = 3.84 * pKF2->GetSigma2(vKeysUn2[(size_t)i].octave);                       // :152
        }
        std::vector<int32_t> taken((size_t)nk, -1), match((size_t)nq);
        if (ci.empty()) ci.push_back(0);
        int nmatches = 0;
        check(uvip_search_lists_epipolar(handle_, TH_LOW, qd.data(), ql.data(), nq, cs.data(), ci.data(), kd.data(), kx.data(), ky.data(), kthr.data(), nk,
                                         taken.data(), match.data(), &nmatches), "uvip_search_lists_epipolar");
        if (mbCheckOrientation) {
            std::vector<float> a1((size_t)nq);
            for (int q = 0; q < nq; q++) a1[q] = vKeysUn1[qidx[q]].angle;
            check(uvip_rot_hist_filter(handle_, match.data(), nq, a1.data(), a2.data(), &nmatches), "uvip_rot_hist_filter");
        }
        std::vector<int> vMatches12(vKeysUn1.size(), -1);
        for (int q = 0; q < nq; q++) if (match[q] >= 0) vMatches12[qidx[q]] = match[q];
        for (size_t i = 0; i < vMatches12.size(); i++) {
            if (vMatches12[i] < 0) continue;
            vMatchedKeys1.push_back(vKeysUn1[i]); vMatchedKeys2.push_back(vKeysUn2[(size_t)vMatches12[i]]);
            vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
        }
        return nmatches;
    }

    // haloc::Utils::ratioMatching (include/utils.h:81-111): brute-force k=2 + ratio test; match[i] = train row or -1
    int RatioMatching(const cv::Mat& descriptors1, const cv::Mat& descriptors2, double ratio, std::vector<int>& match)
    {
        ensure();
        match.assign((size_t)descriptors1.rows, -1);
        if (descriptors1.empty() || descriptors2.empty()) return 0;
        const int nq = descriptors1.rows, nt = descriptors2.rows;
        std::vector<unsigned char> q((size_t)nq * 32), t((size_t)nt * 32);
        for (int i = 0; i < nq; i++) std::memcpy(&q[(size_t)i * 32], descriptors1.ptr(i), 32);
        for (int i = 0; i < nt; i++) std::memcpy(&t[(size_t)i * 32], descriptors2.ptr(i), 32);
        std::vector<int32_t> idx((size_t)nq * 2), dist((size_t)nq * 2), m((size_t)nq);
        check(uvip_knn2(handle_, q.data(), nq, t.data(), nt, idx.data(), dist.data()), "uvip_knn2");
        int n = 0;
        check(uvip_ratio_filter(handle_, idx.data(), dist.data(), nq, ratio, m.data(), &n), "uvip_ratio_filter");
        for (int i = 0; i < nq; i++) match[(size_t)i] = m[(size_t)i];
        return n;
    }

    // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:197-270) for a batch of map points: lists[p] holds the
    // observed descriptors (1 x 32 CV_8U rows) of point p; best[p] = index of the descriptor the reference would keep
    // as mDescriptor (-1 for an empty list).  LocalMapping / LoopClosing call it once per touched map point; gathering
    // the touched points of a keyframe insertion into one call is the batched form.
    void ComputeDistinctiveDescriptors(const std::vector<std::vector<cv::Mat> >& lists, std::vector<int>& best)
    {
        ensure();
        const int np = (int)lists.size();
        best.assign((size_t)np, -1);
        if (np == 0) return;
        std::vector<int32_t> start((size_t)np + 1, 0);
        for (int p = 0; p < np; p++) start[(size_t)p + 1] = start[(size_t)p] + (int)lists[(size_t)p].size();
        std::vector<unsigned char> d((size_t)start[(size_t)np] * 32 + 32);
        for (int p = 0; p < np; p++)
            for (size_t i = 0; i < lists[(size_t)p].size(); i++) std::memcpy(&d[((size_t)start[(size_t)p] + i) * 32], lists[(size_t)p][i].ptr(0), 32);
        std::vector<int32_t> bi((size_t)np);
        check(uvip_distinctive_descriptors(handle_, d.data(), start.data(), np, bi.data(), 0), "uvip_distinctive_descriptors");
        for (int p = 0; p < np; p++) best[(size_t)p] = bi[(size_t)p];
    }

    // rotation-consistency histogram (src/ORBmatcher.cc:232-241, :263-281, :1748-1789) over an index match list
    int CheckOrientation(std::vector<int>& match, const std::vector<float>& angles1, const std::vector<float>& angles2)
    {
        ensure();
        if (!mbCheckOrientation || match.empty()) { int n = 0; for (int v : match) n += v >= 0; return n; }
        std::vector<int32_t> m(match.begin(), match.end());
        int kept = 0;
        check(uvip_rot_hist_filter(handle_, m.data(), (int)m.size(), angles1.data(), angles2.data(), &kept), "uvip_rot_hist_filter");
        for (size_t i = 0; i < match.size(); i++) match[i] = m[i];
        return kept;
    }

public:
    static const int TH_LOW = UVIP_TH_LOW;
    static const int TH_HIGH = UVIP_TH_HIGH;
    static const int HISTO_LENGTH = UVIP_HISTO_LENGTH;

protected:
    float RadiusByViewingCos(const float& viewCos) { return uvip_radius_by_viewing_cos(viewCos); }
    struct ThreadHandle {
        uvip_matcher* h = nullptr; int device = 0;
        ~ThreadHandle() { if (h) uvip_matcher_destroy(h); }
    };
    static ThreadHandle& tls() { static thread_local ThreadHandle t; return t; }
    void ensure()
    {
        ThreadHandle& t = tls();
        if (!t.h && uvip_matcher_create(t.device, &t.h) != UVIP_OK)
            throw std::runtime_error(std::string("uvip_matcher_create: ") + uvip_last_error());
        handle_ = t.h;
    }
    static void check(int rc, const char* what) { if (rc != UVIP_OK) throw std::runtime_error(std::string(what) + ": " + uvip_last_error()); }

    float mfNNratio;
    bool mbCheckOrientation;
    uvip_matcher* handle_ = nullptr;        // the calling thread's handle, refreshed by ensure() at every call (not owned)
};

}  // namespace USLAM
