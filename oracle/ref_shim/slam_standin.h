// slam_standin.h — TEST INFRASTRUCTURE.  Stand-ins for the reference's FrameKTL / KeyFrame / MapPoint classes with
// exactly the members src/ORBmatcher.cc touches, so that the UNMODIFIED reference matcher can be compiled here
// (oracle/Makefile target `ref`: the real include/MapPoint.h, KeyFrame.h and FrameKTL.h pull in g2o, DBoW2, Boost, PCL
// and ROS, so their include guards are pre-defined and this header is force-included instead).  Plain data holders:
// the test driver (ref_matcher_driver.cpp) fills them; all matching logic that runs is the reference's.
// GetFeaturesInArea (src/FrameKTL.cc:359-424, src/KeyFrame.cc:952-992 — not compilable here) forwards to the C oracle's
// restatement.
#ifndef UVIP_SLAM_STANDIN_H
#define UVIP_SLAM_STANDIN_H
#include <limits.h>
#include <map>
#include <set>
#include <vector>
#include "uvip_cv_standin.hpp"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "grid_standin.h"

using namespace std;        // the reference headers rely on it (include/ORBmatcher.h:71 uses an unqualified `pair`)

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace USLAM {

class KeyFrame;
class FrameKTL;

class MapPoint {
public:
    MapPoint() : mbTrackInView(false), mnTrackScaleLevel(0), mTrackProjX(0), mTrackProjY(0), mTrackViewCos(0), mnLastFrameSeen(0), mnId(0),
                 bad(false), minDist(0), maxDist(0), nObs(1), replaced(0) {}
    // tracking fields written by FrameKTL::isInFrustum (include/MapPoint.h:96-101)
    bool mbTrackInView; int mnTrackScaleLevel; float mTrackProjX, mTrackProjY, mTrackViewCos;
    long unsigned int mnLastFrameSeen, mnId;
    bool isBad() { return bad; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
    cv::Mat GetNormal() { return normal.clone(); }
    float GetMinDistanceInvariance() { return minDist; }
    float GetMaxDistanceInvariance() { return maxDist; }
    int Observations() { return nObs; }
    bool IsInKeyFrame(KeyFrame* kf) { return obs.count(kf) != 0; }
    int GetIndexInKeyFrame(KeyFrame* kf) { return obs.count(kf) ? (int)obs[kf] : -1; }
    void AddObservation(KeyFrame* kf, size_t idx) { if (!obs.count(kf)) { obs[kf] = idx; nObs++; } }
    void Replace(MapPoint* p) { replaced = p; bad = true; }
    void ComputeDistinctiveDescriptors() {}
    // data
    bool bad; float minDist, maxDist; int nObs; MapPoint* replaced;
    cv::Mat desc, pos, normal;
    std::map<KeyFrame*, size_t> obs;
};

class FrameKTL {
public:
    FrameKTL() : fx(0), fy(0), cx(0), cy(0), mnMinX(0), mnMaxX(0), mnMinY(0), mnMaxY(0), mfGridElementWidthInv(0), mfGridElementHeightInv(0),
                 mnScaleLevels(0), mnId(0) {}
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    cv::Mat mDescriptors, mTcw;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors;
    DBoW2::FeatureVector mFeatVec;
    float fx, fy, cx, cy;
    float mnMinX, mnMaxX, mnMinY, mnMaxY;       // static members in the reference (include/FrameKTL.h:173-176)
    float mfGridElementWidthInv, mfGridElementHeightInv;     // static members in the reference (include/FrameKTL.h:157-158); read by the drop-in shim
    int mnScaleLevels; long unsigned int mnId;
    GridStandin grid;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const
    { return grid.area(x, y, r, minLevel, maxLevel); }
};

class KeyFrame {
public:
    KeyFrame() : mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS), mfGridElementWidthInv(0), mfGridElementHeightInv(0), fx(0), fy(0), cx(0), cy(0),
                 mnId(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0) {}
    // public members of the reference class (include/KeyFrame.h:186-246)
    int mnGridCols, mnGridRows; float mfGridElementWidthInv, mfGridElementHeightInv;
    float fx, fy, cx, cy; long unsigned int mnId;
    // protected in the reference (include/KeyFrame.h:260-263); the shim needs mnMinX / mnMinY
    int mnMinX, mnMinY, mnMaxX, mnMaxY;
    // data
    std::vector<cv::KeyPoint> keysUn;
    cv::Mat descriptors, Rcw, tcw, Ow;
    std::vector<MapPoint*> mapPoints;
    std::vector<float> scaleFactors, levelSigma2;
    DBoW2::FeatureVector featVec;
    GridStandin grid;
    void set_bounds(int minX, int maxX, int minY, int maxY)
    {
        mnMinX = minX; mnMaxX = maxX; mnMinY = minY; mnMaxY = maxY;
        mfGridElementWidthInv = (float)mnGridCols / (float)(maxX - minX); mfGridElementHeightInv = (float)mnGridRows / (float)(maxY - minY);
        grid.build(keysUn, (float)minX, (float)maxX, (float)minY, (float)maxY);
    }
    std::vector<MapPoint*> GetMapPointMatches() { return mapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mapPoints) if (p && !p->isBad()) s.insert(p); return s; }
    void AddMapPoint(MapPoint* p, const size_t& idx) { mapPoints[idx] = p; }
    DBoW2::FeatureVector GetFeatureVector() { return featVec; }
    cv::Mat GetDescriptor(const size_t& idx) { return descriptors.row((int)idx).clone(); }
    cv::Mat GetDescriptors() { return descriptors.clone(); }
    std::vector<cv::KeyPoint> GetKeyPointsUn() const { return keysUn; }
    cv::KeyPoint GetKeyPointUn(const size_t& idx) const { return keysUn[idx]; }
    int GetKeyPointScaleLevel(const size_t& idx) const { return keysUn[idx].octave; }
    std::vector<float> GetScaleFactors() const { return scaleFactors; }
    float GetScaleFactor(int level = 1) const { return scaleFactors[level]; }
    int GetScaleLevels() const { return (int)scaleFactors.size(); }
    float GetSigma2(int level = 1) const { return levelSigma2[level]; }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }     // src/KeyFrame.cc:994-997
    // src/KeyFrame.cc:952-992: no level filter
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return grid.area(x, y, r, -1, -1); }
};

}  // namespace USLAM
#endif
