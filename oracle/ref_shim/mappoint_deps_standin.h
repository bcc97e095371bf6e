// mappoint_deps_standin.h — TEST INFRASTRUCTURE.  Stand-ins for KeyFrame / Map / FrameKTL with the members the reference's
// src/MapPoint.cc touches, so that the REAL MapPoint class (include/MapPoint.h + src/MapPoint.cc, unmodified) can be compiled
// here and MapPoint::ComputeDistinctiveDescriptors (:197-270) pinned.  Force-included with the include guards of KeyFrame.h,
// Map.h and FrameKTL.h pre-defined (oracle/Makefile).
#ifndef UVIP_MAPPOINT_DEPS_STANDIN_H
#define UVIP_MAPPOINT_DEPS_STANDIN_H
#include <limits.h>
#include <map>
#include <set>
#include <vector>
#include "uvip_cv_standin.hpp"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "grid_standin.h"

using namespace std;

namespace USLAM {

class MapPoint;

class KeyFrame {
public:
    KeyFrame() : mnId(0), mfLogScaleFactor(0.f), mnScaleLevels(0), bad(false), mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS),
                 mfGridElementWidthInv(0), mfGridElementHeightInv(0), fx(0), fy(0), cx(0), cy(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0) {}
    long unsigned int mnId;
    std::vector<float> mvScaleFactors; float mfLogScaleFactor; int mnScaleLevels;
    bool bad; cv::Mat descriptors, Ow, Rcw, tcw; std::vector<cv::KeyPoint> keysUn;
    std::vector<MapPoint*> mapPoints;
    std::vector<float> levelSigma2;
    DBoW2::FeatureVector featVec;
    int mnGridCols, mnGridRows; float mfGridElementWidthInv, mfGridElementHeightInv, fx, fy, cx, cy; int mnMinX, mnMinY, mnMaxX, mnMaxY;
    GridStandin grid;
    bool isBad() { return bad; }
    cv::Mat GetDescriptor(const size_t& idx) { return descriptors.row((int)idx).clone(); }
    cv::Mat GetDescriptors() { return descriptors.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    int GetKeyPointScaleLevel(const size_t& idx) const { return keysUn[idx].octave; }
    std::vector<cv::KeyPoint> GetKeyPointsUn() const { return keysUn; }
    cv::KeyPoint GetKeyPointUn(const size_t& idx) const { return keysUn[idx]; }
    float GetScaleFactor(int level = 1) const { return mvScaleFactors[level]; }
    std::vector<float> GetScaleFactors() const { return mvScaleFactors; }
    int GetScaleLevels() const { return mnScaleLevels; }
    float GetSigma2(int level = 1) const { return levelSigma2[level]; }
    void EraseMapPointMatch(const size_t& idx) { if (idx < mapPoints.size()) mapPoints[idx] = 0; }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* p) { if (idx < mapPoints.size()) mapPoints[idx] = p; }
    std::vector<MapPoint*> GetMapPointMatches() { return mapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mapPoints[idx]; }
    void AddMapPoint(MapPoint* p, const size_t& idx) { mapPoints[idx] = p; }
    std::set<MapPoint*> GetMapPoints();                 // needs the complete MapPoint: defined by the driver
    DBoW2::FeatureVector GetFeatureVector() { return featVec; }
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return grid.area(x, y, r, -1, -1); }
};

class Map {
public:
    std::set<MapPoint*> erased;
    void EraseMapPoint(MapPoint* p) { erased.insert(p); }
};

#ifndef UVIP_REAL_FRAMEKTL            // the mini-frontend build (frame_deps_standin.h) compiles the reference's real FrameKTL
class FrameKTL {
public:
    FrameKTL() : mfLogScaleFactor(0.f), mnScaleLevels(0), mnId(0) {}
    float mfLogScaleFactor; int mnScaleLevels; long unsigned int mnId;
};
#else
class FrameKTL;                      // include/KeyFrame.h forward-declares it for include/MapPoint.h
#endif

}  // namespace USLAM
#endif
