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

using namespace std;

namespace USLAM {

class MapPoint;

class KeyFrame {
public:
    KeyFrame() : mnId(0), mfLogScaleFactor(0.f), mnScaleLevels(0), bad(false) {}
    long unsigned int mnId;
    std::vector<float> mvScaleFactors; float mfLogScaleFactor; int mnScaleLevels;
    bool bad; cv::Mat descriptors, Ow; std::vector<cv::KeyPoint> keysUn;
    std::vector<MapPoint*> mapPoints;
    bool isBad() { return bad; }
    cv::Mat GetDescriptor(const size_t& idx) { return descriptors.row((int)idx).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    int GetKeyPointScaleLevel(const size_t& idx) const { return keysUn[idx].octave; }
    float GetScaleFactor(int level = 1) const { return mvScaleFactors[level]; }
    int GetScaleLevels() const { return mnScaleLevels; }
    void EraseMapPointMatch(const size_t& idx) { if (idx < mapPoints.size()) mapPoints[idx] = 0; }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* p) { if (idx < mapPoints.size()) mapPoints[idx] = p; }
};

class Map {
public:
    std::set<MapPoint*> erased;
    void EraseMapPoint(MapPoint* p) { erased.insert(p); }
};

class FrameKTL {
public:
    FrameKTL() : mfLogScaleFactor(0.f), mnScaleLevels(0), mnId(0) {}
    float mfLogScaleFactor; int mnScaleLevels; long unsigned int mnId;
};

}  // namespace USLAM
#endif
