// frame_deps_standin.h — TEST INFRASTRUCTURE.  What the reference's REAL FrameKTL (include/FrameKTL.h + src/FrameKTL.cc),
// REAL MapPoint (src/MapPoint.cc) and REAL ORBmatcher (src/ORBmatcher.cc) need from the rest of the system in order to be
// compiled together, unmodified, in this image: stand-ins for KeyFrame / Map (mappoint_deps_standin.h), for the IMU types
// (imudata.h, NavState.h, IMUPreintegrator.h pull in Eigen + Sophus), for Converter (g2o) and for the few extra OpenCV
// entry points FrameKTL.cc names.  Force-included with the corresponding include guards pre-defined (oracle/Makefile).
// The pinned functions — FrameKTL::PosInGrid, the grid fill of compute_descriptors (:250-264), GetFeaturesInArea (:359-424),
// isInFrustum (:299-357), MapPoint::PredictScale, ORBmatcher::SearchByProjection — use none of the stand-in IMU code.
#ifndef UVIP_FRAME_DEPS_STANDIN_H
#define UVIP_FRAME_DEPS_STANDIN_H
#include <iomanip>
#include "uvip_cv_standin.hpp"
#include "Eigen/Dense"

namespace Eigen {
struct Vector3d {
    double v[3];
    Vector3d() { v[0] = v[1] = v[2] = 0; }
    static Vector3d Zero() { return Vector3d(); }
    Vector3d operator+(const Vector3d& o) const { Vector3d r; for (int i = 0; i < 3; i++) r.v[i] = v[i] + o.v[i]; return r; }
    Vector3d operator-(const Vector3d& o) const { Vector3d r; for (int i = 0; i < 3; i++) r.v[i] = v[i] - o.v[i]; return r; }
};
struct Matrix3d { double m[9]; Matrix3d() { for (int i = 0; i < 9; i++) m[i] = (i % 4 == 0) ? 1 : 0; } };
template <class T, int R, int C> struct Matrix { T m[R * C]; Matrix() { for (int i = 0; i < R * C; i++) m[i] = 0; } };
}
using namespace Eigen;

namespace cv {
// FrameKTL's image constructor and ComputeImageBounds name these; the tests build frames through the default constructor
inline int buildOpticalFlowPyramid(InputArray, std::vector<Mat>&, Size, int) { throw std::string("cv::buildOpticalFlowPyramid is a stand-in"); }
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) { throw std::string("cv::undistortPoints is a stand-in"); }
}

#include "mappoint_deps_standin.h"          // KeyFrame, Map (FrameKTL is the REAL class in this build: see the guard below)

namespace USLAM {

class IMUData { public: Vector3d wm, am; double timestamp; IMUData() : timestamp(0) {} };
class NavState {
public:
    Vector3d Get_BiasGyr() const { return bg; }
    Vector3d Get_BiasAcc() const { return ba; }
    Vector3d Get_dBias_Gyr() const { return dbg; }
    Vector3d Get_dBias_Acc() const { return dba; }
    void Set_BiasGyr(const Vector3d& x) { bg = x; }
    void Set_BiasAcc(const Vector3d& x) { ba = x; }
    void Set_DeltaBiasGyr(const Vector3d& x) { dbg = x; }
    void Set_DeltaBiasAcc(const Vector3d& x) { dba = x; }
    Matrix3d Get_RotMatrix() const { return Matrix3d(); }
    Vector3d Get_P() const { return Vector3d(); }
private:
    Vector3d bg, ba, dbg, dba;
};
class IMUPreintegrator {
public:
    void reset() {}
    void update(const Vector3d&, const Vector3d&, double) {}
};
class Converter {
public:
    static cv::Mat toCvMat(const Matrix3d& m) { cv::Mat r(3, 3, CV_32F); for (int i = 0; i < 9; i++) ((float*)r.data)[i] = (float)m.m[i]; return r; }
    static cv::Mat toCvMat(const Vector3d& v) { cv::Mat r(3, 1, CV_32F); for (int i = 0; i < 3; i++) ((float*)r.data)[i] = (float)v.v[i]; return r; }
    static void updateNS(NavState&, const IMUPreintegrator&, const Vector3d&) {}
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& d) { std::vector<cv::Mat> v; for (int i = 0; i < d.rows; i++) v.push_back(d.row(i)); return v; }
};

}  // namespace USLAM
#endif
