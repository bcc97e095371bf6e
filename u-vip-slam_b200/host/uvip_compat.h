// uvip_compat.h — minimal stand-ins for the OpenCV / Eigen types that appear in the reference's ORBextractor and
// ORBmatcher signatures, for builds where OpenCV and Eigen headers are not installed (this repository's build
// image).  With -DUVIP_WITH_OPENCV the shim uses the real cv:: / Eigen:: types instead and this file is not used.
// Only the members the shim touches exist; layouts match the originals where the C-ABI depends on them
// (cv::KeyPoint = 28 bytes, Eigen::MatrixXi = column-major int).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#endif

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };

struct KeyPoint {            // OpenCV core/types.hpp
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");


class Mat {                  // 8-bit single-channel only
public:
    int rows = 0, cols = 0; size_t step = 0; unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void* ext, size_t step_) : rows(r), cols(c), step(step_), data(static_cast<unsigned char*>(ext)) {}
    void create(int r, int c, int /*type*/) {
        if (r == rows && c == cols && owner_ && step == (size_t)c) return;
        owner_ = std::shared_ptr<std::vector<unsigned char>>(new std::vector<unsigned char>((size_t)r * c));
        rows = r; cols = c; step = (size_t)c; data = owner_->data();
    }
    void release() { owner_.reset(); rows = cols = 0; step = 0; data = nullptr; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.owner_ = owner_; return m; }
    Mat getMat() const { return *this; }
private:
    std::shared_ptr<std::vector<unsigned char>> owner_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv

namespace Eigen {
class MatrixXi {             // column-major dynamic int matrix
public:
    MatrixXi() {}
    MatrixXi(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0) {}
    static MatrixXi Zero(int r, int c) { return MatrixXi(r, c); }
    int rows() const { return r_; } int cols() const { return c_; }
    int& operator()(int r, int c) { return d_[(size_t)c * r_ + r]; }
    int operator()(int r, int c) const { return d_[(size_t)c * r_ + r]; }
    int* data() { return d_.data(); } const int* data() const { return d_.data(); }
private:
    int r_ = 0, c_ = 0; std::vector<int> d_;
};
}  // namespace Eigen
