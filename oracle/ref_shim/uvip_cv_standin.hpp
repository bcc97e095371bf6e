// uvip_cv_standin.hpp — TEST INFRASTRUCTURE.  A minimal stand-in for the slice of the OpenCV 3.4 C++ API that the
// reference's src/ORBextractor.cc uses, so that the UNMODIFIED reference source can be compiled in this image
// (which has no OpenCV C++ headers) into oracle/_ref/libref_orbextractor.so.  Everything that is the reference's
// own code (cell geometry, empty-cell retry, DistributeOctTree / DivideNode, IC_Angle, computeOrbDescriptor,
// operator() orchestration, occupancy-grid filter) then runs for real; the OpenCV-owned primitives behind
// cv::resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 forward to the C oracle's restatements, which are
// pinned against the cv2 4.13 wheel by tests/golden/.  Not a general OpenCV replacement: only what that file touches.
#ifndef UVIP_CV_STANDIN_HPP
#define UVIP_CV_STANDIN_HPP
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <algorithm>
#include <vector>
#include "../uvip_oracle.h"

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) assert(expr)

// OpenCV 3.4 core/fast_math.hpp: cvRound(double) = lrint, cvRound(float) = lrintf (round half to even), cvFloor, cvCeil
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
// cv::Point_<float> * float multiplies in float (saturate_cast<float>(pt.x * s))
template <typename T> static inline Point_<T> operator*(const Point_<T>& p, float s) { return Point_<T>((T)(p.x * s), (T)(p.y * s)); }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

struct KeyPoint {           // same 28-byte layout as cv::KeyPoint
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float sz, float ang = -1, float resp = 0, int oct = 0, int cid = -1)
        : pt(x, y), size(sz), angle(ang), response(resp), octave(oct), class_id(cid) {}
};

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

struct MatZeros { int rows, cols, type; };

// 8-bit single-channel matrix header over a shared, reference-counted buffer (ROI views share the buffer).  The
// buffer and its counter come from malloc, never from operator new, so that the driver's arena (ref_driver.cpp) only
// ever sees the reference's own allocations.
class Mat {
public:
    struct Step { size_t v; Step() : v(0) {} operator size_t() const { return v; } };
    uchar* data; int rows, cols; Step step;
    Mat() : data(0), rows(0), cols(0), buf(0) {}
    Mat(Size sz, int type) : data(0), rows(0), cols(0), buf(0) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : data(0), rows(0), cols(0), buf(0) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t st) : data((uchar*)ext), rows(r), cols(c), buf(0) { (void)type; step.v = st; }
    Mat(const Mat& m) : data(m.data), rows(m.rows), cols(m.cols), step(m.step), buf(m.buf) { retain(); }
    Mat(const Mat& m, const Rect& r) : data(m.data + (size_t)r.y * m.step.v + r.x), rows(r.height), cols(r.width), step(m.step), buf(m.buf)
    { assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= m.cols && r.y + r.height <= m.rows); retain(); }
    Mat(const MatZeros& z) : data(0), rows(0), cols(0), buf(0) { *this = z; }
    ~Mat() { drop(); }
    Mat& operator=(const Mat& m)
    {
        if (this != &m) { if (m.buf) ++*m.buf; drop(); data = m.data; rows = m.rows; cols = m.cols; step = m.step; buf = m.buf; }
        return *this;
    }
    static MatZeros zeros(int r, int c, int type) { MatZeros z = {r, c, type}; return z; }
    // assigning Mat::zeros to a matrix of the same shape clears it in place (cv::Mat::operator=(const MatExpr&) -> create + setTo)
    Mat& operator=(const MatZeros& z) { create(z.rows, z.cols, z.type); for (int y = 0; y < rows; y++) memset(data + (size_t)y * step.v, 0, (size_t)cols); return *this; }
    void create(int r, int c, int type)
    {
        assert(type == CV_8UC1);
        if (data && r == rows && c == cols) return;
        drop();
        buf = (int*)malloc(64 + (size_t)r * c + 64);
        *buf = 1;
        data = (uchar*)buf + 64; rows = r; cols = c; step.v = (size_t)c;
    }
    void release() { drop(); data = 0; rows = cols = 0; step.v = 0; }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t elemSize1() const { return 1; }
    size_t step1() const { return step.v; }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat operator()(const Rect& r) const { return Mat(*this, r); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step.v))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step.v))[x]; }
    Mat clone() const { Mat m(rows, cols, CV_8UC1); for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), (size_t)cols); return m; }
private:
    int* buf;                       // reference counter at the head of the malloc'ed block, 0 for external data
    void retain() { if (buf) ++*buf; }
    void drop() { if (buf && --*buf == 0) free(buf); buf = 0; }
};

class _InputArray {
public:
    _InputArray() : m(0) {}
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
protected:
    const Mat* m;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& mm) : _InputArray(mm), o(&mm) {}
    void release() const { o->release(); }
    void create(int r, int c, int type) const { o->create(r, c, type); }
    void create(Size sz, int type) const { o->create(sz.height, sz.width, type); }
    Mat getMat() const { return *o; }
    Mat& getMatRef() const { return *o; }
private:
    Mat* o;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline _InputArray noArray() { return _InputArray(); }

template <typename T> class AutoBuffer {
public:
    explicit AutoBuffer(size_t n) : v(n) {}
    operator T*() { return v.data(); }
private:
    std::vector<T> v;
};

struct ORB { enum { HARRIS_SCORE = 0, FAST_SCORE = 1 }; };

static inline float fastAtan2(float y, float x) { return uo_fast_atan2(y, x); }

// cv::FAST(image, keypoints, threshold, nonmaxSuppression): FAST-9/16, KeyPoint(x, y, 7.f, -1, score), row-major order
static inline void FAST(InputArray image, std::vector<KeyPoint>& kps, int threshold, bool nms = true)
{
    const Mat img = image.getMat();
    kps.clear();
    if (img.empty() || img.rows < 7 || img.cols < 7) return;
    const int cap = img.rows * img.cols;
    static thread_local int* scratch = 0; static thread_local int scratch_cap = 0;      // malloc, not operator new (see Mat)
    if (cap > scratch_cap) { free(scratch); scratch = (int*)malloc(sizeof(int) * 3 * (size_t)cap); scratch_cap = cap; }
    int *xs = scratch, *ys = scratch + cap, *sc = scratch + 2 * (size_t)cap;
    const int n = uo_fast9(img.data, (int)img.step.v, img.cols, img.rows, threshold, nms ? 1 : 0, xs, ys, sc, cap);
    kps.reserve(n);
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

// cv::resize, 8-bit INTER_LINEAR only (the mask pyramid of the reference is never built: callers pass an empty mask)
static inline void resize(InputArray src_, OutputArray dst_, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR)
{
    (void)fx; (void)fy;
    assert(interp == INTER_LINEAR);
    const Mat src = src_.getMat();
    dst_.create(dsize, CV_8UC1);
    Mat dst = dst_.getMat();
    uo_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dst.cols, dst.rows, (int)dst.step.v);
}

// cv::copyMakeBorder with BORDER_REFLECT_101 (+ISOLATED); when src is already the interior of dst only the border is written
static inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType)
{
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && top == bottom && top == left && top == right);
    const Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = dst_.getMat();
    uchar* interior = dst.data + (size_t)top * dst.step.v + left;
    if (interior != src.data)
        for (int y = 0; y < src.rows; y++) memmove(interior + (size_t)y * dst.step.v, src.ptr(y), (size_t)src.cols);
    uo_border_reflect101(dst.data, src.cols, src.rows, (int)dst.step.v, top);
}

// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on a ROI whose 3-px surround holds the reflect-101 border (the only
// way the reference calls it); in place
static inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy = 0, int borderType = BORDER_REFLECT_101)
{
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    const Mat src = src_.getMat();
    Mat tmp(src.rows, src.cols, CV_8UC1);
    uo_blur7(src.data, src.cols, src.rows, (int)src.step.v, tmp.data, (int)tmp.step.v);
    dst_.create(src.rows, src.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    for (int y = 0; y < src.rows; y++) memcpy(dst.ptr(y), tmp.ptr(y), (size_t)src.cols);
}

// cv::KeyPointsFilter::retainBest (features2d/keypoint.cpp): nth_element on response, keep ties with the n-th response.
// Only reachable from the reference's dead ComputeKeyPoints path; present so that the file links.
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& kps, int n)
    {
        if (n >= 0 && kps.size() > (size_t)n) {
            if (n == 0) { kps.clear(); return; }
            std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            const float amb = kps[n - 1].response;
            auto e = std::partition(kps.begin() + n, kps.end(), [amb](const KeyPoint& k) { return k.response >= amb; });
            kps.resize(e - kps.begin());
        }
    }
};

}  // namespace cv
#endif
