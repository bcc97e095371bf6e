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
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../uvip_oracle.h"

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
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

// Single-channel matrix header (CV_8U or CV_32F) over a shared, reference-counted buffer (ROI views share the
// buffer).  The buffer and its counter come from malloc, never from operator new, so that the driver's arena
// (ref_driver.cpp) only ever sees the reference's own allocations.
struct MatT;
class Mat {
public:
    struct Step { size_t v; Step() : v(0) {} operator size_t() const { return v; } };
    uchar* data; int rows, cols; Step step;
    Mat() : data(0), rows(0), cols(0), tp(CV_8UC1), buf(0) {}
    Mat(Size sz, int type) : data(0), rows(0), cols(0), tp(CV_8UC1), buf(0) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : data(0), rows(0), cols(0), tp(CV_8UC1), buf(0) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : data((uchar*)ext), rows(r), cols(c), tp(type), buf(0) { step.v = st ? st : (size_t)c * esz(type); }
    Mat(const Mat& m) : data(m.data), rows(m.rows), cols(m.cols), step(m.step), tp(m.tp), buf(m.buf) { retain(); }
    Mat(const Mat& m, const Rect& r) : data(m.data + (size_t)r.y * m.step.v + (size_t)r.x * esz(m.tp)), rows(r.height), cols(r.width), step(m.step), tp(m.tp), buf(m.buf)
    { assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= m.cols && r.y + r.height <= m.rows); retain(); }
    Mat(const MatZeros& z) : data(0), rows(0), cols(0), tp(CV_8UC1), buf(0) { *this = z; }
    inline Mat(const MatT& e);
    ~Mat() { drop(); }
    Mat& operator=(const Mat& m)
    {
        if (this != &m) { if (m.buf) ++*m.buf; drop(); data = m.data; rows = m.rows; cols = m.cols; step = m.step; tp = m.tp; buf = m.buf; }
        return *this;
    }
    static MatZeros zeros(int r, int c, int type) { MatZeros z = {r, c, type}; return z; }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int y = 0; y < r; y++) { memset(m.ptr(y), 0, (size_t)c * esz(type)); if (y < c && type == CV_32F) m.at<float>(y, y) = 1.f; } return m; }
    Mat reshape(int) const { return *this; }      // only named by FrameKTL::ComputeImageBounds' distortion branch, which the tests never take
    // assigning Mat::zeros to a matrix of the same shape clears it in place (cv::Mat::operator=(const MatExpr&) -> create + setTo)
    Mat& operator=(const MatZeros& z) { create(z.rows, z.cols, z.type); for (int y = 0; y < rows; y++) memset(data + (size_t)y * step.v, 0, (size_t)cols * esz(tp)); return *this; }
    void create(int r, int c, int type)
    {
        assert(type == CV_8UC1 || type == CV_32F);
        if (data && r == rows && c == cols && type == tp) return;
        drop();
        buf = (int*)malloc(64 + (size_t)r * c * esz(type) + 64);
        *buf = 1;
        data = (uchar*)buf + 64; rows = r; cols = c; tp = type; step.v = (size_t)c * esz(type);
    }
    void release() { drop(); data = 0; rows = cols = 0; step.v = 0; }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    int type() const { return tp; }
    size_t elemSize() const { return esz(tp); }
    size_t elemSize1() const { return esz(tp); }
    size_t step1() const { return step.v / esz(tp); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }
    Mat col(int x) const { return Mat(*this, Rect(x, 0, 1, rows)); }
    Mat operator()(const Rect& r) const { return Mat(*this, r); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step.v))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step.v))[x]; }
    // single-index access to a row or column vector
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat clone() const { Mat m(rows, cols, tp); for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), (size_t)cols * esz(tp)); return m; }
    void copyTo(Mat& m) const { m.create(rows, cols, tp); for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), (size_t)cols * esz(tp)); }
    void copyTo(const Mat& view) const { assert(view.rows == rows && view.cols == cols && view.tp == tp); for (int y = 0; y < rows; y++) memcpy(const_cast<Mat&>(view).ptr(y), ptr(y), (size_t)cols * esz(tp)); }   // into a temporary ROI view
    // ---- CV_32F algebra (what src/ORBmatcher.cc needs); products and sums in float, left to right, like cv::gemm's
    //      small-matrix path
    inline struct MatT t() const;       // lazy alpha * A^T, see below
    double dot(const Mat& o) const
    {
        assert(tp == CV_32F && o.tp == CV_32F && rows * cols == o.rows * o.cols);
        double s = 0; const int n = rows * cols;      // cv::Mat::dot accumulates float products in double
        for (int i = 0; i < n; i++) s += (double)lin(i) * (double)o.lin(i);
        return s;
    }
    float lin(int i) const { return at<float>(i / cols, i % cols); }
private:
    int tp;
    int* buf;                       // reference counter at the head of the malloc'ed block, 0 for external data
    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    void retain() { if (buf) ++*buf; }
    void drop() { if (buf && --*buf == 0) free(buf); buf = 0; }
};

// cv::Mat products as OpenCV 3.4 evaluates them (core/src/matmul.cpp).  A*B (+C) without transposition and with inner
// size 2..4 takes gemm's small-matrix path: products and sums in float, left to right, then (float)(t*alpha + c*beta) in
// double.  Anything with a transposed operand (e.g. -R.t()*t) takes the generic path, which accumulates in double.
struct MatProd { Mat a, b; inline operator Mat() const; inline MatT t() const; };
struct MatT { Mat a; double alpha; };
inline MatT Mat::t() const { assert(tp == CV_32F); MatT e = {*this, 1.0}; return e; }
inline Mat::Mat(const MatT& e) : data(0), rows(0), cols(0), tp(CV_8UC1), buf(0)
{   // transpose, then scale through convertTo (float product, see mat_scale)
    create(e.a.cols, e.a.rows, CV_32F);
    for (int y = 0; y < e.a.rows; y++) for (int x = 0; x < e.a.cols; x++)
        at<float>(x, y) = e.alpha == 1.0 ? e.a.at<float>(y, x) : e.a.at<float>(y, x) * (float)e.alpha;
}
static inline Mat gemm_small(const Mat& a, const Mat& b, const Mat* c)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    const bool small = a.cols >= 2 && a.cols <= 4;
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < b.cols; x++) {
        const double cc = c ? (double)c->at<float>(y, x) : 0.0;
        if (small) {
            float t = a.at<float>(y, 0) * b.at<float>(0, x);
            for (int k = 1; k < a.cols; k++) t = t + a.at<float>(y, k) * b.at<float>(k, x);
            m.at<float>(y, x) = (float)((double)t * 1.0 + cc * 1.0);
        } else {
            double sacc = 0;
            for (int k = 0; k < a.cols; k++) sacc += (double)a.at<float>(y, k) * (double)b.at<float>(k, x);
            m.at<float>(y, x) = (float)(sacc + cc);
        }
    }
    return m;
}
inline MatProd::operator Mat() const { return gemm_small(a, b, 0); }
inline MatT MatProd::t() const { MatT e = {gemm_small(a, b, 0), 1.0}; return e; }
static inline MatProd operator*(const Mat& a, const Mat& b) { MatProd p = {a, b}; return p; }
static inline Mat operator+(const MatProd& p, const Mat& c) { return gemm_small(p.a, p.b, &c); }
static inline MatT operator-(const MatT& e) { MatT r = {e.a, -e.alpha}; return r; }
static inline MatT operator*(double s, const MatT& e) { MatT r = {e.a, e.alpha * s}; return r; }
static inline Mat operator*(const MatT& e, const Mat& b)
{   // alpha * A^T * B: generic gemm, double accumulation
    assert(e.a.type() == CV_32F && b.type() == CV_32F && e.a.rows == b.rows);
    Mat m(e.a.cols, b.cols, CV_32F);
    for (int y = 0; y < e.a.cols; y++) for (int x = 0; x < b.cols; x++) {
        double sacc = 0;
        for (int k = 0; k < e.a.rows; k++) sacc += (double)e.a.at<float>(k, y) * (double)b.at<float>(k, x);
        m.at<float>(y, x) = (float)(sacc * e.alpha);
    }
    return m;
}
static inline Mat mat_map2(const Mat& a, const Mat& b, float sb)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.at<float>(y, x) = a.at<float>(y, x) + sb * b.at<float>(y, x);
    return m;
}
static inline Mat mat_scale(const Mat& a, double s)
{   // cv::Mat * scalar / scalar go through convertTo(alpha); for 32f -> 32f OpenCV 3.4's cvtScale works in float
    // (core/src/convert.cpp, DEF_CVT_SCALE_FUNC(32f, float, float, float)): dst = src * (float)alpha
    assert(a.type() == CV_32F);
    Mat m(a.rows, a.cols, CV_32F);
    const float sf = (float)s;
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.at<float>(y, x) = a.at<float>(y, x) * sf;
    return m;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return mat_map2(a, b, 1.f); }
static inline Mat operator-(const Mat& a, const Mat& b) { return mat_map2(a, b, -1.f); }
static inline Mat operator-(const Mat& a) { return mat_scale(a, -1.0); }
static inline Mat operator*(const Mat& a, double s) { return mat_scale(a, s); }
static inline Mat operator*(double s, const Mat& a) { return mat_scale(a, s); }
static inline Mat operator/(const Mat& a, double s) { return mat_scale(a, 1.0 / s); }
static inline double norm(const Mat& a) { return sqrt(a.dot(a)); }

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

// cv::FileStorage / cv::FileNode: DBoW2's TemplatedVocabulary has virtual YAML save / load members that must COMPILE
// (virtual members are instantiated with the class); the tests only use its text loader, so these throw when reached.
struct FileNode {
    FileNode operator[](const char*) const { fail(); return FileNode(); }
    FileNode operator[](const std::string&) const { fail(); return FileNode(); }
    FileNode operator[](int) const { fail(); return FileNode(); }
    size_t size() const { fail(); return 0; }
    operator int() const { fail(); return 0; }
    operator double() const { fail(); return 0; }
    operator std::string() const { fail(); return std::string(); }
    static void fail() { throw std::string("cv::FileStorage is a stand-in: YAML vocabularies are not supported by oracle/ref_shim"); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const char*) const { FileNode::fail(); return FileNode(); }
    FileNode operator[](const std::string&) const { FileNode::fail(); return FileNode(); }
    template <class T> FileStorage& operator<<(const T&) { FileNode::fail(); return *this; }
};

}  // namespace cv
#endif
