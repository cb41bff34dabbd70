// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// A small, self-contained stand-in for the subset of the OpenCV C++ API that the reference's hot-path headers use
// (GuidedFilter.h, StereoEnergy.h, CostVolumeEnergy.h, Plane.h, Utilities.hpp, LayerManager.h, Proposer.h), so that
// those headers can be compiled *where they lie* into oracle/_ref/ (see oracle/build_ref.py) without an OpenCV
// installation.  Everything here is written for this repository from the documented behaviour of the OpenCV calls
// (cv::Mat reference-counted views, boxFilter with BORDER_CONSTANT, convertTo/cvtScale, reduce, warpAffine's
// fixed-point sampler, cv::RNG's multiply-with-carry ...); the primitives that matter for the hot path are checked
// against the real library (cv2 4.13) in tests/test_oracle.py and tests/test_ref_pin.py.  Evaluation is eager (no
// MatExpr), one channel-interleaved buffer per matrix, depths 8U / 32S / 32F / 64F.
//
// It also carries the three dialect helpers the MSVC-written reference needs under g++ (isnan<T>/isinf<T>,
// fopen_s/__int32, and the `Reusable& r = Reusable()` default argument).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef int64_t int64;
typedef uint64_t uint64;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
#define CV_PI 3.1415926535897932384626433832795
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

// ---- MSVC dialect helpers -------------------------------------------------------------------------------------------
template <class T> inline bool isnan(T v) { return std::isnan(v); }
template <class T> inline bool isinf(T v) { return std::isinf(v); }
typedef int32_t __int32;
#define fopen_s(pf, name, mode) ((*(pf) = fopen((name), (mode))) == NULL)
// `Reusable& reusable = Reusable()` (StereoEnergy.h:625-626, CostVolumeEnergy.h:55,176) binds a temporary to a
// non-const reference, an MSVC extension.  The token sequence `Reusable()` is mapped to an object that converts to a
// fresh, per-thread default instance of whatever reference type is asked for: same observable behaviour.
struct lexp_msvc_default_arg {
    template <class T> operator T&() const {
        static thread_local T slot;
        slot = T();
        return slot;
    }
};
#define Reusable() lexp_msvc_default_arg()

namespace cv {

[[noreturn]] inline void shim_fail(const char* what) { throw std::runtime_error(std::string("cvshim: ") + what); }
inline int cvRound(double v) { return (int)std::nearbyint(v); }  // round half to even, like lrint/SSE2 cvtsd2si
inline int cvFloor(double v) { return (int)std::floor(v); }

template <class T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

// ---- small value types ----------------------------------------------------------------------------------------------
template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
template <class T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
template <class T> inline bool operator==(const Size_<T>& a, const Size_<T>& b) { return a.width == b.width && a.height == b.height; }
template <class T> inline bool operator!=(const Size_<T>& a, const Size_<T>& b) { return !(a == b); }
typedef Size_<int> Size;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T>& p, const Size_<T>& s) : x(p.x), y(p.y), width(s.width), height(s.height) {}
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    Size_<T> size() const { return Size_<T>(width, height); }
    T area() const { return width * height; }
    bool contains(const Point_<T>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
template <class T> inline bool operator==(const Rect_<T>& a, const Rect_<T>& b) { return a.x == b.x && a.y == b.y && a.width == b.width && a.height == b.height; }
template <class T> inline bool operator!=(const Rect_<T>& a, const Rect_<T>& b) { return !(a == b); }
template <class T> inline Rect_<T> operator+(const Rect_<T>& r, const Point_<T>& p) { return Rect_<T>(r.x + p.x, r.y + p.y, r.width, r.height); }
template <class T> inline Rect_<T> operator-(const Rect_<T>& r, const Point_<T>& p) { return Rect_<T>(r.x - p.x, r.y - p.y, r.width, r.height); }
// intersection; disjoint rectangles give the empty Rect() (all zero), as OpenCV does
template <class T> inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) {
    T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    T w = std::min(a.x + a.width, b.x + b.width) - x1, h = std::min(a.y + a.height, b.y + b.height) - y1;
    if (w <= 0 || h <= 0) return Rect_<T>();
    return Rect_<T>(x1, y1, w, h);
}
typedef Rect_<int> Rect;

template <class T, int n> struct Vec {
    T val[n];
    Vec() { for (int i = 0; i < n; i++) val[i] = T(0); }
    Vec(T a, T b) : Vec() { val[0] = a; if (n > 1) val[1] = b; }
    Vec(T a, T b, T c) : Vec() { val[0] = a; if (n > 1) val[1] = b; if (n > 2) val[2] = c; }
    Vec(T a, T b, T c, T d) : Vec() { val[0] = a; if (n > 1) val[1] = b; if (n > 2) val[2] = c; if (n > 3) val[3] = d; }
    template <class U> Vec(const Vec<U, n>& o) { for (int i = 0; i < n; i++) val[i] = (T)o.val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    double ddot(const Vec& o) const { double s = 0; for (int i = 0; i < n; i++) s += (double)val[i] * o.val[i]; return s; }
};
template <class T, int n> inline Vec<T, n> operator+(const Vec<T, n>& a, const Vec<T, n>& b) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] + b.val[i]); return r; }
template <class T, int n> inline Vec<T, n> operator-(const Vec<T, n>& a, const Vec<T, n>& b) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] - b.val[i]); return r; }
template <class T, int n> inline Vec<T, n> operator*(const Vec<T, n>& a, float s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] * s); return r; }
template <class T, int n> inline Vec<T, n> operator*(const Vec<T, n>& a, double s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] * s); return r; }
template <class T, int n> inline Vec<T, n> operator*(const Vec<T, n>& a, int s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] * s); return r; }
// OpenCV scales by the reciprocal (Matx_ScaleOp with 1./alpha)
template <class T, int n> inline Vec<T, n> operator/(const Vec<T, n>& a, double s) { return a * (1. / s); }
template <class T, int n> inline Vec<T, n> operator/(const Vec<T, n>& a, float s) { return a * (1.f / s); }
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;

template <class T> struct Scalar_ : public Vec<T, 4> {
    Scalar_() {}
    Scalar_(T v0) { this->val[0] = v0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { this->val[0] = v0; this->val[1] = v1; this->val[2] = v2; this->val[3] = v3; }
    template <class U, int n> Scalar_(const Vec<U, n>& v) { for (int i = 0; i < n && i < 4; i++) this->val[i] = (T)v.val[i]; }
    static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
};
typedef Scalar_<double> Scalar;

template <typename _Tp> class DataType {
public:
    typedef _Tp value_type;
    enum { generic_type = 1, depth = -1, channels = 1, fmt = 0, type = -1 };
};
#define CVSHIM_DATATYPE(T, D)                                                                    \
    template <> class DataType<T> {                                                               \
    public:                                                                                       \
        typedef T value_type; typedef T channel_type;                                             \
        enum { generic_type = 0, depth = D, channels = 1, fmt = 0, type = CV_MAKETYPE(D, 1) };    \
    };
CVSHIM_DATATYPE(uchar, CV_8U)
CVSHIM_DATATYPE(int, CV_32S)
CVSHIM_DATATYPE(float, CV_32F)
CVSHIM_DATATYPE(double, CV_64F)
template <class T, int n> class DataType<Vec<T, n> > {
public:
    typedef Vec<T, n> value_type; typedef T channel_type;
    enum { generic_type = 0, depth = DataType<T>::depth, channels = n, fmt = 0, type = CV_MAKETYPE(DataType<T>::depth, n) };
};
template <typename _Tp> class DataDepth {
public:
    enum { value = DataType<_Tp>::depth, fmt = 0 };
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { REDUCE_SUM = 0, REDUCE_AVG = 1, REDUCE_MAX = 2, REDUCE_MIN = 3 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };
enum { COLOR_BGR2GRAY = 6 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };

inline size_t depth_bytes(int depth) {
    switch (depth) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2; case CV_32S: case CV_32F: return 4; case CV_64F: return 8; }
    shim_fail("bad depth");
}

// calls f with a null pointer of the element type of `depth`
template <class F> inline void depth_switch(int depth, F&& f) {
    switch (depth) {
    case CV_8U: f((uchar*)0); break;
    case CV_32S: f((int*)0); break;
    case CV_32F: f((float*)0); break;
    case CV_64F: f((double*)0); break;
    default: shim_fail("unsupported depth");
    }
}
// arithmetic is carried out in float for 32F matrices and in double for everything else
template <class T> struct work_type { typedef double type; };
template <> struct work_type<float> { typedef float type; };

class Mat;
struct OutputArray;

class Mat {
public:
    struct MSize {
        int p[3];
        MSize() { p[0] = p[1] = p[2] = 0; }
        Size operator()() const { return Size(p[1], p[0]); }
    };
    int flags;  // the type (depth + channels)
    int dims;
    int rows, cols;
    uchar* data;
    size_t step;      // bytes between rows (2-D) / between [i][j] rows (3-D)
    size_t step0;     // 3-D: bytes between planes
    MSize size;
    std::shared_ptr<uchar> owner;

    Mat() : flags(0), dims(0), rows(0), cols(0), data(nullptr), step(0), step0(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); *this = s; }
    Mat(Size sz, int type, const Scalar& s) : Mat() { create(sz.height, sz.width, type); *this = s; }
    // external (not owned) storage
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : Mat() {
        flags = type; dims = 2; rows = r; cols = c; data = (uchar*)ext;
        step = step_ ? step_ : (size_t)c * elemSize(); size.p[0] = r; size.p[1] = c;
    }
    Mat(int ndims, const int* sizes, int type, void* ext) : Mat() {
        if (ndims != 3) shim_fail("only 3-D external volumes");
        flags = type; dims = 3; rows = cols = -1; data = (uchar*)ext;
        size.p[0] = sizes[0]; size.p[1] = sizes[1]; size.p[2] = sizes[2];
        step = (size_t)sizes[2] * elemSize(); step0 = step * sizes[1];
    }

    void create(int r, int c, int type) {
        if (dims == 2 && data && rows == r && cols == c && flags == type) return;
        flags = type; dims = 2; rows = r; cols = c; size.p[0] = r; size.p[1] = c; size.p[2] = 0;
        step = (size_t)c * elemSize(); step0 = 0;
        size_t bytes = step * (size_t)r;
        owner.reset(bytes ? (uchar*)std::malloc(bytes) : nullptr, std::free);
        data = owner.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { *this = Mat(); }

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { return depth_bytes(depth()); }
    size_t elemSize() const { return depth_bytes(depth()) * channels(); }
    size_t total() const { return dims == 3 ? (size_t)size.p[0] * size.p[1] * size.p[2] : (size_t)rows * cols; }
    bool empty() const { return data == nullptr || total() == 0; }
    bool isContinuous() const { return dims != 2 || rows <= 1 || step == (size_t)cols * elemSize(); }

    template <class T> T* ptr(int y = 0) { return (T*)(data + step * (size_t)y); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + step * (size_t)y); }
    template <class T> T& at(int y, int x) { return ((T*)(data + step * (size_t)y))[x]; }
    template <class T> const T& at(int y, int x) const { return ((const T*)(data + step * (size_t)y))[x]; }
    template <class T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <class T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    template <class T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    template <class T> T& at(int i0, int i1, int i2) { return ((T*)(data + step0 * (size_t)i0 + step * (size_t)i1))[i2]; }
    template <class T> const T& at(int i0, int i1, int i2) const { return ((const T*)(data + step0 * (size_t)i0 + step * (size_t)i1))[i2]; }

    Mat operator()(const Rect& r) const {
        if (dims != 2 || r.x < 0 || r.y < 0 || r.x + r.width > cols || r.y + r.height > rows || r.width < 0 || r.height < 0) shim_fail("ROI out of range");
        Mat m(*this);
        m.data = data + step * (size_t)r.y + elemSize() * (size_t)r.x;
        m.rows = r.height; m.cols = r.width; m.size.p[0] = m.rows; m.size.p[1] = m.cols;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }

    Mat clone() const;
    void copyTo(const OutputArray& dst) const;
    void copyTo(const OutputArray& dst, const Mat& mask) const;
    void convertTo(const OutputArray& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s, const Mat& mask = Mat());
    Mat& operator=(const Scalar& s) { return setTo(s); }
    Mat reshape(int cn, int newRows = 0) const;
    Mat mul(const Mat& m, double scale = 1) const;
    Mat mul(const Scalar& s, double scale = 1) const;
    double dot(const Mat& m) const;

    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar::all(0)); }
    static Mat zeros(Size s, int type) { return Mat(s, type, Scalar::all(0)); }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
    static Mat ones(Size s, int type) { return Mat(s, type, Scalar(1)); }
};

struct OutputArray {
    Mat* m;
    OutputArray(Mat& x) : m(&x) {}
    OutputArray(const Mat& x) : m(const_cast<Mat*>(&x)) {}  // writing through a view (temporary header), as cv::_OutputArray allows
    // hand `res` over: keep the destination's storage when it already has the right shape (views!), else rebind it
    void assign(const Mat& res) const {
        if (m->dims == 2 && m->data && m->rows == res.rows && m->cols == res.cols && m->flags == res.flags) {
            if (m->data == res.data) return;
            size_t rb = (size_t)res.cols * res.elemSize();
            for (int y = 0; y < res.rows; y++) std::memmove(m->data + m->step * (size_t)y, res.data + res.step * (size_t)y, rb);
        } else
            *m = res;
    }
};

template <class T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { flags = DataType<T>::type; }
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { fill(v); }
    explicit Mat_(Size s) : Mat(s, DataType<T>::type) {}
    Mat_(Size s, const T& v) : Mat(s, DataType<T>::type) { fill(v); }
    Mat_(const Mat& m) : Mat() {
        if (m.empty() || m.type() == DataType<T>::type) { Mat::operator=(m); }
        else if (m.depth() == DataType<T>::depth) { Mat::operator=(m.reshape(DataType<T>::channels, m.rows)); }
        else { Mat t; m.convertTo(t, DataType<T>::type); Mat::operator=(t); }
    }
    void fill(const T& v) { for (int y = 0; y < rows; y++) { T* p = ptr<T>(y); for (int x = 0; x < cols; x++) p[x] = v; } }
    T& operator()(int y, int x) { return at<T>(y, x); }
    const T& operator()(int y, int x) const { return at<T>(y, x); }
    static Mat_ zeros(int r, int c) { return Mat_(r, c, T()); }
    static Mat_ zeros(Size s) { return Mat_(s, T()); }
    static Mat_ ones(int r, int c) { return Mat_(r, c, onev()); }
    static Mat_ ones(Size s) { return Mat_(s, onev()); }
private:
    template <class U = T> static typename std::enable_if<std::is_arithmetic<U>::value, T>::type onev() { return T(1); }
    template <class U = T> static typename std::enable_if<!std::is_arithmetic<U>::value, T>::type onev() { T v; v[0] = 1; return v; }
};

// ---- element-wise machinery -----------------------------------------------------------------------------------------
inline void check_same(const Mat& a, const Mat& b) {
    if (a.dims != 2 || b.dims != 2 || a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) shim_fail("size/type mismatch");
}
template <class Op> inline Mat map1(const Mat& a, Op op) {  // per element, all channels alike
    Mat r(a.rows, a.cols, a.type());
    int n = a.cols * a.channels();
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        typedef typename work_type<T>::type W;
        for (int y = 0; y < a.rows; y++) {
            const T* pa = a.ptr<T>(y); T* pr = r.ptr<T>(y);
            for (int i = 0; i < n; i++) pr[i] = saturate_cast<T>(op((W)pa[i]));
        }
    });
    return r;
}
template <class Op> inline Mat map2(const Mat& a, const Mat& b, Op op) {
    check_same(a, b);
    Mat r(a.rows, a.cols, a.type());
    int n = a.cols * a.channels();
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        typedef typename work_type<T>::type W;
        for (int y = 0; y < a.rows; y++) {
            const T* pa = a.ptr<T>(y); const T* pb = b.ptr<T>(y); T* pr = r.ptr<T>(y);
            for (int i = 0; i < n; i++) pr[i] = saturate_cast<T>(op((W)pa[i], (W)pb[i]));
        }
    });
    return r;
}
// per channel scalar (channel c uses s[c]; the scalar is first converted to the working type, as OpenCV does)
template <class Op> inline Mat mapS(const Mat& a, const Scalar& s, Op op) {
    Mat r(a.rows, a.cols, a.type());
    int cn = a.channels();
    if (cn > 4) shim_fail("scalar op on > 4 channels");
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        typedef typename work_type<T>::type W;
        W sv[4]; for (int c = 0; c < 4; c++) sv[c] = (W)s.val[c];
        for (int y = 0; y < a.rows; y++) {
            const T* pa = a.ptr<T>(y); T* pr = r.ptr<T>(y);
            for (int x = 0; x < a.cols; x++)
                for (int c = 0; c < cn; c++) pr[x * cn + c] = saturate_cast<T>(op((W)pa[x * cn + c], sv[c]));
        }
    });
    return r;
}

inline Mat Mat::clone() const {
    if (dims != 2) shim_fail("clone of n-D");
    Mat r(rows, cols, type());
    size_t rb = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) std::memcpy(r.data + r.step * (size_t)y, data + step * (size_t)y, rb);
    return r;
}
inline void Mat::copyTo(const OutputArray& dst) const {
    if (dst.m->dims == 2 && dst.m->data && dst.m->rows == rows && dst.m->cols == cols && dst.m->flags == flags) dst.assign(*this);
    else *dst.m = clone();
}
inline void Mat::copyTo(const OutputArray& dst, const Mat& mask) const {
    if (mask.empty()) { copyTo(dst); return; }
    Mat& d = *dst.m;
    if (!(d.dims == 2 && d.data && d.rows == rows && d.cols == cols && d.flags == flags)) d = Mat::zeros(rows, cols, type());
    size_t es = elemSize();
    for (int y = 0; y < rows; y++) {
        const uchar* pm = mask.ptr<uchar>(y);
        for (int x = 0; x < cols; x++)
            if (pm[x]) std::memcpy(d.data + d.step * (size_t)y + es * x, data + step * (size_t)y + es * x, es);
    }
}
// dst = saturate(src * alpha + beta); float destinations compute in float with (float)alpha, (float)beta (cvtScale),
// double destinations in double
inline void Mat::convertTo(const OutputArray& dst, int rtype, double alpha, double beta) const {
    int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
    int cn = channels();
    Mat r(rows, cols, CV_MAKETYPE(ddepth, cn));
    int n = cols * cn;
    bool noscale = (alpha == 1 && beta == 0);
    depth_switch(depth(), [&](auto* stag) {
        typedef typename std::remove_pointer<decltype(stag)>::type S;
        depth_switch(ddepth, [&](auto* dtag) {
            typedef typename std::remove_pointer<decltype(dtag)>::type D;
            for (int y = 0; y < rows; y++) {
                const S* ps = ptr<S>(y); D* pd = r.ptr<D>(y);
                if (noscale) for (int i = 0; i < n; i++) pd[i] = saturate_cast<D>((double)ps[i]);
                else if (std::is_same<D, double>::value || std::is_same<S, double>::value || std::is_same<S, int>::value)
                    for (int i = 0; i < n; i++) pd[i] = saturate_cast<D>((double)ps[i] * alpha + beta);
                else { float a = (float)alpha, b = (float)beta; for (int i = 0; i < n; i++) pd[i] = saturate_cast<D>((float)ps[i] * a + b); }
            }
        });
    });
    dst.assign(r);
}
inline Mat& Mat::setTo(const Scalar& s, const Mat& mask) {
    int cn = channels();
    if (cn > 4) shim_fail("setTo on > 4 channels");
    depth_switch(depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        T sv[4]; for (int c = 0; c < 4; c++) sv[c] = saturate_cast<T>(s.val[c]);
        for (int y = 0; y < rows; y++) {
            T* p = ptr<T>(y); const uchar* pm = mask.empty() ? nullptr : mask.ptr<uchar>(y);
            for (int x = 0; x < cols; x++)
                if (!pm || pm[x]) for (int c = 0; c < cn; c++) p[x * cn + c] = sv[c];
        }
    });
    return *this;
}
inline Mat Mat::reshape(int cn, int newRows) const {
    if (cn == 0) cn = channels();
    if (newRows == 0) newRows = rows;
    if (cn == channels() && newRows == rows) return *this;
    if (!isContinuous()) shim_fail("reshape of a non-continuous matrix");
    size_t totalCh = (size_t)rows * cols * channels();
    if (totalCh % ((size_t)newRows * cn)) shim_fail("bad reshape");
    Mat m(*this);
    m.flags = CV_MAKETYPE(depth(), cn);
    m.rows = newRows; m.cols = (int)(totalCh / ((size_t)newRows * cn));
    m.step = (size_t)m.cols * m.elemSize();
    m.size.p[0] = m.rows; m.size.p[1] = m.cols;
    return m;
}
// float: (float)scale * a * b evaluated left to right when scale != 1 (OpenCV's mul_)
inline Mat Mat::mul(const Mat& m, double scale) const {
    if (scale == 1) return map2(*this, m, [](auto a, auto b) { return a * b; });
    return map2(*this, m, [scale](auto a, auto b) { typedef decltype(a) W; return (W)scale * a * b; });
}
inline Mat Mat::mul(const Scalar& s, double scale) const {
    if (scale == 1) return mapS(*this, s, [](auto a, auto b) { return a * b; });
    return mapS(*this, s, [scale](auto a, auto b) { typedef decltype(a) W; return (W)scale * a * b; });
}
inline double Mat::dot(const Mat& m) const {
    check_same(*this, m);
    double s = 0; int n = cols * channels();
    depth_switch(depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int y = 0; y < rows; y++) { const T* a = ptr<T>(y); const T* b = m.ptr<T>(y); for (int i = 0; i < n; i++) s += (double)a[i] * b[i]; }
    });
    return s;
}

// ---- operators (eager) ----------------------------------------------------------------------------------------------
inline Mat operator+(const Mat& a, const Mat& b) { return map2(a, b, [](auto x, auto y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return map2(a, b, [](auto x, auto y) { return x - y; }); }
inline Mat operator+(const Mat& a, const Scalar& s) { return mapS(a, s, [](auto x, auto y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Scalar& s) { return mapS(a, s, [](auto x, auto y) { return x - y; }); }
inline Mat operator+(const Scalar& s, const Mat& a) { return a + s; }
inline Mat operator-(const Scalar& s, const Mat& a) { return mapS(a, s, [](auto x, auto y) { return y - x; }); }
inline Mat operator+(const Mat& a, double s) { return a + Scalar(s); }
inline Mat operator-(const Mat& a, double s) { return a - Scalar(s); }
inline Mat operator+(double s, const Mat& a) { return a + Scalar(s); }
inline Mat operator-(double s, const Mat& a) { return Scalar(s) - a; }
inline Mat operator*(const Mat& a, double s) { Mat r; a.convertTo(r, -1, s); return r; }   // MatExpr a*alpha -> convertTo
inline Mat operator*(double s, const Mat& a) { return a * s; }
inline Mat operator/(const Mat& a, double s) { Mat r; a.convertTo(r, -1, 1. / s); return r; }
inline Mat operator-(const Mat& a) { return a * -1.0; }
// OpenCV 3.x: a zero denominator gives 0
inline Mat operator/(const Mat& a, const Mat& b) { return map2(a, b, [](auto x, auto y) { typedef decltype(x) W; return y != 0 ? x / y : (W)0; }); }
inline Mat operator/(double s, const Mat& b) { return map1(b, [s](auto y) { typedef decltype(y) W; return y != 0 ? (W)s / y : (W)0; }); }
inline Mat& operator/=(Mat& a, const Mat& b) { OutputArray(a).assign(a / b); return a; }
inline Mat& operator+=(Mat& a, const Mat& b) { OutputArray(a).assign(a + b); return a; }
inline Mat& operator-=(Mat& a, const Mat& b) { OutputArray(a).assign(a - b); return a; }
inline Mat& operator*=(Mat& a, double s) { OutputArray(a).assign(a * s); return a; }
inline Mat operator~(const Mat& a) {
    if (a.depth() != CV_8U) shim_fail("~ on non-8U");
    Mat r(a.rows, a.cols, a.type()); int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) { const uchar* p = a.ptr<uchar>(y); uchar* q = r.ptr<uchar>(y); for (int i = 0; i < n; i++) q[i] = (uchar)~p[i]; }
    return r;
}
template <class Cmp> inline Mat compare_(const Mat& a, double s, Cmp cmp) {
    if (a.channels() != 1) shim_fail("compare on multi-channel");
    Mat r(a.rows, a.cols, CV_8UC1);
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int y = 0; y < a.rows; y++) { const T* p = a.ptr<T>(y); uchar* q = r.ptr<uchar>(y); for (int x = 0; x < a.cols; x++) q[x] = cmp((double)p[x], s) ? 255 : 0; }
    });
    return r;
}
inline Mat operator>(const Mat& a, double s) { return compare_(a, s, [](double x, double y) { return x > y; }); }
inline Mat operator<(const Mat& a, double s) { return compare_(a, s, [](double x, double y) { return x < y; }); }
inline Mat operator==(const Mat& a, double s) { return compare_(a, s, [](double x, double y) { return x == y; }); }
// only named by the reference's post-processing / debug output (PMStereoBase.h:165, FastGCStereo.h:165), which the oracle never runs
inline void dilate(const Mat&, const OutputArray&, const Mat&) { shim_fail("dilate: not part of the hot path"); }
inline bool imwrite(const std::string&, const Mat&) { return false; }
inline std::string format(const char* fmt, ...) { char b[1024]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); return b; }
inline Mat operator>(const Mat& a, const Mat& b) {
    check_same(a, b);
    Mat r(a.rows, a.cols, CV_8UC1);
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int y = 0; y < a.rows; y++) { const T* p = a.ptr<T>(y); const T* q = b.ptr<T>(y); uchar* o = r.ptr<uchar>(y); for (int x = 0; x < a.cols; x++) o[x] = p[x] > q[x] ? 255 : 0; }
    });
    return r;
}
// matrix product (single channel)
inline Mat operator*(const Mat& a, const Mat& b) {
    if (a.channels() != 1 || b.channels() != 1 || a.cols != b.rows || a.type() != b.type()) shim_fail("bad matrix product");
    Mat r(a.rows, b.cols, a.type());
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int i = 0; i < a.rows; i++) for (int j = 0; j < b.cols; j++) { double s = 0; for (int k = 0; k < a.cols; k++) s += (double)a.at<T>(i, k) * b.at<T>(k, j); r.at<T>(i, j) = saturate_cast<T>(s); }
    });
    return r;
}

// ---- free functions -------------------------------------------------------------------------------------------------
inline Mat abs(const Mat& a) { return map1(a, [](auto x) { return x < 0 ? -x : x; }); }
inline void absdiff(const Mat& a, const Mat& b, const OutputArray& dst) { dst.assign(map2(a, b, [](auto x, auto y) { return x > y ? x - y : y - x; })); }
inline void add(const Mat& a, const Mat& b, const OutputArray& dst) { dst.assign(a + b); }
inline void exp(const Mat& a, const OutputArray& dst) { dst.assign(map1(a, [](auto x) { return std::exp(x); })); }
inline void sqrt(const Mat& a, const OutputArray& dst) { dst.assign(map1(a, [](auto x) { return std::sqrt(x); })); }
inline void divide(double s, const Mat& b, const OutputArray& dst) { dst.assign(s / b); }
inline Mat max(double s, const Mat& a) { return map1(a, [s](auto x) { typedef decltype(x) W; return x > (W)s ? x : (W)s; }); }
inline Mat max(const Mat& a, double s) { return max(s, a); }
inline double threshold(const Mat& src, const OutputArray& dst, double thresh, double, int type) {
    if (type != THRESH_TRUNC) shim_fail("only THRESH_TRUNC");
    dst.assign(map1(src, [thresh](auto x) { typedef decltype(x) W; return x > (W)thresh ? (W)thresh : x; }));
    return thresh;
}
inline Scalar sum(const Mat& a) {
    Scalar s; int cn = a.channels();
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int y = 0; y < a.rows; y++) { const T* p = a.ptr<T>(y); for (int x = 0; x < a.cols; x++) for (int c = 0; c < cn && c < 4; c++) s.val[c] += (double)p[x * cn + c]; }
    });
    return s;
}
inline int countNonZero(const Mat& a) {
    int n = 0;
    depth_switch(a.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int y = 0; y < a.rows; y++) { const T* p = a.ptr<T>(y); for (int x = 0; x < a.cols; x++) n += p[x] != 0; }
    });
    return n;
}
inline void split(const Mat& m, std::vector<Mat>& out) {
    int cn = m.channels();
    out.resize(cn);
    for (int c = 0; c < cn; c++) out[c].create(m.rows, m.cols, CV_MAKETYPE(m.depth(), 1));
    size_t es = m.elemSize1();
    for (int y = 0; y < m.rows; y++) {
        const uchar* p = m.data + m.step * (size_t)y;
        for (int c = 0; c < cn; c++) { uchar* q = out[c].data + out[c].step * (size_t)y; for (int x = 0; x < m.cols; x++) std::memcpy(q + es * x, p + es * ((size_t)x * cn + c), es); }
    }
}
inline void merge(const std::vector<Mat>& in, const OutputArray& dst) {
    int cn = (int)in.size();
    if (!cn) shim_fail("merge of nothing");
    for (auto& m : in) if (m.channels() != 1 || m.rows != in[0].rows || m.cols != in[0].cols || m.depth() != in[0].depth()) shim_fail("merge: planes differ");
    Mat r(in[0].rows, in[0].cols, CV_MAKETYPE(in[0].depth(), cn));
    size_t es = r.elemSize1();
    for (int y = 0; y < r.rows; y++) {
        uchar* q = r.data + r.step * (size_t)y;
        for (int c = 0; c < cn; c++) { const uchar* p = in[c].data + in[c].step * (size_t)y; for (int x = 0; x < r.cols; x++) std::memcpy(q + es * ((size_t)x * cn + c), p + es * x, es); }
    }
    dst.assign(r);
}
// REDUCE_SUM along dim 1 (each row to one element): sequential accumulation in the source type for 32F/64F
inline void reduce(const Mat& src, const OutputArray& dst, int dim, int rtype, int dtype = -1) {
    if (dim != 1 || rtype != REDUCE_SUM || src.channels() != 1 || dtype >= 0) shim_fail("reduce: only row sums of 1-channel");
    Mat r(src.rows, 1, src.type());
    depth_switch(src.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        typedef typename work_type<T>::type W;
        for (int y = 0; y < src.rows; y++) { const T* p = src.ptr<T>(y); W a = (W)p[0]; for (int x = 1; x < src.cols; x++) a = a + (W)p[x]; r.at<T>(y, 0) = saturate_cast<T>(a); }
    });
    dst.assign(r);
}
inline void copyMakeBorder(const Mat& src, const OutputArray& dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar()) {
    Mat r(src.rows + top + bottom, src.cols + left + right, src.type());
    size_t es = src.elemSize();
    if (borderType == BORDER_CONSTANT) {
        r.setTo(value);
        src.copyTo(r(Rect(left, top, src.cols, src.rows)));
    } else if (borderType == BORDER_REPLICATE) {
        for (int y = 0; y < r.rows; y++) {
            int sy = std::min(std::max(y - top, 0), src.rows - 1);
            for (int x = 0; x < r.cols; x++) { int sx = std::min(std::max(x - left, 0), src.cols - 1); std::memcpy(r.data + r.step * (size_t)y + es * x, src.data + src.step * (size_t)sy + es * sx, es); }
        }
    } else shim_fail("copyMakeBorder: border type");
    dst.assign(r);
}

// Unnormalised box sum, zero outside the matrix (BORDER_CONSTANT); 32F/64F sources accumulate in double, as
// cv::boxFilter does (sum type CV_64F): a sliding row sum followed by a sliding column sum.
inline void boxFilter(const Mat& src, const OutputArray& dst, int ddepth, Size ksize, Point anchor = Point(-1, -1), bool normalize = true, int borderType = BORDER_DEFAULT) {
    if (ksize.width != ksize.height || !(ksize.width & 1) || borderType != BORDER_CONSTANT || src.channels() != 1 || (anchor.x != -1 && anchor.x != ksize.width / 2))
        shim_fail("boxFilter: only odd square kernels, BORDER_CONSTANT, 1 channel");
    if (ddepth >= 0 && ddepth != src.depth()) shim_fail("boxFilter: ddepth");
    const int R = ksize.width / 2, h = src.rows, w = src.cols;
    const double nrm = normalize ? 1.0 / ((double)ksize.width * ksize.height) : 1.0;
    Mat r(h, w, src.type());
    std::vector<double> rowsum((size_t)h * w), col(w);
    depth_switch(src.depth(), [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type T;
        for (int y = 0; y < h; y++) {
            const T* s = src.ptr<T>(y); double* t = &rowsum[(size_t)y * w];
            double run = 0;
            for (int x = 0; x < R && x < w; x++) run += (double)s[x];
            for (int x = 0; x < w; x++) {
                if (x + R < w) run += (double)s[x + R];
                if (x - R - 1 >= 0) run -= (double)s[x - R - 1];
                t[x] = run;
            }
        }
        std::fill(col.begin(), col.end(), 0.0);
        for (int y = 0; y < R && y < h; y++) for (int x = 0; x < w; x++) col[x] += rowsum[(size_t)y * w + x];
        for (int y = 0; y < h; y++) {
            if (y + R < h) { const double* t = &rowsum[(size_t)(y + R) * w]; for (int x = 0; x < w; x++) col[x] += t[x]; }
            if (y - R - 1 >= 0) { const double* t = &rowsum[(size_t)(y - R - 1) * w]; for (int x = 0; x < w; x++) col[x] -= t[x]; }
            T* d = r.ptr<T>(y);
            if (normalize) for (int x = 0; x < w; x++) d[x] = saturate_cast<T>(col[x] * nrm);
            else for (int x = 0; x < w; x++) d[x] = saturate_cast<T>(col[x]);
        }
    });
    dst.assign(r);
}

// BGR -> gray on 32F: B*0.114f + G*0.587f + R*0.299f in float, left to right
inline void cvtColor(const Mat& src, const OutputArray& dst, int code) {
    if (code != COLOR_BGR2GRAY || src.type() != CV_32FC3) shim_fail("cvtColor: only BGR2GRAY on 32FC3");
    Mat r(src.rows, src.cols, CV_32FC1);
    for (int y = 0; y < src.rows; y++) {
        const float* p = src.ptr<float>(y); float* q = r.ptr<float>(y);
        for (int x = 0; x < src.cols; x++) { float t = p[3 * x] * 0.114f + p[3 * x + 1] * 0.587f; q[x] = t + p[3 * x + 2] * 0.299f; }
    }
    dst.assign(r);
}
// Sobel dx=1, dy=0, ksize=1: the 3x1 kernel [-1 0 1] (no smoothing across rows), then * scale, + delta, in float
inline void Sobel(const Mat& src, const OutputArray& dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT) {
    if (src.type() != CV_32FC1 || ddepth != CV_32F || dx != 1 || dy != 0 || ksize != 1 || borderType != BORDER_REPLICATE) shim_fail("Sobel: only the x-derivative, ksize 1, replicate");
    Mat r(src.rows, src.cols, CV_32FC1);
    const float sc = (float)scale, de = (float)delta;
    for (int y = 0; y < src.rows; y++) {
        const float* p = src.ptr<float>(y); float* q = r.ptr<float>(y);
        for (int x = 0; x < src.cols; x++) { float a = p[std::max(x - 1, 0)], b = p[std::min(x + 1, src.cols - 1)]; q[x] = (b - a) * sc + de; }
    }
    dst.assign(r);
}
// Gaussian elimination with partial pivoting in double (cv::solve, DECOMP_LU) for the 6x6 system of getAffineTransform
inline Mat getAffineTransform(const Point2f src[], const Point2f dst[]) {
    double a[6][7];
    for (int i = 0; i < 3; i++) {
        double r0[7] = {src[i].x, src[i].y, 1, 0, 0, 0, dst[i].x}, r1[7] = {0, 0, 0, src[i].x, src[i].y, 1, dst[i].y};
        std::memcpy(a[i], r0, sizeof r0); std::memcpy(a[i + 3], r1, sizeof r1);
    }
    Mat M(2, 3, CV_64FC1);
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (std::fabs(a[piv][c]) < 2.220446049250313e-16) { M.setTo(Scalar(0)); return M; }
        if (piv != c) for (int k = 0; k < 7; k++) std::swap(a[c][k], a[piv][k]);
        double d = -1 / a[c][c];
        for (int r = c + 1; r < 6; r++) { double f = a[r][c] * d; for (int k = c + 1; k < 7; k++) a[r][k] += f * a[c][k]; }
    }
    double x[6];
    for (int r = 5; r >= 0; r--) { double s = a[r][6]; for (int k = r + 1; k < 6; k++) s -= a[r][k] * x[k]; x[r] = s / a[r][r]; }
    for (int i = 0; i < 6; i++) M.at<double>(i / 3, i % 3) = x[i];
    return M;
}
// warpAffine, INTER_LINEAR, BORDER_REPLICATE, 32F sources: M is inverted in double, coordinates are 10-bit fixed point
// (AB_BITS) rounded to 1/32 pixel (INTER_BITS = 5), bilinear weights are the float table values (1-fy)(1-fx) ...
inline void warpAffine(const Mat& src, const OutputArray& dst, const Mat& M0, Size dsize, int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT, const Scalar& = Scalar()) {
    if (flags != INTER_LINEAR || borderMode != BORDER_REPLICATE || src.depth() != CV_32F || M0.rows != 2 || M0.cols != 3) shim_fail("warpAffine: only linear/replicate/32F");
    double M[6];
    for (int i = 0; i < 6; i++) M[i] = M0.depth() == CV_64F ? M0.at<double>(i / 3, i % 3) : (double)M0.at<float>(i / 3, i % 3);
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    const int cn = src.channels(), w = dsize.width, h = dsize.height, W = src.cols, H = src.rows;
    Mat r(h, w, src.type());
    std::vector<int> adelta(w), bdelta(w);
    for (int x = 0; x < w; x++) { adelta[x] = cvRound(M[0] * x * 1024); bdelta[x] = cvRound(M[3] * x * 1024); }
    for (int y = 0; y < h; y++) {
        int X0 = cvRound((M[1] * y + M[2]) * 1024) + 16, Y0 = cvRound((M[4] * y + M[5]) * 1024) + 16;
        float* q = r.ptr<float>(y);
        for (int x = 0; x < w; x++) {
            int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            int sx = X >> 5, sy = Y >> 5;
            float fx = (X & 31) * (1.f / 32), fy = (Y & 31) * (1.f / 32);
            float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
            int x0 = std::min(std::max(sx, 0), W - 1), x1 = std::min(std::max(sx + 1, 0), W - 1);
            int y0 = std::min(std::max(sy, 0), H - 1), y1 = std::min(std::max(sy + 1, 0), H - 1);
            const float *p00 = src.ptr<float>(y0) + x0 * cn, *p01 = src.ptr<float>(y0) + x1 * cn, *p10 = src.ptr<float>(y1) + x0 * cn, *p11 = src.ptr<float>(y1) + x1 * cn;
            for (int c = 0; c < cn; c++) { float v = p00[c] * w00; v = v + p01[c] * w01; v = v + p10[c] * w10; v = v + p11[c] * w11; q[x * cn + c] = v; }
        }
    }
    dst.assign(r);
}
inline void resize(const Mat&, const OutputArray&, Size, double = 0, double = 0, int = INTER_LINEAR) { shim_fail("resize: not on the path"); }
inline bool solve(const Mat&, const Mat&, const OutputArray&, int = DECOMP_LU) { shim_fail("solve: not on the path (RansacProposer)"); }

// ---- cv::RNG (multiply-with-carry) ------------------------------------------------------------------------------------
class RNG {
public:
    uint64 state;
    RNG() : state(0xffffffff) {}
    RNG(uint64 s) : state(s ? s : 0xffffffff) {}
    unsigned next() { state = (uint64)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
    operator unsigned() { return next(); }
    operator float() { return next() * 2.3283064365386962890625e-10f; }
    operator double() { unsigned t = next(); return (((uint64)t << 32) | next()) * 5.4210108624275221700372640043497e-20; }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (b - a) + a); }
    float uniform(float a, float b) { return ((float)*this) * (b - a) + a; }
    double uniform(double a, double b) { return ((double)*this) * (b - a) + a; }
};
inline RNG& theRNG() { static thread_local RNG r; return r; }

}  // namespace cv
