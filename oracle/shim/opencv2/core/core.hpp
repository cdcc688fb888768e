// OpenCV *shim* -- TEST INFRASTRUCTURE (oracle/), not product code.
//
// The reference's hot-path translation unit (/root/reference/src/ORBextractor.cc)
// needs the OpenCV C++ SDK, which this image does not have.  This header declares
// the small subset of cv:: that the TU uses so that the
// UNMODIFIED reference source compiles; oracle/shim/cvshim.cpp implements the
// five external primitives with the arithmetic verified bit-exact against
// cv2 4.13.0.  Written from scratch; shares no code with OpenCV.
#ifndef ORBX_ORACLE_OPENCV_SHIM_CORE_HPP
#define ORBX_ORACLE_OPENCV_SHIM_CORE_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <list>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

static inline int cvRound(double v) { return (int)std::lrint(v); }     // round half to even
static inline int cvRound(float v) { return (int)std::lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return (int)std::floor(v); }
static inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) {
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);   // dead code in the reference
};

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

class Mat;

// Result of Mat::zeros(): assigning it to a Mat of the same shape fills in place
// (what cv::MatExpr does; src/ORBextractor.cc:1037 depends on it).
struct MatZeros {
    int rows, cols, type;
};

// 8-bit single-channel, ref-counted plane with ROI views.
class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;

    Mat() : rows(0), cols(0), data(0), step(0), buf_(0) {}
    Mat(Size sz, int type) : rows(0), cols(0), data(0), step(0), buf_(0) { (void)type; create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), data(0), step(0), buf_(0) { create(r, c, type); }
    // non-owning header over external memory
    Mat(int r, int c, int type, void* ext, size_t _step) : rows(r), cols(c), data((uchar*)ext), step(_step), buf_(0) { (void)type; }
    Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), buf_(m.buf_) { retain(); }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) {
        if (this != &m) {
            Buf* nb = m.buf_;
            if (nb) nb->refs++;
            release();
            rows = m.rows; cols = m.cols; data = m.data; step = m.step; buf_ = nb;
        }
        return *this;
    }
    Mat& operator=(const MatZeros& z) {
        if (!data || rows != z.rows || cols != z.cols) create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols);
        return *this;
    }
    void create(int r, int c, int type) {
        (void)type;
        if (data && rows == r && cols == c) return;
        release();
        rows = r; cols = c; step = (size_t)c;
        size_t bytes = (size_t)r * (size_t)c;
        buf_ = (Buf*)std::malloc(sizeof(Buf) + (bytes ? bytes : 1));
        buf_->refs = 1;
        data = (uchar*)(buf_ + 1);
    }
    void release() {
        if (buf_ && --buf_->refs == 0) std::free(buf_);
        buf_ = 0; data = 0; rows = cols = 0; step = 0;
    }
    Mat operator()(const Rect& r) const {
        Mat v(*this);
        v.data = data + (size_t)r.y * step + r.x;
        v.rows = r.height; v.cols = r.width;
        return v;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat clone() const {
        Mat m(rows, cols, 0);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    int type() const { return CV_8UC1; }
    size_t step1(int i = 0) const { (void)i; return step; }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols; }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + c * sizeof(T)); }
    static MatZeros zeros(int r, int c, int type) { MatZeros z = {r, c, type}; return z; }

private:
    struct Buf { long refs; long pad; };
    Buf* buf_;
    void retain() { if (buf_) buf_->refs++; }
};

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return *m_; }
    bool empty() const { return m_->empty(); }
private:
    const Mat* m_;
};

class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    Mat getMat() const { return *m_; }
    Mat& getMatRef() const { return *m_; }
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
private:
    Mat* m_;
};

typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
float fastAtan2(float y, float x);

}  // namespace cv

#endif
