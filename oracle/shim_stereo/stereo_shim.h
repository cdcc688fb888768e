// OpenCV / ORB-SLAM2 *stub* for the stereo matcher -- TEST INFRASTRUCTURE (oracle/), not product code.
//
// oracle/build_stereo_ref.sh compiles the reference's own lines of Frame::ComputeStereoMatches
// (/root/reference/src/Frame.cc:466-640) and ORBmatcher::DescriptorDistance / TH_HIGH / TH_LOW
// (/root/reference/src/ORBmatcher.cc:37-38, :1647-1663), taken from where they lie at build time, against
// this header: the handful of cv:: types those lines use (8U / 32F Mat views, convertTo, ones, scalar * Mat,
// Mat - Mat, norm L1) and the members of Frame / ORBextractor / ORBmatcher they touch.  Written from scratch.
#ifndef ORBX_ORACLE_STEREO_SHIM_H
#define ORBX_ORACLE_STEREO_SHIM_H
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

using namespace std;      // src/Frame.cc sees std through its includes (ORBmatcher.cc:32) and uses vector/pair/sort unqualified

#define CV_8U 0
#define CV_32F 5

namespace cv {

struct Point2f { float x, y; };
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

enum { NORM_L1 = 2 };

class Mat {
public:
    int rows, cols, depth;          // depth: CV_8U or CV_32F
    unsigned char* data;
    size_t step;                    // bytes
    std::shared_ptr<std::vector<unsigned char>> own;

    Mat() : rows(0), cols(0), depth(CV_8U), data(0), step(0) {}
    Mat(int r, int c, int d) : rows(r), cols(c), depth(d), data(0), step(0) { alloc(); }
    Mat(int r, int c, int d, void* ext, size_t s) : rows(r), cols(c), depth(d), data((unsigned char*)ext), step(s) {}
    size_t esz() const { return depth == CV_32F ? 4 : 1; }
    void alloc() {
        step = (size_t)cols * esz();
        own = std::make_shared<std::vector<unsigned char>>((size_t)rows * step + 16);
        data = own->data();
    }
    Mat sub(int r0, int r1, int c0, int c1) const {
        Mat m(*this);
        m.data = data + (ptrdiff_t)r0 * (ptrdiff_t)step + (ptrdiff_t)c0 * (ptrdiff_t)esz();
        m.rows = r1 - r0;
        m.cols = c1 - c0;
        return m;
    }
    Mat row(int r) const { return sub(r, r + 1, 0, cols); }
    Mat rowRange(int a, int b) const { return sub(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return sub(0, rows, a, b); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    // dst may be *this (IL.convertTo(IL, CV_32F), Frame.cc:566): build the result first, then assign
    void convertTo(Mat& dst, int d) const {
        Mat out(rows, cols, d);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const float v = depth == CV_32F ? at<float>(r, c) : (float)at<unsigned char>(r, c);
                if (d == CV_32F) out.at<float>(r, c) = v; else out.at<unsigned char>(r, c) = (unsigned char)v;
            }
        dst = out;
    }
    static Mat ones(int r, int c, int d) {
        Mat m(r, c, d);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j)
                if (d == CV_32F) m.at<float>(i, j) = 1.f; else m.at<unsigned char>(i, j) = 1;
        return m;
    }
};

static inline Mat operator*(float s, const Mat& m) {            // 32F only (Frame.cc:567, :585)
    Mat o(m.rows, m.cols, CV_32F);
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c) o.at<float>(r, c) = s * m.at<float>(r, c);
    return o;
}
static inline Mat operator-(const Mat& a, const Mat& b) {       // 32F only
    Mat o(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) o.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c);
    return o;
}
static inline double norm(const Mat& a, const Mat& b, int type) {    // NORM_L1 of 32F matrices, accumulated in double
    (void)type;
    double s = 0;
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}

}  // namespace cv

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};

struct ORBextractor {
    std::vector<cv::Mat> mvImagePyramid;
};

class Frame {
public:
    void ComputeStereoMatches();
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mbf, mb;
};

}  // namespace ORB_SLAM2
#endif
