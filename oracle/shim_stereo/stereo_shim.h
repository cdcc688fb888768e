// OpenCV / ORB-SLAM2 *stub* for the Frame-level consumers -- TEST INFRASTRUCTURE (oracle/), not product code.
//
// oracle/build_stereo_ref.sh compiles the reference's own lines of Frame::ComputeStereoMatches
// (/root/reference/src/Frame.cc:466-640), ORBmatcher::DescriptorDistance / TH_HIGH / TH_LOW
// (/root/reference/src/ORBmatcher.cc:37-38, :1647-1663) and Frame::AssignFeaturesToGrid / PosInGrid /
// UndistortKeyPoints / ComputeImageBounds (Frame.cc:230-245, :382-392, :404-434, :436-464), Frame::GetFeaturesInArea
// (:327-380), ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) / ComputeThreeMaxima
// (ORBmatcher.cc:1328-1470, :1601-1642) and ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) /
// RadiusByViewingCos (ORBmatcher.cc:45-129, :131-137), ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (ORBmatcher.cc:159-288),
// ORBmatcher::SearchForInitialization (ORBmatcher.cc:405-520), ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th,
// ORBdist) (ORBmatcher.cc:1472-1599), MapPoint::GetMin/MaxDistanceInvariance / PredictScale(dist, Frame*) (MapPoint.cc:373-383, :402-417)
// and Frame::isInFrustum (Frame.cc:269-325),
// taken from where they lie at
// build time, against this header: the handful of cv:: types those lines use (8U / 32F Mat views, convertTo, ones, scalar * Mat,
// Mat - Mat, norm L1) and the members of Frame / ORBextractor / ORBmatcher they touch.  Written from scratch.
#ifndef ORBX_ORACLE_STEREO_SHIM_H
#define ORBX_ORACLE_STEREO_SHIM_H
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
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

class Mat;
struct MatT;      // Mat::t() with a scale (MatExpr of OpenCV, only as far as ORBmatcher.cc:1340-1348 needs it)
struct MatMul;    // A * B, evaluated when it meets "+ C" or a Mat

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
    Mat col(int c) const { return sub(0, rows, c, c + 1); }
    inline MatT t() const;
    inline Mat(const MatMul& e);
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }                 // vectors (Frame.cc:406)
    template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
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
    Mat reshape(int cn) const { (void)cn; return *this; }      // N x 2 floats either way (Frame.cc:421-423, :449-451)
    // Mat::dot of two CV_32F vectors (Frame.cc:308): OpenCV's dotProd_32f accumulates the float products in double (for 3
    // elements no SIMD block is reached; the float products are exact in double)
    double dot(const Mat& o) const {
        double r = 0;
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) r += (double)at<float>(i, j) * (double)o.at<float>(i, j);
        return r;
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
// cv::norm(m) of a CV_32F matrix: NORM_L2, squares accumulated in double (OpenCV's normL2_32f returns the double sum, then
// std::sqrt) -- used for the 3 x 1 vector PO (ORBmatcher.cc:1511); pinned against the real cv2.norm by tests/test_match_oracle.py
static inline double norm(const Mat& a) {
    double s = 0;
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) { const double v = (double)a.at<float>(r, c); s += v * v; }
    return std::sqrt(s);
}
static inline double norm(const Mat& a, const Mat& b, int type) {    // NORM_L1 of 32F matrices, accumulated in double
    (void)type;
    double s = 0;
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}

// cv::gemm for the 3x3 * 3x1 float products of ORBmatcher.cc:1337-1348, restated from OpenCV's published algorithm and
// pinned bit for bit against the real cv2.gemm (tests/test_match_oracle.py):
//   A * B (+ C), no flags (the len <= 4 special case): t0 = fl32(fl32(a0*b0 + a1*b1) + a2*b2) in float32, then
//                 d = fl32(double(t0) * alpha + double(c) * beta);
//   A.t() * B scaled (GEMM_1_T, general path): products and sum in double, d = fl32(alpha * s).
struct MatT { Mat m; double alpha; };
struct MatMul { Mat a, b; };
inline MatT Mat::t() const { MatT r; r.m = *this; r.alpha = 1.0; return r; }
static inline MatT operator-(const MatT& a) { MatT r = a; r.alpha = -a.alpha; return r; }
static inline MatMul operator*(const Mat& a, const Mat& b) { MatMul r; r.a = a; r.b = b; return r; }
static inline Mat gemm_small(const Mat& a, const Mat& b, const Mat* c) {
    Mat o(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int q = 0; q < b.cols; ++q) {
            float t0 = a.at<float>(r, 0) * b.at<float>(0, q);
            for (int k = 1; k < a.cols; ++k) { const float p = a.at<float>(r, k) * b.at<float>(k, q); t0 = t0 + p; }
            o.at<float>(r, q) = (float)((double)t0 * 1.0 + (c ? (double)c->at<float>(r, q) * 1.0 : 0.0));
        }
    return o;
}
inline Mat::Mat(const MatMul& e) { *this = gemm_small(e.a, e.b, 0); }
static inline Mat operator+(const MatMul& e, const Mat& c) { return gemm_small(e.a, e.b, &c); }
static inline Mat operator*(const MatT& a, const Mat& b) {       // (alpha * A^T) * B
    Mat o(a.m.cols, b.cols, CV_32F);
    for (int r = 0; r < a.m.cols; ++r)
        for (int q = 0; q < b.cols; ++q) {
            double s = 0;
            for (int k = 0; k < a.m.rows; ++k) s += (double)a.m.at<float>(k, r) * (double)b.at<float>(k, q);
            o.at<float>(r, q) = (float)(s * a.alpha);
        }
    return o;
}

// cv::undistortPoints(src, dst, K, D, R = empty, P = K) for N x 2 float points: the iterative inverse of OpenCV's
// distortion model, 5 iterations, evaluated in double (restated from the published algorithm; pinned against the real
// cv2.undistortPoints bit for bit by tests/test_frame_oracle.py).  K, D: CV_32F; D holds k1 k2 p1 p2 [k3].
static inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D, const Mat& R, const Mat& P) {
    (void)R;
    const double fx = K.at<float>(0, 0), fy = K.at<float>(1, 1), cx = K.at<float>(0, 2), cy = K.at<float>(1, 2);
    const double pfx = P.at<float>(0, 0), pfy = P.at<float>(1, 1), pcx = P.at<float>(0, 2), pcy = P.at<float>(1, 2);
    double k[12] = {0};
    const int nd = D.rows * D.cols;
    for (int i = 0; i < nd && i < 12; ++i) k[i] = ((const float*)D.data)[i];
    const double ifx = 1. / fx, ify = 1. / fy;
    Mat out(src.rows, 2, CV_32F);
    for (int i = 0; i < src.rows; ++i) {
        double x = (src.at<float>(i, 0) - cx) * ifx, y = (src.at<float>(i, 1) - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = pfx * x + 0 * y + pcx, yy = 0 * x + pfy * y + pcy, ww = 1. / (0 * x + 0 * y + 1.);
        out.at<float>(i, 0) = (float)(xx * ww);
        out.at<float>(i, 1) = (float)(yy * ww);
    }
    dst = out;
}

}  // namespace cv

#define FRAME_GRID_ROWS 48          // include/Frame.h:39-40
#define FRAME_GRID_COLS 64

namespace DBoW2 {                       // Thirdparty/DBoW2/DBoW2/FeatureVector.h:21-22 is a std::map with one extra method
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;
}

namespace ORB_SLAM2 {

class Frame;
class KeyFrame;

struct MapPoint {                     // the accessors and tracking fields ORBmatcher.cc:45-129, :1328-1470, :1472-1599 touch
    cv::Mat mWorldPos, mDescriptor, mNormalVector;
    int nObs;
    bool mbTrackInView, mbBad;        // include/MapPoint.h:92-97
    float mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos;
    int mnTrackScaleLevel;
    float mfMinDistance, mfMaxDistance;   // include/MapPoint.h:140-141
    std::mutex mMutexPos;
    MapPoint() : nObs(0), mbTrackInView(false), mbBad(false), mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mTrackViewCos(0),
                 mnTrackScaleLevel(0), mfMinDistance(0), mfMaxDistance(0) {}
    MapPoint(const MapPoint& o) : mWorldPos(o.mWorldPos), mDescriptor(o.mDescriptor), mNormalVector(o.mNormalVector), nObs(o.nObs), mbTrackInView(o.mbTrackInView),
                                  mbBad(o.mbBad), mTrackProjX(o.mTrackProjX), mTrackProjY(o.mTrackProjY), mTrackProjXR(o.mTrackProjXR),
                                  mTrackViewCos(o.mTrackViewCos), mnTrackScaleLevel(o.mnTrackScaleLevel),
                                  mfMinDistance(o.mfMinDistance), mfMaxDistance(o.mfMaxDistance) {}
    bool isBad() { return mbBad; }
    cv::Mat GetWorldPos() { return mWorldPos; }
    cv::Mat GetDescriptor() { return mDescriptor; }
    cv::Mat GetNormal() { return mNormalVector; }
    int Observations() { return nObs; }
    float GetMinDistanceInvariance();                       // the reference's own lines (MapPoint.cc:373-383, :402-417)
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, Frame* pF);
};

class ORBmatcher {
public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th = 3);    // include/ORBmatcher.h
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);                                             // include/ORBmatcher.h:69
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                           const int ORBdist);                                                    // include/ORBmatcher.h:56
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

struct ORBextractor {
    std::vector<cv::Mat> mvImagePyramid;
};

class Frame {
public:
    void ComputeStereoMatches();
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    void UndistortKeyPoints();
    void ComputeImageBounds(const cv::Mat& imLeft);
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);                        // include/Frame.h:83
    cv::Mat mRcw, mtcw, mOw;                                                       // include/Frame.h:199-202 (UpdatePoseMatrices)
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                     const int maxLevel = -1) const;               // include/Frame.h:92
    static float fx, fy, cx, cy;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mTcw;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mK, mDistCoef;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    static float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mbf, mb;
    int mnScaleLevels;                  // include/Frame.h:182-184
    float mfLogScaleFactor;
};

class KeyFrame {                        // what ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) reads (src/ORBmatcher.cc:159-288)
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeysUn;
};

}  // namespace ORB_SLAM2
#endif
