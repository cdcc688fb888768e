// Drop-in replacement for the reference's include/ORBextractor.h (yxqc/ORBSLAM2_with_quadrics).
//
// Same namespace, class name, constructor, operator(), getters and public mvImagePyramid as
// reference include/ORBextractor.h:45-85, so Frame, Tracking, KeyFrame and the quadric_slam
// modules compile against it unchanged (they only need source compatibility: the extractor is
// one TU of libORB_SLAM2.so, reference CMakeLists.txt:57).  All work happens on the GPU behind
// the C ABI in include/orbx.h; this class only moves cv:: types in and out.
//
// Differences a maintainer should know about:
//  * ExtractorNode (reference .h:32-43) is gone: nothing outside the old TU used it.
//  * The destructor is out of line (it releases the GPU handle).
//  * mvImagePyramid[l] are non-owning cv::Mat headers over the library's pinned host buffer; as in
//    the reference they are valid until the next operator() on the same object.
//  * Errors (CUDA failure, unsupported geometry) throw std::runtime_error; there is no silent CPU
//    fallback.  A FAST candidate-buffer overflow (pathological texture) is not an error: operator()
//    rebuilds the handle with larger buffers and runs the frame again.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <utility>
#include <vector>
#include <opencv2/core/core.hpp>

struct orbx_handle;
struct orbx_vocabulary;

namespace ORB_SLAM2
{

class ORBextractor
{
public:

    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels,
                 int iniThFAST, int minThFAST);

    ~ORBextractor();

    // Compute the ORB features and descriptors on an image.
    // Mask is ignored, as in the reference implementation.
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints,
      cv::OutputArray descriptors);

    int GetLevels();
    float GetScaleFactor();
    std::vector<float> GetScaleFactors();
    std::vector<float> GetInverseScaleFactors();
    std::vector<float> GetScaleSigmaSquares();
    std::vector<float> GetInverseScaleSigmaSquares();

    std::vector<cv::Mat> mvImagePyramid;

    // --- additions (not in the reference) ---
    // Monocular / RGB-D pipelines never read mvImagePyramid (only Frame::ComputeStereoMatches does,
    // reference src/Frame.cc:563-580); switching the download off keeps the pyramid in HBM and removes
    // the largest device->host copy.  Default: on (exact drop-in behaviour).
    void SetPyramidDownload(bool enable);
    // CUDA device used by extractors created afterwards in this thread's process (default 0).
    static void SetDevice(int device);
    // Colour input: with this set, operator() also accepts CV_8UC3 / CV_8UC4 images -- the frames Tracking::GrabImage*
    // receives (reference src/Tracking.cc:172-255) -- and does the cvtColor(RGB/BGR(A) -> GRAY) on the GPU, bit-identical
    // to cv::cvtColor.  rgb = Tracking::mbRGB.  The cvtColor calls in GrabImage* can then be dropped.
    void SetColorOrder(bool rgb);
    // Frame::ComputeStereoMatches (reference src/Frame.cc:466-640) on the GPU, over the results both extractors still
    // hold in HBM from their last operator(): fills mvuRight / mvDepth exactly as the reference does (-1 = no match).
    // Call it from Frame::ComputeStereoMatches after the two ExtractORB threads have joined (src/Frame.cc:78-81); with
    // it the stereo pipeline can run with SetPyramidDownload(false).  mbf, mb: Frame::mbf, Frame::mb.
    static void ComputeStereoMatches(ORBextractor& left, ORBextractor& right, float mbf, float mb,
                                     std::vector<float>& mvuRight, std::vector<float>& mvDepth);

    // Frame::UndistortKeyPoints + Frame::AssignFeaturesToGrid (reference src/Frame.cc:404-434, :230-245) on the GPU, over
    // the keypoints this extractor still holds in HBM from its last operator(): fills mvKeysUn (= mvKeys with undistorted
    // pt), mGrid[FRAME_GRID_COLS = 64][FRAME_GRID_ROWS = 48] in the reference's push_back order and the image bounds of
    // Frame::ComputeImageBounds (:436-464).  mK, mDistCoef: Frame::mK (3x3 CV_32F), Frame::mDistCoef (4 or 5 x 1 CV_32F).
    void UndistortAndAssignToGrid(const cv::Mat& mK, const cv::Mat& mDistCoef, const std::vector<cv::KeyPoint>& mvKeys,
                                  std::vector<cv::KeyPoint>& mvKeysUn, std::vector<std::size_t> (*mGrid)[48],
                                  float& mnMinX, float& mnMaxX, float& mnMinY, float& mnMaxY);

    // ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (reference
    // src/ORBmatcher.cc:1328-1470) on the GPU: `this` is CurrentFrame.mpORBextractorLeft, whose last operator(),
    // UndistortAndAssignToGrid and (with useStereo) ComputeStereoMatches left mvKeysUn, mDescriptors, mGrid and mvuRight in
    // HBM.  The LastFrame side is gathered by the caller from LastFrame.mvpMapPoints: world position (3 floats),
    // representative descriptor (32 bytes) and Observations() per keypoint, obs < 0 where there is no map point or
    // mvbOutlier is set.  matchedLast[i2] = index into LastFrame of the map point CurrentFrame.mvpMapPoints[i2] receives,
    // -1 = NULL.  Returns nmatches.  CurrentFrame.mvpMapPoints is taken to be all NULL on entry, as at both call sites
    // (src/Tracking.cc:871, :890).
    int SearchByProjection(const std::vector<float>& lastWorldPos, const std::vector<unsigned char>& lastDescriptors,
                           const std::vector<int>& lastObservations, const std::vector<cv::KeyPoint>& lastKeysUn,
                           const cv::Mat& TcwCurrent, const cv::Mat& TcwLast, const cv::Mat& mK, float mbf, float mb, float th,
                           bool bMono, bool checkOrientation, bool useStereo, std::vector<int>& matchedLast);

    // Frame::isInFrustum(MapPoint*, viewingCosLimit) (reference src/Frame.cc:269-325, with MapPoint::PredictScale) for the whole
    // local map on the GPU: the loop of Tracking::SearchLocalPoints (src/Tracking.cc:1165-1178).  Per map point: consider = it
    // is handed to isInFrustum at all (not seen in this frame, !isBad()), world position and mean viewing direction (3 floats
    // each), mfMinDistance and mfMaxDistance.  Tcw = F.mTcw (4 x 4 CV_32F), K = F.mK, bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.
    // Outputs are SearchLocalPoints' inputs: mbTrackInView, (mTrackProjX, mTrackProjY, mTrackProjXR), mnTrackScaleLevel,
    // mTrackViewCos.  Returns nToMatch.  Reads nothing of the last operator(): may be called before it.
    int IsInFrustum(const std::vector<unsigned char>& consider, const std::vector<float>& worldPos, const std::vector<float>& normals,
                    const std::vector<float>& minDistance, const std::vector<float>& maxDistance, const cv::Mat& Tcw,
                    const cv::Mat& mK, float mbf, const float bounds[4], float logScaleFactor, float viewingCosLimit,
                    std::vector<unsigned char>& inView, std::vector<float>& projXYXR, std::vector<int>& scaleLevel,
                    std::vector<float>& viewCos);

    // ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (reference src/ORBmatcher.cc:45-129,
    // the matcher of Tracking::SearchLocalPoints) on the GPU; `this` is F.mpORBextractorLeft as above.  Per map point: the
    // tracking fields Frame::isInFrustum left on it -- inView = mbTrackInView && !isBad(), (mTrackProjX, mTrackProjY,
    // mTrackProjXR) as 3 floats, mnTrackScaleLevel, mTrackViewCos -- its descriptor (32 bytes) and Observations().
    // currentObservations[i2] = -1 where F.mvpMapPoints[i2] is NULL, else that point's Observations().
    // matched[i2] = index into vpMapPoints F.mvpMapPoints[i2] has to be set to, -1 = leave it.  Returns nmatches.
    int SearchLocalPoints(const std::vector<unsigned char>& inView, const std::vector<float>& projXYXR,
                          const std::vector<int>& scaleLevel, const std::vector<float>& viewCos,
                          const std::vector<unsigned char>& descriptors, const std::vector<int>& observations,
                          const std::vector<int>& currentObservations, float th, float nnRatio, bool useStereo,
                          std::vector<int>& matched);

    // Frame::ComputeBoW (reference src/Frame.cc:395-402: mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)) on
    // the GPU, over the descriptors this extractor still holds in HBM from its last operator().  `voc` is the ORB
    // vocabulary uploaded once with orbx_vocabulary_create (include/orbx.h; INTEGRATION.md shows the flattening of
    // DBoW2's m_nodes).  bow: (word id, value) in std::map order, L1-normalised doubles bit-identical to DBoW2's;
    // featVec: (node id, feature index) in std::map / push_back order.
    void ComputeBoW(const orbx_vocabulary* voc, std::vector<std::pair<unsigned int, double> >& bow,
                    std::vector<std::pair<unsigned int, unsigned int> >& featVec, int levelsup = 4);

    // ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (reference src/ORBmatcher.cc:159-288) on the GPU;
    // `this` is F.mpORBextractorLeft after ComputeBoW.  KeyFrame side: descriptors (32 bytes per feature), validity
    // (0 = no map point, 1 = good, 2 = isBad()), pKF->mvKeysUn angles, and pKF->mFeatVec as (node, feature) pairs in map
    // order.  matchedKF[iF] = KeyFrame feature whose map point vpMapPointMatches[iF] receives, -1 = NULL.  Returns nmatches.
    int SearchByBoW(const std::vector<unsigned char>& kfDescriptors, const std::vector<unsigned char>& kfValid,
                    const std::vector<float>& kfAngles, const std::vector<std::pair<unsigned int, unsigned int> >& kfFeatVec,
                    float nnRatio, bool checkOrientation, std::vector<int>& matchedKF);

    // ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
    // (reference src/ORBmatcher.cc:1472-1599, Tracking::Relocalization) on the GPU; `this` is CurrentFrame.mpORBextractorLeft.
    // Per KeyFrame map point the caller stages what the reference computes on the host with the map's own accessors:
    // search = pMP && !isBad() && !sAlreadyFound.count(pMP) && dist3D inside [GetMinDistanceInvariance(),
    // GetMaxDistanceInvariance()], predLevel = pMP->PredictScale(dist3D, &CurrentFrame), plus world position (3 floats),
    // descriptor (32 bytes) and pKF->mvKeysUn[i].angle.  currentHeld[i2] > 0 where CurrentFrame.mvpMapPoints[i2] is not NULL.
    // matched[i2] = index i of the KeyFrame point CurrentFrame.mvpMapPoints[i2] has to be set to, -1 = leave it.  Returns nmatches.
    int SearchByProjectionKF(const std::vector<unsigned char>& search, const std::vector<float>& worldPos,
                             const std::vector<int>& predLevel, const std::vector<unsigned char>& descriptors,
                             const std::vector<float>& kfAngles, const std::vector<int>& currentHeld, const cv::Mat& TcwCurrent,
                             const cv::Mat& mK, float th, int ORBdist, bool checkOrientation, std::vector<int>& matched);

    // ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (reference src/ORBmatcher.cc:405-520,
    // Tracking::MonocularInitialization) on the GPU; `this` is F2.mpORBextractorLeft (the current frame), F1 = mInitialFrame
    // comes as its undistorted keypoints and descriptors (32 bytes per keypoint).  vbPrevMatched is updated as at :515-517.
    int SearchForInitialization(const std::vector<cv::KeyPoint>& keysUn1, const std::vector<unsigned char>& descriptors1,
                                std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, float nnRatio,
                                bool checkOrientation, int windowSize = 10);

private:
    ORBextractor(const ORBextractor&);
    ORBextractor& operator=(const ORBextractor&);
    void Create();

    orbx_handle* handle_;
    int nfeatures_, nlevels_, iniThFAST_, minThFAST_;
    float scaleFactor_;
    bool downloadPyramid_;
    bool rgb_;
    int candidateDivisor_;      // FAST candidate buffers hold w*h / divisor + 1024 corners per level; halved after an overflow
};

} //namespace ORB_SLAM

#endif
