// ORB_SLAM2::ORBextractor on top of the B200 C ABI (include/orbx.h).  Replaces the reference's
// src/ORBextractor.cc in libORB_SLAM2.so (reference CMakeLists.txt:57); see INTEGRATION.md.
#include "ORBextractor.h"

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orbx.h"

namespace ORB_SLAM2
{

static int g_orbx_device = 0;

void ORBextractor::SetDevice(int device) { g_orbx_device = device; }

static void orbx_throw(orbx_handle* h, int rc, const char* where)
{
    std::string msg = std::string("ORBextractor (orbx): ") + where + ": " + orbx_strerror(rc);
    if (h && rc == ORBX_ERR_CUDA) msg += std::string(" -- ") + orbx_last_cuda_error(h);
    throw std::runtime_error(msg);
}

void ORBextractor::Create()
{
    orbx_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.nfeatures = nfeatures_;
    cfg.scale_factor = scaleFactor_;
    cfg.nlevels = nlevels_;
    cfg.ini_th_fast = iniThFAST_;
    cfg.min_th_fast = minThFAST_;
    cfg.device = g_orbx_device;
    cfg.max_batch = 1;
    cfg.download_pyramid = downloadPyramid_ ? 1 : 0;
    cfg.candidate_divisor = candidateDivisor_;
    int rc = orbx_create(&cfg, &handle_);
    if (rc != ORBX_OK) { handle_ = 0; orbx_throw(0, rc, "orbx_create"); }
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels,
                           int _iniThFAST, int _minThFAST):
    handle_(0), nfeatures_(_nfeatures), nlevels_(_nlevels), iniThFAST_(_iniThFAST), minThFAST_(_minThFAST),
    scaleFactor_(_scaleFactor), downloadPyramid_(true), rgb_(false), candidateDivisor_(8)
{
    Create();
    mvImagePyramid.resize(nlevels_);       // reference src/ORBextractor.cc:433
}

ORBextractor::~ORBextractor()
{
    if (handle_) orbx_destroy(handle_);
}

void ORBextractor::SetPyramidDownload(bool enable)
{
    if (enable == downloadPyramid_) return;
    downloadPyramid_ = enable;
    if (handle_) orbx_destroy(handle_);
    handle_ = 0;
    for (size_t l = 0; l < mvImagePyramid.size(); ++l) mvImagePyramid[l] = cv::Mat();
    Create();
}

void ORBextractor::SetColorOrder(bool rgb) { rgb_ = rgb; }

int ORBextractor::GetLevels() { return nlevels_; }

float ORBextractor::GetScaleFactor() { return orbx_get_scale_factor(handle_); }

static std::vector<float> table(orbx_handle* h, int which, int n)
{
    const float* t[4] = {0, 0, 0, 0};
    orbx_scale_tables(h, &t[0], &t[1], &t[2], &t[3]);
    return std::vector<float>(t[which], t[which] + n);
}

std::vector<float> ORBextractor::GetScaleFactors() { return table(handle_, 0, nlevels_); }
std::vector<float> ORBextractor::GetInverseScaleFactors() { return table(handle_, 1, nlevels_); }
std::vector<float> ORBextractor::GetScaleSigmaSquares() { return table(handle_, 2, nlevels_); }
std::vector<float> ORBextractor::GetInverseScaleSigmaSquares() { return table(handle_, 3, nlevels_); }

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                      cv::OutputArray _descriptors)
{
    (void)_mask;                                   // ignored by the reference too (.h:58)
    if(_image.empty())
        return;                                    // outputs untouched (reference :1046-1047)

    cv::Mat image = _image.getMat();
    const int type = image.type();                 // CV_8UC1 = 0, CV_8UC3 = 16, CV_8UC4 = 24
    assert(type == CV_8UC1 || type == 16 || type == 24);

    orbx_result r;
    int rc;
    for (;;)
    {
        if (type == CV_8UC1)
            rc = orbx_extract(handle_, image.data, image.cols, image.rows, (size_t)image.step, &r);
        else
        {
            // colour frame: cvtColor on the device (reference src/Tracking.cc:172-255 does it on the CPU first)
            const unsigned char* p = image.data;
            const size_t step = (size_t)image.step;
            const int fmt = type == 16 ? (rgb_ ? ORBX_RGB8 : ORBX_BGR8) : (rgb_ ? ORBX_RGBA8 : ORBX_BGRA8);
            rc = orbx_extract_batch_color(handle_, 1, &p, image.cols, image.rows, &step, fmt, &r);
        }
        if (rc != ORBX_ERR_CANDIDATE_OVERFLOW || candidateDivisor_ == 1) break;
        // More FAST corners than the candidate buffers were sized for (w*h / divisor + 1024 per level): the reference has
        // no such failure mode, so the handle is rebuilt with larger buffers and the frame is run again.  Divisor 1 holds
        // every pixel of a level, which cannot overflow.
        candidateDivisor_ = candidateDivisor_ > 2 ? candidateDivisor_ / 2 : 1;
        orbx_destroy(handle_);
        handle_ = 0;
        for (size_t l = 0; l < mvImagePyramid.size(); ++l) mvImagePyramid[l] = cv::Mat();
        Create();
    }
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_extract");

    if( r.n == 0 )
        _descriptors.release();                    // (:1064-1065)
    else
    {
        _descriptors.create(r.n, 32, CV_8U);       // (:1068)
        cv::Mat descriptors = _descriptors.getMat();
        for (int i = 0; i < r.n; ++i)
            std::memcpy(descriptors.ptr(i), r.desc + 32 * (size_t)i, 32);
    }

    _keypoints.clear();
    _keypoints.reserve(r.n);
    for (int i = 0; i < r.n; ++i)
    {
        const orbx_keypoint& k = r.kps[i];
        cv::KeyPoint kp;
        kp.pt.x = k.x; kp.pt.y = k.y;
        kp.size = k.size; kp.angle = k.angle; kp.response = k.response;
        kp.octave = k.octave; kp.class_id = k.class_id;
        _keypoints.push_back(kp);
    }

    if (downloadPyramid_)
    {
        for (int level = 0; level < nlevels_; ++level)
        {
            const unsigned char* px = 0; int w = 0, h = 0; size_t step = 0;
            rc = orbx_pyramid_level(handle_, 0, level, &px, &w, &h, &step);
            if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_pyramid_level");
            // non-owning header at offset (19,19) of the padded plane, like temp(Rect(...)) (:1115)
            mvImagePyramid[level] = cv::Mat(h, w, CV_8UC1, const_cast<unsigned char*>(px), step);
        }
    }
}

void ORBextractor::ComputeStereoMatches(ORBextractor& left, ORBextractor& right, float mbf, float mb,
                                        std::vector<float>& mvuRight, std::vector<float>& mvDepth)
{
    orbx_stereo_result r;
    int rc = orbx_stereo_match(left.handle_, right.handle_, 1, 0, 0, mbf, mb, &r);
    if (rc != ORBX_OK) orbx_throw(left.handle_, rc, "orbx_stereo_match");
    mvuRight.assign(r.u_right, r.u_right + r.n);       // (:468-469) N floats each, -1 where unmatched
    mvDepth.assign(r.depth, r.depth + r.n);
}

void ORBextractor::UndistortAndAssignToGrid(const cv::Mat& mK, const cv::Mat& mDistCoef, const std::vector<cv::KeyPoint>& mvKeys,
                                            std::vector<cv::KeyPoint>& mvKeysUn, std::vector<std::size_t> (*mGrid)[48],
                                            float& mnMinX, float& mnMaxX, float& mnMinY, float& mnMaxY)
{
    const float K4[4] = {mK.at<float>(0,0), mK.at<float>(1,1), mK.at<float>(0,2), mK.at<float>(1,2)};
    float dist[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const int nd = mDistCoef.rows * mDistCoef.cols >= 5 ? 5 : 4;
    for (int i = 0; i < nd; ++i)
        dist[i] = mDistCoef.cols == 1 ? mDistCoef.at<float>(i,0) : mDistCoef.at<float>(0,i);
    orbx_grid_result r;
    int rc = orbx_undistort_grid(handle_, 1, 0, K4, dist, nd, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_undistort_grid");
    if ((size_t)r.n != mvKeys.size()) throw std::runtime_error("ORBextractor (orbx): mvKeys is not the result of the last operator()");
    mvKeysUn = mvKeys;                                 // (:427-432) every field but pt is copied
    for (int i = 0; i < r.n; ++i) { mvKeysUn[i].pt.x = r.xy_un[2*i]; mvKeysUn[i].pt.y = r.xy_un[2*i+1]; }
    for (int gx = 0; gx < 64; ++gx)
        for (int gy = 0; gy < 48; ++gy)
        {
            const int c = gx * 48 + gy;
            mGrid[gx][gy].assign(r.cell_items + r.cell_start[c], r.cell_items + r.cell_start[c+1]);
        }
    mnMinX = r.bounds[0]; mnMaxX = r.bounds[1]; mnMinY = r.bounds[2]; mnMaxY = r.bounds[3];
}

int ORBextractor::SearchByProjection(const std::vector<float>& lastWorldPos, const std::vector<unsigned char>& lastDescriptors,
                                     const std::vector<int>& lastObservations, const std::vector<cv::KeyPoint>& lastKeysUn,
                                     const cv::Mat& TcwCurrent, const cv::Mat& TcwLast, const cv::Mat& mK, float mbf, float mb, float th,
                                     bool bMono, bool checkOrientation, bool useStereo, std::vector<int>& matchedLast)
{
    const size_t n = lastKeysUn.size();
    if (lastWorldPos.size() != 3 * n || lastDescriptors.size() != 32 * n || lastObservations.size() != n)
        throw std::runtime_error("ORBextractor (orbx): SearchByProjection needs 3 floats, 32 bytes and 1 int per LastFrame keypoint");
    std::vector<int> octave(n);
    std::vector<float> angle(n);
    for (size_t i = 0; i < n; ++i) { octave[i] = lastKeysUn[i].octave; angle[i] = lastKeysUn[i].angle; }
    orbx_projection_query q;
    q.cur_frame = 0;
    q.n_last = (int)n;
    q.world_pos = lastWorldPos.data();
    q.mp_desc = lastDescriptors.data();
    q.mp_obs = lastObservations.data();
    q.outlier = 0;
    q.octave = octave.data();
    q.angle = angle.data();
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { q.Tcw_cur[4*r+c] = TcwCurrent.at<float>(r,c); q.Tcw_last[4*r+c] = TcwLast.at<float>(r,c); }
    const float K4[4] = {mK.at<float>(0,0), mK.at<float>(1,1), mK.at<float>(0,2), mK.at<float>(1,2)};
    orbx_projection_result r;
    int rc = orbx_search_by_projection(handle_, 1, &q, K4, mbf, mb, th, bMono ? 1 : 0, checkOrientation ? 1 : 0, useStereo ? 1 : 0, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_search_by_projection");
    matchedLast.assign(r.match, r.match + r.n);
    return r.nmatches;
}

int ORBextractor::IsInFrustum(const std::vector<unsigned char>& consider, const std::vector<float>& worldPos,
                              const std::vector<float>& normals, const std::vector<float>& minDistance,
                              const std::vector<float>& maxDistance, const cv::Mat& Tcw, const cv::Mat& mK, float mbf,
                              const float bounds[4], float logScaleFactor, float viewingCosLimit, std::vector<unsigned char>& inView,
                              std::vector<float>& projXYXR, std::vector<int>& scaleLevel, std::vector<float>& viewCos)
{
    const size_t n = minDistance.size();
    if ((!consider.empty() && consider.size() != n) || worldPos.size() != 3 * n || normals.size() != 3 * n || maxDistance.size() != n)
        throw std::runtime_error("ORBextractor (orbx): IsInFrustum needs one entry per map point in every array");
    orbx_frustum_query q;
    q.n_points = (int)n;
    q.consider = consider.empty() ? 0 : consider.data();
    q.world_pos = worldPos.data();
    q.normal = normals.data();
    q.min_dist = minDistance.data();
    q.max_dist = maxDistance.data();
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) q.Tcw[4 * r + c] = Tcw.at<float>(r, c);
    const float K4[4] = {mK.at<float>(0, 0), mK.at<float>(1, 1), mK.at<float>(0, 2), mK.at<float>(1, 2)};
    orbx_frustum_result r;
    int rc = orbx_is_in_frustum(handle_, 1, &q, K4, mbf, bounds, logScaleFactor, viewingCosLimit, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_is_in_frustum");
    inView.assign(r.in_view, r.in_view + n);
    projXYXR.assign(r.proj_xy_xr, r.proj_xy_xr + 3 * n);
    scaleLevel.assign(r.scale_level, r.scale_level + n);
    viewCos.assign(r.view_cos, r.view_cos + n);
    return r.n_in_view;
}

int ORBextractor::SearchLocalPoints(const std::vector<unsigned char>& inView, const std::vector<float>& projXYXR,
                                    const std::vector<int>& scaleLevel, const std::vector<float>& viewCos,
                                    const std::vector<unsigned char>& descriptors, const std::vector<int>& observations,
                                    const std::vector<int>& currentObservations, float th, float nnRatio, bool useStereo,
                                    std::vector<int>& matched)
{
    const size_t n = observations.size();
    if (inView.size() != n || projXYXR.size() != 3 * n || scaleLevel.size() != n || viewCos.size() != n || descriptors.size() != 32 * n)
        throw std::runtime_error("ORBextractor (orbx): SearchLocalPoints needs one entry per map point in every array");
    orbx_local_points_query q;
    q.cur_frame = 0;
    q.n_points = (int)n;
    q.in_view = inView.data();
    q.proj_xy_xr = projXYXR.data();
    q.scale_level = scaleLevel.data();
    q.view_cos = viewCos.data();
    q.mp_desc = descriptors.data();
    q.mp_obs = observations.data();
    q.cur_obs = currentObservations.empty() ? 0 : currentObservations.data();
    orbx_projection_result r;
    int rc = orbx_search_local_points(handle_, 1, &q, th, nnRatio, useStereo ? 1 : 0, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_search_local_points");
    if (!currentObservations.empty() && (size_t)r.n != currentObservations.size())
        throw std::runtime_error("ORBextractor (orbx): currentObservations is not sized like the last operator()'s keypoints");
    matched.assign(r.match, r.match + r.n);
    return r.nmatches;
}

void ORBextractor::ComputeBoW(const orbx_vocabulary* voc, std::vector<std::pair<unsigned int, double> >& bow,
                              std::vector<std::pair<unsigned int, unsigned int> >& featVec, int levelsup)
{
    orbx_bow_result r;
    int rc = orbx_compute_bow(handle_, voc, 1, 0, levelsup, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_compute_bow");
    bow.resize(r.n_words);
    for (int i = 0; i < r.n_words; ++i) bow[i] = std::make_pair(r.word_ids[i], r.word_values[i]);
    featVec.resize(r.n_features);
    for (int i = 0; i < r.n_features; ++i) featVec[i] = std::make_pair(r.fv_nodes[i], r.fv_features[i]);
}

int ORBextractor::SearchByBoW(const std::vector<unsigned char>& kfDescriptors, const std::vector<unsigned char>& kfValid,
                              const std::vector<float>& kfAngles, const std::vector<std::pair<unsigned int, unsigned int> >& kfFeatVec,
                              float nnRatio, bool checkOrientation, std::vector<int>& matchedKF)
{
    const size_t n = kfValid.size();
    if (kfDescriptors.size() != 32 * n || kfAngles.size() != n)
        throw std::runtime_error("ORBextractor (orbx): SearchByBoW needs 32 bytes, one flag and one angle per KeyFrame feature");
    std::vector<uint32_t> nodes(kfFeatVec.size()), feats(kfFeatVec.size());
    for (size_t i = 0; i < kfFeatVec.size(); ++i) { nodes[i] = kfFeatVec[i].first; feats[i] = kfFeatVec[i].second; }
    orbx_bow_match_query q;
    q.cur_frame = 0;
    q.n_kf = (int)n;
    q.kf_desc = kfDescriptors.data();
    q.kf_valid = kfValid.data();
    q.kf_angle = kfAngles.data();
    q.n_kf_fv = (int)nodes.size();
    q.kf_fv_nodes = nodes.data();
    q.kf_fv_features = feats.data();
    orbx_projection_result r;
    int rc = orbx_search_by_bow(handle_, 1, &q, nnRatio, checkOrientation ? 1 : 0, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_search_by_bow");
    matchedKF.assign(r.match, r.match + r.n);
    return r.nmatches;
}

int ORBextractor::SearchByProjectionKF(const std::vector<unsigned char>& search, const std::vector<float>& worldPos,
                                       const std::vector<int>& predLevel, const std::vector<unsigned char>& descriptors,
                                       const std::vector<float>& kfAngles, const std::vector<int>& currentHeld, const cv::Mat& TcwCurrent,
                                       const cv::Mat& mK, float th, int ORBdist, bool checkOrientation, std::vector<int>& matched)
{
    const size_t n = search.size();
    if (worldPos.size() != 3 * n || predLevel.size() != n || descriptors.size() != 32 * n || kfAngles.size() != n)
        throw std::runtime_error("ORBextractor (orbx): SearchByProjectionKF needs one entry per KeyFrame map point in every array");
    orbx_keyframe_projection_query q;
    q.cur_frame = 0;
    q.n_points = (int)n;
    q.search = search.data();
    q.world_pos = worldPos.data();
    q.pred_level = predLevel.data();
    q.mp_desc = descriptors.data();
    q.kf_angle = kfAngles.data();
    q.cur_held = currentHeld.empty() ? 0 : currentHeld.data();
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) q.Tcw_cur[4 * r + c] = TcwCurrent.at<float>(r, c);
    const float K4[4] = {mK.at<float>(0, 0), mK.at<float>(1, 1), mK.at<float>(0, 2), mK.at<float>(1, 2)};
    orbx_projection_result r;
    int rc = orbx_search_by_projection_kf(handle_, 1, &q, K4, th, ORBdist, checkOrientation ? 1 : 0, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_search_by_projection_kf");
    if (!currentHeld.empty() && (size_t)r.n != currentHeld.size())
        throw std::runtime_error("ORBextractor (orbx): currentHeld is not sized like the last operator()'s keypoints");
    matched.assign(r.match, r.match + r.n);
    return r.nmatches;
}

int ORBextractor::SearchForInitialization(const std::vector<cv::KeyPoint>& keysUn1, const std::vector<unsigned char>& descriptors1,
                                          std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, float nnRatio,
                                          bool checkOrientation, int windowSize)
{
    const size_t n = keysUn1.size();
    if (descriptors1.size() != 32 * n || vbPrevMatched.size() != n)
        throw std::runtime_error("ORBextractor (orbx): SearchForInitialization needs one descriptor and one vbPrevMatched entry per F1 keypoint");
    std::vector<int> octave(n);
    std::vector<float> angle(n), prev(2 * n);
    for (size_t i = 0; i < n; ++i) {
        octave[i] = keysUn1[i].octave;
        angle[i] = keysUn1[i].angle;
        prev[2 * i] = vbPrevMatched[i].x;
        prev[2 * i + 1] = vbPrevMatched[i].y;
    }
    orbx_initialization_query q;
    q.cur_frame = 0;
    q.n1 = (int)n;
    q.octave1 = octave.data();
    q.angle1 = angle.data();
    q.desc1 = descriptors1.data();
    q.prev_matched = prev.data();
    orbx_initialization_result r;
    int rc = orbx_search_for_initialization(handle_, 1, &q, nnRatio, checkOrientation ? 1 : 0, windowSize, &r);
    if (rc != ORBX_OK) orbx_throw(handle_, rc, "orbx_search_for_initialization");
    vnMatches12.assign(r.matches12, r.matches12 + r.n1);
    for (size_t i = 0; i < n; ++i) {
        vbPrevMatched[i].x = r.prev_matched[2 * i];
        vbPrevMatched[i].y = r.prev_matched[2 * i + 1];
    }
    return r.nmatches;
}

} //namespace ORB_SLAM
