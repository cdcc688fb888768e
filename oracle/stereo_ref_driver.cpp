// C entry point around the reference's own Frame::ComputeStereoMatches lines (see shim_stereo/stereo_shim.h and
// build_stereo_ref.sh).  TEST INFRASTRUCTURE: the checker of orbx_stereo_match, never on the product path.
#include "stereo_shim.h"

extern "C" int stereoref_match(int nL, const float* kpL /* nL x 7, cv::KeyPoint field order */, const unsigned char* descL,
                               int nR, const float* kpR, const unsigned char* descR, int nlevels,
                               const unsigned char* const* pyrL /* level (0,0) pixel */, const unsigned char* const* pyrR,
                               const int* lw, const int* lh, const size_t* stepL, const size_t* stepR, const float* sf,
                               const float* inv_sf, float mbf, float mb, float* uRight, float* depth) {
    using namespace ORB_SLAM2;
    Frame F;
    ORBextractor EL, ER;
    F.N = nL;
    F.mbf = mbf;
    F.mb = mb;
    F.mpORBextractorLeft = &EL;
    F.mpORBextractorRight = &ER;
    F.mvScaleFactors.assign(sf, sf + nlevels);
    F.mvInvScaleFactors.assign(inv_sf, inv_sf + nlevels);
    for (int l = 0; l < nlevels; ++l) {
        EL.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, (void*)pyrL[l], stepL[l]));
        ER.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, (void*)pyrR[l], stepR[l]));
    }
    auto fill = [](std::vector<cv::KeyPoint>& v, const float* p, int n) {
        v.resize(n);
        for (int i = 0; i < n; ++i) {
            const float* q = p + 7 * i;
            v[i].pt.x = q[0]; v[i].pt.y = q[1]; v[i].size = q[2]; v[i].angle = q[3]; v[i].response = q[4];
            memcpy(&v[i].octave, q + 5, 4); memcpy(&v[i].class_id, q + 6, 4);
        }
    };
    fill(F.mvKeys, kpL, nL);
    fill(F.mvKeysRight, kpR, nR);
    F.mDescriptors = cv::Mat(nL, 32, CV_8U, (void*)descL, 32);
    F.mDescriptorsRight = cv::Mat(nR, 32, CV_8U, (void*)descR, 32);
    F.ComputeStereoMatches();
    for (int i = 0; i < nL; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
    return 0;
}

// Frame::ComputeImageBounds + the grid scale set-up of the Frame constructors (src/Frame.cc:148-156) +
// UndistortKeyPoints + AssignFeaturesToGrid, the reference's own lines.  K = {fx, fy, cx, cy}; D = nd distortion floats.
// Outputs: undistorted (x, y) per keypoint, the 64 x 48 grid as CSR (cell gx * 48 + gy = mGrid[gx][gy], indices in
// push_back order) and the four image bounds.
namespace ORB_SLAM2 {
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
}
extern "C" int frameref_undistort_grid(int n, const float* kp, const float* K, const float* D, int nd, int width, int height,
                                       float* xy_un, int* cell_start, int* cell_items, float* bounds) {
    using namespace ORB_SLAM2;
    Frame F;
    F.N = n;
    F.mvKeys.resize(n);
    for (int i = 0; i < n; ++i) {
        const float* q = kp + 7 * i;
        F.mvKeys[i].pt.x = q[0]; F.mvKeys[i].pt.y = q[1]; F.mvKeys[i].size = q[2]; F.mvKeys[i].angle = q[3];
        F.mvKeys[i].response = q[4];
        memcpy(&F.mvKeys[i].octave, q + 5, 4); memcpy(&F.mvKeys[i].class_id, q + 6, 4);
    }
    F.mK = cv::Mat(3, 3, CV_32F);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F.mK.at<float>(r, c) = r == c ? 1.f : 0.f;
    F.mK.at<float>(0, 0) = K[0]; F.mK.at<float>(1, 1) = K[1]; F.mK.at<float>(0, 2) = K[2]; F.mK.at<float>(1, 2) = K[3];
    F.mDistCoef = cv::Mat(nd, 1, CV_32F);
    for (int i = 0; i < nd; ++i) F.mDistCoef.at<float>(i, 0) = D[i];
    cv::Mat im(height, width, CV_8U);
    F.ComputeImageBounds(im);
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);    // (:155)
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);   // (:156)
    F.UndistortKeyPoints();
    F.AssignFeaturesToGrid();
    for (int i = 0; i < n; ++i) { xy_un[2 * i] = F.mvKeysUn[i].pt.x; xy_un[2 * i + 1] = F.mvKeysUn[i].pt.y; }
    int pos = 0;
    for (int gx = 0; gx < FRAME_GRID_COLS; ++gx)
        for (int gy = 0; gy < FRAME_GRID_ROWS; ++gy) {
            cell_start[gx * FRAME_GRID_ROWS + gy] = pos;
            for (size_t k = 0; k < F.mGrid[gx][gy].size(); ++k) cell_items[pos++] = (int)F.mGrid[gx][gy][k];
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = pos;
    bounds[0] = Frame::mnMinX; bounds[1] = Frame::mnMaxX; bounds[2] = Frame::mnMinY; bounds[3] = Frame::mnMaxY;
    return pos;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (src/ORBmatcher.cc:1328-1470,
// called by Tracking::TrackWithMotionModel, src/Tracking.cc:885/:891, right after the current frame's map points are
// filled with NULL), the reference's own lines, with Frame::GetFeaturesInArea (src/Frame.cc:327-380) and
// ComputeThreeMaxima.  LastFrame side: world positions / representative descriptors / Observations() of its map points
// (mp_obs < 0: no map point), outlier flags, octaves and angles of its keypoints, its pose.  CurrentFrame side:
// undistorted keypoint records, descriptors, mvuRight, the grid as CSR, bounds, K4 = {fx, fy, cx, cy}, pose.
// cur_match[i2] = index of the LastFrame keypoint whose map point CurrentFrame.mvpMapPoints[i2] ends up holding, else -1.
namespace ORB_SLAM2 {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
}
extern "C" int matchref_search_by_projection(int nL, const float* world, const unsigned char* mp_desc, const int* mp_obs,
                                             const unsigned char* outlier, const int* last_octave, const float* last_angle,
                                             const float* Tcw_cur, const float* Tcw_last, int nC, const float* kp_un,
                                             const unsigned char* desc, const float* u_right, const int* cell_start,
                                             const int* cell_items, const float* bounds, const float* K4, float mbf, float mb,
                                             const float* sf, int nlevels, float th, int mono, int check_ori, int* cur_match) {
    using namespace ORB_SLAM2;
    Frame C, L;
    std::vector<MapPoint> mps(nL);
    L.N = nL;
    L.mvKeys.resize(nL);
    L.mvKeysUn.resize(nL);
    L.mvpMapPoints.assign(nL, (MapPoint*)0);
    L.mvbOutlier.assign(nL, false);
    for (int i = 0; i < nL; ++i) {
        L.mvKeys[i].octave = last_octave[i];
        L.mvKeysUn[i].octave = last_octave[i];
        L.mvKeysUn[i].angle = last_angle[i];
        L.mvbOutlier[i] = outlier[i] != 0;
        if (mp_obs[i] >= 0) {
            mps[i].mWorldPos = cv::Mat(3, 1, CV_32F);
            for (int k = 0; k < 3; ++k) mps[i].mWorldPos.at<float>(k, 0) = world[3 * i + k];
            mps[i].mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i), 32);
            mps[i].nObs = mp_obs[i];
            L.mvpMapPoints[i] = &mps[i];
        }
    }
    auto pose = [](const float* T) {
        cv::Mat m(4, 4, CV_32F);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m.at<float>(r, c) = T[4 * r + c];
        return m;
    };
    L.mTcw = pose(Tcw_last);
    C.mTcw = pose(Tcw_cur);
    C.N = nC;
    C.mvKeysUn.resize(nC);
    for (int i = 0; i < nC; ++i) {
        const float* q = kp_un + 7 * i;
        C.mvKeysUn[i].pt.x = q[0]; C.mvKeysUn[i].pt.y = q[1]; C.mvKeysUn[i].size = q[2]; C.mvKeysUn[i].angle = q[3];
        C.mvKeysUn[i].response = q[4];
        memcpy(&C.mvKeysUn[i].octave, q + 5, 4); memcpy(&C.mvKeysUn[i].class_id, q + 6, 4);
    }
    C.mvpMapPoints.assign(nC, (MapPoint*)0);
    C.mvuRight.assign(nC, -1.f);
    if (u_right) C.mvuRight.assign(u_right, u_right + nC);
    C.mDescriptors = cv::Mat(nC, 32, CV_8U, (void*)desc, 32);
    C.mvScaleFactors.assign(sf, sf + nlevels);
    C.mbf = mbf;
    C.mb = mb;
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    for (int gx = 0; gx < FRAME_GRID_COLS; ++gx)
        for (int gy = 0; gy < FRAME_GRID_ROWS; ++gy) {
            const int c = gx * FRAME_GRID_ROWS + gy;
            C.mGrid[gx][gy].assign(cell_items + cell_start[c], cell_items + cell_start[c + 1]);
        }
    ORBmatcher matcher(0.9, check_ori != 0);                       // src/Tracking.cc:868
    const int nmatches = matcher.SearchByProjection(C, L, th, mono != 0);
    for (int i = 0; i < nC; ++i) cur_match[i] = C.mvpMapPoints[i] ? (int)(C.mvpMapPoints[i] - &mps[0]) : -1;
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (src/ORBmatcher.cc:45-129, called by
// Tracking::SearchLocalPoints, src/Tracking.cc:1184-1194, with ORBmatcher(0.8)), the reference's own lines.  Map point
// side: the tracking fields Frame::isInFrustum left on each point (mbTrackInView && !isBad as `in_view`, mTrackProjX/Y/XR,
// mnTrackScaleLevel, mTrackViewCos), descriptor, Observations().  Frame side as above plus the map points the frame
// already holds: cur_obs[i2] < 0 = NULL, else that point's Observations().
// new_match[i2] = index into vpMapPoints of the point F.mvpMapPoints[i2] was SET to by this call, else -1.
extern "C" int matchref_search_local_points(int nP, const unsigned char* in_view, const float* proj_x, const float* proj_y,
                                            const float* proj_xr, const int* scale_level, const float* view_cos,
                                            const unsigned char* mp_desc, const int* mp_obs, int nC, const float* kp_un,
                                            const unsigned char* desc, const float* u_right, const int* cur_obs,
                                            const int* cell_start, const int* cell_items, const float* bounds, const float* sf,
                                            int nlevels, float th, float nnratio, int* new_match) {
    using namespace ORB_SLAM2;
    Frame C;
    std::vector<MapPoint> mps(nP), held(nC);
    std::vector<MapPoint*> vp(nP);
    for (int i = 0; i < nP; ++i) {
        mps[i].mbTrackInView = in_view[i] != 0;
        mps[i].mbBad = false;
        mps[i].mTrackProjX = proj_x[i]; mps[i].mTrackProjY = proj_y[i]; mps[i].mTrackProjXR = proj_xr[i];
        mps[i].mnTrackScaleLevel = scale_level[i];
        mps[i].mTrackViewCos = view_cos[i];
        mps[i].mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i), 32);
        mps[i].nObs = mp_obs[i];
        vp[i] = &mps[i];
    }
    C.N = nC;
    C.mvKeysUn.resize(nC);
    C.mvpMapPoints.assign(nC, (MapPoint*)0);
    for (int i = 0; i < nC; ++i) {
        const float* q = kp_un + 7 * i;
        C.mvKeysUn[i].pt.x = q[0]; C.mvKeysUn[i].pt.y = q[1]; C.mvKeysUn[i].angle = q[3];
        memcpy(&C.mvKeysUn[i].octave, q + 5, 4);
        if (cur_obs[i] >= 0) { held[i].nObs = cur_obs[i]; C.mvpMapPoints[i] = &held[i]; }
    }
    C.mvuRight.assign(nC, -1.f);
    if (u_right) C.mvuRight.assign(u_right, u_right + nC);
    C.mDescriptors = cv::Mat(nC, 32, CV_8U, (void*)desc, 32);
    C.mvScaleFactors.assign(sf, sf + nlevels);
    Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    for (int gx = 0; gx < FRAME_GRID_COLS; ++gx)
        for (int gy = 0; gy < FRAME_GRID_ROWS; ++gy) {
            const int c = gx * FRAME_GRID_ROWS + gy;
            C.mGrid[gx][gy].assign(cell_items + cell_start[c], cell_items + cell_start[c + 1]);
        }
    ORBmatcher matcher(nnratio, true);
    const int nmatches = matcher.SearchByProjection(C, vp, th);
    for (int i = 0; i < nC; ++i) {
        MapPoint* p = C.mvpMapPoints[i];
        new_match[i] = (p && p >= &mps[0] && p < &mps[0] + nP) ? (int)(p - &mps[0]) : -1;
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:159-288, called by
// Tracking::TrackReferenceKeyFrame src/Tracking.cc:771-776 with ORBmatcher(0.7, true) and by Relocalization :1376 with 0.75),
// the reference's own lines.  Both FeatureVectors come as (node, feature) pairs in map / push_back order.
// match[iF] = index of the KeyFrame feature whose map point vpMapPointMatches[iF] holds, else -1.
extern "C" int matchref_search_by_bow(int nKF, const unsigned char* kf_desc, const unsigned char* kf_valid, const float* kf_angle,
                                      int nKFfv, const unsigned int* kf_fv_nodes, const unsigned int* kf_fv_features, int nF,
                                      const unsigned char* f_desc, const float* f_angle, int nFfv, const unsigned int* f_fv_nodes,
                                      const unsigned int* f_fv_features, float nnratio, int check_ori, int* match) {
    using namespace ORB_SLAM2;
    KeyFrame KF;
    Frame F;
    std::vector<MapPoint> mps(nKF > 0 ? nKF : 1);
    KF.mvpMapPoints.assign(nKF, (MapPoint*)0);
    KF.mvKeysUn.resize(nKF);
    for (int i = 0; i < nKF; ++i) {
        mps[i].mbBad = kf_valid[i] == 2;                                 // 0: no map point, 1: good, 2: bad
        if (kf_valid[i]) KF.mvpMapPoints[i] = &mps[i];
        KF.mvKeysUn[i].angle = kf_angle[i];
    }
    KF.mDescriptors = cv::Mat(nKF, 32, CV_8U, (void*)kf_desc, 32);
    for (int i = 0; i < nKFfv; ++i) KF.mFeatVec[kf_fv_nodes[i]].push_back(kf_fv_features[i]);
    F.N = nF;
    F.mvKeys.resize(nF);
    for (int i = 0; i < nF; ++i) F.mvKeys[i].angle = f_angle[i];
    F.mDescriptors = cv::Mat(nF, 32, CV_8U, (void*)f_desc, 32);
    for (int i = 0; i < nFfv; ++i) F.mFeatVec[f_fv_nodes[i]].push_back(f_fv_features[i]);
    ORBmatcher matcher(nnratio, check_ori != 0);
    std::vector<MapPoint*> out;
    const int nmatches = matcher.SearchByBoW(&KF, F, out);
    for (int i = 0; i < nF; ++i) match[i] = out[i] ? (int)(out[i] - &mps[0]) : -1;
    return nmatches;
}

namespace {
// the current frame's side shared by the two entries below: keypoint records, descriptors, grid, bounds
void fill_frame(ORB_SLAM2::Frame& C, int nC, const float* kp_un, const unsigned char* desc, const int* cell_start,
                const int* cell_items, const float* bounds, const float* sf, int nlevels) {
    using namespace ORB_SLAM2;
    C.N = nC;
    C.mvKeysUn.resize(nC);
    for (int i = 0; i < nC; ++i) {
        const float* q = kp_un + 7 * i;
        C.mvKeysUn[i].pt.x = q[0]; C.mvKeysUn[i].pt.y = q[1]; C.mvKeysUn[i].size = q[2]; C.mvKeysUn[i].angle = q[3];
        C.mvKeysUn[i].response = q[4];
        memcpy(&C.mvKeysUn[i].octave, q + 5, 4); memcpy(&C.mvKeysUn[i].class_id, q + 6, 4);
    }
    C.mvpMapPoints.assign(nC, (MapPoint*)0);
    C.mDescriptors = cv::Mat(nC, 32, CV_8U, (void*)desc, 32);
    C.mvScaleFactors.assign(sf, sf + nlevels);
    C.mnScaleLevels = nlevels;
    Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    for (int gx = 0; gx < FRAME_GRID_COLS; ++gx)
        for (int gy = 0; gy < FRAME_GRID_ROWS; ++gy) {
            const int c = gx * FRAME_GRID_ROWS + gy;
            C.mGrid[gx][gy].assign(cell_items + cell_start[c], cell_items + cell_start[c + 1]);
        }
}
}  // namespace

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:1472-1599, called by Tracking::Relocalization, src/Tracking.cc:1452 / :1466, with ORBmatcher(0.9, true)),
// the reference's own lines, with MapPoint::GetMin/MaxDistanceInvariance and PredictScale (src/MapPoint.cc:373-383, :402-417).
// KeyFrame side: valid[i] = 0: no map point, 1: good, 2: isBad(), 3: in sAlreadyFound; world position, descriptor,
// mfMinDistance / mfMaxDistance of the point, angle of the KeyFrame's keypoint.  Frame side as above; cur_held[i2] != 0:
// CurrentFrame.mvpMapPoints[i2] is not NULL at entry.  new_match[i2] = index i of the KeyFrame point the call SET
// CurrentFrame.mvpMapPoints[i2] to (and did not cull), else -1.  pred_level[i] / in_range[i] (may be NULL) receive what the
// reference's own PredictScale and distance test say for point i with the frame's pose (what a caller of the C ABI stages).
extern "C" int matchref_search_by_projection_kf(int nP, const unsigned char* valid, const float* world, const unsigned char* mp_desc,
                                                const float* min_dist, const float* max_dist, const float* kf_angle,
                                                const float* Tcw_cur, int nC, const float* kp_un, const unsigned char* desc,
                                                const unsigned char* cur_held, const int* cell_start, const int* cell_items,
                                                const float* bounds, const float* K4, const float* sf, int nlevels,
                                                float log_scale_factor, float th, int orb_dist, int check_ori, int* new_match,
                                                int* pred_level, unsigned char* in_range) {
    using namespace ORB_SLAM2;
    Frame C;
    KeyFrame KF;
    std::vector<MapPoint> mps(nP > 0 ? nP : 1), held(nC > 0 ? nC : 1);
    std::set<MapPoint*> found;
    KF.mvpMapPoints.assign(nP, (MapPoint*)0);
    KF.mvKeysUn.resize(nP);
    for (int i = 0; i < nP; ++i) {
        KF.mvKeysUn[i].angle = kf_angle[i];
        if (!valid[i]) continue;
        mps[i].mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) mps[i].mWorldPos.at<float>(k, 0) = world[3 * i + k];
        mps[i].mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i), 32);
        mps[i].mbBad = valid[i] == 2;
        mps[i].mfMinDistance = min_dist[i];
        mps[i].mfMaxDistance = max_dist[i];
        KF.mvpMapPoints[i] = &mps[i];
        if (valid[i] == 3) found.insert(&mps[i]);
    }
    fill_frame(C, nC, kp_un, desc, cell_start, cell_items, bounds, sf, nlevels);
    C.mfLogScaleFactor = log_scale_factor;
    for (int i = 0; i < nC; ++i) if (cur_held[i]) C.mvpMapPoints[i] = &held[i];
    C.mTcw = cv::Mat(4, 4, CV_32F);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) C.mTcw.at<float>(r, c) = Tcw_cur[4 * r + c];
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    if (pred_level && in_range) {
        // what the caller of orbx_search_by_projection_kf stages per point (ORBmatcher.cc:1476-1478, :1510-1522), by the same lines' arithmetic
        const cv::Mat Rcw = C.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = C.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -Rcw.t() * tcw;
        for (int i = 0; i < nP; ++i) {
            pred_level[i] = 0;
            in_range[i] = 0;
            if (!valid[i]) continue;
            cv::Mat PO = mps[i].mWorldPos - Ow;
            float dist3D = cv::norm(PO);
            if (dist3D < mps[i].GetMinDistanceInvariance() || dist3D > mps[i].GetMaxDistanceInvariance()) continue;
            in_range[i] = 1;
            pred_level[i] = mps[i].PredictScale(dist3D, &C);
        }
    }
    ORBmatcher matcher(0.9, check_ori != 0);                       // src/Tracking.cc:1437
    const int nmatches = matcher.SearchByProjection(C, &KF, found, th, orb_dist);
    for (int i = 0; i < nC; ++i) {
        MapPoint* p = C.mvpMapPoints[i];
        new_match[i] = (p && p >= &mps[0] && p < &mps[0] + nP) ? (int)(p - &mps[0]) : -1;
    }
    return nmatches;
}

// ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:405-520,
// called by Tracking::MonocularInitialization, src/Tracking.cc:600, with ORBmatcher(0.9, true) and windowSize 100), the reference's
// own lines.  F1 side: octave / angle of its undistorted keypoints, descriptors, vbPrevMatched (n1 x 2, updated in place).
// F2 side: keypoint records, descriptors, grid, bounds.  matches12[i1] = index in F2 or -1.
extern "C" int matchref_search_for_initialization(int n1, const float* kp1_un, const unsigned char* desc1, float* prev_matched, int n2,
                                                  const float* kp2_un, const unsigned char* desc2, const int* cell_start,
                                                  const int* cell_items, const float* bounds, const float* sf, int nlevels,
                                                  float nnratio, int check_ori, int window, int* matches12) {
    using namespace ORB_SLAM2;
    Frame F1, F2;
    F1.N = n1;
    F1.mvKeysUn.resize(n1);
    for (int i = 0; i < n1; ++i) {
        const float* q = kp1_un + 7 * i;
        F1.mvKeysUn[i].pt.x = q[0]; F1.mvKeysUn[i].pt.y = q[1]; F1.mvKeysUn[i].size = q[2]; F1.mvKeysUn[i].angle = q[3];
        F1.mvKeysUn[i].response = q[4];
        memcpy(&F1.mvKeysUn[i].octave, q + 5, 4); memcpy(&F1.mvKeysUn[i].class_id, q + 6, 4);
    }
    F1.mDescriptors = cv::Mat(n1, 32, CV_8U, (void*)desc1, 32);
    fill_frame(F2, n2, kp2_un, desc2, cell_start, cell_items, bounds, sf, nlevels);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; ++i) { prev[i].x = prev_matched[2 * i]; prev[i].y = prev_matched[2 * i + 1]; }
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nmatches = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return nmatches;
}

// Frame::isInFrustum(MapPoint*, viewingCosLimit) (src/Frame.cc:269-325, called per local map point by Tracking::SearchLocalPoints,
// src/Tracking.cc:1165-1178) with MapPoint::PredictScale, the reference's own lines.  The pose split is Frame::UpdatePoseMatrices'
// (src/Frame.cc:260-267: mRcw, mtcw, mOw = -mRcw.t() * mtcw).  consider[i] = 0: the point is not handed to isInFrustum.
// Outputs per point: mbTrackInView, (mTrackProjX, mTrackProjY, mTrackProjXR), mnTrackScaleLevel, mTrackViewCos (0 where not in view).
extern "C" int matchref_is_in_frustum(int nP, const unsigned char* consider, const float* world, const float* normal,
                                      const float* min_dist, const float* max_dist, const float* Tcw, const float* K4,
                                      const float* bounds, float mbf, float log_scale_factor, int nlevels, float cos_limit,
                                      unsigned char* in_view, float* proj, int* level, float* view_cos) {
    using namespace ORB_SLAM2;
    Frame F;
    F.mTcw = cv::Mat(4, 4, CV_32F);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) F.mTcw.at<float>(r, c) = Tcw[4 * r + c];
    F.mRcw = F.mTcw.rowRange(0, 3).colRange(0, 3);
    F.mtcw = F.mTcw.rowRange(0, 3).col(3);
    F.mOw = -F.mRcw.t() * F.mtcw;
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];
    F.mbf = mbf;
    F.mfLogScaleFactor = log_scale_factor;
    F.mnScaleLevels = nlevels;
    int n_in = 0;
    for (int i = 0; i < nP; ++i) {
        in_view[i] = 0; level[i] = 0; view_cos[i] = 0.f;
        proj[3 * i] = proj[3 * i + 1] = proj[3 * i + 2] = 0.f;
        if (!consider[i]) continue;
        MapPoint mp;
        mp.mWorldPos = cv::Mat(3, 1, CV_32F);
        mp.mNormalVector = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) { mp.mWorldPos.at<float>(k, 0) = world[3 * i + k]; mp.mNormalVector.at<float>(k, 0) = normal[3 * i + k]; }
        mp.mfMinDistance = min_dist[i];
        mp.mfMaxDistance = max_dist[i];
        if (F.isInFrustum(&mp, cos_limit)) {
            in_view[i] = 1; ++n_in;
            proj[3 * i] = mp.mTrackProjX; proj[3 * i + 1] = mp.mTrackProjY; proj[3 * i + 2] = mp.mTrackProjXR;
            level[i] = mp.mnTrackScaleLevel;
            view_cos[i] = mp.mTrackViewCos;
        }
    }
    return n_in;
}
