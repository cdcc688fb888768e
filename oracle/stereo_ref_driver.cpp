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
