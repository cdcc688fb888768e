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
