// Test driver for the C++ adapter (orbslam2_with_quadrics_b200/cpp/ORBextractor.{h,cc}) compiled against
// the OpenCV shim.  Mirrors Frame::ExtractORB (reference src/Frame.cc:247-253): calls operator() with a
// cv::Mat image and an empty mask, then dumps keypoints, descriptors and the pyramid for the pytest side.
//   adapter_main <w> <h> <nfeatures> <scale> <nlevels> <ini> <min> <in.raw> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ORBextractor.h"

int main(int argc, char** argv)
{
    if (argc < 10) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    // ORBX_TEST_DEVICE: CUDA device of the extractors (ORBextractor::SetDevice); ORBX_TEST_STRIDE: row step of the cv::Mat
    // handed to operator() (> w: an odd, unaligned stride in ordinary pageable memory, like a ROI of a larger frame)
    if (const char* d = getenv("ORBX_TEST_DEVICE")) ORB_SLAM2::ORBextractor::SetDevice(atoi(d));
    const size_t step = getenv("ORBX_TEST_STRIDE") ? (size_t)atol(getenv("ORBX_TEST_STRIDE")) : (size_t)w;
    if (step < (size_t)w) return 2;
    ORB_SLAM2::ORBextractor ex(atoi(argv[3]), (float)atof(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]));
    std::vector<unsigned char> buf((size_t)w * h);
    FILE* f = fopen(argv[8], "rb");
    if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) return 3;
    fclose(f);
    std::vector<unsigned char> strided(step * (size_t)(h - 1) + (size_t)w + 1, 0xA5);      // +1: so that data + 1 is odd-aligned
    for (int y = 0; y < h; ++y) memcpy(strided.data() + 1 + (size_t)y * step, buf.data() + (size_t)y * w, (size_t)w);
    cv::Mat im(h, w, CV_8UC1, step == (size_t)w ? buf.data() : strided.data() + 1, step);
    std::vector<cv::KeyPoint> keys(3);                 // must be cleared by the call
    cv::Mat desc;
    ex(cv::Mat(), cv::Mat(), keys, desc);              // empty image: outputs untouched
    if (keys.size() != 3) return 4;
    for (int rep = 0; rep < 2; ++rep) ex(im, cv::Mat(), keys, desc);
    FILE* o = fopen(argv[9], "wb");
    int n = (int)keys.size(), nl = ex.GetLevels();
    fwrite(&n, 4, 1, o);
    fwrite(&nl, 4, 1, o);
    for (int i = 0; i < n; ++i) {
        float v[5] = {keys[i].pt.x, keys[i].pt.y, keys[i].size, keys[i].angle, keys[i].response};
        int q[2] = {keys[i].octave, keys[i].class_id};
        fwrite(v, 4, 5, o); fwrite(q, 4, 2, o);
    }
    for (int i = 0; i < n; ++i) fwrite(desc.ptr(i), 1, 32, o);
    std::vector<float> sf = ex.GetScaleFactors(), isf = ex.GetInverseScaleSigmaSquares();
    fwrite(sf.data(), 4, nl, o); fwrite(isf.data(), 4, nl, o);
    for (int l = 0; l < nl; ++l) {                     // padded planes through the public mvImagePyramid views
        const cv::Mat& m = ex.mvImagePyramid[l];
        int dims[2] = {m.cols, m.rows};
        fwrite(dims, 4, 2, o);
        for (int y = -19; y < m.rows + 19; ++y) fwrite(m.data + (long)y * (long)m.step - 19, 1, (size_t)m.cols + 38, o);
    }
    fclose(o);
    // optional stereo leg: <right.raw> <stereo_out.bin> <mbf> <mb> -- two extractors as in src/Frame.cc:78-81, then the
    // GPU ComputeStereoMatches; dumps N, mvuRight, mvDepth
    if (argc >= 14) {
        ORB_SLAM2::ORBextractor exr(atoi(argv[3]), (float)atof(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]));
        std::vector<unsigned char> rb((size_t)w * h);
        FILE* fr = fopen(argv[10], "rb");
        if (!fr || fread(rb.data(), 1, rb.size(), fr) != rb.size()) return 5;
        fclose(fr);
        cv::Mat imr(h, w, CV_8UC1, rb.data(), (size_t)w);
        std::vector<cv::KeyPoint> keysr;
        cv::Mat descr;
        exr(imr, cv::Mat(), keysr, descr);
        std::vector<float> uR, dep;
        ORB_SLAM2::ORBextractor::ComputeStereoMatches(ex, exr, (float)atof(argv[12]), (float)atof(argv[13]), uR, dep);
        FILE* so = fopen(argv[11], "wb");
        int ns = (int)uR.size();
        fwrite(&ns, 4, 1, so);
        fwrite(uR.data(), 4, uR.size(), so);
        fwrite(dep.data(), 4, dep.size(), so);
        fclose(so);
    }
    return 0;
}
