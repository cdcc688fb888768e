// OpenCV shim primitives -- TEST INFRASTRUCTURE (oracle/), not product code.
//
// CPU restatements of the five OpenCV functions that /root/reference/src/ORBextractor.cc
// calls (call sites :809,:814 FAST; :1086 GaussianBlur; :1120 resize; :1122,:1127
// copyMakeBorder; :103 fastAtan2).  OpenCV is an un-vendored dependency of the reference
// (CMakeLists.txt:31-37, "OpenCV 3.0, else 2.4.3", no pinned version); the arithmetic
// below restates the published behaviour of OpenCV 4.x for 8-bit single-channel data
// and is pinned by tests/test_oracle_primitives.py against the cv2 4.13.0 wheel in this
// image.  Build with -ffp-contract=off (canonical rule B-2).
#include "opencv2/core/core.hpp"

#include <cfloat>
#include <cstdio>

namespace {
// Scratch that outlives a call must not come from operator new: oracle/ref_driver.cpp replaces
// it with a per-extraction bump arena (canonical rule B-1) that is rewound on every call.
template <typename T> struct Scratch {
    T* p;
    size_t cap;
    Scratch() : p(0), cap(0) {}
    ~Scratch() { std::free(p); }
    T* get(size_t n) {
        if (n > cap) { std::free(p); p = (T*)std::malloc(n * sizeof(T)); cap = n; }
        return p;
    }
};
}  // namespace

namespace cv {

void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) {
    std::fprintf(stderr, "cvshim: KeyPointsFilter::retainBest is only reachable from dead code\n");
    std::abort();
}

// ---------------------------------------------------------------- fastAtan2 (App. A-4)
float fastAtan2(float y, float x) {
    static const float scale = (float)(180.0 / CV_PI);
    static const float p1 = 0.9997878412794807f * scale;
    static const float p3 = -0.3258083974640975f * scale;
    static const float p5 = 0.1555786518463281f * scale;
    static const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)DBL_EPSILON;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------- copyMakeBorder (App. A-5)
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType) {
    (void)borderType;  // REFLECT_101; an ROI source is treated as isolated (no parent tracking in the shim)
    Mat src = _src.getMat();
    _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = _dst.getMat();
    const int w = src.cols, h = src.rows;
    uchar* inner = dst.data + (size_t)top * dst.step + left;
    if (inner != src.data)
        for (int y = 0; y < h; ++y) std::memmove(inner + (size_t)y * dst.step, src.data + (size_t)y * src.step, (size_t)w);
    for (int y = 0; y < h; ++y) {                 // left/right borders of the interior rows
        uchar* row = inner + (size_t)y * dst.step;
        for (int x = 1; x <= left; ++x) row[-x] = row[reflect101(-x, w)];
        for (int x = 0; x < right; ++x) row[w + x] = row[reflect101(w + x, w)];
    }
    const size_t full = (size_t)(w + left + right);
    for (int y = 1; y <= top; ++y)
        std::memcpy(dst.data + (size_t)(top - y) * dst.step, dst.data + (size_t)(top + reflect101(-y, h)) * dst.step, full);
    for (int y = 0; y < bottom; ++y)
        std::memcpy(dst.data + (size_t)(top + h + y) * dst.step, dst.data + (size_t)(top + reflect101(h + y, h)) * dst.step, full);
}

// ---------------------------------------------------------------- resize INTER_LINEAR 8UC1 (App. A-1)
namespace {
struct LinTab {
    Scratch<int> s_ofs0, s_ofs1;
    Scratch<short> s_c0, s_c1;
    int *ofs0, *ofs1;
    short *c0, *c1;
    void build(int ssize, int dsize) {
        ofs0 = s_ofs0.get(dsize); ofs1 = s_ofs1.get(dsize); c0 = s_c0.get(dsize); c1 = s_c1.get(dsize);
        const double scale = (double)ssize / dsize;
        for (int d = 0; d < dsize; ++d) {
            float f = (float)((d + 0.5) * scale - 0.5);
            int s = (int)std::floor(f);
            f -= s;
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
            long a1 = std::lrintf(f * 2048.f), a0 = std::lrintf((1.f - f) * 2048.f);
            c1[d] = (short)std::max(-32768L, std::min(32767L, a1));
            c0[d] = (short)std::max(-32768L, std::min(32767L, a0));
            ofs0[d] = s;
            ofs1[d] = std::min(s + 1, ssize - 1);
        }
    }
};
}  // namespace

void resize(InputArray _src, OutputArray _dst, Size dsize, double, double, int) {
    Mat src = _src.getMat();
    _dst.create(dsize.height, dsize.width, CV_8UC1);
    Mat dst = _dst.getMat();
    static thread_local LinTab tx, ty;
    static thread_local Scratch<int> rowscratch[2];
    tx.build(src.cols, dsize.width);
    ty.build(src.rows, dsize.height);
    const int dw = dsize.width;
    int* rowbuf[2] = {rowscratch[0].get(dw), rowscratch[1].get(dw)};
    int have[2] = {-1, -1};                                   // which source row each H buffer holds
    for (int dy = 0; dy < dsize.height; ++dy) {
        const int sy[2] = {ty.ofs0[dy], ty.ofs1[dy]};
        const int* H[2];
        for (int k = 0; k < 2; ++k) {
            const int slot = sy[k] & 1;                        // sy[1] is sy[0]+1 (or equal when clamped)
            if (have[slot] != sy[k]) {
                const uchar* s = src.data + (size_t)sy[k] * src.step;
                int* h = rowbuf[slot];
                const int *o0 = tx.ofs0, *o1 = tx.ofs1;
                const short *a0 = tx.c0, *a1 = tx.c1;
                for (int dx = 0; dx < dw; ++dx) h[dx] = (int)s[o0[dx]] * a0[dx] + (int)s[o1[dx]] * a1[dx];
                have[slot] = sy[k];
            }
            H[k] = rowbuf[slot];
        }
        const int b0 = ty.c0[dy], b1 = ty.c1[dy];
        uchar* d = dst.data + (size_t)dy * dst.step;
        const int *h0 = H[0], *h1 = H[1];
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (h0[dx] >> 4)) >> 16) + ((b1 * (h1[dx] >> 4)) >> 16) + 2) >> 2;
            d[dx] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

// ---------------------------------------------------------------- GaussianBlur 7x7 sigma 2, 8U (App. A-2)
void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int) {
    if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2.0 || sigmaY != 2.0) {
        std::fprintf(stderr, "cvshim: GaussianBlur only implements the reference's 7x7 sigma=2 call\n");
        std::abort();
    }
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};      // sums to 256
    Mat src = _src.getMat();
    const int w = src.cols, h = src.rows;
    static thread_local Scratch<unsigned short> s_rowpass;
    static thread_local Scratch<uchar> s_padded;
    unsigned short* rowpass = s_rowpass.get((size_t)w * h);
    uchar* padded = s_padded.get((size_t)w + 6);
    for (int y = 0; y < h; ++y) {
        const uchar* s = src.data + (size_t)y * src.step;
        uchar* p = padded;
        for (int x = -3; x < w + 3; ++x) p[x + 3] = s[reflect101(x, w)];
        unsigned short* r = rowpass + (size_t)y * w;
        for (int x = 0; x < w; ++x)
            r[x] = (unsigned short)(K[0] * (p[x] + p[x + 6]) + K[1] * (p[x + 1] + p[x + 5]) + K[2] * (p[x + 2] + p[x + 4]) + K[3] * p[x + 3]);
    }
    _dst.create(h, w, CV_8UC1);
    Mat dst = _dst.getMat();     // may alias src: the row pass above has consumed src completely
    for (int y = 0; y < h; ++y) {
        const unsigned short* r[7];
        for (int k = 0; k < 7; ++k) r[k] = rowpass + (size_t)reflect101(y + k - 3, h) * w;
        uchar* d = dst.data + (size_t)y * dst.step;
        for (int x = 0; x < w; ++x) {
            unsigned c = K[0] * ((unsigned)r[0][x] + r[6][x]) + K[1] * ((unsigned)r[1][x] + r[5][x]) +
                         K[2] * ((unsigned)r[2][x] + r[4][x]) + K[3] * (unsigned)r[3][x];
            d[x] = (uchar)((c + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------- FAST-9/16 with NMS (App. A-3)
namespace {
const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline bool has_run9(unsigned m) {          // 16-bit circular mask: is there a run of >= 9 set bits?
    unsigned x = m | (m << 16);
    x &= x >> 1;      // runs >= 2
    x &= x >> 2;      // runs >= 4
    x &= x >> 4;      // runs >= 8
    x &= (m | (m << 16)) >> 8;   // runs >= 9
    return (x & 0xffffu) != 0;
}
}  // namespace

void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax) {
    Mat img = _img.getMat();
    keypoints.clear();
    const int w = img.cols, h = img.rows;
    if (w < 7 || h < 7) return;
    const int step = (int)img.step;
    int ofs[25];
    for (int k = 0; k < 25; ++k) ofs[k] = RING_DY[k & 15] * step + RING_DX[k & 15];
    static thread_local Scratch<uchar> s_score;
    uchar* score = s_score.get((size_t)w * h);
    std::memset(score, 0, (size_t)w * h);
    const int t = threshold;
    for (int y = 3; y < h - 3; ++y) {
        const uchar* row = img.data + (size_t)y * step;
        uchar* srow = score + (size_t)y * w;
        for (int x = 3; x < w - 3; ++x) {
            const uchar* p = row + x;
            const int v = p[0];
            const int lo = v - t, hi = v + t;
            // every 9-arc contains one pixel of each opposite pair
            int a = p[ofs[0]], b = p[ofs[8]];
            if (a <= hi && a >= lo && b <= hi && b >= lo) continue;
            a = p[ofs[4]]; b = p[ofs[12]];
            if (a <= hi && a >= lo && b <= hi && b >= lo) continue;
            unsigned bright = 0, dark = 0;
            int d[25];
            for (int k = 0; k < 16; ++k) {
                const int q = p[ofs[k]];
                d[k] = v - q;
                bright |= (unsigned)(q > hi) << k;
                dark |= (unsigned)(q < lo) << k;
            }
            if (!has_run9(bright) && !has_run9(dark)) continue;
            for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
            int A = 0;                       // max over arcs of the min |difference| with a common sign
            for (int k = 0; k < 16; ++k) {
                int mn = d[k], mx = d[k];
                for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
                A = std::max(A, std::max(mn, -mx));
            }
            srow[x] = (uchar)(A - 1);
            if (!nonmax) keypoints.push_back(KeyPoint((float)x, (float)y, 7.f, -1.f, (float)(A - 1)));
        }
    }
    if (!nonmax) return;
    for (int y = 3; y < h - 3; ++y) {
        const uchar* s = score + (size_t)y * w;
        for (int x = 3; x < w - 3; ++x) {
            const int c = s[x];
            if (!c) continue;
            if (c > s[x - 1] && c > s[x + 1] && c > s[x - w - 1] && c > s[x - w] && c > s[x - w + 1] &&
                c > s[x + w - 1] && c > s[x + w] && c > s[x + w + 1])
                keypoints.push_back(KeyPoint((float)x, (float)y, 7.f, -1.f, (float)c));
        }
    }
}

}  // namespace cv

// ---------------------------------------------------------------- C entry points for the primitive KATs
extern "C" {
void cvshim_resize(const uchar* src, int sw, int sh, size_t sstep, uchar* dst, int dw, int dh) {
    cv::Mat s(sh, sw, 0, (void*)src, sstep), d(dh, dw, 0, dst, (size_t)dw);
    cv::resize(s, d, cv::Size(dw, dh), 0, 0, cv::INTER_LINEAR);
}
void cvshim_border(const uchar* src, int w, int h, size_t sstep, uchar* dst, int b) {
    cv::Mat s(h, w, 0, (void*)src, sstep), d(h + 2 * b, w + 2 * b, 0, dst, (size_t)(w + 2 * b));
    cv::copyMakeBorder(s, d, b, b, b, b, cv::BORDER_REFLECT_101);
}
void cvshim_blur(const uchar* src, int w, int h, size_t sstep, uchar* dst) {
    cv::Mat s(h, w, 0, (void*)src, sstep), d(h, w, 0, dst, (size_t)w);
    cv::GaussianBlur(s, d, cv::Size(7, 7), 2, 2, cv::BORDER_REFLECT_101);
}
int cvshim_fast(const uchar* src, int w, int h, size_t sstep, int threshold, int max_out, int* xyr) {
    cv::Mat s(h, w, 0, (void*)src, sstep);
    std::vector<cv::KeyPoint> k;
    cv::FAST(s, k, threshold, true);
    int n = (int)k.size();
    for (int i = 0; i < n && i < max_out; ++i) {
        xyr[3 * i] = (int)k[i].pt.x; xyr[3 * i + 1] = (int)k[i].pt.y; xyr[3 * i + 2] = (int)k[i].response;
    }
    return n;
}
float cvshim_fast_atan2(float y, float x) { return cv::fastAtan2(y, x); }
}
