// Driver around the UNMODIFIED reference translation unit -- TEST INFRASTRUCTURE (oracle/).
//
// oracle/Makefile compiles /root/reference/src/ORBextractor.cc where it lies (never copied
// into this repo) against oracle/shim and links it with this file into
// oracle/_ref/liborbref.so.  The C entry points below are what tests/ and bench.py's
// CPU-baseline / --impl reference legs call through ctypes.
//
// Canonical rule B-1: DistributeOctTree sorts
// pair<int, ExtractorNode*> (src/ORBextractor.cc:684), i.e. ties are ordered by heap
// address.  A monotonic (bump, never reusing) operator new, active for the duration of
// one extraction and private to the calling thread, makes address order == creation
// order, so the unmodified code becomes a pure function of the image.
#include "ORBextractor.h"      // the reference's own header (-I/root/reference/include)

#include <sys/mman.h>
#include <new>

namespace {
struct Arena {
    char* base;
    size_t cap, off;
    int depth;
};
thread_local Arena g_arena = {0, 0, 0, 0};
const size_t kArenaBytes = (size_t)8 << 30;      // virtual reservation; pages are touched lazily

struct ArenaScope {
    ArenaScope() {
        if (!g_arena.base) {
            void* p = mmap(0, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) std::abort();
            g_arena.base = (char*)p;
            g_arena.cap = kArenaBytes;
        }
        if (g_arena.depth++ == 0) g_arena.off = 0;
    }
    ~ArenaScope() { --g_arena.depth; }
};
}  // namespace

void* operator new(size_t n) {
    Arena& a = g_arena;
    if (a.depth > 0) {
        size_t o = (a.off + 15) & ~(size_t)15;
        if (o + n > a.cap) { std::fprintf(stderr, "orbref: arena exhausted\n"); std::abort(); }
        a.off = o + n;
        return a.base + o;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
    const Arena& a = g_arena;
    if (a.base && (char*)p >= a.base && (char*)p < a.base + a.cap) return;
    std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

namespace {
// Derived only to reach the protected stage functions for stage-level parity tests.
class RefExtractor : public ORB_SLAM2::ORBextractor {
public:
    RefExtractor(int n, float s, int l, int a, int b) : ORB_SLAM2::ORBextractor(n, s, l, a, b) {}
    std::vector<cv::KeyPoint> Distribute(const std::vector<cv::KeyPoint>& k, int minX, int maxX, int minY, int maxY, int N) {
        return DistributeOctTree(k, minX, maxX, minY, maxY, N, 0);
    }
    const std::vector<int>& Quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int>& Umax() const { return umax; }
};
}  // namespace

extern "C" {

void* orbref_create(int nfeatures, float scale, int nlevels, int ini, int min) {
    return new RefExtractor(nfeatures, scale, nlevels, ini, min);
}
void orbref_destroy(void* h) { delete (RefExtractor*)h; }

// Runs ORBextractor::operator() (src/ORBextractor.cc:1043).  kps: max_kp x 7 x 4 bytes in
// cv::KeyPoint field order; desc: max_kp x 32.  Returns the keypoint count (outputs are
// filled only when it fits), -1 for "outputs untouched" (empty image).
int orbref_extract(void* h, const unsigned char* img, int w, int hgt, size_t stride, int max_kp, void* kps, unsigned char* desc) {
    RefExtractor* ex = (RefExtractor*)h;
    ArenaScope scope;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, stride);
    cv::Mat mask, d;
    std::vector<cv::KeyPoint> k;
    k.push_back(cv::KeyPoint());                 // sentinel: must be cleared by a non-empty run
    (*ex)(image, mask, k, d);
    if (image.empty()) return -1;
    int n = (int)k.size();
    if (n <= max_kp) {
        if (n) std::memcpy(kps, &k[0], (size_t)n * sizeof(cv::KeyPoint));
        for (int i = 0; i < n; ++i) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    }
    return n;
}

// Copies padded pyramid plane `level` ((h+38) x (w+38), step w+38) of the last extraction.
int orbref_pyramid_level(void* h, int level, unsigned char* out, int* w, int* hgt) {
    RefExtractor* ex = (RefExtractor*)h;
    if (level < 0 || level >= (int)ex->mvImagePyramid.size()) return -1;
    const cv::Mat& m = ex->mvImagePyramid[level];
    if (m.empty()) return -1;
    *w = m.cols; *hgt = m.rows;
    if (out) {
        const unsigned char* base = m.data - 19 * m.step - 19;
        for (int y = 0; y < m.rows + 38; ++y) std::memcpy(out + (size_t)y * (m.cols + 38), base + (size_t)y * m.step, (size_t)m.cols + 38);
    }
    return 0;
}

// ORBextractor::DistributeOctTree (src/ORBextractor.cc:539-763) on a caller-supplied candidate list.
// xyr: M x 3 ints (x, y, response) in box coordinates; out_idx receives the index of each kept
// candidate (found by matching the returned keypoints back to the input); returns the count.
int orbref_distribute(void* h, const int* xyr, int M, int minX, int maxX, int minY, int maxY, int N, int* out_idx, int max_out) {
    RefExtractor* ex = (RefExtractor*)h;
    ArenaScope scope;
    std::vector<cv::KeyPoint> in((size_t)M);
    for (int i = 0; i < M; ++i) {
        in[i] = cv::KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2]);
        in[i].class_id = i;                     // carried through the octree untouched
    }
    std::vector<cv::KeyPoint> out = ex->Distribute(in, minX, maxX, minY, maxY, N);
    int n = (int)out.size();
    for (int i = 0; i < n && i < max_out; ++i) out_idx[i] = out[i].class_id;
    return n;
}

int orbref_tables(void* h, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2, int* quotas, int* umax16) {
    RefExtractor* ex = (RefExtractor*)h;
    int n = ex->GetLevels();
    std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(), d = ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; ++i) { sf[i] = a[i]; inv_sf[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; quotas[i] = ex->Quotas()[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = ex->Umax()[i];
    return n;
}

float orbref_scale_factor(void* h) { return ((RefExtractor*)h)->GetScaleFactor(); }

}  // extern "C"
