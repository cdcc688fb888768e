// Hand-written sm_100a kernels for ORB_SLAM2::ORBextractor::operator().
//
// One kernel per stage of the reference path (src/ORBextractor.cc), every launch covers
// ALL frames of a batch with a persistent / grid-stride loop:
//   pyr_level_kernel      ComputePyramid           (:1107-1132)  resize + REFLECT_101 border
//   fast_cells_kernel     ComputeKeyPointsOctTree  (:765-829)    per-cell FAST-9 + NMS + retry
//   octree_kernel         DistributeOctTree        (:539-763)    quadtree split, exact order
//   orient_kernel         IC_Angle                 (:77-104)     intensity centroid + fastAtan2
//   blur_kernel           GaussianBlur 7x7 s=2     (:1086)       separable fixed point
//   desc_kernel           computeOrbDescriptor     (:108-147)    rotated BRIEF, packed stores
// No stage is a dense contraction: everything is integer/byte work bounded by HBM/L2 and
// instruction issue, so there is no tensor-core code here by design.
//
// Exactness rules: integer stages are bit-exact restatements of the
// OpenCV 4.x fixed-point arithmetic; float stages use __f*_rn intrinsics so nothing is
// contracted into FMAs.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "orbx_kernels.h"

namespace orbx {

// One-time per-device kernel configuration (dynamic shared memory limits, occupancy queries) in the launch wrappers is
// guarded by this mutex: handles are driven from several host threads (one handle per thread).
static std::mutex g_config_mutex;

// In global (not __constant__) memory on purpose: lane i reads ITS 32 bytes, i.e. 32 different
// addresses per warp, which the constant cache would serialise.
static __device__ __align__(16) signed char g_pattern[1024] = {
#include "orb_pattern_31.inc"
};

// Programmatic dependent launch (opt-in, ORBX_PDL=1): kernels are launched with programmatic stream serialization, so
// their blocks may be scheduled while the previous kernel of the stream drains; this wait (first statement of every
// kernel, a no-op without the attribute) blocks until that kernel has completed and its writes are visible.
// Measured on B200: device time of a single 1080p frame 0.163 -> 0.154 ms (VGA 0.119 -> 0.100), but the end-to-end
// single-frame latency through orbx_extract (multi-stream H2D / kernels / D2H) got worse (0.234 -> 0.252 ms), and
// batched throughput is unchanged -- hence off by default.
#define ORBX_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool no_pdl = getenv("ORBX_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// -DORBX_BOUNDS_CHECK (the liborbx_boundscheck.so build, tests/test_gpu_bounds.py): every shared-memory tile / queue /
// score-map / bitmap index of the FAST and describe kernels is checked on the device and a violation traps (the launch
// fails with cudaErrorLaunchFailure instead of silently reading a neighbour's bytes).  compute-sanitizer is closed on the
// GPU pool this was developed on; this build is the tool-checked substitute.  Compiled out of the product library.
#ifdef ORBX_BOUNDS_CHECK
#define ORBX_BC(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define ORBX_BC(cond) do { } while (0)
#endif

__device__ __forceinline__ int reflect_clamp(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return max(0, min(i, n - 1));
}

__device__ __forceinline__ const uint8_t* level_px(const uint8_t* slab, const OrbxLevel& L, int x, int y) {
    return slab + L.plane_off + (size_t)(y + ORBX_EDGE) * L.pitch + ORBX_XO + x;
}

// =====================================================================================
// ComputePyramid (:1107-1132).  Every padded pixel is produced directly as
// f(reflect(x), reflect(y)), so the REFLECT_101 border (copyMakeBorder of the just-resized
// level, :1122) needs no second pass.
//
// pyr_level0_kernel: level 0 = the input image plus border (:1127); aligned 32-bit copies in
// the interior, byte gathers only where the reflection applies.
//
// pyr_resize_kernel: level l from level l-1, cv::resize INTER_LINEAR 8U (11-bit fixed point,
// SURVEY App. A-1).  One lane owns 4 adjacent plane columns (one 32-bit store) and walks down
// 16 plane rows.  Its column taps live in registers: per pixel the packed weights (a0, a1) for
// one IDP.2A and the byte offset of its source pair inside three aligned words of the source
// row.  H[sy][dx] >> 4 is cached for the two most recent source rows, so going down one output
// row costs ~1.2 row passes; the column pass is two IMAD.HI per pixel.
// =====================================================================================
// One work item = 32 x 8 threads' worth of level 0 (32 lanes x 16 bytes wide, 8 plane rows); (tx, ty) = the thread's place in it.
// A 16-byte vector is "interior" when it is a plain (re-aligned) copy of source bytes; the few vectors per row that touch the
// reflected border are byte gathers (~200 instructions), and a warp that holds one of them pays for it with all 32 lanes.
// They are therefore left out here and given to pyr_level0_border_item, where every lane of a warp is such a vector
// (lane = (row, border vector)): 20.5 M -> 7 M warp instructions per 32 x 1080p.
struct Lv0Cols { int lo, hiv, nL, cR, nb; };
__host__ __device__ __forceinline__ Lv0Cols pyr_level0_cols(int w, int aligned16) {
    Lv0Cols c;
    c.lo = aligned16 ? 0 : 16;                                     // interior: lo <= dx0 <= hiv (dx0 = source x of byte 0, multiple of 16)
    c.hiv = aligned16 ? w - 16 : w - 20;
    c.nL = (ORBX_XO + c.lo) / 16;                                  // border vectors left of the interior
    c.cR = c.hiv >= c.lo ? ORBX_XO + (c.hiv / 16 + 1) * 16 : ORBX_XO + c.lo;      // first plane column right of it
    const int ncols16 = (ORBX_XO + w + ORBX_EDGE + 15) / 16;
    c.nb = c.nL + max(ncols16 - c.cR / 16, 0);
    return c;
}

// (a thread copies its vector of PYR_L0_ROWS plane rows, 8 rows apart: all loads are issued before the first store --
// the copy is bound by bytes in flight, not by instructions)
#define PYR_L0_ROWS 4
__device__ __forceinline__ void pyr_level0_item(const OrbxPlan* __restrict__ plan, const uint8_t* __restrict__ imgs, size_t img_pitch,
                                                size_t img_frame_stride, int aligned16, uint8_t* pyr, int bx, int by, int frame,
                                                int tx, int ty) {
    const OrbxLevel& L = plan->lv[0];
    const int w = L.w, h = L.h;
    const int c = (bx * 32 + tx) * 16;                             // plane column, multiple of 16 (XO is too)
    const int row0 = by * (8 * PYR_L0_ROWS) + ty;                  // first plane row
    const int dx0 = c - ORBX_XO;
    const Lv0Cols lc = pyr_level0_cols(w, aligned16);
    if (row0 >= L.rows || dx0 < lc.lo || dx0 > lc.hiv) return;     // border vectors: pyr_level0_border_item
    const uint8_t* src0 = imgs + (size_t)frame * img_frame_stride + dx0;
    uint4 out[PYR_L0_ROWS];
    if (aligned16) {
#pragma unroll
        for (int k = 0; k < PYR_L0_ROWS; ++k) {
            const int dy = reflect_clamp(min(row0 + 8 * k, L.rows - 1) - ORBX_EDGE, h);
            out[k] = __ldg(reinterpret_cast<const uint4*>(src0 + (size_t)dy * img_pitch));
        }
    } else {
        // source rows of arbitrary alignment (e.g. a 1241-byte stride): five aligned words, re-aligned with funnel
        // shifts; the reads stay inside this image row
#pragma unroll
        for (int k = 0; k < PYR_L0_ROWS; ++k) {
            const int dy = reflect_clamp(min(row0 + 8 * k, L.rows - 1) - ORBX_EDGE, h);
            const uint8_t* ps = src0 + (size_t)dy * img_pitch;
            const int a = (int)(reinterpret_cast<uintptr_t>(ps) & 3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(ps - a);
            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3), w4 = __ldg(wp + 4);
            out[k] = make_uint4(__funnelshift_r(w0, w1, 8 * a), __funnelshift_r(w1, w2, 8 * a), __funnelshift_r(w2, w3, 8 * a),
                                __funnelshift_r(w3, w4, 8 * a));
        }
    }
    uint8_t* dst = pyr + (size_t)frame * plan->slab_bytes + L.plane_off + (size_t)row0 * L.pitch + c;
#pragma unroll
    for (int k = 0; k < PYR_L0_ROWS; ++k)
        if (row0 + 8 * k < L.rows) *reinterpret_cast<uint4*>(dst + (size_t)(8 * k) * L.pitch) = out[k];
}

// item = 256 (plane row, border vector) pairs
__device__ __forceinline__ void pyr_level0_border_item(const OrbxPlan* __restrict__ plan, const uint8_t* __restrict__ imgs, size_t img_pitch,
                                                       size_t img_frame_stride, int aligned16, uint8_t* pyr, int item, int frame, int tid) {
    const OrbxLevel& L = plan->lv[0];
    const int w = L.w, h = L.h;
    const Lv0Cols lc = pyr_level0_cols(w, aligned16);
    const int p = item * 256 + tid;
    const int row = p / lc.nb, j = p - row * lc.nb;
    if (row >= L.rows) return;
    const int c = j < lc.nL ? 16 * j : lc.cR + 16 * (j - lc.nL);
    const int dx0 = c - ORBX_XO;
    const int dy = reflect_clamp(row - ORBX_EDGE, h);
    const uint8_t* src = imgs + (size_t)frame * img_frame_stride + (size_t)dy * img_pitch;
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i >> 2] |= (uint32_t)__ldg(src + reflect_clamp(dx0 + i, w)) << (8 * (i & 3));
    *reinterpret_cast<uint4*>(pyr + (size_t)frame * plan->slab_bytes + L.plane_off + (size_t)row * L.pitch + c) = make_uint4(o[0], o[1], o[2], o[3]);
}

// grid (gx, gy_main + gy_border, frames): block rows below gy_main copy the interior, the rest take the border vectors
__global__ void __launch_bounds__(256) pyr_level0_kernel(const OrbxPlan* __restrict__ plan,
                                                         const uint8_t* __restrict__ imgs, size_t img_pitch,
                                                         size_t img_frame_stride, int aligned16,
                                                         uint8_t* __restrict__ pyr, int gy_main) {
    ORBX_PDL_WAIT();
    if ((int)blockIdx.y < gy_main)
        pyr_level0_item(plan, imgs, img_pitch, img_frame_stride, aligned16, pyr, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, threadIdx.y);
    else
        pyr_level0_border_item(plan, imgs, img_pitch, img_frame_stride, aligned16, pyr,
                               ((int)blockIdx.y - gy_main) * (int)gridDim.x + (int)blockIdx.x, blockIdx.z, threadIdx.y * 32 + threadIdx.x);
}

// =====================================================================================
// cvtColor(RGB/BGR(A) -> GRAY) of Tracking::GrabImage* (reference src/Tracking.cc:172-197, :212-225, :242-255): the step
// in front of the path.  OpenCV 4.x 8U arithmetic, verified against cv2 4.13 on 16 M random
// pixels: gray = (B * 3735 + G * 19235 + R * 9798 + 16384) >> 15.  A thread converts 16 adjacent pixels: 3 or 4 aligned
// 128-bit loads, per pixel one funnel shift and two IDP.2A (the fourth byte meets a zero weight), one 128-bit store.
// =====================================================================================
template <int CH>
__global__ void __launch_bounds__(256) cvt_gray_kernel(const uint8_t* __restrict__ src, size_t src_pitch, size_t src_frame_stride,
                                                       int w, int h, uint32_t c01, uint32_t c2, int aligned16,
                                                       uint8_t* __restrict__ dst, size_t dst_pitch, size_t dst_frame_stride) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 16, y = blockIdx.y * 8 + threadIdx.y, frame = blockIdx.z;
    if (x0 >= w || y >= h) return;
    const uint8_t* row = src + (size_t)frame * src_frame_stride + (size_t)y * src_pitch;
    uint8_t* out = dst + (size_t)frame * dst_frame_stride + (size_t)y * dst_pitch + x0;
    uint32_t o[4] = {0, 0, 0, 0};
    if (aligned16 && x0 + 16 <= w) {
        const uint4* p = reinterpret_cast<const uint4*>(row + (size_t)x0 * CH);
        uint32_t wd[CH * 4 + 1];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const uint4 v = __ldg(p + k);
            wd[4 * k] = v.x; wd[4 * k + 1] = v.y; wd[4 * k + 2] = v.z; wd[4 * k + 3] = v.w;
        }
        wd[CH * 4] = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = (i * CH) >> 2, sh = 8 * ((i * CH) & 3);
            const uint32_t px = CH == 4 ? wd[i] : __funnelshift_r(wd[j], wd[j + 1], sh);
            const uint32_t g = (__dp2a_lo(c01, px, __dp2a_hi(c2, px, 16384u))) >> 15;
            o[i >> 2] |= g << (8 * (i & 3));
        }
        *reinterpret_cast<uint4*>(out) = make_uint4(o[0], o[1], o[2], o[3]);      // dst rows are 16-byte aligned (staging buffer)
    } else {
        for (int i = 0; i < 16 && x0 + i < w; ++i) {
            const uint8_t* q = row + (size_t)(x0 + i) * CH;
            const uint32_t px = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
            out[i] = (uint8_t)((__dp2a_lo(c01, px, __dp2a_hi(c2, px, 16384u))) >> 15);
        }
    }
}

#ifndef PYR_RY
#define PYR_RY 16
#endif
template <bool WIDE>
__device__ __forceinline__ void pyr_resize_item(const OrbxPlan* __restrict__ plan, int l, int RY, uint8_t* pyr,
                                                const OrbxTap* __restrict__ taps, int bx, int by, int frame, int tid) {
    const OrbxLevel& L = plan->lv[l];
    const OrbxLevel& S = plan->lv[l - 1];
    const int w = L.w, h = L.h;
    const int c = 12 + (bx * 32 + (tid & 31)) * 4;                             // plane column, multiple of 4
    const int row0 = (by * 4 + (tid >> 5)) * RY;                              // first plane row of this warp
    if (row0 >= L.rows || c >= ORBX_XO + w + ORBX_EDGE) return;
    uint8_t* slab = pyr + (size_t)frame * plan->slab_bytes;
    // ---- column taps of the 4 pixels
    int sx[4];
    uint32_t coef[4];
    int minsx = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const OrbxTap t = taps[L.xtab_off + reflect_clamp(c + j - ORBX_XO, w)];
        sx[j] = t.ofs;
        coef[j] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
        minsx = min(minsx, t.ofs);
    }
    const int wb = minsx & ~3;                                                // first source column of word 0
    int sh[4], wofs[4];
    bool hiw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = sx[j] - wb;                                             // <= 7 unless WIDE (scale factors > ~1.3)
        hiw[j] = o >= 4;
        sh[j] = 8 * (o & 3);
        wofs[j] = o >> 2;
    }
    const uint8_t* sbase = slab + S.plane_off + (size_t)ORBX_EDGE * S.pitch + ORBX_XO + wb;   // row 0 of the source level
    const int spitch = S.pitch;
    uint32_t HA[4] = {0, 0, 0, 0}, HB[4] = {0, 0, 0, 0};
    int rowA = -1, rowB = -1;
    // the three source words of a row (non-WIDE); loaded one output row ahead so the DRAM/L2 latency of the
    // next row pass overlaps the arithmetic of the current one
    struct Words { uint32_t w0, w1, w2; };
    auto load_row = [&](int sy) -> Words {
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(sbase + (unsigned)(sy * spitch));
        Words q;
        q.w0 = rp[0]; q.w1 = rp[1]; q.w2 = rp[2];
        return q;
    };
    auto row_pass_words = [&](const Words& q, uint32_t* H) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t pair = __funnelshift_r(hiw[j] ? q.w1 : q.w0, hiw[j] ? q.w2 : q.w1, sh[j]);
            H[j] = __dp2a_lo(coef[j], pair, 0u) >> 4;                         // (S[sx]*a0 + S[sx+1]*a1) >> 4
        }
    };
    auto row_pass = [&](int sy, uint32_t* H) {
        if (WIDE) {
            const uint32_t* rp = reinterpret_cast<const uint32_t*>(sbase + (unsigned)(sy * spitch));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t pair = __funnelshift_r(rp[wofs[j]], rp[wofs[j] + 1], sh[j]);
                H[j] = __dp2a_lo(coef[j], pair, 0u) >> 4;
            }
        } else {
            row_pass_words(load_row(sy), H);
        }
    };
    const int row_end = min(row0 + RY, L.rows);
    const int hs1 = S.h - 1;
    uint8_t* dst = slab + L.plane_off + (size_t)row0 * L.pitch + c;
    const OrbxTap* ytab = taps + L.ytab_off;
    OrbxTap ty = ytab[reflect_clamp(row0 - ORBX_EDGE, h)];
    Words pre = {0, 0, 0};
    if (!WIDE) pre = load_row(min(ty.ofs + 1, hs1));                          // row r1 of the first output row
    for (int row = row0; row < row_end; ++row, dst += L.pitch) {
        // prefetch for the next output row: its tap entry and (speculatively) its second source row
        const OrbxTap tyn = ytab[reflect_clamp(min(row + 1, L.rows - 1) - ORBX_EDGE, h)];
        Words pren = {0, 0, 0};
        if (!WIDE) pren = load_row(min(tyn.ofs + 1, hs1));
        const int r0 = ty.ofs, r1 = min(r0 + 1, hs1);
        if (r0 == rowB) {                                                     // the usual step: one new source row
#pragma unroll
            for (int j = 0; j < 4; ++j) HA[j] = HB[j];
            rowA = rowB;
            if (r1 != r0) {
                if (WIDE) row_pass(r1, HB); else row_pass_words(pre, HB);
                rowB = r1;
            }
        } else if (!(r0 == rowA && (r1 == rowB || r1 == r0))) {
            row_pass(r0, HA);
            rowA = r0;
            if (r1 != r0) {
                if (WIDE) row_pass(r1, HB); else row_pass_words(pre, HB);
                rowB = r1;
            } else {
                rowB = -1;
            }
        }
        // when the source row is clamped (r1 == r0) the table's second weight is 0, so a stale HB is harmless
        const uint32_t b0 = (uint32_t)ty.c0 << 16, b1 = (uint32_t)ty.c1 << 16;
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // <= 255: the weights sum to 2048 and H >> 4 <= 255 * 128
            const uint32_t v = (__umulhi(b0, HA[j]) + __umulhi(b1, HB[j]) + 2u) >> 2;
            out |= v << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(dst) = out;
        ty = tyn;
        pre = pren;
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(128) pyr_resize_kernel(const OrbxPlan* __restrict__ plan, int l, int RY, uint8_t* pyr,
                                                         const OrbxTap* __restrict__ taps) {
    ORBX_PDL_WAIT();
    pyr_resize_item<WIDE>(plan, l, RY, pyr, taps, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
}

// pyr_resize8_kernel: the same arithmetic with 8 adjacent plane columns per lane (one 64-bit store per row) and every
// column decision moved into a host-built OrbxCol8 record.  A half (4 pixels) reads 3 aligned source words; one funnel
// shift pair leaves the 8 bytes from the half's smallest tap offset on in (U, V), a pixel PAIR's two source byte
// pairs are one PRMT with a constant selector, a pixel's row sum one IDP.2A.  The two cached rows H0/H1 never move:
// which of them is the upper row is tracked by their source row ids and only the two row weights swap.  The row-tap
// table is indexed by PLANE row (border reflection folded in).  ~11 thread instructions per output pixel against 26.
struct Row6 { uint32_t a0, a1, a2, b0, b1, b2; };

#ifndef ORBX_R8_MINB
#define ORBX_R8_MINB 1
#endif
// Cache-line prefetch of the source rows ORBX_R8_PF rows beyond the register prefetch (costs one instruction and no
// register per output row).  Measured at 64 x 1080p, pyramid ms per step: off 0.425, 2 rows 0.388, 3 rows 0.393, 4 rows
// 0.394, 6 rows 0.400; into L1 or only into L2 makes no difference (0.3925 / 0.3928 at 3 rows).
#ifndef ORBX_R8_PF
#define ORBX_R8_PF 2
#endif
#ifndef ORBX_R8_PF_L1
#define ORBX_R8_PF_L1 1         // 1: prefetch.global.L1, 0: prefetch.global.L2
#endif
__device__ __forceinline__ void pyr_resize8_item(const OrbxPlan* __restrict__ plan, int l, int RY, uint8_t* pyr,
                                                 const OrbxTap* __restrict__ taps, int bx, int by, int frame, int tid) {
    const OrbxLevel& L = plan->lv[l];
    const OrbxLevel& S = plan->lv[l - 1];
    const int g = bx * 32 + (tid & 31);                                        // group of 8 plane columns from column 8 on
    const int row0 = (by * 4 + (tid >> 5)) * RY;                              // first plane row of this warp
    if (row0 >= L.rows || g >= L.ngroups8) return;
    uint8_t* slab = pyr + (size_t)frame * plan->slab_bytes;
    const uint4* cgp = reinterpret_cast<const uint4*>(taps + L.col8_off) + 4 * g;
    const uint4 c0 = __ldg(cgp), c1 = __ldg(cgp + 1), c2 = __ldg(cgp + 2), c3 = __ldg(cgp + 3);
    const uint8_t* sbase = slab + S.plane_off + (size_t)ORBX_EDGE * S.pitch + ORBX_XO;     // source pixel (0, 0)
    const uint8_t* plo = sbase + (int)c0.x;
    const uint8_t* phi = sbase + (int)c0.y;
    const uint32_t shl = c0.z & 31u, shh = (c0.z >> 8) & 31u;
    const uint32_t sel0 = c1.x, sel1 = c1.y, sel2 = c1.z, sel3 = c1.w;
    const uint32_t k0 = c2.x, k1 = c2.y, k2 = c2.z, k3 = c2.w, k4 = c3.x, k5 = c3.y, k6 = c3.z, k7 = c3.w;
    const int spitch = S.pitch;
    auto load_row = [&](int sy) -> Row6 {
        const unsigned o = (unsigned)(sy * spitch);
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(plo + o);
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(phi + o);
        Row6 q;
        q.a0 = pa[0]; q.a1 = pa[1]; q.a2 = pa[2];
        q.b0 = pb[0]; q.b1 = pb[1]; q.b2 = pb[2];
        return q;
    };
    auto pass = [&](const Row6& q, uint32_t* H) {                              // H[j] = (S[sx]*a0 + S[sx+1]*a1) >> 4
        const uint32_t U = __funnelshift_r(q.a0, q.a1, shl), V = __funnelshift_r(q.a1, q.a2, shl);
        const uint32_t X = __funnelshift_r(q.b0, q.b1, shh), Y = __funnelshift_r(q.b1, q.b2, shh);
        const uint32_t p0 = __byte_perm(U, V, sel0), p1 = __byte_perm(U, V, sel1);
        const uint32_t p2 = __byte_perm(X, Y, sel2), p3 = __byte_perm(X, Y, sel3);
        H[0] = __dp2a_lo(k0, p0, 0u) >> 4; H[1] = __dp2a_hi(k1, p0, 0u) >> 4;
        H[2] = __dp2a_lo(k2, p1, 0u) >> 4; H[3] = __dp2a_hi(k3, p1, 0u) >> 4;
        H[4] = __dp2a_lo(k4, p2, 0u) >> 4; H[5] = __dp2a_hi(k5, p2, 0u) >> 4;
        H[6] = __dp2a_lo(k6, p3, 0u) >> 4; H[7] = __dp2a_hi(k7, p3, 0u) >> 4;
    };
    const int nrow = min(RY, L.rows - row0);
    const int hs1 = S.h - 1;
    const uint2* yt = reinterpret_cast<const uint2*>(taps + L.yrow_off + row0);     // (source row, c0 | c1 << 16)
    uint8_t* dst = slab + L.plane_off + (size_t)row0 * L.pitch + 8 + 8 * g;
    uint32_t H0[8], H1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) H0[j] = H1[j] = 0;
    int id0 = -1, id1 = -1;
    uint2 ty = __ldg(yt);
    int pre_id = min((int)ty.x + 1, hs1);                                     // speculative: the lower row of the first output row
    Row6 pre = load_row(pre_id);
    for (int i = 0; i < nrow; ++i, dst += L.pitch) {
        // one output row ahead: its tap entry and (speculatively) its lower source row
        const uint2 tyn = __ldg(yt + min(i + 1, nrow - 1));
        const int pren_id = min((int)tyn.x + 1, hs1);
        const Row6 pren = load_row(pren_id);
#if ORBX_R8_PF > 0
        {   // the source rows ORBX_R8_PF output rows further down: a cache-line prefetch costs no register
            const unsigned o = (unsigned)(min(pren_id + ORBX_R8_PF, hs1) * spitch);
#if ORBX_R8_PF_L1
            asm volatile("prefetch.global.L1 [%0];" ::"l"(plo + o));
#else
            asm volatile("prefetch.global.L2 [%0];" ::"l"(plo + o));
#endif
        }
#endif
        const int r0 = (int)ty.x, r1 = min(r0 + 1, hs1);
        uint32_t b0, b1;                                                      // row weights of H0, H1, << 16
        if (r0 == id1) {                                                      // the usual step: last row's lower row is the upper one
            if (r1 != r0 && id0 != r1) {
                if (pre_id == r1) pass(pre, H0); else pass(load_row(r1), H0);
                id0 = r1;
            }
            b0 = ty.y & 0xffff0000u; b1 = ty.y << 16;
        } else {
            if (r0 != id0) { pass(load_row(r0), H0); id0 = r0; }
            if (r1 != r0 && id1 != r1) {
                if (pre_id == r1) pass(pre, H1); else pass(load_row(r1), H1);
                id1 = r1;
            }
            b0 = ty.y << 16; b1 = ty.y & 0xffff0000u;
        }
        // when the source row is clamped (r1 == r0) the second weight is 0, so a stale lower row is harmless
        uint32_t s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = __umulhi(b0, H0[j]) + __umulhi(b1, H1[j]);     // <= 1020: weights sum to 2048
        // (s + 2) >> 2 for two pixels at a time (10-bit fields), then one PRMT per four pixels
        const uint32_t t01 = (s[0] + (s[1] << 16) + 0x00020002u) >> 2, t23 = (s[2] + (s[3] << 16) + 0x00020002u) >> 2;
        const uint32_t t45 = (s[4] + (s[5] << 16) + 0x00020002u) >> 2, t67 = (s[6] + (s[7] << 16) + 0x00020002u) >> 2;
        *reinterpret_cast<uint2*>(dst) = make_uint2(__byte_perm(t01, t23, 0x6420), __byte_perm(t45, t67, 0x6420));
        ty = tyn;
        pre = pren;
        pre_id = pren_id;
    }
}

__global__ void __launch_bounds__(128, ORBX_R8_MINB) pyr_resize8_kernel(const OrbxPlan* __restrict__ plan, int l, int RY, uint8_t* pyr,
                                                                        const OrbxTap* __restrict__ taps) {
    ORBX_PDL_WAIT();
    pyr_resize8_item(plan, l, RY, pyr, taps, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
}

// =====================================================================================
// FAST-9/16 per 30-px cell (cv::FAST(window, t, true)), one warp per cell,
// persistent warps with a dynamic work counter.
//
// Staging: a work item is a strip of NC horizontally adjacent cell windows (default 2), fetched by
// ONE elected lane with a TMA tile load (cp.async.bulk.tensor.3d over a per-level
// {pitch, rows, frame} tensor map, mbarrier complete_tx) into shared memory (NB = 1 or 2 tile
// buffers per warp; with 2 the next strip is in flight while the current one is processed --
// measured equal, so the default is 1 and more resident warps).  TMA needs the inner coordinate
// on a 16-byte boundary, so the BW x BH box starts at the 16-aligned column at or before
// (window x0 - 1) and a window sits `delta` bytes into the tile: tile column = window x + 1 +
// delta.  Words are re-aligned with funnel shifts.
//   phase 1  4 pixels per lane, SIMD-in-word: |ring - centre| for the compass points 0/8 and
//            4/12 with VABSDIFF4; a 9-arc needs one pixel of every opposite pair beyond the
//            threshold, so pixels failing either pair are dropped.  8 warp iterations of tests
//            are issued back to back, their survivor counts scanned together (packed 8-bit
//            counts), survivors queued in row-major order.
//   phase 2  per survivor: exact score A = max over the 16 arcs of min(+-(ring - centre)) with
//            packed 16-bit min/max (two arcs per instruction); corner iff A > t, score = A - 1.
//   phase 3  strict 3x3 NMS on a zero-framed score map, count, then ordered emission.
// The iniThFAST pass is repeated with minThFAST iff it produced no keypoint (:812).
// =====================================================================================
struct FastMaps {
    CUtensorMap m[ORBX_MAXL];
};

// ceil(65536 / G): (n * c_recip16[G]) >> 16 == n / G for n <= 32, G <= 32
static __constant__ int c_recip16[33] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042,
                                         4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428,
                                         2341, 2260, 2185, 2115, 2048};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// pyr_resize8_tile_kernel: the same lanes, tables and arithmetic as pyr_resize8_kernel, but the source pixels of a CTA's
// 256 x 64 output block (<= 307 x 79 source pixels + alignment) arrive as ONE TMA tile in shared memory (u32 tensor map over
// the source plane: the 336-byte box is wider than a u8 box may be), so a source row costs six LDS instead of six dependent
// global loads and no register prefetch is needed.  Used for the launches whose warps take full PYR_RY-row strips (the
// large levels of a batch).  A CTA always covers 64 plane rows: with shorter strips (RY = 8 / 4 / 2 rows per warp, the smaller
// levels) it simply has 8 / 16 / 32 warps sharing the one tile.  (A persistent variant -- a CTA walking down its column strip
// with two tile buffers, the next tile in flight during the computation -- was measured slower: 0.415 vs 0.343 ms per
// 64 x 1080p; the 54 KB of tiles leave 4 CTAs per SM, and this kernel lives on resident warps.)
#define PYR_TILE_W 336                       // bytes per tile row: 256 * 1.2 + 15 (alignment) + 12 (three words per half) <= 336
#define PYR_TILE_H 80                        // source rows of 64 output rows: 64 * 1.2 + 2 <= 80
template <int RY>
__global__ void __launch_bounds__(64 / RY * 32) pyr_resize8_tile_kernel(const __grid_constant__ CUtensorMap src_map, const OrbxPlan* __restrict__ plan,
                                                                        int l, int frame0, uint8_t* pyr, const OrbxTap* __restrict__ taps) {
    // (`pyr` is the slab of the sub-batch's first frame, the tensor map covers the handle's whole pyramid: frame0 = that frame)
    ORBX_PDL_WAIT();
    constexpr int NW = 64 / RY;                                               // warps per CTA
    __shared__ __align__(128) uint8_t s_tile[PYR_TILE_W * PYR_TILE_H];
    __shared__ uint64_t s_bar;
    __shared__ int s_org[2];
    const OrbxLevel& L = plan->lv[l];
    const OrbxLevel& S = plan->lv[l - 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.z;
    const int g = blockIdx.x * 32 + lane;                                      // group of 8 plane columns from column 8 on
    const int gc = min(g, L.ngroups8 - 1);
    const int row0 = (blockIdx.y * NW + warp) * RY;                            // first plane row of this warp
    const uint4* cgp = reinterpret_cast<const uint4*>(taps + L.col8_off) + 4 * gc;
    const uint4 c0 = __ldg(cgp), c1 = __ldg(cgp + 1), c2 = __ldg(cgp + 2), c3 = __ldg(cgp + 3);
    const int hs1 = S.h - 1;
    // tile origin: smallest source column of the CTA's 32 groups (16-byte aligned, plane coordinates) and smallest source
    // row of its 64 plane rows (border rows reflect, so neither is simply the first entry)
    if (warp == 0) {
        int mn = min((int)c0.x, (int)c0.y);
        mn = __reduce_min_sync(0xffffffffu, mn);
        const int cta_row0 = blockIdx.y * 64;
        int r = INT_MAX;
        for (int i = lane; i < 64; i += 32)
            if (cta_row0 + i < L.rows) r = min(r, (int)__ldg(reinterpret_cast<const uint2*>(taps + L.yrow_off + cta_row0 + i)).x);
        r = __reduce_min_sync(0xffffffffu, r);
        if (lane == 0) {
            const int tx0 = (mn + ORBX_XO) & ~15, ty0 = r + ORBX_EDGE;
            s_org[0] = tx0;
            s_org[1] = ty0;
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_bar, (uint32_t)(PYR_TILE_W * PYR_TILE_H));
            tma_load_3d(s_tile, &src_map, &s_bar, tx0 >> 2, ty0, frame0 + frame);
        }
    }
    __syncthreads();
    const int tx0 = s_org[0], ty0 = s_org[1];
    mbar_wait(&s_bar, 0);
    if (row0 >= L.rows || g >= L.ngroups8) return;
    uint8_t* slab = pyr + (size_t)frame * plan->slab_bytes;
    const uint8_t* tlo = s_tile + ((int)c0.x + ORBX_XO - tx0) + (ORBX_EDGE - ty0) * PYR_TILE_W;      // source pixel (0, 0) of each half
    const uint8_t* thi = s_tile + ((int)c0.y + ORBX_XO - tx0) + (ORBX_EDGE - ty0) * PYR_TILE_W;
    const uint32_t shl = c0.z & 31u, shh = (c0.z >> 8) & 31u;
    const uint32_t sel0 = c1.x, sel1 = c1.y, sel2 = c1.z, sel3 = c1.w;
    const uint32_t k0 = c2.x, k1 = c2.y, k2 = c2.z, k3 = c2.w, k4 = c3.x, k5 = c3.y, k6 = c3.z, k7 = c3.w;
    auto pass = [&](int sy, uint32_t* H) {                                     // H[j] = (S[sx]*a0 + S[sx+1]*a1) >> 4 of source row sy
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(tlo + sy * PYR_TILE_W);
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(thi + sy * PYR_TILE_W);
        ORBX_BC(reinterpret_cast<const uint8_t*>(pa) >= s_tile && reinterpret_cast<const uint8_t*>(pa + 3) <= s_tile + sizeof(s_tile) &&
                reinterpret_cast<const uint8_t*>(pb) >= s_tile && reinterpret_cast<const uint8_t*>(pb + 3) <= s_tile + sizeof(s_tile));
        const uint32_t a0 = pa[0], a1 = pa[1], a2 = pa[2], b0 = pb[0], b1 = pb[1], b2 = pb[2];
        const uint32_t U = __funnelshift_r(a0, a1, shl), V = __funnelshift_r(a1, a2, shl);
        const uint32_t X = __funnelshift_r(b0, b1, shh), Y = __funnelshift_r(b1, b2, shh);
        const uint32_t p0 = __byte_perm(U, V, sel0), p1 = __byte_perm(U, V, sel1);
        const uint32_t p2 = __byte_perm(X, Y, sel2), p3 = __byte_perm(X, Y, sel3);
        H[0] = __dp2a_lo(k0, p0, 0u) >> 4; H[1] = __dp2a_hi(k1, p0, 0u) >> 4;
        H[2] = __dp2a_lo(k2, p1, 0u) >> 4; H[3] = __dp2a_hi(k3, p1, 0u) >> 4;
        H[4] = __dp2a_lo(k4, p2, 0u) >> 4; H[5] = __dp2a_hi(k5, p2, 0u) >> 4;
        H[6] = __dp2a_lo(k6, p3, 0u) >> 4; H[7] = __dp2a_hi(k7, p3, 0u) >> 4;
    };
    const int nrow = min(RY, L.rows - row0);
    const uint2* yt = reinterpret_cast<const uint2*>(taps + L.yrow_off + row0);     // (source row, c0 | c1 << 16)
    uint8_t* dst = slab + L.plane_off + (size_t)row0 * L.pitch + 8 + 8 * g;
    uint32_t H0[8], H1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) H0[j] = H1[j] = 0;
    int id0 = -1, id1 = -1;
    uint2 ty = __ldg(yt);
    for (int i = 0; i < nrow; ++i, dst += L.pitch) {
        const uint2 tyn = __ldg(yt + min(i + 1, nrow - 1));                   // one output row ahead
        const int r0 = (int)ty.x, r1 = min(r0 + 1, hs1);
        uint32_t b0, b1;                                                      // row weights of H0, H1, << 16
        if (r0 == id1) {                                                      // the usual step: last row's lower row is the upper one
            if (r1 != r0 && id0 != r1) { pass(r1, H0); id0 = r1; }
            b0 = ty.y & 0xffff0000u; b1 = ty.y << 16;
        } else {
            if (r0 != id0) { pass(r0, H0); id0 = r0; }
            if (r1 != r0 && id1 != r1) { pass(r1, H1); id1 = r1; }
            b0 = ty.y << 16; b1 = ty.y & 0xffff0000u;
        }
        uint32_t sv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sv[j] = __umulhi(b0, H0[j]) + __umulhi(b1, H1[j]);    // <= 1020: weights sum to 2048
        const uint32_t t01 = (sv[0] + (sv[1] << 16) + 0x00020002u) >> 2, t23 = (sv[2] + (sv[3] << 16) + 0x00020002u) >> 2;
        const uint32_t t45 = (sv[4] + (sv[5] << 16) + 0x00020002u) >> 2, t67 = (sv[6] + (sv[7] << 16) + 0x00020002u) >> 2;
        *reinterpret_cast<uint2*>(dst) = make_uint2(__byte_perm(t01, t23, 0x6420), __byte_perm(t45, t67, 0x6420));
        ty = tyn;
    }
}

struct FastStrip {
    int frame, l, ci, cj0;     // NC horizontally adjacent cells starting at column cj0 of cell row ci
};

// Items of one launch: table entries [first, first + spf) of every frame, frame-major.  An entry of the strip table
// (built with the plan) is level | cell row << 4 | first cell << 16.
__device__ __forceinline__ FastStrip fast_decode(const uint32_t* __restrict__ strip_tab, unsigned item, unsigned spf, int first) {
    FastStrip c;
    c.frame = (int)(item / spf);
    const uint32_t e = __ldg(strip_tab + first + (int)(item - (unsigned)c.frame * spf));
    c.l = (int)(e & 15u);
    c.ci = (int)((e >> 4) & 0xfffu);
    c.cj0 = (int)(e >> 16);
    return c;
}

// VIMNMX3.S16x2 issues on the quarter-rate XU pipe on sm_100 (measured: XU 85 % busy with the score network
// written entirely in 3-input min/max), two VIMNMX.S16x2 on the full-rate ALU pipe.  ORBX_MNMX3_MODE picks which
// chains use which: 0 = all 3-input, 1 = all 2-input pairs, 2 = min chain 3-input / max chain pairs.
#ifndef ORBX_MNMX3_MODE
#define ORBX_MNMX3_MODE 2
#endif
__device__ __forceinline__ uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
#if ORBX_MNMX3_MODE == 1
    return __vmins2(__vmins2(a, b), c);
#else
    return __vimin3_s16x2(a, b, c);
#endif
}
__device__ __forceinline__ uint32_t max3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
#if ORBX_MNMX3_MODE == 0
    return __vimax3_s16x2(a, b, c);
#else
    return __vmaxs2(__vmaxs2(a, b), c);
#endif
}

// exact FAST score of the pixel at tile byte p (pitch BW): A - 1 if A > t else 0, with
//   A = max over the 16 arcs of 9 contiguous ring pixels of max(min(ring) - centre, centre - max(ring))
// (cv::FAST's cornerScore<16>: the largest threshold for which the pixel is still a corner, + 1).
// ORBX_SCORE_NET 1 (product): Z[k] = (ring k, 255 - ring k) as s16x2, so ONE min network yields both polarities --
// low half: min over the arc, high half: 255 - max over the arc.  16 + 16 three-input minima (arcs of 3, then of 9) and
// 8 three-input maxima: 40 VIMNMX3 per pixel.  ORBX_SCORE_NET 0 (round 1, kept for A/B): Z[k] = (ring k, ring k + 8),
// a min network and a max network over 8 arcs each (26 + 26 three-input equivalents).
#ifndef ORBX_SCORE_NET
#define ORBX_SCORE_NET 1
#endif
__device__ __forceinline__ int fast_score_packed(const uint8_t* __restrict__ p, int BW, int t) {
    const int v = p[0];
    uint32_t q[16];
    q[0] = p[3 * BW];        q[1] = p[3 * BW + 1];    q[2] = p[2 * BW + 2];    q[3] = p[BW + 3];
    q[4] = p[3];             q[5] = p[-BW + 3];       q[6] = p[-2 * BW + 2];   q[7] = p[-3 * BW + 1];
    q[8] = p[-3 * BW];       q[9] = p[-3 * BW - 1];   q[10] = p[-2 * BW - 2];  q[11] = p[-BW - 3];
    q[12] = p[-3];           q[13] = p[BW - 3];       q[14] = p[2 * BW - 2];   q[15] = p[3 * BW - 1];
#if ORBX_SCORE_NET == 1
    uint32_t Z[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) Z[k] = q[k] * 0xFFFF0001u + 0x00FF0000u;      // q | (255 - q) << 16, one IMAD
    uint32_t n3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) n3[k] = min3_s16x2(Z[k], Z[(k + 1) & 15], Z[(k + 2) & 15]);
    uint32_t n9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) n9[k] = min3_s16x2(n3[k], n3[(k + 3) & 15], n3[(k + 6) & 15]);   // arc k .. k + 8
    uint32_t m0 = max3_s16x2(n9[0], n9[1], n9[2]), m1 = max3_s16x2(n9[3], n9[4], n9[5]);
    uint32_t m2 = max3_s16x2(n9[6], n9[7], n9[8]), m3 = max3_s16x2(n9[9], n9[10], n9[11]);
    uint32_t m4 = max3_s16x2(n9[12], n9[13], n9[14]);
    m0 = max3_s16x2(m0, m1, m2);
    m3 = max3_s16x2(m3, m4, n9[15]);
    const uint32_t M = __vmaxs2(m0, m3);
    const int Ab = (int)(M & 0xffffu) - v;                               // brighter arc: min(ring) - centre
    const int Ad = v + (int)(M >> 16) - 255;                             // darker arc: centre - max(ring)
    const int A = max(Ab, Ad);
    return A > t ? A - 1 : 0;
#else
    // Z[k] = (ring k, ring k + 8) as s16x2; Z[k + 8] = halves swapped.  min / max commute with subtracting the centre,
    // so the network runs on the raw ring values (0 .. 255) and the centre is subtracted once at the end.
    uint32_t Z[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        Z[k] = q[k] + (q[k + 8] << 16);
        Z[k + 8] = q[k + 8] + (q[k] << 16);
    }
    uint32_t n3[14], x3[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) {
        n3[k] = min3_s16x2(Z[k], Z[k + 1], Z[k + 2]);
        x3[k] = max3_s16x2(Z[k], Z[k + 1], Z[k + 2]);
    }
    uint32_t n9[8], x9[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        n9[k] = min3_s16x2(n3[k], n3[k + 3], n3[k + 6]);      // min over ring k .. k+8 (lo) and k+8 .. k+16 (hi)
        x9[k] = max3_s16x2(x3[k], x3[k + 3], x3[k + 6]);
    }
    uint32_t bm = max3_s16x2(n9[0], n9[1], n9[2]);
    bm = max3_s16x2(bm, n9[3], n9[4]);
    bm = max3_s16x2(bm, n9[5], n9[6]);
    bm = __vmaxs2(bm, n9[7]);
    uint32_t dm = min3_s16x2(x9[0], x9[1], x9[2]);
    dm = min3_s16x2(dm, x9[3], x9[4]);
    dm = min3_s16x2(dm, x9[5], x9[6]);
    dm = __vmins2(dm, x9[7]);
    const int Ab = max((int)(bm & 0xffffu), (int)(bm >> 16)) - v;        // brighter arc: min(ring) - centre
    const int Ad = v - min((int)(dm & 0xffffu), (int)(dm >> 16));        // darker arc: centre - max(ring)
    const int A = max(Ab, Ad);
    return A > t ? A - 1 : 0;
#endif
}

#ifndef ORBX_FAST_MINB
#define ORBX_FAST_MINB 3
#endif

// What a warp needs to emit the corners of one cell (all frame-relative pointers already resolved).
struct FastEmit {
    uint32_t* cand;            // this frame's candidate region of the level
    uint2* cell_rec;           // this frame's cell records of the level
    int* level_count;          // candidates of (frame, level) so far
    int* status;               // this frame's status word
    int* retry_count;          // (:812) statistics of (frame, level)
    int cand_cap;
};

// One cell window, the way cv::FAST sees it (:805-816): passes [pass0, 2) with iniThFAST / minThFAST until one leaves a
// keypoint, then NMS and the ordered emission.  `win32` / `sh` address the tile word holding the byte in front of window
// pixel (0, 0), `tile` is the byte of window pixel (0, 0); `sc` is a zero score map addressed as y * SP + x + 1 whose
// entries around the window's scoring region are zero and stay zero; it is all-zero again on return.  `queue` holds at
// least (ww - 6) * (wh - 6) entries.
template <int BW_T>
__device__ __forceinline__ void fast_cell_path(const OrbxPlan* __restrict__ plan, const uint8_t* __restrict__ tile,
                                               const uint32_t* __restrict__ tile32, int sh, int BW, int ww, int wh,
                                               uint8_t* __restrict__ sc, int SP, uint16_t* __restrict__ queue, int pass0,
                                               const FastEmit& em, int cell_index, int ox, int oy, int lane, int qcap,
                                               int tile_bytes_left) {
    // qcap, tile_bytes_left: queue capacity and bytes from `tile` to the end of the tile buffer (bounds-check build only)
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int BW4 = BW >> 2;
    int count = 0, cn = 0;
    ORBX_BC((ww - 6) * (wh - 6) <= qcap || ww < 7 || wh < 7);
    ORBX_BC((wh - 1) * BW + ww + 8 <= tile_bytes_left);                  // last window row + the 2 words of read-ahead
    (void)qcap; (void)tile_bytes_left;
    if (ww >= 7 && wh >= 7) {
        const int ew = ww - 6;                                       // emission width
        const int G = (ew + 3) >> 2;                                 // 4-pixel groups per row
        const int mg = c_recip16[min(G, 32)];                       // (n * mg) >> 16 == n / G for n <= 32
        const int RPI = (32 * mg) >> 16;                             // rows per warp iteration
        const int ry = (lane * mg) >> 16, g = lane - ry * G;
        const int nvalid = min(max(ew - 4 * g, 0), 4);
        const uint32_t vmask = (ry < RPI && nvalid > 0) ? (0x80808080u >> (8 * (4 - nvalid))) : 0u;
        for (int pass = pass0; pass < 2 && count == 0; ++pass) {
            const int t = pass == 0 ? plan->ini_th : plan->min_th;
            if (pass == 1 && lane == 0) atomicAdd(em.retry_count, 1);    // (:812) statistics only
            const uint32_t C = (uint32_t)(0x7f - min(t, 0x7f)) * 0x01010101u;
            int qn = 0;
            {
                // ---- phase 1, in chunks of 8 warp iterations (RPI rows each): first all pre-tests (independent
                //      loads and SIMD math, nothing serialises), then ONE pair of packed warp scans for the chunk
                //      (four 8-bit counts per register), then the ordered queue writes.
                for (int yc = 3; yc < wh - 3; yc += 8 * RPI) {
                    uint32_t m[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int y = yc + k * RPI + ry;
                        // rows past the window are clamped (always inside the tile) and masked out: no divergence
                        const uint32_t* r = tile32 + min(y, wh - 4) * BW4 + g;    // raw word holding tile column 4g + (delta & ~3)
                        const uint32_t c1 = __funnelshift_r(r[1], r[2], sh);      // pixels x .. x+3, x = 3 + 4g
                        const uint32_t c0 = __funnelshift_r(r[0], r[1], sh);
                        const uint32_t c2 = __funnelshift_r(r[2], r[3], sh);
                        const uint32_t up = __funnelshift_r(r[1 - 3 * BW4], r[2 - 3 * BW4], sh);
                        const uint32_t dn = __funnelshift_r(r[1 + 3 * BW4], r[2 + 3 * BW4], sh);
                        const uint32_t r4 = __funnelshift_r(c1, c2, 24);        // pixels x+3 .. x+6
                        const uint32_t r12 = __funnelshift_r(c0, c1, 8);        // pixels x-3 .. x
                        const uint32_t a0 = __vabsdiffu4(dn, c1), a8 = __vabsdiffu4(up, c1);
                        const uint32_t a4 = __vabsdiffu4(r4, c1), a12 = __vabsdiffu4(r12, c1);
                        // bit 7 of a byte of ((a & 0x7f) + C) | a  <=>  a > t
                        const uint32_t s0 = (a0 & 0x7f7f7f7fu) + C, s8 = (a8 & 0x7f7f7f7fu) + C;
                        const uint32_t s4 = (a4 & 0x7f7f7f7fu) + C, s12 = (a12 & 0x7f7f7f7fu) + C;
                        m[k] = ((s0 | a0) | (s8 | a8)) & ((s4 | a4) | (s12 | a12)) & (y < wh - 3 ? vmask : 0u);
                    }
                    // per-lane counts are <= 4 and a warp total is <= 128, so four counts fit one register
                    uint32_t pa = (uint32_t)__popc(m[0]) | ((uint32_t)__popc(m[1]) << 8) | ((uint32_t)__popc(m[2]) << 16) | ((uint32_t)__popc(m[3]) << 24);
                    uint32_t pb = (uint32_t)__popc(m[4]) | ((uint32_t)__popc(m[5]) << 8) | ((uint32_t)__popc(m[6]) << 16) | ((uint32_t)__popc(m[7]) << 24);
                    const uint32_t ca = pa, cb = pb;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t ta = __shfl_up_sync(0xffffffffu, pa, o), tb = __shfl_up_sync(0xffffffffu, pb, o);
                        if (lane >= o) { pa += ta; pb += tb; }
                    }
                    const uint32_t tota = __shfl_sync(0xffffffffu, pa, 31), totb = __shfl_sync(0xffffffffu, pb, 31);
                    pa -= ca;                                                    // exclusive prefixes
                    pb -= cb;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t pk = ((k < 4 ? pa : pb) >> (8 * (k & 3))) & 0xffu;
                        const uint32_t tk = ((k < 4 ? tota : totb) >> (8 * (k & 3))) & 0xffu;
                        if (m[k]) {
                            uint16_t* wq = queue + qn + pk;
                            const int e = ((yc + k * RPI + ry) << 8) | (3 + 4 * g);
                            if (m[k] & 0x80u) *wq++ = (uint16_t)e;
                            if (m[k] & 0x8000u) *wq++ = (uint16_t)(e + 1);
                            if (m[k] & 0x800000u) *wq++ = (uint16_t)(e + 2);
                            if (m[k] & 0x80000000u) *wq++ = (uint16_t)(e + 3);
                        }
                        qn += (int)tk;
                    }
                }
            }
            __syncwarp();
            // ---- phase 2: exact score; corners compacted in place (order kept), scores to the map
            cn = 0;
            for (int i0 = 0; i0 < qn; i0 += 32) {
                const int i = i0 + lane;
                const int e = queue[min(i, qn - 1)];                         // clamped: every lane scores a real pixel
                ORBX_BC((e >> 8) >= 3 && (e >> 8) < wh - 3 && (e & 0xff) >= 3 && (e & 0xff) < ww - 3 && qn <= qcap);
                int s = fast_score_packed(tile + (e >> 8) * BW + (e & 0xff), BW, t);
                if (i >= qn) s = 0;
                const uint32_t bal = __ballot_sync(0xffffffffu, s > 0);
                if (s > 0) {
                    sc[(e >> 8) * SP + (e & 0xff) + 1] = (uint8_t)s;
                    queue[cn + __popc(bal & lt_mask)] = (uint16_t)e;
                }
                cn += __popc(bal);
            }
            __syncwarp();
            // ---- phase 3a: strict 3x3 NMS (the window frame and non-corners score 0); mark + count
            for (int i0 = 0; i0 < cn; i0 += 32) {
                const int i = i0 + lane;
                bool keep = false;
                if (i < cn) {
                    const int e = queue[i];
                    const uint8_t* mp = sc + (e >> 8) * SP + (e & 0xff) + 1;
                    const int s = mp[0];
                    keep = s > mp[-1] && s > mp[1] && s > mp[-SP - 1] && s > mp[-SP] && s > mp[-SP + 1] &&
                           s > mp[SP - 1] && s > mp[SP] && s > mp[SP + 1];
                    if (keep) queue[i] = (uint16_t)(e | 0x8000);
                }
                count += __popc(__ballot_sync(0xffffffffu, keep));
            }
            __syncwarp();
            if (count == 0) {                                                // clear the map before the retry
                for (int i = lane; i < cn; i += 32) sc[(queue[i] >> 8) * SP + (queue[i] & 0xff) + 1] = 0;
                __syncwarp();
                cn = 0;
            }
        }
    }
    // ---- emit: claim a contiguous block of the level's candidate region, write in row-major order
    int gbase = 0;
    bool overflow = false;
    if (count > 0) {
        if (lane == 0) gbase = atomicAdd(em.level_count, count);
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        if (gbase + count > em.cand_cap) {
            if (lane == 0) atomicOr(em.status, ORBX_DEV_CAND_OVERFLOW);
            overflow = true;
        }
    }
    uint32_t* dst = em.cand + gbase;
    int w = 0;
    for (int i0 = 0; i0 < cn; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        int x = 0, y = 0, s = 0;
        if (i < cn) {
            const int e = queue[i];
            keep = (e & 0x8000) != 0;
            y = (e >> 8) & 0x7f;
            x = e & 0xff;
            s = sc[y * SP + x + 1];
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep && !overflow) dst[w + __popc(bal & lt_mask)] = ORBX_PACK(x + ox, y + oy, s);      // (:822-823)
        if (i < cn) sc[y * SP + x + 1] = 0;                               // leave the score map all-zero (NMS is done)
        w += __popc(bal);
    }
    if (lane == 0) em.cell_rec[cell_index] = make_uint2((uint32_t)gbase, overflow ? 0u : (uint32_t)count);
    __syncwarp();
}

// Legacy schedule (ORBX_FAST_LEGACY=1, kept for A/B measurements): one cell at a time, unaligned 4-pixel groups.
// BW_T: tile pitch known at compile time (ring offsets become immediates); 0 = read it from the plan.
template <int BW_T>
__global__ void __launch_bounds__(ORBX_FAST_WARPS * 32, ORBX_FAST_MINB)
fast_cells_kernel(const __grid_constant__ FastMaps maps, const OrbxPlan* __restrict__ plan, const uint32_t* __restrict__ strip_tab,
                  int frame0, int nframes, int first_strip, int nstrips, uint32_t* __restrict__ cand, uint2* __restrict__ cell_rec,
                  int* __restrict__ level_counts, int* __restrict__ work_counter, int* __restrict__ status,
                  int* __restrict__ retry_counts) {
    ORBX_PDL_WAIT();
    extern __shared__ uint8_t fast_smem_raw[];
    __shared__ uint64_t s_bar[ORBX_FAST_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int BW = BW_T ? BW_T : plan->cells_bw, BH = plan->cells_bh;
    const int NB = plan->fast_nb;                                        // tile buffers per warp
    const int TB = (BW * BH + 127) & ~127;                               // tile bytes
    const int QN = (plan->max_cell_w - 6) * (plan->max_cell_h - 6);      // queue entries (u16)
    const int SP = (plan->max_cell_w + 2 + 3) & ~3;                      // score-map pitch; column = window x + 1
    const int SB = (SP * BH + 127) & ~127;
    const int per_warp = NB * TB + SB + ((QN * 2 + 127) & ~127);
    uint8_t* base = fast_smem_raw + ((128 - (smem_u32(fast_smem_raw) & 127)) & 127) + (size_t)warp * per_warp;
    uint8_t* sc = base + NB * TB;                                        // zero-framed score map
    uint16_t* queue = reinterpret_cast<uint16_t*>(base + NB * TB + SB);  // entries (y << 8) | x, window coordinates
    const int nlevels = plan->nlevels;
    const unsigned spf = (unsigned)nstrips;
    const unsigned total = (unsigned)nframes * spf;

    for (int i = lane; i < SB / 4; i += 32) reinterpret_cast<uint32_t*>(sc)[i] = 0;
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    auto fetch = [&]() -> unsigned {
        int v = 0;
        if (lane == 0) v = atomicAdd(work_counter, 1);
        return (unsigned)__shfl_sync(0xffffffffu, v, 0);
    };
    auto issue = [&](const FastStrip& c, int b) {
        if (lane == 0) {
            const OrbxLevel& L = plan->lv[c.l];
            mbar_expect_tx(&s_bar[warp][b], (uint32_t)(BW * BH));
            tma_load_3d(base + b * TB, &maps.m[c.l], &s_bar[warp][b], (ORBX_XO + ORBX_BOX + c.cj0 * L.wCell - 1) & ~15,
                        ORBX_EDGE + ORBX_BOX + c.ci * L.hCell, frame0 + c.frame);
        }
    };

    // A work item is a strip of NC cells fetched as ONE tile (fewer, wider TMA rows).  The work counter is read
    // one item further ahead than the tile prefetch, so the atomic's round trip overlaps a whole strip.
    unsigned cur = fetch();
    unsigned nxt = fetch();
    FastStrip cc, nc;
    if (cur < total) { cc = fast_decode(strip_tab, cur, spf, first_strip); issue(cc, 0); }
    uint32_t phase[2] = {0, 0};
    int b = 0;
    while (cur < total) {
        if (NB == 2 && nxt < total) { nc = fast_decode(strip_tab, nxt, spf, first_strip); issue(nc, b ^ 1); }
        const unsigned nxt2 = nxt < total ? fetch() : nxt;
        mbar_wait(&s_bar[warp][b], phase[b]);
        phase[b] ^= 1;

        const OrbxLevel& L = plan->lv[cc.l];
        const int delta0 = (ORBX_XO + ORBX_BOX + cc.cj0 * L.wCell - 1) & 15;
        const int iniY = ORBX_BOX + cc.ci * L.hCell;
        const int wh = min(iniY + L.hCell + 6, L.maxBY) - iniY;
        const int ncell = min(L.strip_nc, L.nColsV - cc.cj0);
        FastEmit em;
        em.cand = cand + (size_t)cc.frame * plan->cand_per_frame + L.cand_off;
        em.cell_rec = cell_rec + (size_t)cc.frame * plan->cells_per_frame + L.cell_base;
        em.level_count = &level_counts[cc.frame * nlevels + cc.l];
        em.status = &status[cc.frame];
        em.retry_count = &retry_counts[cc.frame * nlevels + cc.l];
        em.cand_cap = L.cand_cap;
        for (int cix = 0; cix < ncell; ++cix) {
            const int cj = cc.cj0 + cix;
            const int iniX = ORBX_BOX + cj * L.wCell;
            const int ww = min(iniX + L.wCell + 6, L.maxBX) - iniX;
            const int delta = delta0 + cix * L.wCell;                        // byte offset of (window x0 - 1) inside the tile
            fast_cell_path<BW_T>(plan, base + b * TB + delta + 1, reinterpret_cast<const uint32_t*>(base + b * TB) + (delta >> 2),
                                 (delta & 3) * 8, BW, ww, wh, sc, SP, queue, 0, em, cc.ci * L.nColsV + cj, cj * L.wCell,
                                 cc.ci * L.hCell, lane, QN, TB - delta - 1);
        }
        if (NB == 1 && nxt < total) { nc = fast_decode(strip_tab, nxt, spf, first_strip); issue(nc, 0); }
        cur = nxt;
        nxt = nxt2;
        cc = nc;
        if (NB == 2) b ^= 1;
    }
}

// =====================================================================================
// fast_strips_kernel -- the product schedule of ComputeKeyPointsOctTree's cell loop (:789-829).
//
// The scoring regions of horizontally adjacent cells tile a cell row without gaps (cv::FAST scores the window minus a
// 3-px frame, and windows overlap by 6), and the iniThFAST pre-test does not depend on the cell.  A CTA therefore takes
// a strip of strip_nc (4) cells as ONE dense region whose 4-pixel words are aligned with the TMA tile.  Everything a
// strip needs that does not depend on the frame comes from a host-built 32-byte record (OrbxStripRec), and every
// shared-memory structure (survivor queue entries aside) is indexed by the TILE BYTE OFFSET of the pixel, so no stage
// converts coordinates:
//   phase 1  warp = group of 8 rows, lane = tile word column: 5 aligned LDS, 2 funnel shifts, 4 VABSDIFF4 and the
//            threshold logic per 4 pixels (every offset is an immediate).  The survivor bits of 8 rows x 4 pixels are
//            packed into one register.
//   phase 1b one warp scan of the per-lane survivor counts + the warp totals, then every lane appends its own
//            survivors to the strip's queue (order is irrelevant here: the emission order comes from a bitmap).
//   phase 2  exact score in full 32-lane batches over the whole strip (warp w takes batches w, w + W, ...); a corner's
//            score goes to the score map and its tile offset is compacted into the queue slots of the warp's own
//            consumed batches.
//   phase 3  strict 3x3 NMS per corner.  cv::FAST's NMS treats everything outside the window's scoring region as score
//            0, so a corner in the first / last scoring column of its cell ignores the neighbours in the adjacent cell.
//            Kept corners set a bit in a bitmap.
//   phase 4  warp = cell, lane = row: extract the row's bits, warp scan, claim the cell's block of the candidate
//            region, write (x, y, score) in row-major order = cv::FAST's keypoint order.
// A cell whose iniThFAST pass leaves no keypoint is redone with minThFAST (:812-816) by a second pass over the same tile
// restricted to those cells' columns (8.8 % of the cells of a 1080p cluttered frame, and they are the flat ones: few
// survivors).  Strips with more survivors than the queue holds (noise) and levels whose cells are larger than 32 px go
// through fast_cell_path.
// =====================================================================================
#ifndef ORBX_FS_MINB
#define ORBX_FS_MINB (32 / ORBX_FS_WARPS)
#endif
#define ORBX_FS_CPW (4 / ORBX_FS_WARPS)      // cells per warp (phase 4)
#ifndef ORBX_FS_P2U
#define ORBX_FS_P2U 1                        // score batches per iteration of phase 2 (1 or 2; 2 measured no faster: 0.801 vs 0.806 ms)
#endif

template <int BW_T>
__device__ __forceinline__ uint32_t fast_pretest_word(const uint32_t* __restrict__ p, uint32_t C, uint32_t colmask) {
    constexpr int BW4 = BW_T / 4;
    const uint32_t c1 = p[0], c0 = p[-1], c2 = p[1], up = p[-3 * BW4], dn = p[3 * BW4];
    const uint32_t r4 = __funnelshift_r(c1, c2, 24);        // pixels x+3 .. x+6
    const uint32_t r12 = __funnelshift_r(c0, c1, 8);        // pixels x-3 .. x
    const uint32_t a0 = __vabsdiffu4(dn, c1), a8 = __vabsdiffu4(up, c1);
    const uint32_t a4 = __vabsdiffu4(r4, c1), a12 = __vabsdiffu4(r12, c1);
    // bit 7 of a byte of ((a & 0x7f) + C) | a  <=>  a > t
    const uint32_t s0 = (a0 & 0x7f7f7f7fu) + C, s8 = (a8 & 0x7f7f7f7fu) + C;
    const uint32_t s4 = (a4 & 0x7f7f7f7fu) + C, s12 = (a12 & 0x7f7f7f7fu) + C;
    // (a third and fourth opposite pair -- the diagonals 2/10 and 6/14, tested as (a | a') > t with the rows reused across
    // the 8-row group -- drops the survivors of the bench scene from 15.9 % to 11.7 % of the pixels but costs as much as it
    // saves: FAST 0.774 -> 0.763 ms per 64 x 1080p, the whole step unchanged at 1.372 ms; not kept)
    return ((s0 | a0) | (s8 | a8)) & ((s4 | a4) | (s12 | a12)) & colmask;
}

#define ORBX_FS_NONE 0xffffffffu

// NG: 8-row groups of a strip.  4 = cells of <= 32 scoring rows (nearly every level); 5 = cells of 33 .. 40 rows (the
// "tall" cells of small levels: hCell = ceil(H / floor(H / 30)) reaches 40 when a level has 3 cell rows), own launch.
template <int BW_T, int NG>
__global__ void __launch_bounds__(ORBX_FS_WARPS * 32, ORBX_FS_MINB)
fast_strips_kernel(const __grid_constant__ FastMaps maps, const OrbxPlan* __restrict__ plan, const uint4* __restrict__ strip_rec,
                   int frame0, int nframes, int first_strip, int nstrips, uint32_t* __restrict__ cand, uint2* __restrict__ cell_rec,
                   int* __restrict__ level_counts, int* __restrict__ work_counter, int* __restrict__ status,
                   int* __restrict__ retry_counts) {
    ORBX_PDL_WAIT();
    // Everything a strip needs lives at fixed shared-memory addresses (sizes are compile-time: the plan only sends levels
    // with cells <= 32 x 32 here), so every LDS / STS below takes an immediate offset.
    constexpr int BW = BW_T, BW4 = BW_T / 4;
    constexpr int BH = 8 * NG + 6;                                       // tile rows: the scoring rows + the 6-px frame
    constexpr int GPW = (NG + ORBX_FS_WARPS - 1) / ORBX_FS_WARPS;        // 8-row groups per warp (warp w takes groups w, w + W, ...)
    constexpr int NR = (8 * NG + 31) / 32;                               // phase 4: rounds of 32 scoring rows
    constexpr int TB = (BW * BH + 127) & ~127;                           // tile bytes
    constexpr int MOFF = 2 * BW;                                         // score map / bitmap index = tile byte offset - MOFF
    constexpr int SBYTES = (BH - 4) * BW;                                // map rows = tile rows 2 .. BH - 3
    constexpr int KBW = SBYTES / 32 + 1;                                 // kept-corner bitmap (+ 1: funnel read-ahead)
    constexpr int QCAP = ORBX_FS_QCAP > 256 * NG ? ORBX_FS_QCAP : 256 * NG;   // one cell (32 x 8 NG pixels) always fits
    constexpr int NT = ORBX_FS_WARPS * 32;
    static_assert(BW_T % 32 == 0, "bitmap rows are whole words");
    __shared__ __align__(128) uint8_t s_tile[ORBX_FS_NBUF * TB];
    __shared__ __align__(16) uint8_t sc[SBYTES];
    __shared__ uint32_t kb[KBW];
    __shared__ __align__(16) uint16_t queue[QCAP];
    __shared__ uint64_t s_bar[2];
    __shared__ int s_wtot[ORBX_FS_WARPS];
    __shared__ __align__(16) uint4 s_item[3][2];     // ring of strip records: this strip, the next one, the one being fetched
    __shared__ unsigned s_redo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nlevels = plan->nlevels;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // thread 0: next work item -> its record (b.z = frame or ORBX_FS_NONE)
    auto fetch = [&](int slot) {
        const unsigned spf = (unsigned)nstrips;
        const unsigned v = (unsigned)atomicAdd(work_counter, 1);
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, ORBX_FS_NONE, 0u);
        if (v < (unsigned)nframes * spf) {
            const unsigned f = v / spf;
            const uint4* r = strip_rec + 2 * (size_t)(first_strip + (int)(v - f * spf));
            a = __ldg(r);
            b = __ldg(r + 1);
            b.z = f;
        }
        s_item[slot][0] = a;
        s_item[slot][1] = b;
    };
    auto issue = [&](int slot, int bi) {                 // thread 0 only (reads its own earlier writes)
        const uint4 a = s_item[slot][0];
        const unsigned f = s_item[slot][1].z;
        if (f == ORBX_FS_NONE) return;
        mbar_expect_tx(&s_bar[bi], (uint32_t)(BW * BH));
        tma_load_3d(s_tile + bi * TB, &maps.m[a.y & 15u], &s_bar[bi], (int)(a.x & 0xffffu), (int)(a.x >> 16), frame0 + (int)f);
    };
    // corner j of this warp lives in the queue slots of the warp's own consumed batches
    auto cslot = [&](int j) { return (warp + ORBX_FS_WARPS * (j >> 5)) * 32 + (j & 31); };

    for (int i = tid; i < SBYTES / 4; i += NT) reinterpret_cast<uint32_t*>(sc)[i] = 0;
    for (int i = tid; i < KBW; i += NT) kb[i] = 0;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fetch(0);
        fetch(1);
        s_redo = 0;
        issue(0, 0);
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    int b = 0, slot = 0;
    for (;;) {
        const uint4 ra = s_item[slot][0], rb = s_item[slot][1];
        if (rb.z == ORBX_FS_NONE) break;                                     // uniform
        const int slot1 = slot == 2 ? 0 : slot + 1, slot2 = slot1 == 2 ? 0 : slot1 + 1;
        if (tid == 0) {
            if (ORBX_FS_NBUF == 2) issue(slot1, b ^ 1);                      // buffer b ^ 1 was released by the last barrier of the previous strip
            fetch(slot2);                                                    // slot2 held the previous strip: every thread is past reading it
        }
        const int frame = (int)rb.z;
        const int lvl = ra.y & 15, ncell = (ra.y >> 4) & 15, hr = (ra.y >> 8) & 0xff, delta0 = (ra.y >> 16) & 0xff, W0 = ra.y >> 24;
        const int lo = ra.z & 0xffff, hi = ra.z >> 16;                       // tile byte columns of the strip's scoring pixels: [lo, hi)
        const int wCell = ra.w & 0xffff, wrecip = ra.w >> 16;
        mbar_wait(&s_bar[b], phase[b]);
        phase[b] ^= 1;

        uint8_t* const tbuf = s_tile + b * TB;
        ORBX_BC(hr <= 8 * NG && ncell >= 1 && ncell <= 4 && lo == delta0 + 4 && hi <= BW - 4 && W0 == (lo >> 2));
        // (:805-816) iniThFAST for every cell, then minThFAST for the cells that came back empty.  A task = one threshold
        // applied to a set of cells (all of them unless their survivors exceed the queue: then the set is halved and the
        // pre-test redone for fewer columns -- noise only; one cell always fits: 32 x 40 <= QCAP).
        unsigned pend_ini = (1u << ncell) - 1u, pend_min = 0;                // cells waiting for their iniThFAST / minThFAST pass
        int limit = 4;                                                       // cells per task
        static_assert(QCAP >= 32 * 8 * NG, "a single cell's survivors must fit the queue");
        if (hr >= 1) {                                                       // (the record says hr = 0 for windows smaller than 7 x 7)
            const int bcol0 = 4 * (W0 + lane);                               // tile byte column of byte 0 of this lane's word
            const uint32_t* rp0 = reinterpret_cast<const uint32_t*>(tbuf) + 3 * BW4 + min(W0 + lane, BW4 - 2);
            while ((pend_ini | pend_min) != 0) {                             // uniform
                const int pass = pend_ini != 0 ? 0 : 1;
                unsigned todo = pass == 0 ? pend_ini : pend_min;
                while (__popc(todo) > limit) todo &= ~(0x80000000u >> __clz(todo));      // keep the lowest `limit` cells
                const int t = pass == 0 ? plan->ini_th : plan->min_th;
                const uint32_t C = (uint32_t)(0x7f - min(t, 0x7f)) * 0x01010101u;
                // ---- phase 1: aligned SIMD pre-test; a warp takes ORBX_FS_GPW groups of 8 rows, lane = word column
                uint32_t colmask = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int bc = bcol0 + j;
                    if (bc >= lo && bc < hi && ((todo >> (((bc - lo) * wrecip) >> 16)) & 1u)) colmask |= 0xffu << (8 * j);
                }
                uint32_t acc[GPW];
                int cnt = 0;
#pragma unroll
                for (int gg = 0; gg < GPW; ++gg) {
                    const int g = warp + ORBX_FS_WARPS * gg;
                    const int nv = hr - 8 * g;                               // rows of this group
                    uint32_t a = 0;
                    if (nv > 0) {
                        const uint32_t* rp = rp0 + 8 * g * BW4;
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            // rows past the window (last group of a short strip) read bytes that exist in shared memory and are masked below
                            uint32_t m = fast_pretest_word<BW_T>(rp + r * BW4, C, colmask);
                            // row r of the group goes to the lane 4r further down: a lane collects 8 different word columns, so
                            // an edge (the typical run of survivors) is spread over many lanes before the per-lane append loop
                            if (r) m = __shfl_sync(0xffffffffu, m, (lane + 4 * r) & 31);
                            a |= (r == 7 ? m : (m >> (7 - r))) & (0x01010101u << r);
                        }
                        if (nv < 8) a &= ((1u << nv) - 1u) * 0x01010101u;   // byte j = pixel column j, bit r = row 8g + r
                    }
                    acc[gg] = a;
                    cnt += __popc(a);
                }
                // ---- phase 1b: scan, then every lane appends its survivors (any order)
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) s_wtot[warp] = incl;
                __syncthreads();                                             // B1
                int wbase = 0, qn = 0;
#pragma unroll
                for (int w = 0; w < ORBX_FS_WARPS; ++w) {
                    const int v = s_wtot[w];
                    if (w < warp) wbase += v;
                    qn += v;
                }
                if (qn > QCAP) {                                             // noise (uniform branch): fewer cells per task
                    ORBX_BC(limit > 1);
                    limit >>= 1;
                    __syncthreads();                                         // s_wtot is rewritten by the next task
                    continue;
                }
                {
                    uint16_t* wq = queue + wbase + (incl - cnt);
#pragma unroll
                    for (int gg = 0; gg < GPW; ++gg) {
                        const int e0 = ((warp + ORBX_FS_WARPS * gg) << 10) | (lane << 5);  // entry: row group | lane | bit of acc
                        uint32_t a = acc[gg];
                        while (a) {
                            const int bit = __ffs((int)a) - 1;
                            a &= a - 1;
                            ORBX_BC(wq >= queue && wq < queue + QCAP);
                            *wq++ = (uint16_t)(e0 | bit);
                        }
                    }
                }
                __syncthreads();                                             // B2
                // ---- phase 2: exact score; corners to the score map, their tile offsets compacted into this warp's batches
                int cn = 0;                                                  // corners of this warp
                const int off0 = 3 * BW + 4 * W0;                            // tile byte of (scoring row 0, this strip's word column 0, byte 0)
                auto entry_off = [&](int i) {                                // queue entry -> tile byte offset
                    const int e = queue[min(i, qn - 1)];                     // clamped: every lane scores a real pixel
                    const int r = e & 7;
                    const int wc = ((e >> 5) + 4 * r) & 31;                  // word column (undo the lane rotation of phase 1)
                    const int off = off0 + (((e >> 7) & 0x38) + r) * BW + 4 * wc + ((e >> 3) & 3);
                    ORBX_BC(off >= 3 * BW + lo && off < (3 + hr) * BW && off % BW >= lo && off % BW < hi);
                    return off;
                };
                auto put_corners = [&](int off, int s, int i0) {             // i0: first queue index of the batch just consumed
                    const uint32_t bal = __ballot_sync(0xffffffffu, s > 0);
                    if (s > 0) {
                        sc[off - MOFF] = (uint8_t)s;
                        const int j = cn + __popc(bal & lt_mask);
                        ORBX_BC(cslot(j) < i0 + 32);                         // in place: never ahead of the read position
                        (void)i0;
                        queue[cslot(j)] = (uint16_t)off;
                    }
                    cn += __popc(bal);
                };
                int i0 = warp * 32;
#if ORBX_FS_P2U == 2
                // two batches per iteration: their queue reads, ring loads and score networks are independent chains
                for (; i0 + NT < qn; i0 += 2 * NT) {
                    const int ia = i0 + lane, ib = ia + NT;
                    const int offa = entry_off(ia), offb = entry_off(ib);
                    int sa = fast_score_packed(tbuf + offa, BW, t);
                    int sb = fast_score_packed(tbuf + offb, BW, t);
                    if (ib >= qn) sb = 0;
                    put_corners(offa, sa, i0);
                    put_corners(offb, sb, i0 + NT);
                }
#endif
                for (; i0 < qn; i0 += NT) {
                    const int i = i0 + lane;
                    const int off = entry_off(i);
                    int s = fast_score_packed(tbuf + off, BW, t);
                    if (i >= qn) s = 0;
                    put_corners(off, s, i0);
                }
                __syncthreads();                                             // B3
                // ---- phase 3: strict 3x3 NMS inside the corner's own cell; kept corners -> bitmap
                for (int j0 = 0; j0 < cn; j0 += 32) {
                    const int j = j0 + lane;
                    if (j < cn) {
                        const int off = queue[cslot(j)];
                        const int y = off / BW;
                        const int c = off - y * BW - lo;                     // scoring column of the strip
                        const int xin = c - ((c * wrecip) >> 16) * wCell;    // ... of the cell
                        ORBX_BC(c >= 0 && xin >= 0 && xin < wCell && off - MOFF > BW && off - MOFF < SBYTES - BW - 1);
                        const uint8_t* mp = sc + (off - MOFF);
                        int m8 = max((int)mp[-BW], (int)mp[BW]);
                        if (xin != 0) m8 = max(m8, max(max((int)mp[-BW - 1], (int)mp[-1]), (int)mp[BW - 1]));
                        if (xin != wCell - 1) m8 = max(m8, max(max((int)mp[-BW + 1], (int)mp[1]), (int)mp[BW + 1]));
                        if ((int)mp[0] > m8) atomicOr(&kb[(off - MOFF) >> 5], 1u << (off & 31));
                    }
                }
                __syncthreads();                                             // B4
                // ---- phase 4: warp = cell, lane = scoring row; ordered emission
                const OrbxLevel& L = plan->lv[lvl];
#pragma unroll
                for (int kk = 0; kk < ORBX_FS_CPW; ++kk)
                if ((todo >> (warp + kk * ORBX_FS_WARPS)) & 1u) {
                    const int k = warp + kk * ORBX_FS_WARPS;
                    const int x0 = lo + k * wCell;                           // tile byte column of the cell's first scoring column
                    const uint32_t cmask = wCell >= 32 ? 0xffffffffu : ((1u << wCell) - 1u);
                    uint32_t bits[NR];
                    int excl[NR];
                    int tot = 0;
#pragma unroll
                    for (int q = 0; q < NR; ++q) {                           // scoring row 32 q + lane
                        const int mrow = (32 * q + lane + 1) * BW + x0;      // map index of (that row, x0)
                        bits[q] = 0;
                        if (32 * q + lane < hr) {
                            ORBX_BC((mrow >> 5) + 1 < KBW && mrow + wCell <= SBYTES);
                            bits[q] = __funnelshift_r(kb[mrow >> 5], kb[(mrow >> 5) + 1], x0 & 31) & cmask;
                        }
                        const int c = __popc(bits[q]);
                        int inc = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += v;
                        }
                        excl[q] = tot + inc - c;
                        tot += __shfl_sync(0xffffffffu, inc, 31);
                    }
                    if (pass == 1 && lane == 0) atomicAdd(&retry_counts[frame * nlevels + lvl], 1);    // (:812) statistics only
                    uint2* rec = cell_rec + (size_t)frame * plan->cells_per_frame + rb.x + k;
                    if (tot == 0) {
                        if (lane == 0) {
                            if (pass == 0) atomicOr(&s_redo, 1u << k);
                            else *rec = make_uint2(0u, 0u);
                        }
                    } else {
                        int gbase = 0;
                        if (lane == 0) gbase = atomicAdd(&level_counts[frame * nlevels + lvl], tot);
                        gbase = __shfl_sync(0xffffffffu, gbase, 0);
                        const bool overflow = gbase + tot > L.cand_cap;
                        if (overflow && lane == 0) atomicOr(&status[frame], ORBX_DEV_CAND_OVERFLOW);
                        uint32_t* const dst0 = cand + (size_t)frame * plan->cand_per_frame + L.cand_off + gbase;
                        const int ox = (int)(rb.y & 0xffffu) + k * wCell;    // (:822-823)
#pragma unroll
                        for (int q = 0; q < NR; ++q) {
                            uint32_t* dst = dst0 + excl[q];
                            const uint8_t* srow = sc + (32 * q + lane + 1) * BW + x0;
                            const int oy = (int)(rb.y >> 16) + 32 * q + lane;
                            uint32_t bb = overflow ? 0u : bits[q];
                            while (bb) {
                                const int x = __ffs((int)bb) - 1;
                                bb &= bb - 1;
                                *dst++ = ORBX_PACK(x + ox, oy, srow[x]);
                            }
                        }
                        if (lane == 0) *rec = make_uint2((uint32_t)gbase, overflow ? 0u : (uint32_t)tot);
                    }
                }
                __syncthreads();                                             // B5
                // leave the score map and the bitmap all-zero
                for (int j = lane; j < cn; j += 32) sc[queue[cslot(j)] - MOFF] = 0;
                for (int i = tid; i < KBW; i += NT) kb[i] = 0;
                if (pass == 0) {
                    pend_ini &= ~todo;
                    pend_min |= s_redo;                                      // (cumulative over this strip's iniThFAST tasks)
                } else {
                    pend_min &= ~todo;
                }
            }
        } else if (tid < ncell) {
            cell_rec[(size_t)frame * plan->cells_per_frame + rb.x + tid] = make_uint2(0u, 0u);     // cv::FAST returns nothing for such a window
        }
        __syncthreads();                                                     // B6: map / bitmap zero, s_redo consumed, tile buffer b free
        if (tid == 0) s_redo = 0;
        if (ORBX_FS_NBUF == 1) {
            if (tid == 0) issue(slot1, 0);                                   // every warp is past its last read of the tile (B3 / B7)
        } else {
            b ^= 1;
        }
        slot = slot1;
    }
}

// =====================================================================================
// DistributeOctTree (:539-763) + DivideNode (:481-537), one CTA per (frame, level).
//
// The reference keeps a std::list of nodes; every new child is push_front'ed, so the list
// is always in descending creation order and the final output order is the list order.
// Each pass expands a set of nodes in a processing order:
//   phase A (:606-665)  every multi-key node, in list order;
//   phase B (:676-737)  multi-key nodes sorted by (key count, creation sequence)
//                       descending (canonical rule B-1), stopping right after the
//                       expansion that brings the list to >= N nodes (:730).
// With the list stored as an array in list order, "creation sequence descending" is
// "array index ascending", the children of the r-th processed node land in front of
// those of the (r-1)-th, and survivors keep their relative order behind all children:
// every pass is counting (smem atomics per key), one sort (phase B), prefix sums and a
// scatter.  Keys stay in HBM/L2 as packed (x, y, response) plus a 16-bit node id.
// =====================================================================================
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* s_warp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int res = inc - v + s_warp[wid];
    *total = s_warp[32];
    __syncthreads();
    return res;
}

struct Rect4 {
    short ulx, urx, uly, bry;
};

__global__ void __launch_bounds__(ORBX_OT_THREADS, 1)
octree_kernel(const OrbxPlan* __restrict__ plan, int nframes, const uint32_t* __restrict__ cand,
              const uint2* __restrict__ cell_rec, uint32_t* __restrict__ cand_sorted,
              uint16_t* __restrict__ key_node, int* __restrict__ sorted_counts, uint32_t* __restrict__ kept,
              int* __restrict__ kept_counts, int* __restrict__ status) {
    ORBX_PDL_WAIT();
    extern __shared__ __align__(16) unsigned char ot_smem[];
    __shared__ int s_warp[33];
    __shared__ int s_misc[4];
    const int tid = threadIdx.x;
    const int T = (int)blockDim.x;             // 512, or 1024 for very large levels (chosen at launch)
    const int l = blockIdx.x / nframes, frame = blockIdx.x - l * nframes;     // big levels first
    const OrbxLevel& L = plan->lv[l];
    const int CAP = plan->node_cap;
    int SORTN = 1;
    while (SORTN < CAP) SORTN <<= 1;
    // ---- shared-memory carve-up
    unsigned char* sp = ot_smem;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)SORTN * 8;
    Rect4* rect0 = reinterpret_cast<Rect4*>(sp); sp += (size_t)CAP * 8;
    Rect4* rect1 = reinterpret_cast<Rect4*>(sp); sp += (size_t)CAP * 8;
    int* cnt0 = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 4;
    int* cnt1 = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 4;
    int* cc = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 16;          // child key counts [node][4]
    int* cp = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 16;          // child positions  [node][4]
    int* procrank = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 4;     // node -> processing rank or -1
    int* proc = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 4;         // rank -> node
    int* ne_incl = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 4;      // inclusive sum of non-empty children by rank
    int* surv_pos = reinterpret_cast<int*>(sp); sp += (size_t)CAP * 4;     // node -> new index if it survives
    uint32_t* keysS = reinterpret_cast<uint32_t*>(sp); sp += (size_t)ORBX_OT_KEYCAP * 4;
    uint16_t* knodeS = reinterpret_cast<uint16_t*>(sp); sp += (size_t)ORBX_OT_KEYCAP * 2;

    const size_t cbase = (size_t)frame * plan->cand_per_frame + L.cand_off;
    const uint32_t* in = cand + cbase;
    uint32_t* keys = cand_sorted + cbase;
    uint16_t* knode = key_node + cbase;

    // ---- prologue: put the per-cell blocks into the reference's cell order (:789-826)
    const int ncell = L.nRowsV * L.nColsV;
    const uint2* rec = cell_rec + (size_t)frame * plan->cells_per_frame + L.cell_base;
    int M = 0;
    for (int b0 = 0; b0 < ncell; b0 += T) {
        const int c = b0 + tid;
        const uint2 r = c < ncell ? rec[c] : make_uint2(0, 0);
        int tot;
        const int e = block_excl_scan((int)r.y, &tot, s_warp);
        for (uint32_t i = 0; i < r.y; ++i) {
            const uint32_t v = in[r.x + i];
            keys[M + e + i] = v;                                               // reference push order (stage dump, final pick)
            if (M + e + (int)i < ORBX_OT_KEYCAP) keysS[M + e + i] = v;
        }
        M += tot;
    }
    if (tid == 0) sorted_counts[frame * plan->nlevels + l] = M;
    __syncthreads();

    // Keys (packed x, y, response) and their node ids are swept twice per pass.  When a level's candidates fit
    // (<= ORBX_OT_KEYCAP) they live in shared memory for the whole kernel; otherwise the sweeps go to HBM/L2.
    auto run = [&](const uint32_t* K, uint16_t* KN) {
    // ---- roots (:543-585)
    const int N = L.quota;
    const int nIni = L.nIni;
    const float hX = L.hX;
    const int boxH = L.maxBY - ORBX_BOX;
    int* root_map = cp;
    for (int i = tid; i < nIni; i += T) {
        Rect4 r;
        r.ulx = (short)(int)__fmul_rn(hX, (float)i);
        r.urx = (short)(int)__fmul_rn(hX, (float)(i + 1));
        r.uly = 0;
        r.bry = (short)boxH;
        rect0[i] = r;
        cnt0[i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < M; k += T) {
        int r = (int)__fdiv_rn((float)ORBX_PX(K[k]), hX);                 // (:569) truncation
        r = min(r, nIni - 1);
        KN[k] = (uint16_t)r;
        atomicAdd(&cnt0[r], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int i = 0; i < nIni; ++i) {
            if (cnt0[i] > 0) {
                root_map[i] = n;
                rect1[n] = rect0[i];
                cnt1[n] = cnt0[i];
                ++n;
            } else {
                root_map[i] = -1;
            }
        }
        s_misc[0] = n;
    }
    __syncthreads();
    int n_nodes = s_misc[0];
    for (int k = tid; k < M; k += T) KN[k] = (uint16_t)root_map[KN[k]];
    __syncthreads();

    Rect4* rc = rect1; Rect4* rn = rect0;
    int* cntc = cnt1; int* cntn = cnt0;
    bool phaseB = false;
    bool failed = false;

    while (true) {
        const int S0 = n_nodes;
        for (int i = tid; i < S0 * 4; i += T) cc[i] = 0;
        if (tid == 0) { s_misc[1] = 0; s_misc[2] = 0x7fffffff; }
        __syncthreads();
        // ---- count keys per child quadrant (DivideNode :510-526)
        for (int k = tid; k < M; k += T) {
            const int nd = KN[k];
            if (cntc[nd] > 1) {
                const Rect4 r = rc[nd];
                const int mx = r.ulx + ((r.urx - r.ulx + 1) >> 1);       // ceil(float(d)/2) (:483-484)
                const int my = r.uly + ((r.bry - r.uly + 1) >> 1);
                const uint32_t p = K[k];
                const int q = (ORBX_PX(p) < mx ? 0 : 1) + (ORBX_PY(p) < my ? 0 : 2);
                atomicAdd(&cc[nd * 4 + q], 1);
            }
        }
        __syncthreads();
        // ---- processing order
        int E = 0;
        if (!phaseB) {
            for (int b0 = 0; b0 < S0; b0 += T) {
                const int i = b0 + tid;
                const int f = (i < S0 && cntc[i] > 1) ? 1 : 0;
                int tot;
                const int e = block_excl_scan(f, &tot, s_warp);
                if (i < S0) procrank[i] = f ? E + e : -1;
                if (f) proc[E + e] = i;
                E += tot;
            }
        } else {
            // order by (key count, creation sequence) descending = (count desc, list index asc): rule B-1.  Keys are
            // distinct, so a node's processing rank is the number of larger keys -- an all-pairs count over <= node_cap
            // shared-memory keys (broadcast reads, two barriers) instead of a 45-barrier bitonic sort.
            for (int i = tid; i < S0; i += T) {
                unsigned long long key = 0;
                if (cntc[i] > 1) key = ((unsigned long long)(uint32_t)cntc[i] << 32) | (uint32_t)(0xffffffffu - (uint32_t)i);
                skey[i] = key;
                procrank[i] = -1;
            }
            __syncthreads();
            int e_local = 0;
            for (int i = tid; i < S0; i += T) {
                const unsigned long long key = skey[i];
                if (key != 0) {
                    int r = 0;
                    for (int j = 0; j < S0; ++j) r += skey[j] > key;
                    proc[r] = i;
                    procrank[i] = r;
                    ++e_local;
                }
            }
            atomicAdd(&s_misc[1], e_local);
            __syncthreads();
            E = s_misc[1];
            __syncthreads();
            if (tid == 0) s_misc[1] = 0;
        }
        __syncthreads();
        // ---- list size after each expansion; phase B stops at the first size >= N (:730)
        {
            int carry = 0;
            for (int b0 = 0; b0 < E; b0 += T) {
                const int r = b0 + tid;
                int ne = 0;
                if (r < E) {
                    const int i = proc[r];
                    ne = (cc[i * 4] > 0) + (cc[i * 4 + 1] > 0) + (cc[i * 4 + 2] > 0) + (cc[i * 4 + 3] > 0);
                }
                int tot;
                const int e = block_excl_scan(ne, &tot, s_warp);
                if (r < E) {
                    const int incl = carry + e + ne;
                    ne_incl[r] = incl;
                    if (phaseB && S0 + incl - (r + 1) >= N) atomicMin(&s_misc[2], r + 1);
                }
                carry += tot;
            }
        }
        __syncthreads();
        const int P = min(E, s_misc[2]);
        const int C = P > 0 ? ne_incl[P - 1] : 0;
        // ---- survivors keep their order behind all new children
        int S1 = C;
        for (int b0 = 0; b0 < S0; b0 += T) {
            const int i = b0 + tid;
            int f = 0;
            if (i < S0) {
                const int pr = procrank[i];
                f = (pr >= 0 && pr < P) ? 0 : 1;
            }
            int tot;
            const int e = block_excl_scan(f, &tot, s_warp);
            if (i < S0) surv_pos[i] = f ? S1 + e : -1;
            S1 += tot;
        }
        if (S1 > CAP) { failed = true; break; }
        // ---- children: the r-th processed node's block sits in front of the (r-1)-th's; inside a
        //      block the order is n4, n3, n2, n1 (push_front of n1..n4, :621-660)
        int nexp = 0;
        for (int r = tid; r < P; r += T) {
            const int i = proc[r];
            const Rect4 pr = rc[i];
            const int mx = pr.ulx + ((pr.urx - pr.ulx + 1) >> 1);
            const int my = pr.uly + ((pr.bry - pr.uly + 1) >> 1);
            int pos = C - ne_incl[r];
#pragma unroll
            for (int q = 3; q >= 0; --q) {
                const int kc = cc[i * 4 + q];
                if (kc > 0) {
                    Rect4 ch;
                    ch.ulx = (short)((q & 1) ? mx : pr.ulx);
                    ch.urx = (short)((q & 1) ? pr.urx : mx);
                    ch.uly = (short)((q & 2) ? my : pr.uly);
                    ch.bry = (short)((q & 2) ? pr.bry : my);
                    rn[pos] = ch;
                    cntn[pos] = kc;
                    cp[i * 4 + q] = pos;
                    nexp += kc > 1;
                    ++pos;
                } else {
                    cp[i * 4 + q] = -1;
                }
            }
        }
        if (nexp) atomicAdd(&s_misc[1], nexp);
        for (int i = tid; i < S0; i += T) {
            const int sp2 = surv_pos[i];
            if (sp2 >= 0) { rn[sp2] = rc[i]; cntn[sp2] = cntc[i]; }
        }
        __syncthreads();
        // ---- move keys to their new nodes
        for (int k = tid; k < M; k += T) {
            const int nd = KN[k];
            const int pr = procrank[nd];
            if (pr >= 0 && pr < P) {
                const Rect4 r = rc[nd];
                const int mx = r.ulx + ((r.urx - r.ulx + 1) >> 1);
                const int my = r.uly + ((r.bry - r.uly + 1) >> 1);
                const uint32_t p = K[k];
                const int q = (ORBX_PX(p) < mx ? 0 : 1) + (ORBX_PY(p) < my ? 0 : 2);
                KN[k] = (uint16_t)cp[nd * 4 + q];
            } else {
                KN[k] = (uint16_t)surv_pos[nd];
            }
        }
        const int nToExpand = s_misc[1];
        __syncthreads();
        { Rect4* t = rc; rc = rn; rn = t; }
        { int* t = cntc; cntc = cntn; cntn = t; }
        n_nodes = S1;
        if (S1 >= N || S1 == S0) break;                                   // (:669, :734)
        if (!phaseB && S1 + 3 * nToExpand > N) phaseB = true;             // (:673)
    }

    // ---- best key per node: max response, first in list order wins (:741-760)
    int* best = cc;
    if (failed) {
        if (tid == 0) { atomicOr(&status[frame], ORBX_DEV_NODE_OVERFLOW); kept_counts[frame * plan->nlevels + l] = 0; }
        return;
    }
    for (int i = tid; i < n_nodes; i += T) best[i] = 0;
    __syncthreads();
    for (int k = tid; k < M; k += T)
        atomicMax(reinterpret_cast<unsigned int*>(&best[KN[k]]),
                  ((uint32_t)ORBX_PR(K[k]) << 24) | (0xffffffu - (uint32_t)k));
    __syncthreads();
    const int out_n = min(n_nodes, L.kept_cap);
    uint32_t* out = kept + (size_t)frame * plan->kept_per_frame + L.kept_off;
    for (int i = tid; i < out_n; i += T) {
        const uint32_t k = 0xffffffu - ((uint32_t)best[i] & 0xffffffu);
        const uint32_t p = K[k];
        out[i] = ORBX_PACK(ORBX_PX(p) + ORBX_BOX, ORBX_PY(p) + ORBX_BOX, ORBX_PR(p));      // (:843-844)
    }
    if (tid == 0) {
        kept_counts[frame * plan->nlevels + l] = out_n;
        if (n_nodes > L.kept_cap) atomicOr(&status[frame], ORBX_DEV_NODE_OVERFLOW);
    }
    };
    if (M <= ORBX_OT_KEYCAP) run(keysS, knodeS);
    else run(keys, knode);
}

// =====================================================================================
// cv::fastAtan2's polynomial without FMA (App. A-4), used by describe_kernel for IC_Angle (:103).
// =====================================================================================
__device__ __forceinline__ float fast_atan2_deg(float y, float x, const OrbxPlan* plan) {
    const float eps = 2.2204460492503131e-16f;                    // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(plan->atan_p7, c2), plan->atan_p5), c2), plan->atan_p3), c2), plan->atan_p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(plan->atan_p7, c2), plan->atan_p5), c2), plan->atan_p3), c2), plan->atan_p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ int level_of_slot(const OrbxPlan* plan, int s) {
    int l = 0;
    while (l + 1 < plan->nlevels && s >= plan->lv[l + 1].kept_off) ++l;
    return l;
}

// =====================================================================================
// describe_kernel: everything the reference does per kept keypoint, one warp per keypoint:
//   IC_Angle (:77-104) on the UN-blurred level, GaussianBlur 7x7 sigma 2 (:1086) of the
//   37 x 37 patch the rotated pattern can reach (|coordinate| <= 18) instead of
//   the whole level, computeOrbDescriptor (:108-147) on that blurred patch, and the final
//   cv::KeyPoint record (pt scaled by mvScaleFactor[level] AFTER sampling, :1095-1101).
// A blurred pixel is a pure function of its 7 x 7 neighbourhood, so blurring only the
// patches gives the same bytes as blurring the level; the plane's REFLECT_101 border is the
// blur's own border mode on the border-less clone (App. B-9).
//
// Staging: one elected lane fetches the 64 x 43 raw window (16-byte aligned start, rows
// ky-21 .. ky+21) with ONE TMA tile load; the next keypoint's window is requested as soon as
// the row pass has consumed the current one, so the load overlaps column pass + sampling.
//   phase 1  intensity centroid: lane = column u, 31 conflict-free byte reads, warp reduction,
//            cv::fastAtan2's polynomial and float(cos/sin in double) (canonical rule B-2);
//   phase 2  blur row pass: item = (row pair, 8 columns); byte windows by funnel shift, two
//            IDP.4A per pixel; the 16-bit sums of two consecutive rows are packed in one word;
//   phase 3  blur column pass: item = (column, 8 rows), four IDP.2A per pixel over the row
//            pairs, (sum + 32768) >> 16, stored transposed (4 rows per 32-bit store);
//   phase 4  512 rotated pattern samples (16 per lane) from the blurred patch, packed stores.
// =====================================================================================
#define KP_RAW_W 64
#define KP_RAW_H 43
#define KP_RAW_BYTES 2816                      // 64 * 43 = 2752, rounded up to 128
#define KP_HP_COLS 40
#define KP_HP_BYTES (24 * KP_HP_COLS * 4)     // 24 row pairs (22 written; the column pass of rows 37..39 reads 2 more)
#define KP_BL_PITCH 40
#define KP_BL_BYTES (KP_HP_COLS * KP_BL_PITCH)
#define KP_WARP_BYTES (KP_RAW_BYTES + KP_HP_BYTES + KP_BL_BYTES + 64)      // 8320 = 65 * 128
#define KP_WARPS 8

struct KpItem {
    unsigned it;               // slot index (frame * kept_per_frame + slot); >= total: none
    int frame, l, sl;
    uint32_t p;                // packed x | y << 12 | response << 24, level coordinates
};

__global__ void __launch_bounds__(KP_WARPS * 32, 3)
describe_kernel(const __grid_constant__ FastMaps maps, const OrbxPlan* __restrict__ plan, int frame0, int nframes,
                const uint32_t* __restrict__ kept, const int* __restrict__ kept_counts, float* __restrict__ angles,
                float* __restrict__ out_kp, uint8_t* __restrict__ out_desc) {
    ORBX_PDL_WAIT();
    extern __shared__ uint8_t kp_smem_raw[];
    __shared__ uint64_t s_bar[KP_WARPS];
    // umax (:454-469) is a function of HALF_PATCH_SIZE only; the host-computed copy in the plan is
    // checked against this table when the plan is built.
    constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* base = kp_smem_raw + ((128 - (smem_u32(kp_smem_raw) & 127)) & 127) + (size_t)warp * KP_WARP_BYTES;
    const uint8_t* raw = base;
    uint32_t* Hp = reinterpret_cast<uint32_t*>(base + KP_RAW_BYTES);
    uint8_t* blT = base + KP_RAW_BYTES + KP_HP_BYTES;
    uint64_t* bar = &s_bar[warp];

    // this lane's 16 pattern points (descriptor byte `lane`, :123-144)
    float px[16], py[16];
    {
        const uint4 lo = reinterpret_cast<const uint4*>(g_pattern)[lane * 2], hi = reinterpret_cast<const uint4*>(g_pattern)[lane * 2 + 1];
        const uint32_t wds[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t wd = wds[j >> 1] >> (16 * (j & 1));
            px[j] = (float)(int)(signed char)(wd & 0xff);
            py[j] = (float)(int)(signed char)((wd >> 8) & 0xff);
        }
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int kpf = plan->kept_per_frame;
    const int nl = plan->nlevels;
    __shared__ int s_koff[ORBX_MAXL + 2];                                  // first output slot of every level, INT_MAX behind the last
    if (threadIdx.x < ORBX_MAXL + 2) s_koff[threadIdx.x] = (int)threadIdx.x < nl ? plan->lv[threadIdx.x].kept_off : INT_MAX;
    __syncthreads();
    // IC_Angle lane roles (phase 1): lane = (row group icg, word icw); word icw holds patch columns u = -16 + 4 icw .. + 3,
    // row group icg takes rows v = -15 + icg + 4 i, i = 0 .. 7.  icmask[i]: bytes of the word inside the circular patch
    // (|u| <= umax[|v|], :86-101), icwt: the four u as signed bytes.
    const int icw = lane & 7, icg = lane >> 3;
    uint32_t icmask[8], icwt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) icwt |= (uint32_t)((-16 + 4 * icw + j) & 0xff) << (8 * j);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int v = -15 + icg + 4 * i;
        const int um = v <= 15 ? plan->umax[v < 0 ? -v : v] : -1;        // (== UMAX, checked when the plan is built)
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = -16 + 4 * icw + j;
            if ((u < 0 ? -u : u) <= um) m |= 0xffu << (8 * j);
        }
        icmask[i] = m;
    }
    const unsigned total = (unsigned)nframes * (unsigned)kpf;
    const unsigned nwarps = gridDim.x * KP_WARPS;
    const uint32_t K0123 = 18u | (34u << 8) | (48u << 16) | (56u << 24);
    const uint32_t K456 = 48u | (34u << 8) | (18u << 16);
    const uint32_t KA = K0123;                                            // even row: pairs m (lo), m+1 (hi)
    const uint32_t KB = K456;                                             //           pairs m+2 (lo), m+3 (hi)
    const uint32_t KC = (18u << 8) | (34u << 16) | (48u << 24);           // odd row:  pairs m (lo), m+1 (hi)
    const uint32_t KD = 56u | (48u << 8) | (34u << 16) | (18u << 24);     //           pairs m+2 (lo), m+3 (hi)

    // next occupied slot at or after the cursor (slots past a level's kept count are empty); the cursor carries (frame, slot)
    // along, so a step costs neither a division nor a walk over the plan's level records
    unsigned c_it = blockIdx.x * KP_WARPS + warp;
    int c_frame = (int)(c_it / (unsigned)kpf), c_s = (int)(c_it - (unsigned)c_frame * (unsigned)kpf);
    auto next_item = [&]() -> KpItem {
        KpItem k;
        k.frame = k.l = k.sl = 0;
        k.p = 0;
        for (; c_it < total;) {
            int l = 0;
            while (c_s >= s_koff[l + 1]) ++l;
            const int sl = c_s - s_koff[l];
            if (sl < kept_counts[c_frame * nl + l]) {
                k.frame = c_frame; k.l = l; k.sl = sl;
                k.p = kept[c_it];
                break;
            }
            c_it += nwarps; c_s += (int)nwarps;
            while (c_s >= kpf) { c_s -= kpf; ++c_frame; }
        }
        k.it = c_it;
        c_it += nwarps; c_s += (int)nwarps;                                // the cursor rests on the slot after the returned one
        while (c_s >= kpf) { c_s -= kpf; ++c_frame; }
        return k;
    };
    auto issue = [&](const KpItem& k) {
        if (lane == 0) {
            mbar_expect_tx(bar, KP_RAW_W * KP_RAW_H);
            tma_load_3d(base, &maps.m[k.l], bar, (ORBX_XO + ORBX_PX(k.p) - 21) & ~15, ORBX_EDGE + ORBX_PY(k.p) - 21,
                        frame0 + k.frame);
        }
    };

    KpItem cur = next_item();
    if (cur.it < total) issue(cur);
    uint32_t phase = 0;
    while (cur.it < total) {
        const KpItem nxt = next_item();
        const OrbxLevel& L = plan->lv[cur.l];
        const int kx = ORBX_PX(cur.p), ky = ORBX_PY(cur.p);
        const int a16 = (ORBX_XO + kx - 21) & 15;                         // byte column of patch column -21 inside the window
        mbar_wait(bar, phase);
        phase ^= 1;

        // ---- phase 1: IC_Angle (:77-104).  Integer moments, so any summation order gives the reference's m_10 / m_01: a lane
        //      sums 4 columns of 8 rows with two byte dot products per row (row sum, and the sum weighted by u)
        int m10 = 0, m01 = 0;
        {
            const int ub = a16 + 5;                                        // window byte column of patch column u = -16
            const int sh = (ub & 3) * 8;
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw) + (6 + icg) * (KP_RAW_W / 4) + (ub >> 2) + icw;
            int s0 = 0, s1 = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t px = __funnelshift_r(rw[i * KP_RAW_W], rw[i * KP_RAW_W + 1], sh) & icmask[i];
                const int rs = (int)__dp4a(px, 0x01010101u, 0u);
                asm("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(m10) : "r"(icwt), "r"(px));
                s0 += rs;
                s1 += i * rs;
            }
            m01 = (icg - 15) * s0 + 4 * s1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        // every lane evaluates the same scalars (no broadcast needed)
        const float angle = fast_atan2_deg((float)m01, (float)m10, plan);
        float a, b;                                                       // (:111-113) under canonical rule B-2
        {
            const float rad = __fmul_rn(angle, plan->factor_pi);
            double sn, cs;
            sincos((double)rad, &sn, &cs);
            a = (float)cs;
            b = (float)sn;
        }

        // ---- phase 2: blur row pass.  Column index B = first-tap byte column - 4 * wbase; blurred patch column
        //      i = B - (a16 & 3), i in [0, 37).  Row pair rp = window rows 2rp, 2rp+1.
        const int wbase = a16 >> 2;
#pragma unroll 1
        for (int q = lane; q < 22 * 5; q += 32) {
            const int rp = (q * 13108) >> 16;                             // q / 5
            const int g = q - rp * 5;
            const uint32_t* r0 = reinterpret_cast<const uint32_t*>(raw + 2 * rp * KP_RAW_W) + wbase + 2 * g;
            const uint32_t* r1 = r0 + (rp < 21 ? KP_RAW_W / 4 : 0);       // window row 43 does not exist (and is never used)
            uint32_t hs[2][8];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t* r = e ? r1 : r0;
                const uint32_t w0 = r[0], w1 = r[1], w2 = r[2], w3 = r[3];
                uint32_t W[12];
                W[0] = w0; W[4] = w1; W[8] = w2;
#pragma unroll
                for (int j = 1; j < 4; ++j) {
                    W[j] = __funnelshift_r(w0, w1, 8 * j);
                    W[4 + j] = __funnelshift_r(w1, w2, 8 * j);
                    W[8 + j] = __funnelshift_r(w2, w3, 8 * j);
                }
#pragma unroll
                for (int x = 0; x < 8; ++x) hs[e][x] = __dp4a(W[x], K0123, __dp4a(W[x + 4], K456, 0u));
            }
            uint4 o0, o1;
            o0.x = __byte_perm(hs[0][0], hs[1][0], 0x5410); o0.y = __byte_perm(hs[0][1], hs[1][1], 0x5410);
            o0.z = __byte_perm(hs[0][2], hs[1][2], 0x5410); o0.w = __byte_perm(hs[0][3], hs[1][3], 0x5410);
            o1.x = __byte_perm(hs[0][4], hs[1][4], 0x5410); o1.y = __byte_perm(hs[0][5], hs[1][5], 0x5410);
            o1.z = __byte_perm(hs[0][6], hs[1][6], 0x5410); o1.w = __byte_perm(hs[0][7], hs[1][7], 0x5410);
            uint4* dst = reinterpret_cast<uint4*>(Hp + rp * KP_HP_COLS + 8 * g);
            dst[0] = o0;
            dst[1] = o1;
        }
        __syncwarp();
        // the raw window is consumed: fetch the next keypoint's while this one is blurred and sampled
        if (nxt.it < total) issue(nxt);

        // ---- phase 3: blur column pass.  item = (patch column ci, chunk of 8 rows); transposed store blT[B][y]
        const int a4 = a16 & 3;
#pragma unroll 1
        for (int q = lane; q < 37 * 5; q += 32) {
            const int ch = (q * 1772) >> 16;                              // q / 37 (exact for q < 185)
            const int B = a4 + q - ch * 37;
            const uint32_t* hp = Hp + 4 * ch * KP_HP_COLS + B;
            uint32_t pr[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) pr[k] = hp[k * KP_HP_COLS];
            uint32_t ve[4], vo[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                ve[m] = __dp2a_lo(pr[m], KA, __dp2a_hi(pr[m + 1], KA, __dp2a_lo(pr[m + 2], KB, __dp2a_hi(pr[m + 3], KB, 32768u))));
                vo[m] = __dp2a_lo(pr[m], KC, __dp2a_hi(pr[m + 1], KC, __dp2a_lo(pr[m + 2], KD, __dp2a_hi(pr[m + 3], KD, 32768u))));
            }
            uint2 o;
            o.x = __byte_perm(__byte_perm(ve[0], vo[0], 0x0062), __byte_perm(ve[1], vo[1], 0x0062), 0x5410);
            o.y = __byte_perm(__byte_perm(ve[2], vo[2], 0x0062), __byte_perm(ve[3], vo[3], 0x0062), 0x5410);
            *reinterpret_cast<uint2*>(blT + B * KP_BL_PITCH + 8 * ch) = o;
        }
        __syncwarp();

        // ---- phase 4: descriptor byte `lane` (:115-144) from the blurred patch; sample (row rr, column cc)
        //      relative to the keypoint sits at blT[a4 + cc + 18][rr + 18]
        const uint8_t* bc = blT + (a4 + 18) * KP_BL_PITCH + 18;
        uint32_t byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float x = px[2 * j + e], y = py[2 * j + e];
                const int rr = __float2int_rn(__fadd_rn(__fmul_rn(x, b), __fmul_rn(y, a)));    // (:119)
                const int cc = __float2int_rn(__fsub_rn(__fmul_rn(x, a), __fmul_rn(y, b)));    // (:120)
                ORBX_BC(rr >= -18 && rr <= 18 && cc >= -18 && cc <= 18 && a4 >= 0 && a4 + cc + 18 < KP_HP_COLS);
                v[e] = bc[cc * KP_BL_PITCH + rr];
            }
            byte |= (uint32_t)(v[0] < v[1]) << j;
        }
        uint32_t word = byte << (8 * (lane & 3));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);                     // lanes 4k..4k+3 hold word k
        const int half = lane >> 4;
        uint4 v4;
        v4.x = __shfl_sync(0xffffffffu, word, 16 * half + 0);
        v4.y = __shfl_sync(0xffffffffu, word, 16 * half + 4);
        v4.z = __shfl_sync(0xffffffffu, word, 16 * half + 8);
        v4.w = __shfl_sync(0xffffffffu, word, 16 * half + 12);
        const int* kc = kept_counts + cur.frame * nl;
        int oidx = cur.sl;
        for (int q = 0; q < cur.l; ++q) oidx += kc[q];                    // levels concatenated (:1076-1104)
        const size_t orow = (size_t)cur.frame * kpf + oidx;
        if ((lane & 15) == 0) *reinterpret_cast<uint4*>(out_desc + orow * 32 + 16 * half) = v4;
        if (lane < 7) {
            float f;
            if (lane == 0) f = cur.l ? __fmul_rn((float)kx, L.scale) : (float)kx;
            else if (lane == 1) f = cur.l ? __fmul_rn((float)ky, L.scale) : (float)ky;
            else if (lane == 2) f = L.kp_size;
            else if (lane == 3) f = angle;
            else if (lane == 4) f = (float)ORBX_PR(cur.p);
            else if (lane == 5) f = __int_as_float(cur.l);
            else f = __int_as_float(-1);
            out_kp[orow * 7 + lane] = f;
        }
        if (lane == 7) angles[cur.it] = angle;
        __syncwarp();
        cur = nxt;
    }
}

// =====================================================================================
// GaussianBlur 7x7 sigma 2 (:1086): separable 8.8 fixed point [18,34,48,56,48,34,18]/256
// (SURVEY App. A-2).  The padded plane's REFLECT_101 border is exactly the blur's own border
// mode on the border-less clone, so taps simply read the plane.
// One warp owns a 128 x 32 output tile, one lane a 4-pixel-wide column strip, and walks down
// the rows with everything in registers (no shared memory, no barriers):
//   row pass     3 aligned 32-bit loads give pixels x-4 .. x+7; each of the 4 row sums is two
//                IDP.4A (byte dot products) on funnel-shifted words;
//   column pass  the 16-bit row sums of two consecutive rows are packed into one register,
//                so each output is four IDP.2A over a ring of 4 row pairs;
//   store        (sum + 32768) >> 16, four pixels per 32-bit store.
// =====================================================================================
#define BL_TW 128
#define BL_TH 32
__global__ void __launch_bounds__(128) blur_kernel(const OrbxPlan* __restrict__ plan, int nframes,
                                                   const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
    ORBX_PDL_WAIT();
    const int lane = threadIdx.x & 31;
    const int tpf = plan->blur_tiles_per_frame;
    const long long total = (long long)nframes * tpf;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const uint32_t K0123 = 18u | (34u << 8) | (48u << 16) | (56u << 24);
    const uint32_t K456 = 48u | (34u << 8) | (18u << 16);
    const uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // even output: pairs m-3 (lo), m-2 (hi)
    const uint32_t KB = 48u | (34u << 8) | (18u << 16);                   //              pairs m-1 (lo), m (hi)
    const uint32_t KC = (18u << 8) | (34u << 16) | (48u << 24);           // odd output:  pairs m-3 (lo), m-2 (hi)
    const uint32_t KD = 56u | (48u << 8) | (34u << 16) | (18u << 24);     //              pairs m-1 (lo), m (hi)
    for (long long it = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < total; it += nwarps) {
        const int frame = (int)(it / tpf);
        const int t = (int)(it - (long long)frame * tpf);
        int l = 0;
        while (l + 1 < plan->nlevels && t >= plan->lv[l + 1].blur_tile_base) ++l;
        const OrbxLevel& L = plan->lv[l];
        const int tl = t - L.blur_tile_base;
        const int ty = tl / L.blur_tiles_x, tx = tl - ty * L.blur_tiles_x;
        const int x = tx * BL_TW + 4 * lane, y0 = ty * BL_TH;
        if (x >= L.w) continue;
        const int pitch = L.pitch;
        // Row pointers advance by one pitch per row (one 64-bit add instead of a multiply per access); the last valid
        // plane row is level row h+18, rows past it (partial bottom tiles) re-read it and their outputs are not stored.
        const int rows_ok = L.h + ORBX_EDGE - (y0 - 3);                    // readable rows from the first tap row on
        const uint8_t* srow = pyr + (size_t)frame * plan->slab_bytes + L.plane_off + (size_t)(ORBX_EDGE + y0 - 3) * pitch + ORBX_XO + x - 4;
        uint8_t* drow = blur + (size_t)frame * plan->slab_bytes + L.plane_off + (size_t)(ORBX_EDGE + y0) * pitch + ORBX_XO + x;
        uint32_t P[4][4];                                                 // ring of row pairs x 4 columns
#pragma unroll
        for (int m = 0; m < (BL_TH + 6) / 2; ++m) {
            uint32_t hs[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t* rp = reinterpret_cast<const uint32_t*>(srow);
                const uint32_t w0 = __ldg(rp), w1 = __ldg(rp + 1), w2 = __ldg(rp + 2);
                if (2 * m + e + 1 < rows_ok) srow += pitch;
                hs[e][0] = __dp4a(__funnelshift_r(w0, w1, 8), K0123, __dp4a(__funnelshift_r(w1, w2, 8), K456, 0u));
                hs[e][1] = __dp4a(__funnelshift_r(w0, w1, 16), K0123, __dp4a(__funnelshift_r(w1, w2, 16), K456, 0u));
                hs[e][2] = __dp4a(__funnelshift_r(w0, w1, 24), K0123, __dp4a(__funnelshift_r(w1, w2, 24), K456, 0u));
                hs[e][3] = __dp4a(w1, K0123, __dp4a(w2, K456, 0u));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) P[m & 3][c] = __byte_perm(hs[0][c], hs[1][c], 0x5410);
            if (m >= 3) {
                const int oe = y0 + 2 * m - 6;                            // even output row (level coordinates)
                uint32_t ve[4], vo[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t p0 = P[(m - 3) & 3][c], p1 = P[(m - 2) & 3][c], p2 = P[(m - 1) & 3][c], p3 = P[m & 3][c];
                    ve[c] = __dp2a_lo(p0, KA, __dp2a_hi(p1, KA, __dp2a_lo(p2, KB, __dp2a_hi(p3, KB, 32768u))));
                    vo[c] = __dp2a_lo(p0, KC, __dp2a_hi(p1, KC, __dp2a_lo(p2, KD, __dp2a_hi(p3, KD, 32768u))));
                }
                if (oe < L.h)
                    *reinterpret_cast<uint32_t*>(drow) =
                        __byte_perm(__byte_perm(ve[0], ve[1], 0x0062), __byte_perm(ve[2], ve[3], 0x0062), 0x5410);
                if (oe + 1 < L.h)
                    *reinterpret_cast<uint32_t*>(drow + pitch) =
                        __byte_perm(__byte_perm(vo[0], vo[1], 0x0062), __byte_perm(vo[2], vo[3], 0x0062), 0x5410);
                drow += 2 * pitch;
            }
        }
    }
}

// =====================================================================================
// Frame::ComputeStereoMatches (reference src/Frame.cc:466-640) on the device-resident
// outputs of a left and a right extraction (the immediate consumer
// of the path in the stereo configurations; doing it here removes the pyramid D2H).
//
// stereo_match_kernel, one warp per left keypoint:
//   1. best right keypoint (:506-546): every right keypoint whose row band floor(y - r) ..
//      ceil(y + r), r = 2 * scale[octave] (:486-492) holds row (size_t)vL, with octave within
//      +-1 and uR in [uL - maxD, uL] is compared by Hamming distance (ORBmatcher.cc:1647-1663);
//      lanes stride over the right keypoints, the warp minimum of (dist << 16 | iR) is the
//      reference's "first strictly smaller wins" in iR order;
//   2. 11x11 SAD over incR = -5..5 on the keypoint's pyramid level (:552-596), each window
//      minus its own centre -- integers, so the reference's float L1 norm is exact;
//   3. parabola, re-scaling, disparity test (:598-626) in float32 without contraction.
// stereo_filter_kernel, one CTA per pair: the (size/2)-th smallest accepted SAD by rank
// counting, thDist = 1.5f * 1.4f * median, matches with SAD >= thDist dropped (:629-640).
// =====================================================================================
struct StereoSide {
    const uint8_t* pyr;        // frame 0 of the handle's pyramid slabs
    const float* kp;           // frame 0 of the 7-float keypoint records
    const uint8_t* desc;
    const int* kept_counts;    // [frame][level]
};

// Row table of the right image (vRowIndices, :476-496), one CTA per pair: right keypoint j is listed in every row
// of its band floor(y - r) .. ceil(y + r), r = 2 * scale[octave].  Counting sort by row: per-row counts in shared
// memory, an exclusive scan gives row_start (height + 1 entries per right frame), then the keypoint indices are
// scattered into their rows' lists.  The order inside a row is arbitrary: the matcher takes the minimum of
// (distance << 16 | index), which is the reference's "first strictly smaller wins" whatever the visiting order.
#define ST_MAX_SPAN 24          // rows per keypoint: 2 * ceil(2 * 5.16) + 2 at the 10th level of a 1.2 pyramid; checked at launch
__global__ void __launch_bounds__(256)
stereo_rows_kernel(const OrbxPlan* __restrict__ plan, StereoSide SR, const int* __restrict__ pairs, int* __restrict__ row_start,
                   uint16_t* __restrict__ bucket) {
    extern __shared__ int sr_cnt[];                                        // height + 1 counters, then cursors
    __shared__ int s_warp[33];
    const int kpf = plan->kept_per_frame, nl = plan->nlevels, H = plan->height;
    const int fR = pairs[2 * blockIdx.x + 1];
    int Nr = 0;
    for (int l = 0; l < nl; ++l) Nr += SR.kept_counts[fR * nl + l];
    for (int y = threadIdx.x; y <= H; y += blockDim.x) sr_cnt[y] = 0;
    __syncthreads();
    const float* kr0 = SR.kp + (size_t)fR * kpf * 7;
    auto band = [&](int j, int& minr, int& maxr) {
        const float yr = kr0[(size_t)j * 7 + 1];
        const float r = __fmul_rn(2.0f, plan->lv[__float_as_int(kr0[(size_t)j * 7 + 5])].scale);
        maxr = min((int)ceilf(__fadd_rn(yr, r)), H - 1);                   // rows outside the image are never looked up
        minr = max((int)floorf(__fsub_rn(yr, r)), 0);
        maxr = min(maxr, minr + ST_MAX_SPAN - 1);
    };
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        int minr, maxr;
        band(j, minr, maxr);
        for (int y = minr; y <= maxr; ++y) atomicAdd(&sr_cnt[y], 1);
    }
    __syncthreads();
    // exclusive scan over the rows: thread t owns rows [t * per, (t + 1) * per)
    const int per = (H + 1 + (int)blockDim.x - 1) / (int)blockDim.x;
    const int y0 = threadIdx.x * per, y1 = min(y0 + per, H + 1);
    int local = 0;
    for (int y = y0; y < y1; ++y) local += sr_cnt[y];
    int total;
    int run = block_excl_scan(local, &total, s_warp);
    int* rs = row_start + (size_t)fR * (H + 1);
    for (int y = y0; y < y1; ++y) {
        const int c = sr_cnt[y];
        sr_cnt[y] = run;                                                   // cursor
        rs[y] = run;
        run += c;
    }
    __syncthreads();
    uint16_t* bk = bucket + (size_t)fR * kpf * ST_MAX_SPAN;
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        int minr, maxr;
        band(j, minr, maxr);
        for (int y = minr; y <= maxr; ++y) bk[atomicAdd(&sr_cnt[y], 1)] = (uint16_t)j;
    }
}

#define ST_WARPS 8
__global__ void __launch_bounds__(ST_WARPS * 32)
stereo_match_kernel(const OrbxPlan* __restrict__ plan, StereoSide SL, StereoSide SR, const int* __restrict__ pairs, int npairs,
                    const int* __restrict__ row_start, const uint16_t* __restrict__ bucket, float mbf, float mb,
                    float* __restrict__ u_right, float* __restrict__ depth, int* __restrict__ sad_out) {
    __shared__ int s_sad[ST_WARPS][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    const unsigned total = (unsigned)npairs * (unsigned)kpf;
    const float maxD = __fdiv_rn(mbf, mb);                                 // (:499-501)
    for (unsigned it = blockIdx.x * ST_WARPS + warp; it < total; it += gridDim.x * ST_WARPS) {
        const int p = (int)(it / (unsigned)kpf), iL = (int)(it - (unsigned)p * (unsigned)kpf);
        const int fL = pairs[2 * p], fR = pairs[2 * p + 1];
        int N = 0;
        for (int l = 0; l < nl; ++l) N += SL.kept_counts[fL * nl + l];
        if (iL >= N) continue;
        const size_t oL = (size_t)fL * kpf + iL;
        const float* kl = SL.kp + oL * 7;
        const float uL = kl[0], vL = kl[1];
        const int levelL = __float_as_int(kl[5]);
        float res_u = -1.f, res_d = -1.f;
        int res_sad = -1;
        const int row = (int)vL;                                           // vRowIndices[vL] (:512)
        const float minU = __fsub_rn(uL, maxD), maxU = uL;                 // minD = 0 (:517-518)
        uint32_t best = ((uint32_t)100 << 16) | 0xffffu;                   // TH_HIGH (:523)
        if (!(maxU < 0.f)) {
            const uint4* dl = reinterpret_cast<const uint4*>(SL.desc + oL * 32);
            const uint4 a0 = __ldg(dl), a1 = __ldg(dl + 1);
            const float* kr0 = SR.kp + (size_t)fR * kpf * 7;
            const uint8_t* dr0 = SR.desc + (size_t)fR * kpf * 32;
            const int* rs = row_start + (size_t)fR * (plan->height + 1);
            const int cbeg = row < plan->height ? rs[row] : 0, cend = row < plan->height ? rs[row + 1] : 0;
            const uint16_t* bk = bucket + (size_t)fR * kpf * ST_MAX_SPAN;
            for (int c = cbeg + lane; c < cend; c += 32) {                 // vRowIndices[vL] (:512)
                const int j = bk[c];
                const float* kr = kr0 + (size_t)j * 7;
                const float xr = kr[0];
                const int oc = __float_as_int(kr[5]);
                if (oc >= levelL - 1 && oc <= levelL + 1 && xr >= minU && xr <= maxU) {
                    const uint4* dr = reinterpret_cast<const uint4*>(dr0 + (size_t)j * 32);
                    const uint4 b0 = __ldg(dr), b1 = __ldg(dr + 1);
                    const uint32_t d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                       __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                    best = min(best, (d << 16) | (uint32_t)j);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
        const int bestDist = (int)(best >> 16);
        if (bestDist < (100 + 50) / 2) {                                   // thOrbDist (:471, :549)
            const int bestR = (int)(best & 0xffffu);
            const float uR0 = SR.kp[((size_t)fR * kpf + bestR) * 7];
            const OrbxLevel& Lv = plan->lv[levelL];
            const float inv = __fdiv_rn(1.0f, Lv.scale);                   // mvInvScaleFactors (src/ORBextractor.cc:425-431)
            const int su = (int)roundf(__fmul_rn(uL, inv)), sv = (int)roundf(__fmul_rn(vL, inv));
            const int sr = (int)roundf(__fmul_rn(uR0, inv));
            // (:575-578) iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1.  The windows themselves reach 10 columns
            // left of scaleduR0: OpenCV throws on a negative colRange; here such a keypoint simply gets no match, and any
            // window leaving the padded plane is refused so no read can go out of bounds.
            const bool ok = !(sr < 0 || sr + 11 >= Lv.w) && sr - 10 >= -ORBX_EDGE && sr + 10 < Lv.w + ORBX_EDGE &&
                            su - 5 >= -ORBX_EDGE && su + 5 < Lv.w + ORBX_EDGE && sv - 5 >= -ORBX_EDGE && sv + 5 < Lv.h + ORBX_EDGE;
            if (ok) {
                if (lane < 12) s_sad[warp][lane] = 0;
                __syncwarp();
                const uint8_t* PL = SL.pyr + (size_t)fL * plan->slab_bytes + Lv.plane_off + (size_t)(sv - 5 + ORBX_EDGE) * Lv.pitch + ORBX_XO + su - 5;
                const uint8_t* PR = SR.pyr + (size_t)fR * plan->slab_bytes + Lv.plane_off + (size_t)(sv - 5 + ORBX_EDGE) * Lv.pitch + ORBX_XO + sr - 10;
                const int cL = PL[5 * Lv.pitch + 5];
                for (int t = lane; t < 121; t += 32) {
                    const int i = (t * 5958) >> 16, rw = t - i * 11;      // incR = i - 5, window row rw
                    const int cR = PR[5 * Lv.pitch + 5 + i];
                    const uint8_t* a = PL + (size_t)rw * Lv.pitch;
                    const uint8_t* b = PR + (size_t)rw * Lv.pitch + i;
                    int acc = 0;
#pragma unroll
                    for (int c = 0; c < 11; ++c) acc += abs(((int)a[c] - cL) - ((int)b[c] - cR));
                    atomicAdd(&s_sad[warp][i], acc);
                }
                __syncwarp();
                int bd = 0x7fffffff, bi = 0;
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const int d = s_sad[warp][i];
                    if (d < bd) { bd = d; bi = i; }
                }
                __syncwarp();
                if (bi != 0 && bi != 10) {                                 // (:598-599)
                    const float d1 = (float)s_sad[warp][bi - 1], d2 = (float)bd, d3 = (float)s_sad[warp][bi + 1];
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                    if (!(deltaR < -1.f || deltaR > 1.f)) {                // a NaN passes, as in the reference (:608)
                        float bestuR = __fmul_rn(Lv.scale, __fadd_rn(__fadd_rn((float)sr, (float)(bi - 5)), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0.f && disparity < maxD) {
                            if (disparity <= 0.f) {
                                disparity = 0.01f;
                                bestuR = (float)((double)uL - 0.01);       // (:620) evaluated in double
                            }
                            res_d = __fdiv_rn(mbf, disparity);
                            res_u = bestuR;
                            res_sad = bd;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            u_right[oL] = res_u;
            depth[oL] = res_d;
            sad_out[oL] = res_sad;
        }
    }
}

__global__ void __launch_bounds__(1024)
stereo_filter_kernel(const OrbxPlan* __restrict__ plan, const int* __restrict__ left_counts, const int* __restrict__ pairs,
                     float* __restrict__ u_right, float* __restrict__ depth, const int* __restrict__ sad) {
    extern __shared__ int sf_d[];                                          // accepted SADs, compacted
    __shared__ int s_n, s_median;
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    const int fL = pairs[2 * blockIdx.x];
    int N = 0;
    for (int l = 0; l < nl; ++l) N += left_counts[fL * nl + l];
    const size_t o = (size_t)fL * kpf;
    if (threadIdx.x == 0) { s_n = 0; s_median = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int d = sad[o + i];
        if (d >= 0) sf_d[atomicAdd(&s_n, 1)] = d;                         // order is irrelevant for a rank
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;                                                    // vDistIdx[0] of an empty vector in the reference
    const int k = n / 2;                                                   // (:630)
    // sorted[k] = the smallest v with #(d <= v) > k: bisection on the value (SADs are < 2^17), one block-wide
    // population count per step
    int lo = 0, hi = (1 << 17) - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) c += sf_d[i] <= mid;
        const int tot = __syncthreads_count(c & 1) + 2 * __syncthreads_count(c & 2) + 4 * __syncthreads_count(c & 4) +
                        8 * __syncthreads_count(c & 8);                    // c <= 15: at most 15 elements per thread (checked at launch)
        if (tot > k) hi = mid; else lo = mid + 1;
    }
    if (threadIdx.x == 0) s_median = lo;
    __syncthreads();
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_median);   // (:631)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int d = sad[o + i];
        if (d >= 0 && !((float)d < thDist)) { u_right[o + i] = -1.f; depth[o + i] = -1.f; }
    }
}

// =====================================================================================
// Frame::UndistortKeyPoints + Frame::AssignFeaturesToGrid (reference src/Frame.cc:404-434, :230-245, PosInGrid :382-392;
// SURVEY.md §8(f) row 3), one CTA per frame on the keypoints the extraction left in HBM.
//   1. cv::undistortPoints(K, D, R = I, P = K): the 5-iteration inverse of OpenCV's distortion model in double with
//      individually rounded operations (bit-identical to cv2 4.13 on 60 000 test points), result narrowed to float;
//      k1 == 0 copies the points (:406-410).
//   2. PosInGrid: round((x - mnMinX) * mfGridElementWidthInv) in float32, cells outside 64 x 48 dropped.
//   3. mGrid[gx][gy] as CSR: per-cell counts (shared-memory atomics), exclusive scan, and a STABLE fill -- a keypoint's
//      slot inside its cell is the number of earlier keypoints of the same cell, i.e. the reference's push_back order.
// =====================================================================================
struct UndistortParams {
    double fx, fy, cx, cy, ifx, ify;
    double k[5];                    // k1 k2 p1 p2 k3
    int distorted;                  // mDistCoef(0) != 0
    float min_x, min_y, winv, hinv; // mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv
};
#define UG_COLS 64
#define UG_ROWS 48

__global__ void __launch_bounds__(1024)
undistort_grid_kernel(const OrbxPlan* __restrict__ plan, const float* __restrict__ kp, const int* __restrict__ kept_counts,
                      const int* __restrict__ frames, UndistortParams P, float* __restrict__ xy_un, int* __restrict__ cell_start,
                      int* __restrict__ cell_items) {
    extern __shared__ int ug_smem[];
    __shared__ int s_warp[33];
    int* cnt = ug_smem;                                  // UG_COLS * UG_ROWS counters, then running starts
    int* cell = ug_smem + UG_COLS * UG_ROWS;             // cell of every keypoint (-1: outside the grid)
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    const int f = frames[blockIdx.x];
    int N = 0;
    for (int l = 0; l < nl; ++l) N += kept_counts[f * nl + l];
    for (int c = threadIdx.x; c < UG_COLS * UG_ROWS; c += blockDim.x) cnt[c] = 0;
    __syncthreads();
    const float* k0 = kp + (size_t)f * kpf * 7;
    float* out = xy_un + (size_t)f * kpf * 2;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float ux = k0[(size_t)i * 7], uy = k0[(size_t)i * 7 + 1];
        if (P.distorted) {
            double x = __dmul_rn(__dsub_rn((double)ux, P.cx), P.ifx), y = __dmul_rn(__dsub_rn((double)uy, P.cy), P.ify);
            const double x0 = x, y0 = y;
#pragma unroll 1
            for (int j = 0; j < 5; ++j) {
                const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
                // k4..k6 = 0: the numerator polynomial is 1 + ((0*r2 + 0)*r2 + 0)*r2 = 1
                const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(P.k[4], r2), P.k[1]), r2), P.k[0]), r2));
                const double icdist = __ddiv_rn(1.0, den);
                const double xy2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, P.k[2]), x), y);                   // 2*p1*x*y
                const double dX = __dadd_rn(__dadd_rn(__dadd_rn(xy2, __dmul_rn(P.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
                                                      __dmul_rn(0.0, r2)), __dmul_rn(__dmul_rn(0.0, r2), r2));
                const double yx2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, P.k[3]), x), y);                   // 2*p2*x*y
                const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), yx2),
                                                      __dmul_rn(0.0, r2)), __dmul_rn(__dmul_rn(0.0, r2), r2));
                x = __dmul_rn(__dsub_rn(x0, dX), icdist);
                y = __dmul_rn(__dsub_rn(y0, dY), icdist);
            }
            const double xx = __dadd_rn(__dadd_rn(__dmul_rn(P.fx, x), __dmul_rn(0.0, y)), P.cx);
            const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(P.fy, y)), P.cy);
            ux = (float)xx;                                                // ww = 1 / (0*x + 0*y + 1) = 1
            uy = (float)yy;
        }
        out[2 * i] = ux;
        out[2 * i + 1] = uy;
        const int px = (int)roundf(__fmul_rn(__fsub_rn(ux, P.min_x), P.winv));          // PosInGrid (:384-385)
        const int py = (int)roundf(__fmul_rn(__fsub_rn(uy, P.min_y), P.hinv));
        int c = -1;
        if (px >= 0 && px < UG_COLS && py >= 0 && py < UG_ROWS) {
            c = px * UG_ROWS + py;
            atomicAdd(&cnt[c], 1);
        }
        cell[i] = c;
    }
    __syncthreads();
    // exclusive scan over the 3072 cells: thread t owns 3 consecutive cells
    const int per = (UG_COLS * UG_ROWS + (int)blockDim.x - 1) / (int)blockDim.x;
    const int c0 = threadIdx.x * per, c1 = min(c0 + per, UG_COLS * UG_ROWS);
    int local = 0;
    for (int c = c0; c < c1; ++c) local += cnt[c];
    int total;
    int run = block_excl_scan(local, &total, s_warp);
    int* cs = cell_start + (size_t)f * (UG_COLS * UG_ROWS + 1);
    for (int c = c0; c < c1; ++c) {
        const int v = cnt[c];
        cnt[c] = run;
        cs[c] = run;
        run += v;
    }
    if (threadIdx.x == 0) cs[UG_COLS * UG_ROWS] = total;
    __syncthreads();
    int* items = cell_items + (size_t)f * kpf;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int c = cell[i];
        if (c < 0) continue;
        int rank = 0;
        for (int j = 0; j < i; ++j) rank += cell[j] == c;                  // push_back order (:243)
        items[cnt[c] + rank] = i;
    }
}

// =====================================================================================
// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (src/ORBmatcher.cc:1328-1470)
// with Frame::GetFeaturesInArea (src/Frame.cc:327-380), DescriptorDistance (src/ORBmatcher.cc:1647-1663) and
// ComputeThreeMaxima (:1601-1642): the per-frame descriptor consumer of Tracking::TrackWithMotionModel.  One CTA per
// (LastFrame, CurrentFrame) query, a warp per LastFrame map point.  The current frame's undistorted keypoints,
// descriptors, mGrid (CSR) and mvuRight are what the extraction, orbx_undistort_grid and orbx_stereo_match left in HBM.
//
// The reference loop is sequential: a current keypoint claimed by a map point with Observations() > 0 is skipped by
// every LATER map point (:1401-1403), one claimed by a point without observations is overwritten by a later one.  The
// choice of point i therefore depends only on the choices of points j < i, so the sequential result is the unique
// fixed point of   match[i] = best candidate not claimed by {j < i : obs[j] > 0 and match[j] = that candidate},
// reached by Jacobi iteration (every point re-evaluated against the previous round's claims; round k fixes at least
// the first k points, in practice 3-6 rounds).  Candidates are visited in CSR position order, which IS the reference's
// visiting order (cells ix-major, iy, push_back order), so "first strictly smaller distance wins" (:1418) is the
// minimum of (distance << 16 | position).  Afterwards: rotation histogram, three maxima, culling (:1425-1466); the
// holder of a keypoint is the last point that matched it, a keypoint with ANY matcher in a culled bin ends up NULL and
// nmatches counts overwritten matches exactly as the reference does.
// =====================================================================================
struct SpParams {
    float fx, fy, cx, cy;                 // Frame::fx ... (static members)
    float min_x, max_x, min_y, max_y;     // mnMinX, mnMaxX, mnMinY, mnMaxY
    float winv, hinv;                     // mfGridElementWidthInv / HeightInv
    float mbf, th;
    float nnratio;                        // ORBmatcher::mfNNratio (the local-points variant)
    int check_ori, cap, use_stereo;
    int list_th;                          // largest distance a candidate-list entry may have
    int match_th;                         // projection variants: best distance that still matches (TH_HIGH, :1425; ORBdist, :1555)
    int check_z;                          // 1: points behind the camera are skipped (:1370-1371); the KeyFrame variant has no such test
};
struct SpQuery {
    float R[9], t[3];                     // Rcw, tcw of the CURRENT frame
    int n_last, frame, fwd, bwd;
};
#define SP_TH_HIGH 100                    // ORBmatcher::TH_HIGH (src/ORBmatcher.cc:37)
#define SP_HISTO 30                       // ORBmatcher::HISTO_LENGTH (:39)

__device__ __forceinline__ int sp_rot_bin(float last_angle, float cur_angle) {
    float rot = __fsub_rn(last_angle, cur_angle);                           // (:1432-1437)
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, __fdiv_rn(1.0f, (float)SP_HISTO)));
    if (bin == SP_HISTO) bin = 0;
    return min(max(bin, 0), SP_HISTO - 1);                                  // the reference asserts the range
}

// Per-frame views of what the extraction, orbx_undistort_grid and orbx_stereo_match left in HBM.
struct SpFrame {
    const float* K0;       // keypoint records (7 floats)
    const uint4* D0;       // descriptors
    const float* XY;       // mvKeysUn[].pt
    const int* CS;         // mGrid as CSR
    const int* IT;
    const float* UR;       // mvuRight or nullptr
};
__device__ __forceinline__ SpFrame sp_frame(const OrbxPlan* __restrict__ plan, const SpParams& P, int f, const float* kp,
                                            const uint8_t* desc, const float* xy_un, const int* cell_start, const int* cell_items,
                                            const float* u_right) {
    const int kpf = plan->kept_per_frame;
    SpFrame F;
    F.K0 = kp + (size_t)f * kpf * 7;
    F.D0 = reinterpret_cast<const uint4*>(desc + (size_t)f * kpf * 32);
    F.XY = xy_un + (size_t)f * kpf * 2;
    F.CS = cell_start + (size_t)f * (UG_COLS * UG_ROWS + 1);
    F.IT = cell_items + (size_t)f * kpf;
    F.UR = P.use_stereo ? u_right + (size_t)f * kpf : nullptr;
    return F;
}

// Search window of one map point in the current frame.
struct SpWin {
    float u, v, radius, ur;               // centre, GetFeaturesInArea's r, the expected right coordinate
    int minL, maxL;
};

// LastFrame variant (src/ORBmatcher.cc:1362-1393) and KeyFrame variant (:1497-1526, fwd = bwd = 0, LO = the level
// MapPoint::PredictScale gave the caller): project the map point with the current pose.  W = world positions.
__device__ __forceinline__ bool sp_project(const OrbxPlan* __restrict__ plan, const SpQuery& q, const SpParams& P,
                                           const float* __restrict__ W, const int* __restrict__ LO, int i, SpWin& w) {
    const float wx = W[3 * i], wy = W[3 * i + 1], wz = W[3 * i + 2];
    // x3Dc = Rcw * x3Dw + tcw: cv::gemm's 3x3 path -- float products and sums, the "+ C" in double
    float c3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(q.R[3 * r], wx), __fmul_rn(q.R[3 * r + 1], wy)), __fmul_rn(q.R[3 * r + 2], wz));
        c3[r] = __double2float_rn(__dadd_rn((double)t0, (double)q.t[r]));
    }
    const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)c3[2]));           // (:1368)
    w.u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c3[0]), invzc), P.cx);                // (:1373-1374)
    w.v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c3[1]), invzc), P.cy);
    // written so that NaN fails (the reference would index the grid with an undefined int cast)
    if (!(!(P.check_z && invzc < 0.f) && w.u >= P.min_x && w.u <= P.max_x && w.v >= P.min_y && w.v <= P.max_y)) return false;
    const int lo = LO[i];
    w.radius = __fmul_rn(P.th, plan->lv[lo].scale);                                 // (:1384)
    w.minL = q.fwd ? lo : q.bwd ? 0 : lo - 1;                                       // (:1388-1393)
    w.maxL = q.fwd ? -1 : q.bwd ? lo : lo + 1;
    w.ur = __fsub_rn(w.u, __fmul_rn(P.mbf, invzc));                                 // (:1407)
    return true;
}

// Local-map variant (src/ORBmatcher.cc:58-70): the window comes from the tracking fields Frame::isInFrustum left on the
// point.  W = (mTrackProjX, mTrackProjY, mTrackProjXR) triples, LO = mnTrackScaleLevel, VC = mTrackViewCos.
__device__ __forceinline__ bool sp_track_window(const OrbxPlan* __restrict__ plan, const SpParams& P, const float* __restrict__ W,
                                                const int* __restrict__ LO, const float* __restrict__ VC, int i, SpWin& w) {
    w.u = W[3 * i];
    w.v = W[3 * i + 1];
    w.ur = W[3 * i + 2];
    if (!(w.u == w.u && w.v == w.v)) return false;                                  // NaN: undefined in the reference
    const int lvl = LO[i];
    float r = (double)VC[i] > 0.998 ? 2.5f : 4.0f;                                  // RadiusByViewingCos (:131-137)
    if (P.th != 1.0f) r = __fmul_rn(r, P.th);                                       // bFactor (:49, :65-66)
    w.radius = __fmul_rn(r, plan->lv[lvl].scale);
    w.minL = lvl - 1;
    w.maxL = lvl;
    return true;
}

// Frame::GetFeaturesInArea (src/Frame.cc:327-380) + the mvuRight test + DescriptorDistance over the candidates of one
// window, executed by a whole warp; every lane returns the warp's best key (distance << 16 | CSR position), 0xffffffff
// when nothing qualifies, and *second receives the runner-up.
//   s_claim != nullptr : candidates claimed by an earlier point with observations are skipped (:96-98, :1401-1403);
//   buf != nullptr     : every candidate with distance <= P.list_th is appended to buf (first `lc` of them), *total counts them.
//   s_dist != nullptr  : SearchForInitialization's rule -- a candidate is skipped while vMatchedDistance <= distance (:447-448);
//                        s_dist is indexed by CSR position.
__device__ __forceinline__ unsigned sp_window(const SpParams& P, const SpFrame& F, const SpWin& w, const uint4* __restrict__ QD, int i,
                                              int lane, const int* s_claim, unsigned* buf, int lc, int* total, unsigned* second,
                                              const int* s_dist = nullptr) {
    const float u = w.u, v = w.v, radius = w.radius;
    const int c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, P.min_x), radius), P.winv)));
    const int c1 = min(UG_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, P.min_x), radius), P.winv)));
    const int r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, P.min_y), radius), P.hinv)));
    const int r1 = min(UG_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, P.min_y), radius), P.hinv)));
    if (second) *second = 0xffffffffu;
    if (!(c0 < UG_COLS && c1 >= 0 && r0 < UG_ROWS && r1 >= 0)) return 0xffffffffu;
    unsigned best = 0xffffffffu, best2 = 0xffffffffu;
    const bool check = w.minL > 0 || w.maxL >= 0;
    const uint4 qa = QD[2 * i], qc = QD[2 * i + 1];
    int tot = 0;
    for (int ix = c0; ix <= c1; ++ix) {
        const int p0 = F.CS[ix * UG_ROWS + r0], p1 = F.CS[ix * UG_ROWS + r1 + 1];   // cells (ix, r0..r1) are contiguous
        for (int pb = p0; pb < p1; pb += 32) {
            const int p = pb + lane;
            unsigned key = 0xffffffffu;
            if (p < p1) {
                const int k = F.IT[p];
                const int oct = __float_as_int(F.K0[(size_t)k * 7 + 5]);
                bool ok = !(check && (oct < w.minL || (w.maxL >= 0 && oct > w.maxL)));
                const float dx = __fsub_rn(F.XY[2 * k], u), dy = __fsub_rn(F.XY[2 * k + 1], v);
                ok = ok && fabsf(dx) < radius && fabsf(dy) < radius;
                if (ok && s_claim) ok = !(s_claim[k] < i);
                if (ok && F.UR) {
                    const float urk = F.UR[k];
                    ok = !(urk > 0.f && fabsf(__fsub_rn(w.ur, urk)) > radius);        // (:100-105, :1405-1411)
                }
                if (ok) {
                    const uint4 da = F.D0[2 * k], dc = F.D0[2 * k + 1];
                    const unsigned dist = __popc(qa.x ^ da.x) + __popc(qa.y ^ da.y) + __popc(qa.z ^ da.z) + __popc(qa.w ^ da.w) +
                                          __popc(qc.x ^ dc.x) + __popc(qc.y ^ dc.y) + __popc(qc.z ^ dc.z) + __popc(qc.w ^ dc.w);
                    key = (dist << 16) | (unsigned)p;
                    if (s_dist && s_dist[p] <= (int)dist) key = 0xffffffffu;
                }
            }
            best2 = min(best2, max(best, key));
            best = min(best, key);
            if (buf) {
                const bool pass = (key >> 16) <= (unsigned)P.list_th;
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const int slot = tot + __popc(m & ((1u << lane) - 1u));
                    if (slot < lc) buf[slot] = key;
                }
                tot += __popc(m);
            }
        }
    }
    if (total) *total = tot;
    const unsigned m1 = __reduce_min_sync(0xffffffffu, best);
    if (second) *second = __reduce_min_sync(0xffffffffu, best == m1 ? best2 : best);     // keys are unique (positions differ)
    return m1;
}

// Decision of the local-map variant from the best and second-best (distance, level) (src/ORBmatcher.cc:117-127).
__device__ __forceinline__ bool sp_accept(const SpParams& P, unsigned d1, int lvl1, unsigned d2, int lvl2) {
    if (d1 > SP_TH_HIGH) return false;
    return !(lvl1 == lvl2 && (float)d1 > __fmul_rn(P.nnratio, (float)d2));
}

// Pass 1, state-free and GPU-wide: a warp per (query, map point) lists the point's candidates of interest
// (distance <= list_th) in the reference's preference order -- ascending (distance, visiting position) -- as
// distance << 16 | keypoint index.  Whatever the claims turn out to be, the point's decision only looks at the first
// unclaimed entries of this list, so the sequential part (pass 2) never touches a descriptor again.
// cand_count > lc marks a list that did not fit.
#define SP_LIST_WARPS 8
template <bool LOCAL>
__global__ void __launch_bounds__(SP_LIST_WARPS * 32)
search_projection_list_kernel(const OrbxPlan* __restrict__ plan, const SpQuery* __restrict__ queries, SpParams P, int nq, int lc,
                              const float* __restrict__ world, const uint4* __restrict__ mp_desc, const int* __restrict__ mp_obs,
                              const int* __restrict__ last_octave, const float* __restrict__ last_angle, const float* __restrict__ kp,
                              const uint8_t* __restrict__ desc, const float* __restrict__ xy_un, const int* __restrict__ cell_start,
                              const int* __restrict__ cell_items, const float* __restrict__ u_right, uint32_t* __restrict__ cand_list,
                              int* __restrict__ cand_count) {
    __shared__ unsigned s_buf[SP_LIST_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long slot = (long long)blockIdx.x * SP_LIST_WARPS + warp;
    if (slot >= (long long)nq * P.cap) return;
    const int qi = (int)(slot / P.cap), i = (int)(slot % P.cap);
    const SpQuery q = queries[qi];
    if (i >= q.n_last) return;
    const size_t qb = (size_t)qi * P.cap;
    int total = 0;
    if (mp_obs[qb + i] >= 0) {                           // has a map point, no outlier (:1356-1360) / in view, not bad (:54-58)
        const SpFrame F = sp_frame(plan, P, q.frame, kp, desc, xy_un, cell_start, cell_items, u_right);
        SpWin w;
        const bool ok = LOCAL ? sp_track_window(plan, P, world + qb * 3, last_octave + qb, last_angle + qb, i, w)
                              : sp_project(plan, q, P, world + qb * 3, last_octave + qb, i, w);
        if (ok) sp_window(P, F, w, mp_desc + qb * 2, i, lane, nullptr, s_buf[warp], lc, &total, nullptr);
        __syncwarp();
        if (total > 0 && total <= lc) {
            // bitonic sort of the (at most 32) keys across the warp, ascending
            unsigned v = lane < total ? s_buf[warp][lane] : 0xffffffffu;
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const unsigned o = __shfl_xor_sync(0xffffffffu, v, j);
                    const bool up = (lane & k) == 0, lower = (lane & j) == 0;
                    v = (lower == up) ? min(v, o) : max(v, o);
                }
            if (lane < total) cand_list[(qb + i) * (size_t)lc + lane] = (v & 0xffff0000u) | (unsigned)F.IT[v & 0xffffu];
        }
    }
    if (lane == 0) cand_count[qb + i] = total;
}

// Pass 2: one CTA per query resolves the claims (see the header comment).  LastFrame variant: then histogram, maxima
// and culling.  Local-map variant (LOCAL): keypoints that already hold an observed map point (cur_obs > 0) are claimed
// from the start, the decision takes the two best unclaimed candidates (ratio test, :117-121), no orientation check.
template <bool LOCAL>
__global__ void __launch_bounds__(1024)
search_projection_kernel(const OrbxPlan* __restrict__ plan, const SpQuery* __restrict__ queries, SpParams P, int lc,
                         const float* __restrict__ world, const uint4* __restrict__ mp_desc, const int* __restrict__ mp_obs,
                         const int* __restrict__ last_octave, const float* __restrict__ last_angle,
                         const float* __restrict__ kp, const uint8_t* __restrict__ desc, const int* __restrict__ kept_counts,
                         const float* __restrict__ xy_un, const int* __restrict__ cell_start, const int* __restrict__ cell_items,
                         const float* __restrict__ u_right, const int* __restrict__ cur_obs, const uint32_t* __restrict__ cand_list,
                         const int* __restrict__ cand_count, int* __restrict__ match_out, int* __restrict__ stats_out) {
    extern __shared__ int sp_smem[];
    __shared__ int s_hist[SP_HISTO];
    __shared__ int s_changed, s_nm, s_ncull, s_nover;
    __shared__ unsigned s_keep;
    const SpQuery q = queries[blockIdx.x];
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    const int f = q.frame, nL = q.n_last;
    int N = 0;
    for (int l = 0; l < nl; ++l) N += kept_counts[f * nl + l];
    int* s_match = sp_smem;                   // [cap]  current keypoint chosen by map point i, -1: none
    int* s_claim = sp_smem + P.cap;           // [kpf]  first j with obs > 0 that matched the keypoint; later: its holder
    const size_t qb = (size_t)blockIdx.x * P.cap;
    const float* W = world + qb * 3;
    const uint4* QD = mp_desc + qb * 2;
    const int* OBS = mp_obs + qb;
    const int* LO = last_octave + qb;
    const float* LA = last_angle + qb;
    const uint32_t* CL = cand_list + qb * (size_t)lc;
    const int* CC = cand_count + qb;
    const int* CO = cur_obs ? cur_obs + (size_t)blockIdx.x * kpf : nullptr;     // what the frame's keypoints hold at entry (local-map and KeyFrame variants)
    const SpFrame F = sp_frame(plan, P, f, kp, desc, xy_un, cell_start, cell_items, u_right);
    const float* K0 = F.K0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    if (threadIdx.x == 0) { s_changed = 0; s_nm = 0; s_ncull = 0; s_keep = 0xffffffffu; s_nover = 0; }
    if (threadIdx.x < SP_HISTO) s_hist[threadIdx.x] = 0;
    __syncthreads();
    int over = 0;
    for (int i = threadIdx.x; i < nL; i += blockDim.x) { s_match[i] = -1; over += CC[i] > lc; }
    if (over) atomicAdd(&s_nover, over);
    for (int k = threadIdx.x; k < kpf; k += blockDim.x) s_claim[k] = (CO && k < N && CO[k] > 0) ? -1 : INT_MAX;   // (:96-98, :1540-1541)
    __syncthreads();
    const int nover = s_nover;

    int rounds = 0;
    for (;;) {
        for (int i = threadIdx.x; i < nL; i += blockDim.x) {               // a thread per point: first unclaimed list entries
            const int c = CC[i];
            if (c > lc) continue;
            int newm = -1;
            if (!LOCAL) {
                for (int j = 0; j < c; ++j) {
                    const unsigned e = CL[(size_t)i * lc + j];
                    const int k = (int)(e & 0xffffu);
                    if (!(s_claim[k] < i)) { newm = (e >> 16) <= (unsigned)P.match_th ? k : -1; break; }
                }
            } else {
                unsigned e1 = 0xffffffffu, e2 = 0xffffffffu;
                for (int j = 0; j < c; ++j) {
                    const unsigned e = CL[(size_t)i * lc + j];
                    if (s_claim[e & 0xffffu] < i) continue;
                    if (e1 == 0xffffffffu) e1 = e; else { e2 = e; break; }
                }
                if (e1 != 0xffffffffu) {
                    const int k1 = (int)(e1 & 0xffffu);
                    // a runner-up outside the list has distance > list_th >= TH_HIGH / mfNNratio and cannot reject
                    const unsigned d2 = e2 == 0xffffffffu ? 0xffffu : (e2 >> 16);
                    const int lvl2 = e2 == 0xffffffffu ? -1 : __float_as_int(K0[(size_t)(e2 & 0xffffu) * 7 + 5]);
                    if (sp_accept(P, e1 >> 16, __float_as_int(K0[(size_t)k1 * 7 + 5]), d2, lvl2)) newm = k1;
                }
            }
            if (s_match[i] != newm) { s_match[i] = newm; s_changed = 1; }
        }
        if (nover)                                                          // lists that did not fit: the full search, a warp per point
            for (int i = warp; i < nL; i += nwarps) {
                if (CC[i] <= lc) continue;
                SpWin w;
                const bool ok = LOCAL ? sp_track_window(plan, P, W, LO, LA, i, w) : sp_project(plan, q, P, W, LO, i, w);
                unsigned best = 0xffffffffu, second = 0xffffffffu;
                if (ok) best = sp_window(P, F, w, QD, i, lane, s_claim, nullptr, 0, nullptr, &second);
                int newm = -1;
                if (best != 0xffffffffu) {
                    const int k1 = F.IT[best & 0xffffu];
                    if (!LOCAL) newm = (best >> 16) <= (unsigned)P.match_th ? k1 : -1;         // (:1425, :1555)
                    else {
                        // no runner-up: bestDist2 stays 256 and bestLevel2 -1 (:77-81)
                        const unsigned d2 = second == 0xffffffffu ? 256u : (second >> 16);
                        const int lvl2 = second == 0xffffffffu ? -1 : __float_as_int(K0[(size_t)F.IT[second & 0xffffu] * 7 + 5]);
                        if (sp_accept(P, best >> 16, __float_as_int(K0[(size_t)k1 * 7 + 5]), d2, lvl2)) newm = k1;
                    }
                }
                if (lane == 0 && s_match[i] != newm) { s_match[i] = newm; s_changed = 1; }
            }
        __syncthreads();
        const int changed = s_changed;
        ++rounds;
        __syncthreads();
        if (!changed) break;
        if (threadIdx.x == 0) s_changed = 0;
        for (int k = threadIdx.x; k < kpf; k += blockDim.x) s_claim[k] = (CO && k < N && CO[k] > 0) ? -1 : INT_MAX;
        __syncthreads();
        for (int i = threadIdx.x; i < nL; i += blockDim.x) {
            const int m = s_match[i];
            if (m >= 0 && OBS[i] > 0) atomicMin(&s_claim[m], i);
        }
        __syncthreads();
    }

    // ---- rotation histogram and three maxima (:1428-1441, :1446-1453, :1601-1642)
    const int check_ori = LOCAL ? 0 : P.check_ori;
    int mine = 0;
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int m = s_match[i];
        if (m < 0) continue;
        ++mine;
        if (check_ori) atomicAdd(&s_hist[sp_rot_bin(LA[i], K0[(size_t)m * 7 + 3])], 1);
    }
    if (mine) atomicAdd(&s_nm, mine);
    for (int k = threadIdx.x; k < kpf; k += blockDim.x) s_claim[k] = -1;      // from here on: the holder of keypoint k
    __syncthreads();
    if (threadIdx.x == 0 && check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int b = 0; b < SP_HISTO; ++b) {
            const int s = s_hist[b];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
            else if (s > max3) { max3 = s; ind3 = b; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        unsigned keep = 0;
        if (ind1 >= 0) keep |= 1u << ind1;
        if (ind2 >= 0) keep |= 1u << ind2;
        if (ind3 >= 0) keep |= 1u << ind3;
        s_keep = keep;
    }
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int m = s_match[i];
        if (m >= 0) atomicMax(&s_claim[m], i);                                 // the last writer holds the keypoint (:124, :1427)
    }
    __syncthreads();
    if (check_ori) {
        const unsigned keep = s_keep;
        int culled = 0;
        for (int i = threadIdx.x; i < nL; i += blockDim.x) {
            const int m = s_match[i];
            if (m < 0) continue;
            if (!((keep >> sp_rot_bin(LA[i], K0[(size_t)m * 7 + 3])) & 1u)) { s_claim[m] = -1; ++culled; }   // (:1455-1463)
        }
        if (culled) atomicAdd(&s_ncull, culled);
    }
    __syncthreads();
    int* out = match_out + (size_t)blockIdx.x * kpf;
    for (int k = threadIdx.x; k < N; k += blockDim.x) out[k] = s_claim[k];
    if (threadIdx.x == 0) {
        stats_out[2 * blockIdx.x] = s_nm - s_ncull;
        stats_out[2 * blockIdx.x + 1] = rounds;
    }
}

// =====================================================================================
// ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:405-520),
// the matcher of Tracking::MonocularInitialization (src/Tracking.cc:600).  F2 is the handle's device-resident frame
// (keypoints, descriptors, mvKeysUn / mGrid from orbx_undistort_grid), F1 -- the initial frame the caller keeps -- comes
// staged: octave, angle and descriptor of its undistorted keypoints and vbPrevMatched.
//
// The reference loop over F1's level-0 keypoints is sequential in a stronger way than the projection matchers: a later
// keypoint STEALS an F2 keypoint from an earlier one when its distance is strictly smaller (vMatchedDistance, :447-448,
// :471-478), and the ratio test only sees the candidates that survive that filter.  So:
//   init_list_kernel    state-free, GPU-wide, a warp per F1 keypoint: every candidate of its window (Frame::GetFeaturesInArea
//                       at level 0, :428) with its Hamming distance, in the reference's visiting order, as distance << 16 |
//                       CSR position;
//   init_resolve_kernel one warp per query replays the loop in order against vMatchedDistance / vnMatches21 in shared memory
//                       (indexed by CSR position): per keypoint one or two 32-entry list chunks, no descriptor is touched
//                       again; the next keypoint's chunk is in flight while the current one is decided.  Lists that did not
//                       fit are re-scanned from the grid.  Then rotation histogram, three maxima, culling (:489-512) and
//                       the vbPrevMatched update (:515-517).
// =====================================================================================
struct InitQuery {
    int n1, frame;
};
#define SP_TH_LOW 50                      // ORBmatcher::TH_LOW (src/ORBmatcher.cc:38)

__global__ void __launch_bounds__(SP_LIST_WARPS * 32)
init_list_kernel(const OrbxPlan* __restrict__ plan, const InitQuery* __restrict__ queries, SpParams P, int nq, int lc,
                 const float* __restrict__ prev, const uint4* __restrict__ desc1, const int* __restrict__ octave1,
                 const float* __restrict__ kp, const uint8_t* __restrict__ desc, const float* __restrict__ xy_un,
                 const int* __restrict__ cell_start, const int* __restrict__ cell_items, uint32_t* __restrict__ cand_list,
                 int* __restrict__ cand_count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long slot = (long long)blockIdx.x * SP_LIST_WARPS + warp;
    if (slot >= (long long)nq * P.cap) return;
    const int qi = (int)(slot / P.cap), i = (int)(slot % P.cap);
    const InitQuery q = queries[qi];
    if (i >= q.n1) return;
    const size_t qb = (size_t)qi * P.cap;
    int total = 0;
    if (octave1[qb + i] <= 0) {                                           // (:424-426) level1 > 0: continue
        const SpFrame F = sp_frame(plan, P, q.frame, kp, desc, xy_un, cell_start, cell_items, nullptr);
        SpWin w;
        w.u = prev[2 * (qb + i)];
        w.v = prev[2 * (qb + i) + 1];
        w.radius = P.th;                                                  // windowSize, converted to float at the call (:428)
        w.ur = 0.f;
        w.minL = w.maxL = octave1[qb + i];                                // (level1, level1)
        if (w.u == w.u && w.v == w.v)                                     // NaN: undefined in the reference, no candidates here
            sp_window(P, F, w, desc1 + qb * 2, i, lane, nullptr, cand_list + (qb + i) * (size_t)lc, lc, &total, nullptr);
    }
    if (lane == 0) cand_count[qb + i] = total;
}

__global__ void __launch_bounds__(32)
init_resolve_kernel(const OrbxPlan* __restrict__ plan, const InitQuery* __restrict__ queries, SpParams P, int lc,
                    const float* __restrict__ prev, const uint4* __restrict__ desc1, const int* __restrict__ octave1,
                    const float* __restrict__ angle1, const float* __restrict__ kp, const uint8_t* __restrict__ desc,
                    const int* __restrict__ kept_counts, const float* __restrict__ xy_un, const int* __restrict__ cell_start,
                    const int* __restrict__ cell_items, const uint32_t* __restrict__ cand_list, const int* __restrict__ cand_count,
                    int* __restrict__ match_out, float* __restrict__ prev_out, uint8_t* __restrict__ bin_scratch,
                    int* __restrict__ stats_out) {
    extern __shared__ int init_smem[];
    __shared__ int s_hist[SP_HISTO];
    const InitQuery q = queries[blockIdx.x];
    const int kpf = plan->kept_per_frame;
    const int n1 = q.n1, lane = threadIdx.x;
    int* s_dist = init_smem;                 // [kpf] vMatchedDistance, by CSR position
    int* s_owner = init_smem + kpf;          // [kpf] vnMatches21
    const size_t qb = (size_t)blockIdx.x * P.cap;
    const uint32_t* CL = cand_list + qb * (size_t)lc;
    const int* CC = cand_count + qb;
    int* M12 = match_out + qb;               // vnMatches12 as CSR positions until the end
    uint8_t* BIN = bin_scratch + qb;
    const SpFrame F = sp_frame(plan, P, q.frame, kp, desc, xy_un, cell_start, cell_items, nullptr);
    for (int k = lane; k < kpf; k += 32) { s_dist[k] = INT_MAX; s_owner[k] = -1; }
    if (lane < SP_HISTO) s_hist[lane] = 0;
    int* PACC = reinterpret_cast<int*>(prev_out + 2 * qb);    // [2 i1]: F2 keypoint (CSR position) i1 had when it was accepted, -1: never
    for (int i = lane; i < n1; i += 32) { M12[i] = -1; BIN[i] = 255; PACC[2 * i] = -1; }
    __syncwarp();
    int nmatches = 0;
    // The loop is a chain of short iterations, so what it waits for is memory latency: the candidate counts are fetched 32
    // keypoints at a time (lane k holds the count of keypoint 32 j + k) and the first list chunk of a keypoint is requested
    // INIT_AHEAD iterations before it is needed.
    constexpr int INIT_AHEAD = 4;
    uint32_t ering[INIT_AHEAD];
#pragma unroll
    for (int a = 0; a < INIT_AHEAD; ++a) ering[a] = a < n1 ? CL[(size_t)a * lc + lane] : 0u;
    int cnt32 = 0;
    for (int i1 = 0; i1 < n1; ++i1) {
        if ((i1 & 31) == 0) cnt32 = i1 + lane < n1 ? CC[i1 + lane] : 0;
        const int cnt = __shfl_sync(0xffffffffu, cnt32, i1 & 31);
        const uint32_t e0 = ering[0];
#pragma unroll
        for (int a = 0; a + 1 < INIT_AHEAD; ++a) ering[a] = ering[a + 1];
        ering[INIT_AHEAD - 1] = i1 + INIT_AHEAD < n1 ? CL[(size_t)(i1 + INIT_AHEAD) * lc + lane] : 0u;
        if (cnt == 0) continue;                                           // higher octave or empty window (:424-431)
        unsigned m1, m2;
        if (cnt <= lc) {
            unsigned best = 0xffffffffu, best2 = 0xffffffffu;
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int j = j0 + lane;
                unsigned key = 0xffffffffu;
                if (j < cnt) {
                    const uint32_t e = j0 == 0 ? e0 : CL[(size_t)i1 * lc + j];
                    if (s_dist[e & 0xffffu] > (int)(e >> 16)) key = e;    // (:447-448)
                }
                best2 = min(best2, max(best, key));
                best = min(best, key);
            }
            m1 = __reduce_min_sync(0xffffffffu, best);
            m2 = __reduce_min_sync(0xffffffffu, best == m1 ? best2 : best);
        } else {                                                          // the list did not fit: scan the window again
            SpWin w;
            w.u = prev[2 * (qb + i1)];
            w.v = prev[2 * (qb + i1) + 1];
            w.radius = P.th;
            w.ur = 0.f;
            w.minL = w.maxL = octave1[qb + i1];
            m1 = sp_window(P, F, w, desc1 + qb * 2, i1, lane, nullptr, nullptr, 0, nullptr, &m2, s_dist);
        }
        if (m1 == 0xffffffffu) continue;
        const int d1 = (int)(m1 >> 16), p = (int)(m1 & 0xffffu);
        // bestDist2 stays INT_MAX without a runner-up (:439); (float)bestDist2 * mfNNratio in float (:469)
        const float lim = __fmul_rn(m2 == 0xffffffffu ? (float)INT_MAX : (float)(int)(m2 >> 16), P.nnratio);
        if (d1 <= SP_TH_LOW && (float)d1 < lim) {
            if (lane == 0) {
                const int old = s_owner[p];
                if (old >= 0) M12[old] = -1;                              // (:471-475)
                M12[i1] = p;
                s_owner[p] = i1;
                s_dist[p] = d1;
                PACC[2 * i1] = p;                                         // what rotHist receives (:481-490) is settled after the loop:
            }                                                             // its two dependent global loads would sit on this chain
            __syncwarp();                                                 // (nmatches is counted at the end: every steal is -1 + 1)
        }
    }
    __syncwarp();
    __threadfence_block();
    unsigned keep = 0xffffffffu;
    if (P.check_ori) {                                                    // rotHist: every keypoint that was ever accepted stays in its bin (:481-490)
        for (int i = lane; i < n1; i += 32) {
            const int p = PACC[2 * i];
            if (p < 0) continue;
            const int b = sp_rot_bin(angle1[qb + i], F.K0[(size_t)F.IT[p] * 7 + 3]);
            atomicAdd(&s_hist[b], 1);
            BIN[i] = (uint8_t)b;
        }
        __syncwarp();
    }
    if (P.check_ori) {                                                    // ComputeThreeMaxima (:1601-1642), every lane the same scalars
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int b = 0; b < SP_HISTO; ++b) {
            const int sz = s_hist[b];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = b; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = b; }
            else if (sz > max3) { max3 = sz; ind3 = b; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        keep = 0;
        if (ind1 >= 0) keep |= 1u << ind1;
        if (ind2 >= 0) keep |= 1u << ind2;
        if (ind3 >= 0) keep |= 1u << ind3;
    }
    for (int i = lane; i < n1; i += 32) {
        int p = M12[i];
        if (p >= 0 && P.check_ori && BIN[i] < SP_HISTO && !((keep >> BIN[i]) & 1u)) p = -1;     // (:497-511)
        const int k = p >= 0 ? F.IT[p] : -1;
        M12[i] = k;
        float px = prev[2 * (qb + i)], py = prev[2 * (qb + i) + 1];
        if (k >= 0) { px = F.XY[2 * k]; py = F.XY[2 * k + 1]; ++nmatches; }                       // (:515-517)
        prev_out[2 * (qb + i)] = px;
        prev_out[2 * (qb + i) + 1] = py;
    }
    nmatches = __reduce_add_sync(0xffffffffu, nmatches);
    if (lane == 0) stats_out[blockIdx.x] = nmatches;
}

// =====================================================================================
// Frame::ComputeBoW (src/Frame.cc:395-402) = DBoW2 TemplatedVocabulary::transform(features, BowVector&, FeatureVector&,
// levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194, descent :1217-1259, FORB::distance FORB.cpp:81-101,
// BowVector::addWeight / normalize BowVector.cpp:34-46, :62-84, FeatureVector::addFeature FeatureVector.cpp:31-45) for
// the ORB vocabulary's TF-IDF weighting and L1 scoring, on the descriptors the extraction left in HBM.
//   bow_descend_kernel : 16 lanes per descriptor walk the tree; at each node the lanes take the children, Hamming
//                        distance by POPC, "first strictly smaller wins" (:1244-1248) = minimum of (distance << 20 | child
//                        position).  Output: the leaf node and the node at level L - levelsup.
//   bow_reduce_kernel  : one CTA per frame.  Bitonic sort of (word id, leaf) in shared memory -> distinct words and
//                        their multiplicities; a word's value is its weight added c times IN SEQUENCE (what c addWeight
//                        calls produce), the L1 norm is the sequential sum in word order (map order), both in double
//                        with __dadd_rn so the bits are the reference's; then (node, feature) keys are sorted for the
//                        FeatureVector (map order, push_back order).  Stopped words (weight 0) drop out of both (:1157).
// =====================================================================================
struct BowVoc {
    const int* child_start;
    const int* child_items;
    const uint4* node_desc;
    const double* node_weight;
    const int* node_word;
    int n_nodes, L;
};
#define BOW_GROUP 16
#define BOW_THREADS 256

__global__ void __launch_bounds__(BOW_THREADS)
bow_descend_kernel(const OrbxPlan* __restrict__ plan, BowVoc V, const int* __restrict__ frames, int nframes, int levelsup,
                   const uint8_t* __restrict__ desc, const int* __restrict__ kept_counts, int* __restrict__ leaf_out,
                   int* __restrict__ nid_out) {
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    const long long g = ((long long)blockIdx.x * BOW_THREADS + threadIdx.x) / BOW_GROUP;
    if (g >= (long long)nframes * kpf) return;                               // whole groups leave together
    const int fi = (int)(g / kpf), i = (int)(g % kpf);
    const int f = frames[fi];
    int N = 0;
    for (int l = 0; l < nl; ++l) N += kept_counts[f * nl + l];
    if (i >= N) return;
    const int sub = threadIdx.x & (BOW_GROUP - 1);
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);
    const uint4* D = reinterpret_cast<const uint4*>(desc + ((size_t)f * kpf + i) * 32);
    const uint4 qa = D[0], qc = D[1];
    const int nid_level = V.L - levelsup;
    int nid = 0, final_id = 0, level = 0;                                    // root when nid_level <= 0 (:1227)
    for (int guard = 0; guard < 64; ++guard) {
        ++level;
        const int cs = V.child_start[final_id], ce = V.child_start[final_id + 1];
        unsigned best = 0xffffffffu;
        for (int c = cs + sub; c < ce; c += BOW_GROUP) {
            const int id = V.child_items[c];
            const uint4 da = V.node_desc[2 * (size_t)id], dc = V.node_desc[2 * (size_t)id + 1];
            const unsigned dist = __popc(qa.x ^ da.x) + __popc(qa.y ^ da.y) + __popc(qa.z ^ da.z) + __popc(qa.w ^ da.w) +
                                  __popc(qc.x ^ dc.x) + __popc(qc.y ^ dc.y) + __popc(qc.z ^ dc.z) + __popc(qc.w ^ dc.w);
            best = min(best, (dist << 20) | (unsigned)(c - cs));
        }
#pragma unroll
        for (int o = BOW_GROUP / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(gmask, best, o, BOW_GROUP));
        final_id = V.child_items[cs + (int)(best & 0xfffffu)];
        if (level == nid_level) nid = final_id;
        if (V.child_start[final_id + 1] == V.child_start[final_id]) break;   // isLeaf (:1254)
    }
    if (sub == 0) {
        leaf_out[(size_t)fi * kpf + i] = final_id;
        nid_out[(size_t)fi * kpf + i] = nid;
    }
}

__device__ __forceinline__ void bow_bitonic_sort(unsigned long long* keys, int n /* power of two */) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const unsigned long long a = keys[t], b = keys[p];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { keys[t] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(1024)
bow_reduce_kernel(const OrbxPlan* __restrict__ plan, BowVoc V, const int* __restrict__ frames, int sort_n,
                  const int* __restrict__ kept_counts, const int* __restrict__ leaf_in, const int* __restrict__ nid_in,
                  unsigned* __restrict__ word_ids, double* __restrict__ word_values, unsigned* __restrict__ fv_nodes,
                  unsigned* __restrict__ fv_features, int* __restrict__ counts_out) {
    extern __shared__ unsigned long long bow_keys[];
    __shared__ int s_warp[33];
    __shared__ int s_nwords, s_nfeat;
    __shared__ double s_norm;
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    const int fi = blockIdx.x, f = frames[fi];
    int N = 0;
    for (int l = 0; l < nl; ++l) N += kept_counts[f * nl + l];
    // sort only as many slots as this frame needs (kept_per_frame is often just above a power of two)
    int need = 32;
    while (need < N) need <<= 1;
    sort_n = min(sort_n, need);
    const int* leaf = leaf_in + (size_t)fi * kpf;
    const int* nid = nid_in + (size_t)fi * kpf;
    unsigned* wid = word_ids + (size_t)fi * kpf;
    double* wval = word_values + (size_t)fi * kpf;
    unsigned* fn = fv_nodes + (size_t)fi * kpf;
    unsigned* ff = fv_features + (size_t)fi * kpf;
    const unsigned long long NONE = ~0ull;

    // ---- BowVector: distinct words with multiplicities
    for (int i = threadIdx.x; i < sort_n; i += blockDim.x) {
        unsigned long long key = NONE;
        if (i < N) {
            const int lf = leaf[i];
            if (V.node_weight[lf] > 0.0) key = ((unsigned long long)(unsigned)V.node_word[lf] << 32) | (unsigned)lf;   // (:1157)
        }
        bow_keys[i] = key;
    }
    __syncthreads();
    bow_bitonic_sort(bow_keys, sort_n);
    // heads of runs -> output slots (block scan over chunks of blockDim)
    int base = 0;
    for (int i0 = 0; i0 < sort_n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const unsigned long long key = i < sort_n ? bow_keys[i] : NONE;
        const int head = key != NONE && (i == 0 || bow_keys[i - 1] != key);
        int tot;
        const int e = block_excl_scan(head, &tot, s_warp);
        if (head) {
            int c = 1;
            while (i + c < sort_n && bow_keys[i + c] == key) ++c;              // multiplicity (runs are short)
            const double w = V.node_weight[(unsigned)(key & 0xffffffffu)];
            double s = w;
            for (int r = 1; r < c; ++r) s = __dadd_rn(s, w);                    // c addWeight calls (BowVector.cpp:40)
            wid[base + e] = (unsigned)(key >> 32);
            wval[base + e] = s;
        }
        base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) s_nwords = base;
    __syncthreads();
    const int nwords = s_nwords;
    if (threadIdx.x == 0) {                                                     // BowVector::normalize(L1): map order, in sequence
        double norm = 0.0;
        for (int j = 0; j < nwords; ++j) norm = __dadd_rn(norm, fabs(wval[j]));
        s_norm = norm;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0)
        for (int j = threadIdx.x; j < nwords; j += blockDim.x) wval[j] = __ddiv_rn(wval[j], norm);

    // ---- FeatureVector: (node at level L - levelsup, feature index) in map / push_back order
    __syncthreads();
    for (int i = threadIdx.x; i < sort_n; i += blockDim.x) {
        unsigned long long key = NONE;
        if (i < N && V.node_weight[leaf[i]] > 0.0) key = ((unsigned long long)(unsigned)nid[i] << 32) | (unsigned)i;
        bow_keys[i] = key;
    }
    __syncthreads();
    bow_bitonic_sort(bow_keys, sort_n);
    int nfeat = 0;
    for (int i = threadIdx.x; i < sort_n; i += blockDim.x) {
        const unsigned long long key = bow_keys[i];
        if (key != NONE) { fn[i] = (unsigned)(key >> 32); ff[i] = (unsigned)(key & 0xffffffffu); ++nfeat; }   // valid keys sort first
    }
    if (threadIdx.x == 0) s_nfeat = 0;
    __syncthreads();
    if (nfeat) atomicAdd(&s_nfeat, nfeat);
    __syncthreads();
    if (threadIdx.x == 0) { counts_out[2 * fi] = nwords; counts_out[2 * fi + 1] = s_nfeat; }
}

// =====================================================================================
// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (src/ORBmatcher.cc:159-288): matching restricted to
// features under the same vocabulary node.  F's FeatureVector is what orbx_compute_bow left in HBM, its descriptors and
// angles what the extraction left; the KeyFrame side (descriptors, map-point validity, angles, FeatureVector) comes from
// the caller.  The reference's claim rule -- a frame feature that has been matched is skipped by every later KeyFrame
// feature (:213-214) -- only couples features of ONE node, so a warp owns a node: it walks the node's KeyFrame features
// in order (sequential, as the reference), the lanes take the node's unclaimed frame features in parallel (8 POPC
// each), best / second best by (distance, position) = the reference's strict "<" updates, TH_LOW and the mfNNratio
// test (:232-236), claim.  Nodes are independent, so warps and CTAs (one per query) run concurrently.  Then the
// rotation histogram, three maxima and culling (:240-284) as in SearchByProjection.
// =====================================================================================
struct BowMatchQuery {
    int frame, slot, n_kf, n_kf_fv;
};
#define SB_TH_LOW 50                      // ORBmatcher::TH_LOW (src/ORBmatcher.cc:38)

__global__ void __launch_bounds__(1024)
search_bow_kernel(const OrbxPlan* __restrict__ plan, const BowMatchQuery* __restrict__ queries, int cap, float nnratio, int check_ori,
                  const uint4* __restrict__ kf_desc, const uint8_t* __restrict__ kf_valid, const float* __restrict__ kf_angle,
                  const unsigned* __restrict__ kf_fv_nodes, const unsigned* __restrict__ kf_fv_features,
                  const float* __restrict__ kp, const uint8_t* __restrict__ desc, const int* __restrict__ kept_counts,
                  const unsigned* __restrict__ f_fv_nodes, const unsigned* __restrict__ f_fv_features, const int* __restrict__ bow_counts,
                  int* __restrict__ match_out, int* __restrict__ stats_out) {
    extern __shared__ int sb_state[];           // per position of F's FeatureVector: -1 unclaimed, else the match's rotation bin (30: none)
    __shared__ int s_hist[SP_HISTO];
    __shared__ int s_nm, s_ncull;
    __shared__ unsigned s_keep;
    const BowMatchQuery q = queries[blockIdx.x];
    const int kpf = plan->kept_per_frame, nl = plan->nlevels;
    int N = 0;
    for (int l = 0; l < nl; ++l) N += kept_counts[q.frame * nl + l];
    const int nF = bow_counts[2 * q.slot + 1];
    const unsigned* FN = f_fv_nodes + (size_t)q.slot * kpf;
    const unsigned* FF = f_fv_features + (size_t)q.slot * kpf;
    const size_t qb = (size_t)blockIdx.x * cap;
    const uint4* KD = kf_desc + qb * 2;
    const uint8_t* KV = kf_valid + qb;
    const float* KA = kf_angle + qb;
    const unsigned* KN = kf_fv_nodes + qb;
    const unsigned* KF = kf_fv_features + qb;
    const float* K0 = kp + (size_t)q.frame * kpf * 7;
    const uint4* D0 = reinterpret_cast<const uint4*>(desc + (size_t)q.frame * kpf * 32);
    int* out = match_out + (size_t)blockIdx.x * kpf;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    for (int p = threadIdx.x; p < kpf; p += blockDim.x) sb_state[p] = -1;
    for (int k = threadIdx.x; k < N; k += blockDim.x) out[k] = -1;            // vector<MapPoint*>(F.N, NULL) (:163)
    if (threadIdx.x < SP_HISTO) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_nm = 0; s_ncull = 0; s_keep = 0xffffffffu; }
    __syncthreads();

    int nm = 0;
    for (int i = warp; i < q.n_kf_fv; i += nwarps) {
        const unsigned node = KN[i];
        if (i > 0 && KN[i - 1] == node) continue;                             // not the head of a node's run
        int lo = 0, hi = nF;                                                   // lower_bound of the node in F's FeatureVector
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (FN[mid] < node) lo = mid + 1; else hi = mid; }
        const int a = lo;
        if (a >= nF || FN[a] != node) continue;                                // (:275-282) no common node
        hi = nF;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (FN[mid] <= node) lo = mid + 1; else hi = mid; }
        const int b = lo;
        // each lane keeps the descriptor of its first frame feature of the node in registers for the whole run (nodes rarely
        // hold more than 32 frame features), and the next KeyFrame descriptor is requested before the current one is used
        uint4 ca = make_uint4(0, 0, 0, 0), cc = ca;
        if (a + lane < b) { const unsigned k = FF[a + lane]; ca = D0[2 * (size_t)k]; cc = D0[2 * (size_t)k + 1]; }
        int ikf_n = (int)KF[i];
        uint4 na = KD[2 * ikf_n], nc = KD[2 * ikf_n + 1];
        for (int j = i; j < q.n_kf_fv && KN[j] == node; ++j) {
            const int ikf = ikf_n;
            const uint4 qa = na, qc = nc;
            if (j + 1 < q.n_kf_fv) { ikf_n = (int)KF[j + 1]; na = KD[2 * ikf_n]; nc = KD[2 * ikf_n + 1]; }
            if (KV[ikf] != 1) continue;                                        // no map point, or a bad one (:196-202)
            unsigned best = 0xffffffffu, best2 = 0xffffffffu;
            for (int pb = a; pb < b; pb += 32) {
                const int p = pb + lane;
                unsigned key = 0xffffffffu;
                if (p < b && sb_state[p] < 0) {                                // vpMapPointMatches[realIdxF] still NULL (:213-214)
                    uint4 da = ca, dc = cc;
                    if (pb != a) { const unsigned k = FF[p]; da = D0[2 * (size_t)k]; dc = D0[2 * (size_t)k + 1]; }
                    const unsigned dist = __popc(qa.x ^ da.x) + __popc(qa.y ^ da.y) + __popc(qa.z ^ da.z) + __popc(qa.w ^ da.w) +
                                          __popc(qc.x ^ dc.x) + __popc(qc.y ^ dc.y) + __popc(qc.z ^ dc.z) + __popc(qc.w ^ dc.w);
                    key = (dist << 16) | (unsigned)p;
                }
                best2 = min(best2, max(best, key));
                best = min(best, key);
            }
            const unsigned m1 = __reduce_min_sync(0xffffffffu, best);
            const unsigned m2 = __reduce_min_sync(0xffffffffu, best == m1 ? best2 : best);
            if (m1 != 0xffffffffu) {
                const unsigned d1 = m1 >> 16, d2 = m2 == 0xffffffffu ? 256u : (m2 >> 16);
                if (d1 <= SB_TH_LOW && (float)d1 < __fmul_rn(nnratio, (float)d2)) {     // (:232-236)
                    const int p1 = (int)(m1 & 0xffffu);
                    const int iF = (int)FF[p1];
                    const int bin = check_ori ? sp_rot_bin(KA[ikf], K0[(size_t)iF * 7 + 3]) : SP_HISTO;
                    if (lane == 0) {
                        sb_state[p1] = bin;
                        out[iF] = ikf;
                        if (check_ori) atomicAdd(&s_hist[bin], 1);
                        ++nm;
                    }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && nm) atomicAdd(&s_nm, nm);
    __syncthreads();
    if (check_ori) {
        if (threadIdx.x == 0) {
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int b = 0; b < SP_HISTO; ++b) {
                const int s = s_hist[b];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
                else if (s > max3) { max3 = s; ind3 = b; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
            unsigned keep = 0;
            if (ind1 >= 0) keep |= 1u << ind1;
            if (ind2 >= 0) keep |= 1u << ind2;
            if (ind3 >= 0) keep |= 1u << ind3;
            s_keep = keep;
        }
        __syncthreads();
        const unsigned keep = s_keep;
        int culled = 0;
        for (int p = threadIdx.x; p < nF; p += blockDim.x) {
            const int bin = sb_state[p];
            if (bin >= 0 && bin < SP_HISTO && !((keep >> bin) & 1u)) { out[FF[p]] = -1; ++culled; }     // (:273-281)
        }
        if (culled) atomicAdd(&s_ncull, culled);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats_out[2 * blockIdx.x] = s_nm - s_ncull;
        stats_out[2 * blockIdx.x + 1] = 1;
    }
}

// =====================================================================================
// launch wrappers (called from orbx_api.cu)
// =====================================================================================
// format: 1 BGR8, 2 RGB8, 3 BGRA8, 4 RGBA8 (orbx_pixel_format)
cudaError_t launch_cvt_gray(const uint8_t* src, size_t src_pitch, size_t src_frame_stride, int w, int h, int nframes, int format,
                            uint8_t* dst, size_t dst_pitch, size_t dst_frame_stride, cudaStream_t st) {
    const bool rgb = format == 2 || format == 4;
    const int ch = format >= 3 ? 4 : 3;
    const uint32_t cB = 3735u, cG = 19235u, cR = 9798u;                 // OpenCV's 15-bit weights (sum 32768)
    const uint32_t c01 = (rgb ? cR : cB) | (cG << 16), c2 = rgb ? cB : cR;
    const int aligned16 = ((reinterpret_cast<uintptr_t>(src) | src_pitch | src_frame_stride | reinterpret_cast<uintptr_t>(dst) |
                            dst_pitch | dst_frame_stride) & 15) == 0;
    dim3 block(32, 8), grid(((w + 15) / 16 + 31) / 32, (h + 7) / 8, nframes);
    if (ch == 4)
        return launch_k(cvt_gray_kernel<4>, grid, block, 0, st, src, src_pitch, src_frame_stride, w, h, c01, c2, aligned16, dst, dst_pitch,
                        dst_frame_stride);
    return launch_k(cvt_gray_kernel<3>, grid, block, 0, st, src, src_pitch, src_frame_stride, w, h, c01, c2, aligned16, dst, dst_pitch,
                    dst_frame_stride);
}

static bool force_resize4() {          // ORBX_RESIZE4=1: A/B switch back to the 4-pixel-per-lane resize kernel
    static const bool v = getenv("ORBX_RESIZE4") != nullptr;
    return v;
}

void launch_pyr_level(const OrbxPlan* d_plan, const OrbxPlan& hp, int l, int nframes, int num_sms, const uint8_t* imgs,
                      size_t img_pitch, size_t img_frame_stride, uint8_t* pyr, const OrbxTap* taps,
                      cudaStream_t st, const void* tile_maps, int frame0);

// Block geometry of one pyramid level, shared by the per-level launches and the chained kernel.
struct PyrGeom { int gx, gy, ry, kind; };
static PyrGeom pyr_geometry(const OrbxPlan& hp, int l, int nframes, int num_sms) {
    const OrbxLevel& L = hp.lv[l];
    PyrGeom g;
    if (l == 0) {
        const int cols16 = (ORBX_XO + L.w + ORBX_EDGE + 15) / 16;          // 16-byte chunks from plane column 0
        g.gx = (cols16 + 31) / 32; g.gy = (L.rows + 8 * PYR_L0_ROWS - 1) / (8 * PYR_L0_ROWS); g.ry = 0; g.kind = 0;
    } else if (L.resize8_ok && !force_resize4()) {
        g.gx = (L.ngroups8 + 31) / 32;
        static const int min_ry = getenv("ORBX_PYR_MINRY") ? atoi(getenv("ORBX_PYR_MINRY")) : 2;      // 2: single 640x480 frame 50 -> 38 us
        int RY = PYR_RY;
        while (RY > min_ry && (long long)g.gx * ((L.rows + RY - 1) / RY) * nframes < (long long)num_sms * 32) RY >>= 1;
        g.ry = RY; g.gy = (L.rows + 4 * RY - 1) / (4 * RY); g.kind = 1;
    } else {
        // rows per warp: long strips reuse row passes (1 + 1/RY... per row) but small levels need warps
        const int cols4 = (ORBX_XO + L.w + ORBX_EDGE - 12 + 3) / 4;
        g.gx = (cols4 + 31) / 32;
        int RY = PYR_RY;
        while (RY > 4 && (long long)g.gx * ((L.rows + RY - 1) / RY) * nframes < (long long)num_sms * 32) RY >>= 1;
        g.ry = RY; g.gy = (L.rows + 4 * RY - 1) / (4 * RY); g.kind = L.resize_wide ? 3 : 2;
    }
    return g;
}

void launch_pyr_level(const OrbxPlan* d_plan, const OrbxPlan& hp, int l, int nframes, int num_sms, const uint8_t* imgs,
                      size_t img_pitch, size_t img_frame_stride, uint8_t* pyr, const OrbxTap* taps,
                      cudaStream_t st, const void* tile_maps, int frame0) {
    const PyrGeom g = pyr_geometry(hp, l, nframes, num_sms);
    static const bool no_tile = getenv("ORBX_PYR_NO_TILE") != nullptr;         // A/B switch
    static const int min_tile_ry = getenv("ORBX_PYR_TILE_MINRY") ? atoi(getenv("ORBX_PYR_TILE_MINRY")) : 8;   // strips shorter than this keep pyr_resize8_kernel (measured: 64 x 1080p 1.425 -> 1.378 ms with 8, 1.376 with 4; 64 x 640x480 0.438 / 0.439 / 0.446 ms without / 8 / 4; 2-row strips = 1024-thread CTAs, a single 1080p frame 0.120 vs 0.098 ms)
    if (g.kind == 0) {
        const int aligned16 = ((reinterpret_cast<uintptr_t>(imgs) | img_pitch | img_frame_stride) & 15) == 0;
        const Lv0Cols lc = pyr_level0_cols(hp.lv[0].w, aligned16);
        const int border_items = (hp.lv[0].rows * lc.nb + 255) / 256;                 // 256 (row, border vector) pairs each
        const int gy_border = (border_items + g.gx - 1) / g.gx;
        launch_k(pyr_level0_kernel, dim3(g.gx, g.gy + gy_border, nframes), dim3(32, 8), 0, st, d_plan, imgs, img_pitch, img_frame_stride, aligned16, pyr, g.gy);
    } else if (g.kind == 1 && g.ry >= min_tile_ry && g.ry <= 16 && hp.lv[l].resize_tile_ok && tile_maps && !no_tile) {
        const CUtensorMap& m = reinterpret_cast<const FastMaps*>(tile_maps)->m[l - 1];
        const dim3 grid(g.gx, (hp.lv[l].rows + 63) / 64, nframes);
        if (g.ry == 16) launch_k(pyr_resize8_tile_kernel<16>, grid, dim3(128), 0, st, m, d_plan, l, frame0, pyr, taps);
        else if (g.ry == 8) launch_k(pyr_resize8_tile_kernel<8>, grid, dim3(256), 0, st, m, d_plan, l, frame0, pyr, taps);
        else if (g.ry == 4) launch_k(pyr_resize8_tile_kernel<4>, grid, dim3(512), 0, st, m, d_plan, l, frame0, pyr, taps);
        else launch_k(pyr_resize8_tile_kernel<2>, grid, dim3(1024), 0, st, m, d_plan, l, frame0, pyr, taps);
    } else if (g.kind == 1) {
        launch_k(pyr_resize8_kernel, dim3(g.gx, g.gy, nframes), dim3(128), 0, st, d_plan, l, g.ry, pyr, taps);
    } else if (g.kind == 3) {
        launch_k(pyr_resize_kernel<true>, dim3(g.gx, g.gy, nframes), dim3(128), 0, st, d_plan, l, g.ry, pyr, taps);
    } else {
        launch_k(pyr_resize_kernel<false>, dim3(g.gx, g.gy, nframes), dim3(128), 0, st, d_plan, l, g.ry, pyr, taps);
    }
}

// u32 tensor maps {pitch / 4, rows, frames} of the level planes, box = pyr_resize8_tile_kernel's source tile
int build_pyr_tile_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, void* out_maps) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !fn)
            return -1;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    FastMaps* fm = reinterpret_cast<FastMaps*>(out_maps);
    memset(fm, 0, sizeof(FastMaps));
    for (int l = 0; l < hp.nlevels; ++l) {
        const OrbxLevel& L = hp.lv[l];
        cuuint64_t dims[3] = {(cuuint64_t)(L.pitch / 4), (cuuint64_t)L.rows, (cuuint64_t)max_frames};
        cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)hp.slab_bytes};
        cuuint32_t box[3] = {PYR_TILE_W / 4, PYR_TILE_H, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&fm->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d_pyr + L.plane_off, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -(int)r - 100;
    }
    return 0;
}

// dynamic shared memory of fast_cells_kernel (fast_strips_kernel's is static)
size_t fast_smem_bytes(const OrbxPlan& hp) {
    const size_t TB = ((size_t)hp.cells_bw * hp.cells_bh + 127) & ~(size_t)127;
    const size_t SP = (size_t)((hp.max_cell_w + 2 + 3) & ~3);
    const size_t SB = (SP * hp.cells_bh + 127) & ~(size_t)127;
    const size_t QB = ((size_t)(hp.max_cell_w - 6) * (hp.max_cell_h - 6) * 2 + 127) & ~(size_t)127;
    return ((size_t)hp.fast_nb * TB + SB + QB) * hp.fast_warps + 128;
}

// One {pitch, rows, frames} u8 tensor map per level over the pyramid slabs; box = bw x bh bytes of one frame.
static int build_tile_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, int bw, int bh, void* out_maps, bool fast_boxes = false) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !fn)
            return -1;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    FastMaps* fm = reinterpret_cast<FastMaps*>(out_maps);
    memset(fm, 0, sizeof(FastMaps));
    for (int l = 0; l < hp.nlevels; ++l) {
        const OrbxLevel& L = hp.lv[l];
        cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)max_frames};
        cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)hp.slab_bytes};
        if (fast_boxes) {              // strips of <= 32-px cells -> fast_strips_kernel's box, everything else -> fast_cells_kernel's
            bw = L.strip_ok ? hp.fast_bw : hp.cells_bw;
            bh = L.strip_ok ? (L.strip_tall ? ORBX_FS_BH_TALL : ORBX_FS_BH) : hp.cells_bh;
        }
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&fm->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_pyr + L.plane_off, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -(int)r - 100;
    }
    return 0;
}

// box = one strip of FAST cell windows
int build_fast_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, void* out_maps) {
    return build_tile_maps(hp, d_pyr, max_frames, hp.fast_bw, hp.fast_bh, out_maps, true);
}

// box = the raw window of one keypoint (describe_kernel)
int build_describe_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, void* out_maps) {
    return build_tile_maps(hp, d_pyr, max_frames, KP_RAW_W, KP_RAW_H, out_maps);
}

size_t fast_maps_bytes() { return sizeof(FastMaps); }

// One segment of the strip table (hp.seg_first / hp.seg_count: strips of levels 0-1, strips of levels 2+, big-cell levels
// 0-1, big-cell levels 2+, tall-cell strips of levels 0-1, of levels 2+) for nframes frames: segments 0-1 and 4-5 go to
// fast_strips_kernel (4 / 5 row groups), 2-3 (levels whose cells are wider than 32 or taller than 40, or everything under
// ORBX_FAST_LEGACY=1) to fast_cells_kernel.  nseg = 2 launches two adjacent segments
// as one.  Each segment has its own work counter (work_counters[seg]).
cudaError_t launch_fast(const OrbxPlan* d_plan, const OrbxPlan& hp, const void* maps, const OrbxTap* taps, int frame0, int nframes,
                        int seg, int nseg, int num_sms, uint32_t* cand, uint2* cell_rec, int* level_counts, int* work_counters,
                        int* status, int* retry_counts, cudaStream_t st) {
    typedef void (*fast_fn)(const FastMaps, const OrbxPlan*, const uint32_t*, int, int, int, int, uint32_t*, uint2*, int*, int*, int*, int*);
    typedef void (*strips_fn)(const FastMaps, const OrbxPlan*, const uint4*, int, int, int, int, uint32_t*, uint2*, int*, int*, int*, int*);
    static const strips_fn all_strips[6] = {fast_strips_kernel<96, 4>, fast_strips_kernel<128, 4>, fast_strips_kernel<160, 4>,
                                            fast_strips_kernel<96, 5>, fast_strips_kernel<128, 5>, fast_strips_kernel<160, 5>};
    static const fast_fn all_cells[4] = {fast_cells_kernel<64>, fast_cells_kernel<96>, fast_cells_kernel<128>, fast_cells_kernel<0>};
    const int first = hp.seg_first[seg];
    int nstrips = 0;
    for (int i = 0; i < nseg; ++i) nstrips += hp.seg_count[seg + i];
    if (nstrips == 0) return cudaSuccess;
    const bool strips = seg < 2 || seg >= 4;
    int which;
    if (strips) {
        which = hp.fast_bw == 96 ? 0 : hp.fast_bw == 128 ? 1 : hp.fast_bw == 160 ? 2 : -1;
        if (which < 0) return cudaErrorInvalidValue;
        if (seg >= 4) which += 3;                                                // tall cells: 5 row groups
    } else {
        which = 6 + (hp.cells_bw == 64 ? 0 : hp.cells_bw == 96 ? 1 : hp.cells_bw == 128 ? 2 : 3);
    }
    const void* fn = strips ? reinterpret_cast<const void*>(all_strips[which]) : reinterpret_cast<const void*>(all_cells[which - 6]);
    const size_t smem = strips ? 0 : fast_smem_bytes(hp);
    const int W = strips ? ORBX_FS_WARPS : hp.fast_warps;
    // per device and kernel: the dynamic shared-memory limit that was configured and the resident CTAs per SM it gives
    static size_t configured[64][10] = {};
    static int per_sm_cache[64][10] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    int per_sm_eff;
    {
        std::lock_guard<std::mutex> config_lock(g_config_mutex);
        if (smem + 1 != configured[dev][which]) {
            if (smem) {
                cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
            }
            configured[dev][which] = smem + 1;
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, W * 32, smem);
            per_sm_cache[dev][which] = per_sm < 1 ? 1 : per_sm;
        }
        per_sm_eff = per_sm_cache[dev][which];
    }
    const long long total = (long long)nframes * nstrips;
    long long blocks = strips ? total : (total + W - 1) / W;                  // a CTA per strip / a warp per strip
    // ORBX_FAST_CTAS_PER_SM (tuning): fewer resident FAST CTAs leave registers / shared memory for the kernels of the
    // other half-batch's stream to co-run
    static const int env_cap = getenv("ORBX_FAST_CTAS_PER_SM") ? atoi(getenv("ORBX_FAST_CTAS_PER_SM")) : 0;
    if (env_cap > 0 && env_cap < per_sm_eff) per_sm_eff = env_cap;
    // (several waves of shorter-lived CTAs -- 2 / 4 / 8 x this grid, so that the block scheduler can slip the CTAs of other
    // streams' kernels in between -- measured 1.309 / 1.318 / 1.348 ms per 64 x 1080p against 1.308 with two handles: the
    // streams share issue capacity, not residency)
    const long long cap = (long long)num_sms * per_sm_eff;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    FastMaps fm;
    memcpy(&fm, maps, sizeof fm);
    if (strips)
        return launch_k(all_strips[which], dim3((unsigned)blocks), dim3(W * 32), smem, st, fm, d_plan,
                        reinterpret_cast<const uint4*>(taps + hp.strip_rec_off), frame0, nframes, first, nstrips, cand, cell_rec,
                        level_counts, work_counters + seg, status, retry_counts);
    return launch_k(all_cells[which - 6], dim3((unsigned)blocks), dim3(W * 32), smem, st, fm, d_plan,
                    reinterpret_cast<const uint32_t*>(taps + hp.strip_tab_off), frame0, nframes, first, nstrips, cand, cell_rec,
                    level_counts, work_counters + seg, status, retry_counts);
}

size_t octree_smem_bytes(const OrbxPlan& hp) {
    int sortn = 1;
    while (sortn < hp.node_cap) sortn <<= 1;
    return (size_t)sortn * 8 + (size_t)hp.node_cap * (8 + 8 + 4 + 4 + 16 + 16 + 4 + 4 + 4 + 4) + (size_t)ORBX_OT_KEYCAP * 6;
}

cudaError_t launch_octree(const OrbxPlan* d_plan, const OrbxPlan& hp, int nframes, const uint32_t* cand,
                          const uint2* cell_rec, uint32_t* cand_sorted, uint16_t* key_node, int* sorted_counts,
                          uint32_t* kept, int* kept_counts, int* status, cudaStream_t st) {
    const size_t smem = octree_smem_bytes(hp);
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> config_lock(g_config_mutex);
    if (smem > configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = smem;
    }
    // 1024 threads halve the key sweeps of huge levels (4K: ~28k candidates) but cost residency on small ones
    static const int env_threads = getenv("ORBX_OT_THREADS") ? atoi(getenv("ORBX_OT_THREADS")) : 0;     // tuning override
    // ... and of a lone big frame (latency mode: 1080p single frame 42.8 -> 37.8 us, while a batch of 32 gets slower)
    const long long area0 = (long long)hp.lv[0].w * hp.lv[0].h;
    const int threads = env_threads == 512 || env_threads == 1024 ? env_threads
                        : (area0 >= 4000000LL || (nframes <= 2 && area0 >= 1000000LL)) ? 1024 : 512;
    const cudaError_t le = launch_k(octree_kernel, dim3(nframes * hp.nlevels), dim3(threads), smem, st, d_plan, nframes, cand,
                                    cell_rec, cand_sorted, key_node, sorted_counts, kept, kept_counts, status);
    if (le != cudaSuccess) return le;
    return cudaSuccess;
}

void launch_blur(const OrbxPlan* d_plan, const OrbxPlan& hp, int nframes, int num_sms, const uint8_t* pyr,
                 uint8_t* blur, cudaStream_t st) {
    long long blocks = ((long long)nframes * hp.blur_tiles_per_frame + 3) / 4;      // one warp per 128 x 32 tile
    const long long cap = (long long)num_sms * 12;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_k(blur_kernel, dim3((unsigned)blocks), dim3(128), 0, st, d_plan, nframes, pyr, blur);
}

cudaError_t launch_describe(const OrbxPlan* d_plan, const OrbxPlan& hp, const void* maps, int frame0, int nframes, int num_sms,
                            const uint32_t* kept, const int* kept_counts, float* angles, float* out_kp, uint8_t* out_desc,
                            cudaStream_t st) {
    const size_t smem = (size_t)KP_WARPS * KP_WARP_BYTES + 128;
    static bool configured[64] = {false};
    static int per_sm_cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> config_lock(g_config_mutex);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, describe_kernel, KP_WARPS * 32, smem);
        per_sm_cache[dev & 63] = per_sm < 1 ? 1 : per_sm;
        configured[dev & 63] = true;
    }
    long long blocks = ((long long)nframes * hp.kept_per_frame + KP_WARPS - 1) / KP_WARPS;
    const long long cap = (long long)num_sms * per_sm_cache[dev & 63];
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    FastMaps fm;
    memcpy(&fm, maps, sizeof fm);
    return launch_k(describe_kernel, dim3((unsigned)blocks), dim3(KP_WARPS * 32), smem, st, fm, d_plan, frame0, nframes, kept,
                    kept_counts, angles, out_kp, out_desc);
}

cudaError_t launch_undistort_grid(const OrbxPlan* d_plan, const OrbxPlan& hp, const float* kp, const int* kept_counts,
                                  const int* d_frames, int nframes, const double* cam /* fx fy cx cy k1 k2 p1 p2 k3 */, int distorted,
                                  const float* grid /* mnMinX mnMinY winv hinv */, float* xy_un, int* cell_start, int* cell_items,
                                  cudaStream_t st) {
    UndistortParams P;
    P.fx = cam[0]; P.fy = cam[1]; P.cx = cam[2]; P.cy = cam[3];
    P.ifx = 1. / cam[0]; P.ify = 1. / cam[1];
    for (int i = 0; i < 5; ++i) P.k[i] = cam[4 + i];
    P.distorted = distorted;
    P.min_x = grid[0]; P.min_y = grid[1]; P.winv = grid[2]; P.hinv = grid[3];
    const size_t smem = (size_t)(UG_COLS * UG_ROWS + hp.kept_per_frame) * sizeof(int);
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> config_lock(g_config_mutex);
    if (smem > 48 * 1024 && smem > configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(undistort_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = smem;
    }
    return launch_k(undistort_grid_kernel, dim3((unsigned)nframes), dim3(1024), smem, st, d_plan, kp, kept_counts, d_frames, P, xy_un,
                    cell_start, cell_items);
}

size_t search_projection_query_bytes() { return sizeof(SpQuery); }

void search_projection_fill_query(void* dst, const float* Rcw, const float* tcw, int n_last, int frame, int fwd, int bwd) {
    SpQuery q;
    for (int i = 0; i < 9; ++i) q.R[i] = Rcw[i];
    for (int i = 0; i < 3; ++i) q.t[i] = tcw[i];
    q.n_last = n_last; q.frame = frame; q.fwd = fwd; q.bwd = bwd;
    memcpy(dst, &q, sizeof q);
}

// local = 0: SearchByProjection(CurrentFrame, LastFrame) -- world = positions, last_octave / last_angle of the LastFrame keypoints;
// local = 1: SearchByProjection(F, vpMapPoints, th) -- world = (mTrackProjX, mTrackProjY, mTrackProjXR), last_octave =
//            mnTrackScaleLevel, last_angle = mTrackViewCos, cur_obs = Observations() of what the frame's keypoints hold.
// mode 0: LastFrame variant (:1328-1470), 1: local-map variant (:45-129), 2: KeyFrame variant (:1472-1599, match_th = ORBdist)
cudaError_t launch_search_projection(const OrbxPlan* d_plan, const OrbxPlan& hp, int mode, int nq, const void* d_queries, const float* K4,
                                     const float* bounds, float mbf, float th, float nnratio, int check_ori, int match_th, int cap, int list_cap,
                                     const float* world, const uint8_t* mp_desc, const int* mp_obs, const int* last_octave,
                                     const float* last_angle, const float* kp, const uint8_t* desc, const int* kept_counts,
                                     const float* xy_un, const int* cell_start, const int* cell_items, const float* u_right,
                                     const int* cur_obs, uint32_t* cand_list, int* cand_count, int* match_out, int* stats_out,
                                     cudaStream_t st) {
    SpParams P;
    P.fx = K4[0]; P.fy = K4[1]; P.cx = K4[2]; P.cy = K4[3];
    P.min_x = bounds[0]; P.max_x = bounds[1]; P.min_y = bounds[2]; P.max_y = bounds[3];
    P.winv = 64.f / (bounds[1] - bounds[0]);
    P.hinv = 48.f / (bounds[3] - bounds[2]);
    P.mbf = mbf; P.th = th; P.nnratio = nnratio; P.check_ori = check_ori; P.cap = cap; P.use_stereo = u_right != nullptr;
    // entries worth listing: the match itself needs distance <= TH_HIGH; a runner-up can only reject (best > ratio * second)
    // while second < TH_HIGH / ratio
    const bool local = mode == 1;
    P.match_th = mode == 2 ? match_th : SP_TH_HIGH;
    P.check_z = mode == 0;
    P.list_th = P.match_th;
    if (local) P.list_th = nnratio > 0.4f ? (int)(SP_TH_HIGH / nnratio) + 1 : 256;
    const size_t smem = (size_t)(cap + hp.kept_per_frame) * sizeof(int);
    static size_t configured[64][2] = {{0}};
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> config_lock(g_config_mutex);
        if (smem > 48 * 1024 && smem > configured[dev & 63][local ? 1 : 0]) {
            cudaError_t e = local ? cudaFuncSetAttribute(search_projection_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(search_projection_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev & 63][local ? 1 : 0] = smem;
        }
    }
    const long long warps = (long long)nq * cap;
    const dim3 lgrid((unsigned)((warps + SP_LIST_WARPS - 1) / SP_LIST_WARPS)), lblock(SP_LIST_WARPS * 32);
    cudaError_t e;
    if (local)
        e = launch_k(search_projection_list_kernel<true>, lgrid, lblock, 0, st, d_plan, (const SpQuery*)d_queries, P, nq, list_cap, world,
                     (const uint4*)mp_desc, mp_obs, last_octave, last_angle, kp, desc, xy_un, cell_start, cell_items, u_right, cand_list,
                     cand_count);
    else
        e = launch_k(search_projection_list_kernel<false>, lgrid, lblock, 0, st, d_plan, (const SpQuery*)d_queries, P, nq, list_cap, world,
                     (const uint4*)mp_desc, mp_obs, last_octave, last_angle, kp, desc, xy_un, cell_start, cell_items, u_right, cand_list,
                     cand_count);
    if (e != cudaSuccess) return e;
    if (local)
        return launch_k(search_projection_kernel<true>, dim3((unsigned)nq), dim3(1024), smem, st, d_plan, (const SpQuery*)d_queries, P,
                        list_cap, world, (const uint4*)mp_desc, mp_obs, last_octave, last_angle, kp, desc, kept_counts, xy_un, cell_start,
                        cell_items, u_right, cur_obs, (const uint32_t*)cand_list, (const int*)cand_count, match_out, stats_out);
    return launch_k(search_projection_kernel<false>, dim3((unsigned)nq), dim3(1024), smem, st, d_plan, (const SpQuery*)d_queries, P,
                    list_cap, world, (const uint4*)mp_desc, mp_obs, last_octave, last_angle, kp, desc, kept_counts, xy_un, cell_start,
                    cell_items, u_right, cur_obs, (const uint32_t*)cand_list, (const int*)cand_count, match_out, stats_out);
}

cudaError_t launch_compute_bow(const OrbxPlan* d_plan, const OrbxPlan& hp, const int* voc_child_start, const int* voc_child_items,
                               const uint8_t* voc_desc, const double* voc_weight, const int* voc_word, int n_nodes, int L,
                               const int* d_frames, int nframes, int levelsup, const uint8_t* desc, const int* kept_counts, int* leaf,
                               int* nid, unsigned* word_ids, double* word_values, unsigned* fv_nodes, unsigned* fv_features,
                               int* counts_out, cudaStream_t st) {
    BowVoc V;
    V.child_start = voc_child_start; V.child_items = voc_child_items; V.node_desc = reinterpret_cast<const uint4*>(voc_desc);
    V.node_weight = voc_weight; V.node_word = voc_word; V.n_nodes = n_nodes; V.L = L;
    const long long groups = (long long)nframes * hp.kept_per_frame;
    const long long blocks = (groups * BOW_GROUP + BOW_THREADS - 1) / BOW_THREADS;
    cudaError_t e = launch_k(bow_descend_kernel, dim3((unsigned)blocks), dim3(BOW_THREADS), 0, st, d_plan, V, d_frames, nframes, levelsup,
                             desc, kept_counts, leaf, nid);
    if (e != cudaSuccess) return e;
    int sort_n = 1024;
    while (sort_n < hp.kept_per_frame) sort_n <<= 1;
    const size_t smem = (size_t)sort_n * sizeof(unsigned long long);
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> config_lock(g_config_mutex);
        if (smem > 48 * 1024 && smem > configured[dev & 63]) {
            e = cudaFuncSetAttribute(bow_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev & 63] = smem;
        }
    }
    return launch_k(bow_reduce_kernel, dim3((unsigned)nframes), dim3(1024), smem, st, d_plan, V, d_frames, sort_n, kept_counts,
                    (const int*)leaf, (const int*)nid, word_ids, word_values, fv_nodes, fv_features, counts_out);
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:405-520): d_queries = nq x {n1, frame} (8 bytes each)
cudaError_t launch_search_init(const OrbxPlan* d_plan, const OrbxPlan& hp, int nq, const void* d_queries, const float* bounds, float window,
                               float nnratio, int check_ori, int cap, int list_cap, const float* prev, const uint8_t* desc1,
                               const int* octave1, const float* angle1, const float* kp, const uint8_t* desc, const int* kept_counts,
                               const float* xy_un, const int* cell_start, const int* cell_items, uint32_t* cand_list, int* cand_count,
                               uint8_t* bin_scratch, int* match_out, float* prev_out, int* stats_out, cudaStream_t st) {
    SpParams P;
    memset(&P, 0, sizeof P);
    P.min_x = bounds[0]; P.max_x = bounds[1]; P.min_y = bounds[2]; P.max_y = bounds[3];
    P.winv = 64.f / (bounds[1] - bounds[0]);
    P.hinv = 48.f / (bounds[3] - bounds[2]);
    P.th = window; P.nnratio = nnratio; P.check_ori = check_ori; P.cap = cap;
    P.list_th = 256;                                                          // every candidate is listed
    const size_t smem = (size_t)hp.kept_per_frame * 2 * sizeof(int);
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> config_lock(g_config_mutex);
        if (smem > 48 * 1024 && smem > configured[dev & 63]) {
            cudaError_t e = cudaFuncSetAttribute(init_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev & 63] = smem;
        }
    }
    const long long warps = (long long)nq * cap;
    const dim3 lgrid((unsigned)((warps + SP_LIST_WARPS - 1) / SP_LIST_WARPS)), lblock(SP_LIST_WARPS * 32);
    cudaError_t e = launch_k(init_list_kernel, lgrid, lblock, 0, st, d_plan, (const InitQuery*)d_queries, P, nq, list_cap, prev,
                             (const uint4*)desc1, octave1, kp, desc, xy_un, cell_start, cell_items, cand_list, cand_count);
    if (e != cudaSuccess) return e;
    return launch_k(init_resolve_kernel, dim3((unsigned)nq), dim3(32), smem, st, d_plan, (const InitQuery*)d_queries, P, list_cap, prev,
                    (const uint4*)desc1, octave1, angle1, kp, desc, kept_counts, xy_un, cell_start, cell_items,
                    (const uint32_t*)cand_list, (const int*)cand_count, match_out, prev_out, bin_scratch, stats_out);
}

size_t search_bow_query_bytes() { return sizeof(BowMatchQuery); }
void search_bow_fill_query(void* dst, int frame, int slot, int n_kf, int n_kf_fv) {
    BowMatchQuery q;
    q.frame = frame; q.slot = slot; q.n_kf = n_kf; q.n_kf_fv = n_kf_fv;
    memcpy(dst, &q, sizeof q);
}

cudaError_t launch_search_bow(const OrbxPlan* d_plan, const OrbxPlan& hp, int nq, const void* d_queries, int cap, float nnratio,
                              int check_ori, const uint8_t* kf_desc, const uint8_t* kf_valid, const float* kf_angle,
                              const unsigned* kf_fv_nodes, const unsigned* kf_fv_features, const float* kp, const uint8_t* desc,
                              const int* kept_counts, const unsigned* f_fv_nodes, const unsigned* f_fv_features, const int* bow_counts,
                              int* match_out, int* stats_out, cudaStream_t st) {
    const size_t smem = (size_t)hp.kept_per_frame * sizeof(int);
    return launch_k(search_bow_kernel, dim3((unsigned)nq), dim3(1024), smem, st, d_plan, (const BowMatchQuery*)d_queries, cap, nnratio,
                    check_ori, (const uint4*)kf_desc, kf_valid, kf_angle, kf_fv_nodes, kf_fv_features, kp, desc, kept_counts, f_fv_nodes,
                    f_fv_features, bow_counts, match_out, stats_out);
}

size_t stereo_bucket_entries(const OrbxPlan& hp) { return (size_t)hp.kept_per_frame * ST_MAX_SPAN; }

cudaError_t launch_stereo(const OrbxPlan* d_plan, const OrbxPlan& hp, int num_sms, const uint8_t* pyrL, const float* kpL,
                          const uint8_t* descL, const int* countsL, const uint8_t* pyrR, const float* kpR, const uint8_t* descR,
                          const int* countsR, const int* d_pairs, int npairs, float mbf, float mb, float* u_right, float* depth,
                          int* sad, int* row_start, uint16_t* bucket, cudaStream_t st) {
    StereoSide SL = {pyrL, kpL, descL, countsL}, SR = {pyrR, kpR, descR, countsR};
    {
        const size_t smem_rows = (size_t)(hp.height + 1) * sizeof(int);
        cudaError_t e0 = launch_k(stereo_rows_kernel, dim3((unsigned)npairs), dim3(256), smem_rows, st, d_plan, SR, d_pairs, row_start, bucket);
        if (e0 != cudaSuccess) return e0;
    }
    long long blocks = ((long long)npairs * hp.kept_per_frame + ST_WARPS - 1) / ST_WARPS;
    const long long cap = (long long)num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaError_t e = launch_k(stereo_match_kernel, dim3((unsigned)blocks), dim3(ST_WARPS * 32), 0, st, d_plan, SL, SR, d_pairs, npairs,
                             (const int*)row_start, (const uint16_t*)bucket, mbf, mb, u_right, depth, sad);
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)hp.kept_per_frame * sizeof(int);
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> config_lock(g_config_mutex);
    if (smem > 48 * 1024 && smem > configured[dev & 63]) {
        e = cudaFuncSetAttribute(stereo_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = smem;
    }
    return launch_k(stereo_filter_kernel, dim3((unsigned)npairs), dim3(1024), smem, st, d_plan, countsL, d_pairs, u_right, depth,
                    (const int*)sad);
}

// =====================================================================================
// Frame::isInFrustum(MapPoint*, viewingCosLimit) (reference src/Frame.cc:269-325) with MapPoint::PredictScale(dist, Frame*)
// (src/MapPoint.cc:402-417) and GetMin/MaxDistanceInvariance (:373-383): the producer of the tracking fields that
// ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) consumes (Tracking::SearchLocalPoints, src/Tracking.cc:1165-1194).
// One thread per (query, map point); every float / double operation is the reference's, individually rounded:
//   Pc = mRcw * P + mtcw        cv::gemm's 3x3 path: float products and sums, the "+ C" in double (as sp_project)
//   invz = 1.0f / PcZ           float division (the matchers use 1.0 / z in double; this function does not)
//   dist = cv::norm(P - mOw)    squares and sum in double, sqrt, rounded to float
//   viewCos = PO.dot(Pn)/dist   Mat::dot: float products accumulated in double; double / float -> double -> float
//   PredictScale                ratio = mfMaxDistance / dist in float; ceil(logf(ratio) / mfLogScaleFactor) goes through the
//                               HOST's libm, so the host hands over, per level k, the smallest float ratio T[k] whose quotient
//                               reaches k (found by bisection over the float's bits with the host's own logf: the function is
//                               monotone), and the level is the number of thresholds the ratio reaches.  A non-finite or
//                               non-positive ratio (the reference then converts inf / NaN to int: undefined, INT_MIN -> 0 on
//                               x86) gives level 0.
// =====================================================================================
struct FrustumParams {
    float fx, fy, cx, cy, min_x, max_x, min_y, max_y, mbf, cos_limit;
    int nlevels;
    float T[ORBX_MAXL];                   // T[k], k = 1 .. nlevels - 1
};
struct FrustumQuery {                     // 80 bytes: the staging area holds them back to back at 16-byte granularity
    float R[9], t[3], Ow[3];
    int n, off, pad[3];
};
static_assert(sizeof(FrustumQuery) % 16 == 0, "queries are staged as an array");

__global__ void __launch_bounds__(256)
frustum_kernel(const FrustumParams P, const FrustumQuery* __restrict__ queries, const uint8_t* __restrict__ consider,
               const float* __restrict__ world, const float* __restrict__ normal, const float* __restrict__ min_dist,
               const float* __restrict__ max_dist, uint8_t* __restrict__ in_view, float* __restrict__ proj, int* __restrict__ level,
               float* __restrict__ view_cos) {
    const FrustumQuery& q = queries[blockIdx.y];
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= q.n) return;
    const size_t i = (size_t)q.off + k;
    in_view[i] = 0;                                                              // pMP->mbTrackInView = false (:271)
    proj[3 * i] = proj[3 * i + 1] = proj[3 * i + 2] = 0.f;
    level[i] = 0;
    view_cos[i] = 0.f;
    if (!consider[i]) return;
    const float wx = world[3 * i], wy = world[3 * i + 1], wz = world[3 * i + 2];
    float c3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {                                                // (:277)
        const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(q.R[3 * r], wx), __fmul_rn(q.R[3 * r + 1], wy)), __fmul_rn(q.R[3 * r + 2], wz));
        c3[r] = __double2float_rn(__dadd_rn((double)t0, (double)q.t[r]));
    }
    if (c3[2] < 0.0f) return;                                                    // (:283)
    const float invz = __fdiv_rn(1.0f, c3[2]);                                   // (:287)
    const float u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c3[0]), invz), P.cx);
    const float v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c3[1]), invz), P.cy);
    if (u < P.min_x || u > P.max_x) return;                                      // (:291-294) NaN passes, as in the reference
    if (v < P.min_y || v > P.max_y) return;
    const float maxd = __fmul_rn(1.2f, max_dist[i]), mind = __fmul_rn(0.8f, min_dist[i]);      // (src/MapPoint.cc:373-383)
    const float px = __fsub_rn(wx, q.Ow[0]), py = __fsub_rn(wy, q.Ow[1]), pz = __fsub_rn(wz, q.Ow[2]);     // (:299)
    const double dx = (double)px, dy = (double)py, dz = (double)pz;
    const float dist = __double2float_rn(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))));
    if (dist < mind || dist > maxd) return;                                      // (:302-303)
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(dx, (double)normal[3 * i]), __dmul_rn(dy, (double)normal[3 * i + 1])),
                                 __dmul_rn(dz, (double)normal[3 * i + 2]));
    const float vc = __double2float_rn(__ddiv_rn(dot, (double)dist));            // (:308)
    if (vc < P.cos_limit) return;                                                // (:310-311)
    const float ratio = __fdiv_rn(max_dist[i], dist);                            // (src/MapPoint.cc:407)
    int lvl = 0;
    if (ratio > 0.f && ratio <= 3.402823466e38f)                                 // finite and positive
        for (int j = 1; j < P.nlevels; ++j) lvl += ratio >= P.T[j];
    in_view[i] = 1;                                                              // (:317-322)
    proj[3 * i] = u;
    proj[3 * i + 1] = v;
    proj[3 * i + 2] = __fsub_rn(u, __fmul_rn(P.mbf, invz));
    level[i] = lvl;
    view_cos[i] = vc;
}

size_t frustum_query_bytes() { return sizeof(FrustumQuery); }
void frustum_fill_query(void* dst, const float* Rcw, const float* tcw, const float* Ow, int n, int off) {
    FrustumQuery q;
    memcpy(q.R, Rcw, sizeof q.R);
    memcpy(q.t, tcw, sizeof q.t);
    memcpy(q.Ow, Ow, sizeof q.Ow);
    q.n = n;
    q.off = off;
    q.pad[0] = q.pad[1] = q.pad[2] = 0;
    memcpy(dst, &q, sizeof q);
}
cudaError_t launch_frustum(int nq, int max_n, const void* d_queries, const float* K4, const float* bounds, float mbf, float cos_limit,
                           int nlevels, const float* thresholds, const uint8_t* consider, const float* world, const float* normal,
                           const float* min_dist, const float* max_dist, uint8_t* in_view, float* proj, int* level, float* view_cos,
                           cudaStream_t st) {
    if (nq < 1 || max_n < 1) return cudaSuccess;
    FrustumParams P;
    P.fx = K4[0]; P.fy = K4[1]; P.cx = K4[2]; P.cy = K4[3];
    P.min_x = bounds[0]; P.max_x = bounds[1]; P.min_y = bounds[2]; P.max_y = bounds[3];
    P.mbf = mbf; P.cos_limit = cos_limit; P.nlevels = nlevels;
    for (int k = 0; k < ORBX_MAXL; ++k) P.T[k] = k < nlevels ? thresholds[k] : 0.f;
    frustum_kernel<<<dim3((unsigned)((max_n + 255) / 256), (unsigned)nq), 256, 0, st>>>(
        P, reinterpret_cast<const FrustumQuery*>(d_queries), consider, world, normal, min_dist, max_dist, in_view, proj, level, view_cos);
    return cudaGetLastError();
}

}  // namespace orbx
