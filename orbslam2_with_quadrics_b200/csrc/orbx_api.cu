// C-ABI layer of the B200 ORB front end (include/orbx.h): handle, plan, HBM/pinned buffers,
// and the launch sequence that replaces ORB_SLAM2::ORBextractor::operator()
// (reference src/ORBextractor.cc:1043-1105).  Host code only prepares tables and enqueues
// kernels; every pixel, keypoint and descriptor is produced on the GPU (orbx_kernels.cu).
#include <cuda_runtime.h>

#include <ctype.h>
#include <math.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <new>
#include <algorithm>
#include <vector>

#include "../../include/orbx.h"
#include "orbx_kernels.h"

namespace {

enum { ST_PYRAMID = 0, ST_FAST, ST_OCTREE, ST_DESCRIBE, ST_COUNT };
const char* kStageNames[ST_COUNT] = {"pyramid", "fast_cells", "octree", "describe"};
const int kTimingRing = 64;
const int kMaxChunks = 8;         // sub-batches of one extract_batch call that overlap H2D, kernels and D2H

template <typename T> T round_up(T v, T a) { return (v + a - 1) / a * a; }

}  // namespace

struct orbx_handle {
    orbx_config cfg;
    int num_sms;
    cudaStream_t stream;          // kernels
    cudaStream_t ks[4];           // kernel streams (ks[0] == stream): sub-batches alternate so latency-bound tails overlap
    cudaStream_t aux[4];          // side stream of each kernel stream: FAST of levels 0-1 runs beside the pyramid tail
    cudaEvent_t ev_low[4], ev_fast_low[4], ev_join[4];
    int nks;                      // kernel streams in use (ORBX_KERNEL_STREAMS, default 2)
    cudaStream_t h2d_stream, d2h_stream;
    cudaEvent_t ev_h2d[8], ev_done[8], ev_clear;
    std::string last_error;
    long long launches;

    // ctor tables (src/ORBextractor.cc:415-469)
    float sf[ORBX_MAXL], inv_sf[ORBX_MAXL], sigma2[ORBX_MAXL], inv_sigma2[ORBX_MAXL];
    int quotas[ORBX_MAXL];
    int umax[16];

    // geometry-dependent state
    bool have_plan;
    OrbxPlan plan;
    std::vector<OrbxTap> taps;
    OrbxPlan* d_plan;
    OrbxTap* d_taps;
    int in_pitch;                // capacity of a staged input row (bytes)
    uint8_t *d_input, *d_pyr;
    uint8_t *d_color, *h_color;  // colour staging of orbx_extract_batch_color (allocated on first use)
    size_t color_bytes;
    uint8_t* d_blur;             // ONE frame's blurred slab, allocated on the first ORBX_STAGE_BLURRED dump (diagnostics only)
    uint32_t *d_cand, *d_cand_sorted, *d_kept;
    uint16_t* d_key_node;
    uint2* d_cell_rec;
    int* d_counters;             // [level_counts B*L][sorted_counts B*L][kept_counts B*L][status B][work counters][retry_counts B*L]
    std::vector<unsigned char> fast_maps;   // per-level TMA descriptors of the pyramid slabs (box = strip of FAST windows)
    std::vector<unsigned char> desc_maps;   // the same planes with box = one keypoint's raw window
    std::vector<unsigned char> pyr_maps;    // the same planes as u32 elements, box = pyr_resize8_tile_kernel's source tile
    float* d_angles;
    float* d_out_kp;
    uint8_t* d_out_desc;
    // pinned host mirrors
    int* h_counters;
    float* h_out_kp;
    uint8_t* h_out_desc;
    uint8_t* h_pyr;
    uint8_t* h_input;
    // CUDA graphs of the per-sub-batch launch sequence (see enqueue_frames_graphed)
    struct GraphEntry {
        int f0, n, chunk, si;
        size_t pitch, frame_stride;
        const uint8_t* d_imgs;
        cudaGraphExec_t exec;     // 0 until the sequence has been seen twice
        int launches;
    };
    std::vector<GraphEntry> graphs;
    bool use_graphs;

    // stereo matcher (orbx_stereo_match; allocated on first use, owned by the LEFT handle)
    float *d_st_u, *d_st_depth, *h_st;        // h_st: [u_right B*kpf][depth B*kpf]
    int *d_st_sad, *d_st_pairs, *h_st_pairs;
    int st_right_cap, st_pairs_n;
    int* d_st_rows;           // row table of the RIGHT handle's frames (vRowIndices): row_start, (height + 1) per frame
    uint16_t* d_st_bucket;    // ... and the keypoint indices listed per row
    // undistort + grid (orbx_undistort_grid; allocated on first use)
    float *d_un_xy, *h_un_xy;
    int *d_un_start, *d_un_items, *d_un_frames, *h_un_start, *h_un_items, *h_un_frames;
    float un_bounds[4];       // mnMinX, mnMaxX, mnMinY, mnMaxY of the last orbx_undistort_grid
    // SearchByProjection (orbx_search_by_projection; staging grows on demand)
    unsigned char *d_sp, *h_sp;
    size_t sp_bytes;
    int *d_sp_out, *h_sp_out;  // [match nq * kpf][stats nq * 2]
    size_t sp_out_ints;
    // Frame::isInFrustum (orbx_is_in_frustum; staging grows on demand, independent of the image geometry)
    unsigned char *d_fr, *h_fr;
    size_t fr_bytes;
    // ComputeBoW (orbx_compute_bow; allocated on first use): [leaf B*kpf][nid B*kpf][word ids B*kpf][fv nodes B*kpf]
    // [fv features B*kpf][counts 2B][frames B] as 32-bit words, and the word values as doubles
    unsigned *d_bow, *h_bow;
    double *d_bow_val, *h_bow_val;
    std::vector<int> bow_slot;  // frame -> slot of the last orbx_compute_bow (-1: none)
    // extract generation: bumped by every extraction; the side results (mvKeysUn + mGrid, mvuRight, BoW) remember the
    // generation they were computed for, and their consumers refuse stale ones instead of matching against an old frame
    unsigned long long gen = 0, un_gen = ~0ull, st_gen = ~0ull, bow_gen = ~0ull;
    cudaEvent_t ev_stereo;
    int last_n;
    bool pyramid_valid;

    // stage timing
    bool timing;
    cudaEvent_t ev[kTimingRing][ST_COUNT + 1];
    bool ev_created;
    int ev_head, ev_pending;
    double stage_ms[ST_COUNT];
    int stage_launches[ST_COUNT];

    int counters_count() const { return cfg.max_batch * (4 * plan.nlevels + 1) + ORBX_MAXL * kMaxChunks; }
    int* d_level_counts() const { return d_counters; }
    int* d_sorted_counts() const { return d_counters + cfg.max_batch * plan.nlevels; }
    int* d_kept_counts() const { return d_counters + 2 * cfg.max_batch * plan.nlevels; }
    int* d_status() const { return d_counters + 3 * cfg.max_batch * plan.nlevels; }
    int* d_work_counter() const { return d_counters + cfg.max_batch * (3 * plan.nlevels + 1); }
    int* d_retry_counts() const { return d_counters + cfg.max_batch * (3 * plan.nlevels + 1) + ORBX_MAXL * kMaxChunks; }
    const int* h_retry_counts() const { return h_counters + cfg.max_batch * (3 * plan.nlevels + 1) + ORBX_MAXL * kMaxChunks; }
    const int* h_level_counts() const { return h_counters; }
    const int* h_sorted_counts() const { return h_counters + cfg.max_batch * plan.nlevels; }
    const int* h_kept_counts() const { return h_counters + 2 * cfg.max_batch * plan.nlevels; }
    const int* h_status() const { return h_counters + 3 * cfg.max_batch * plan.nlevels; }
};

namespace {

int cuda_fail(orbx_handle* h, cudaError_t e, const char* what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    if (h) h->last_error = buf;
    return e == cudaErrorMemoryAllocation ? ORBX_ERR_OUT_OF_MEMORY : ORBX_ERR_CUDA;
}
#define CK(h, call)                                                  \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail((h), e__, #call);   \
    } while (0)

// ORBextractor::ORBextractor (src/ORBextractor.cc:410-470), float/double semantics kept as written there.
void build_ctor_tables(orbx_handle* h) {
    const int n = h->cfg.nlevels;
    const double scaleFactor = (double)h->cfg.scale_factor;          // member is double, argument float (.h:98)
    h->sf[0] = 1.0f;
    h->sigma2[0] = 1.0f;
    for (int i = 1; i < n; ++i) {
        h->sf[i] = (float)(h->sf[i - 1] * scaleFactor);
        h->sigma2[i] = h->sf[i] * h->sf[i];
    }
    for (int i = 0; i < n; ++i) {
        h->inv_sf[i] = 1.0f / h->sf[i];
        h->inv_sigma2[i] = 1.0f / h->sigma2[i];
    }
    const float factor = (float)(1.0f / scaleFactor);
    const float one_minus = 1 - factor;
    const float denom = 1 - (float)pow((double)factor, (double)n);
    float nDesired = (float)h->cfg.nfeatures * one_minus / denom;
    int sum = 0;
    for (int l = 0; l < n - 1; ++l) {
        h->quotas[l] = (int)lrintf(nDesired);
        sum += h->quotas[l];
        nDesired *= factor;
    }
    h->quotas[n - 1] = h->cfg.nfeatures - sum > 0 ? h->cfg.nfeatures - sum : 0;
    // umax (:454-469)
    const int HP = 15;
    const float half_diag = HP * sqrtf(2.f) / 2;
    const int vmax = (int)floor((double)(half_diag + 1));
    const int vmin = (int)ceil((double)half_diag);
    const double hp2 = HP * HP;
    for (int v = 0; v < 16; ++v) h->umax[v] = 0;
    for (int v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (int v = HP, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0;
        ++v0;
    }
}

// cv::resize INTER_LINEAR 8U tap table for one axis
void build_taps(std::vector<OrbxTap>& out, int ssize, int dsize) {
    const double scale = (double)ssize / dsize;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        long c1 = lrintf(f * 2048.f), c0 = lrintf((1.f - f) * 2048.f);
        OrbxTap t;
        t.ofs = s;
        t.c0 = (short)(c0 > 32767 ? 32767 : c0);
        t.c1 = (short)(c1 > 32767 ? 32767 : c1);
        out.push_back(t);
    }
}

int reflect_clamp_host(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}

// Tables of pyr_resize8_kernel for level L (source level S): one OrbxCol8 per group of 8 plane columns and one row-tap
// entry per PLANE row (the REFLECT_101 border of copyMakeBorder, :1122, folded into the index), both appended to the
// tap buffer (8-byte units).  Needs L.xtab_off / L.ytab_off.
void build_resize8_tables(std::vector<OrbxTap>& taps, const OrbxLevel& S, OrbxLevel& L) {
    static_assert(sizeof(OrbxCol8) == 8 * sizeof(OrbxTap), "OrbxCol8 is 8 tap units");
    while (taps.size() % 8) taps.push_back(OrbxTap{0, 0, 0});               // 64-byte alignment of the records
    L.ngroups8 = (ORBX_XO + L.w + ORBX_EDGE - 8 + 7) / 8;
    L.resize8_ok = 1;
    std::vector<OrbxCol8> cols((size_t)L.ngroups8);
    for (int g = 0; g < L.ngroups8; ++g) {
        OrbxCol8& C = cols[g];
        memset(&C, 0, sizeof C);
        int ofs[8];
        for (int j = 0; j < 8; ++j) {
            const OrbxTap& t = taps[L.xtab_off + reflect_clamp_host(8 + 8 * g + j - ORBX_XO, L.w)];
            ofs[j] = t.ofs;
            C.coef[j] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
        }
        for (int hf = 0; hf < 2; ++hf) {
            int mn = ofs[4 * hf];
            for (int j = 1; j < 4; ++j) mn = ofs[4 * hf + j] < mn ? ofs[4 * hf + j] : mn;
            C.base[hf] = mn & ~3;
            C.sh |= (uint32_t)(8 * (mn & 3)) << (8 * hf);
            for (int pr = 0; pr < 2; ++pr) {
                const int o0 = ofs[4 * hf + 2 * pr] - mn, o1 = ofs[4 * hf + 2 * pr + 1] - mn;
                if (o0 + 1 > 7 || o1 + 1 > 7) L.resize8_ok = 0;
                C.sel[2 * hf + pr] = (uint32_t)(o0 & 7) | ((uint32_t)((o0 + 1) & 7) << 4) | ((uint32_t)(o1 & 7) << 8) |
                                     ((uint32_t)((o1 + 1) & 7) << 12);
            }
        }
    }
    L.col8_off = (int)taps.size();
    const OrbxTap* raw = reinterpret_cast<const OrbxTap*>(cols.data());
    taps.insert(taps.end(), raw, raw + (size_t)L.ngroups8 * 8);
    L.yrow_off = (int)taps.size();
    for (int r = 0; r < L.rows; ++r) taps.push_back(taps[L.ytab_off + reflect_clamp_host(r - ORBX_EDGE, L.h)]);
    // pyr_resize8_tile_kernel: does every CTA's source footprint (32 groups x 64 plane rows) fit its 336 x 80 byte TMA tile?
    L.resize_tile_ok = L.resize8_ok;
    for (int g0 = 0; g0 < L.ngroups8 && L.resize_tile_ok; g0 += 32) {
        int mn = 1 << 30, mx = 0;
        for (int g = g0; g < g0 + 32 && g < L.ngroups8; ++g)
            for (int hf = 0; hf < 2; ++hf) {
                mn = std::min(mn, cols[g].base[hf]);
                mx = std::max(mx, cols[g].base[hf] + 12);
            }
        if (mx - ((mn + ORBX_XO) & ~15) + ORBX_XO > 336) L.resize_tile_ok = 0;
    }
    for (int r0 = 0; r0 < L.rows && L.resize_tile_ok; r0 += 64) {
        int mn = 1 << 30, mx = 0;
        for (int r = r0; r < r0 + 64 && r < L.rows; ++r) {
            const int sy = taps[L.yrow_off + r].ofs;
            mn = std::min(mn, sy);
            mx = std::max(mx, std::min(sy + 1, S.h - 1));
        }
        if (mx - mn + 1 > 80) L.resize_tile_ok = 0;
    }
}

void level_size(const orbx_handle* h, int w, int hgt, int l, int* lw, int* lh) {
    const float scale = h->inv_sf[l];                                  // (:1111-1112)
    *lw = (int)lrintf((float)w * scale);
    *lh = (int)lrintf((float)hgt * scale);
}

int build_plan(orbx_handle* h, int w, int hgt, OrbxPlan* P, std::vector<OrbxTap>* taps) {
    memset(P, 0, sizeof *P);
    if (w <= 0 || hgt <= 0) return ORBX_ERR_BAD_ARGS;
    if (w > 4096 || hgt > 4096) return ORBX_ERR_BAD_GEOMETRY;
    const int n = h->cfg.nlevels;
    P->nlevels = n;
    P->width = w;
    P->height = hgt;
    P->ini_th = h->cfg.ini_th_fast;
    P->min_th = h->cfg.min_th_fast;
    const int div = h->cfg.candidate_divisor > 0 ? h->cfg.candidate_divisor : 8;
    long long off = 0;
    int cells = 0, cand = 0, kept = 0, tiles = 0;
    taps->clear();
    for (int l = 0; l < n; ++l) {
        OrbxLevel& L = P->lv[l];
        level_size(h, w, hgt, l, &L.w, &L.h);
        if (L.w < 62 || L.h < 62) return ORBX_ERR_BAD_GEOMETRY;     // nCols/nRows would be 0 (App. B-7b)
        L.pitch = round_up(ORBX_XO + L.w + ORBX_EDGE, 64);
        L.rows = L.h + 2 * ORBX_EDGE;
        L.plane_off = off;
        off += round_up((long long)L.pitch * L.rows, 256LL);
        // cell grid (:771-787)
        L.maxBX = L.w - ORBX_EDGE + 3;
        L.maxBY = L.h - ORBX_EDGE + 3;
        const float width = (float)(L.maxBX - ORBX_BOX), height = (float)(L.maxBY - ORBX_BOX);
        const float W = 30;
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        L.wCell = (int)ceilf(width / nCols);
        L.hCell = (int)ceilf(height / nRows);
        L.nColsV = 0;
        for (int j = 0; j < nCols; ++j)
            if (ORBX_BOX + j * L.wCell < L.maxBX - 6) ++L.nColsV;     // (:803)
        L.nRowsV = 0;
        for (int i = 0; i < nRows; ++i)
            if (ORBX_BOX + i * L.hCell < L.maxBY - 3) ++L.nRowsV;     // (:794)
        L.cell_base = cells;
        cells += L.nColsV * L.nRowsV;
        const int cw = (L.wCell + 6 < (int)width) ? L.wCell + 6 : (int)width;
        const int ch = (L.hCell + 6 < (int)height) ? L.hCell + 6 : (int)height;
        if (cw > P->max_cell_w) P->max_cell_w = cw;
        if (ch > P->max_cell_h) P->max_cell_h = ch;
        L.cand_off = cand;
        L.cand_cap = (int)((long long)L.w * L.h / div) + 1024;
        if (L.cand_cap > (1 << 24) - 1) L.cand_cap = (1 << 24) - 1;
        cand += round_up(L.cand_cap, 64);
        L.quota = h->quotas[l];
        // DistributeOctTree roots (:543-545)
        const float ratio = (float)(L.maxBX - ORBX_BOX) / (L.maxBY - ORBX_BOX);
        L.nIni = (int)roundf(ratio);
        if (L.nIni < 1 || L.nIni > ORBX_MAX_ROOTS) return ORBX_ERR_BAD_GEOMETRY;   // division by zero in the reference (App. B-7)
        L.hX = (float)(L.maxBX - ORBX_BOX) / L.nIni;
        L.kept_off = kept;
        L.kept_cap = L.quota + 4 * L.nIni + 8;
        kept += L.kept_cap;
        if (L.kept_cap > P->node_cap) P->node_cap = L.kept_cap;
        L.scale = h->sf[l];
        L.kp_size = (float)(int)(31 * h->sf[l]);                         // (:837)
        if (l > 0) {
            L.xtab_off = (int)taps->size();
            build_taps(*taps, P->lv[l - 1].w, L.w);
            L.ytab_off = (int)taps->size();
            build_taps(*taps, P->lv[l - 1].h, L.h);
            // does any lane's group of 4 plane columns need more than 8 consecutive source bytes?
            L.resize_wide = 0;
            for (int c = 12; c < ORBX_XO + L.w + ORBX_EDGE; c += 4) {
                int lo = 1 << 30, hi = 0;
                for (int j = 0; j < 4; ++j) {
                    int dx = c + j - ORBX_XO;
                    if (dx < 0) dx = -dx;
                    if (dx >= L.w) dx = 2 * (L.w - 1) - dx;
                    dx = dx < 0 ? 0 : (dx > L.w - 1 ? L.w - 1 : dx);
                    const int s = (*taps)[L.xtab_off + dx].ofs;
                    lo = s < lo ? s : lo;
                    hi = s > hi ? s : hi;
                }
                if (hi - (lo & ~3) > 7) L.resize_wide = 1;
            }
            build_resize8_tables(*taps, P->lv[l - 1], L);
        }
        L.blur_tile_base = tiles;
        L.blur_tiles_x = (L.w + 127) / 128;
        tiles += L.blur_tiles_x * ((L.h + 31) / 32);
    }
    if (P->node_cap < 64) P->node_cap = 64;
    if (P->node_cap > 60000) return ORBX_ERR_BAD_ARGS;
    P->node_cap = round_up(P->node_cap, 32);
    if (P->max_cell_w < 7) P->max_cell_w = 7;
    if (P->max_cell_h < 7) P->max_cell_h = 7;
    // FAST tiling (env overrides are for tuning / A-B runs).  Product: fast_strips_kernel, a CTA takes a strip of up to 4
    // cells whose scoring pixels fit the 32 word columns of a warp's lanes (wCell <= 31: 4 cells, 32: 3).  Levels whose cells
    // are larger than 32 x 32 (tiny levels of unusual geometries) and everything under ORBX_FAST_LEGACY=1 go to
    // fast_cells_kernel (round 1: a warp per cell), 2 cells per tile (1 for latency-mode handles), NB tile buffers.
    const char* e_nc = getenv("ORBX_FAST_NC"); const char* e_nb = getenv("ORBX_FAST_NB"); const char* e_w = getenv("ORBX_FAST_WARPS");
    P->fast_legacy = getenv("ORBX_FAST_LEGACY") != nullptr;
    P->fast_nc = e_nc ? atoi(e_nc) : (P->fast_legacy ? (h->cfg.max_batch <= 2 ? 1 : 2) : 4);
    P->fast_nb = (e_nb && P->fast_legacy) ? atoi(e_nb) : 1;
    P->fast_warps = e_w ? atoi(e_w) : 8;                                          // fast_cells_kernel: warps per CTA
    if (P->fast_nc < 1 || P->fast_nc > (P->fast_legacy ? 8 : 4) || P->fast_nb < 1 || P->fast_nb > 2 || P->fast_warps < 1 ||
        P->fast_warps > ORBX_FAST_WARPS)
        return ORBX_ERR_BAD_ARGS;
    int bw = 0, cbw = 64, cbh = 7;
    for (int l = 0; l < n; ++l) {
        OrbxLevel& L = P->lv[l];
        L.wcell_recip = 65536 / L.wCell + 1;
        // latency-mode handles (max_batch <= 2) keep tall cells on the per-cell kernel: a single frame's tall strips would be
        // one more 17-us launch behind the per-cell one on the side stream (640 x 480: 0.087 vs 0.072 ms per frame)
        L.strip_ok = !P->fast_legacy && L.wCell <= 32 && L.hCell <= (h->cfg.max_batch <= 2 ? 32 : 40);
        L.strip_tall = L.strip_ok && L.hCell > 32;
        L.strip_nc = L.strip_ok ? std::max(1, std::min(P->fast_nc, 125 / L.wCell))
                                : (P->fast_legacy ? P->fast_nc : (h->cfg.max_batch <= 2 ? 1 : 2));
        L.strips_x = (L.nColsV + L.strip_nc - 1) / L.strip_nc;
        // 16-aligned TMA start (delta <= 15), 1-byte shift, the strip's cell steps + the 6-px overlap, 2 words of read-ahead
        if (L.strip_ok) {
            bw = std::max(bw, L.strip_nc * L.wCell + 6 + 24);
        } else {
            cbw = std::max(cbw, L.strip_nc * L.wCell + 6 + 24);
            cbh = std::max(cbh, std::min(L.hCell + 6, L.maxBY - ORBX_BOX));
        }
    }
    {   // strip table of a frame, stored behind the resize tap tables (8-byte units), in four segments so that a launch
        // covers a contiguous range: (strip levels | big-cell levels) x (levels 0-1, whose FAST starts early | levels 2+)
        std::vector<uint32_t> tab;
        for (int seg = 0; seg < 6; ++seg) {
            P->seg_first[seg] = (int)tab.size();
            for (int l = 0; l < n; ++l) {
                const int kind = !P->lv[l].strip_ok ? 1 : (P->lv[l].strip_tall ? 2 : 0);     // segment pair of the level
                if (kind != seg / 2 || (l < 2) != ((seg & 1) == 0)) continue;
                for (int i = 0; i < P->lv[l].nRowsV; ++i)
                    for (int j = 0; j < P->lv[l].strips_x; ++j)
                        tab.push_back((uint32_t)l | ((uint32_t)i << 4) | ((uint32_t)(j * P->lv[l].strip_nc) << 16));
            }
            P->seg_count[seg] = (int)tab.size() - P->seg_first[seg];
        }
        P->strips_per_frame = (int)tab.size();
        // fast_strips_kernel's records, one per table entry
        std::vector<OrbxStripRec> recs(tab.size());
        for (size_t s = 0; s < tab.size(); ++s) {
            const int l = (int)(tab[s] & 15u), ci = (int)((tab[s] >> 4) & 0xfffu), cj0 = (int)(tab[s] >> 16);
            const OrbxLevel& L = P->lv[l];
            OrbxStripRec& r = recs[s];
            memset(&r, 0, sizeof r);
            if (!L.strip_ok) continue;
            const int x0 = ORBX_XO + ORBX_BOX + cj0 * L.wCell - 1;               // plane column of (strip window x0 - 1)
            const int delta0 = x0 & 15;
            const int iniY = ORBX_BOX + ci * L.hCell, iniX0 = ORBX_BOX + cj0 * L.wCell;
            const int wh = std::min(iniY + L.hCell + 6, L.maxBY) - iniY;
            const int ncell = std::min(L.strip_nc, L.nColsV - cj0);
            const int sw = std::min(iniX0 + ncell * L.wCell + 6, L.maxBX) - iniX0;
            const int hr = (wh >= 7 && sw >= 7) ? wh - 6 : 0;
            const int lo = delta0 + 4, hi = std::max(lo, delta0 + sw - 2);
            if (hr > (L.strip_tall ? 40 : 32) || ncell < 1 || ncell > 4 || hi > 160 - 4 || L.wcell_recip > 0xffff) return ORBX_ERR_BAD_GEOMETRY;
            r.tile_xy = (uint32_t)(x0 & ~15) | ((uint32_t)(ORBX_EDGE + iniY) << 16);
            r.shape = (uint32_t)l | ((uint32_t)ncell << 4) | ((uint32_t)hr << 8) | ((uint32_t)delta0 << 16) | ((uint32_t)(lo >> 2) << 24);
            r.cols = (uint32_t)lo | ((uint32_t)hi << 16);
            r.cellw = (uint32_t)L.wCell | ((uint32_t)L.wcell_recip << 16);
            r.cell0 = (uint32_t)(L.cell_base + ci * L.nColsV + cj0);
            r.origin = (uint32_t)(cj0 * L.wCell + 3) | ((uint32_t)(ci * L.hCell + 3) << 16);
            r.window = (uint32_t)sw | ((uint32_t)wh << 16);
        }
        if (tab.size() & 1) tab.push_back(0);
        P->strip_tab_off = (int)taps->size();
        const OrbxTap* raw = reinterpret_cast<const OrbxTap*>(tab.data());
        taps->insert(taps->end(), raw, raw + tab.size() / 2);
        if (taps->size() & 1) taps->push_back(OrbxTap());                        // records are read as 16-byte vectors
        P->strip_rec_off = (int)taps->size();
        static_assert(sizeof(OrbxStripRec) == 4 * sizeof(OrbxTap), "record = 4 tap units");
        const OrbxTap* rraw = reinterpret_cast<const OrbxTap*>(recs.data());
        taps->insert(taps->end(), rraw, rraw + recs.size() * 4);
    }
    P->fast_bw = bw <= 96 ? 96 : bw <= 128 ? 128 : 160;                           // fast_strips_kernel instantiations
    P->fast_bh = ORBX_FS_BH;
    if (bw > 160) return ORBX_ERR_BAD_GEOMETRY;                                   // 4 cells of <= 31 px: cannot happen
    cbw = round_up(cbw, 16);
    P->cells_bw = cbw <= 64 ? 64 : cbw <= 96 ? 96 : cbw <= 128 ? 128 : cbw;       // fast_cells_kernel<64 / 96 / 128 / any>
    P->cells_bh = cbh;
    if (P->cells_bw > 256 || P->cells_bh > 127 || P->max_cell_w > 250) return ORBX_ERR_BAD_GEOMETRY;
    P->cells_per_frame = cells;
    P->cand_per_frame = cand;
    P->kept_per_frame = kept;
    P->blur_tiles_per_frame = tiles;
    P->slab_bytes = off;
    // cv::fastAtan2 coefficients as OpenCV forms them: float products (App. A-4)
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    P->atan_p1 = 0.9997878412794807f * scale;
    P->atan_p3 = -0.3258083974640975f * scale;
    P->atan_p5 = 0.1555786518463281f * scale;
    P->atan_p7 = -0.04432655554792128f * scale;
    P->factor_pi = (float)(3.1415926535897932384626433832795 / 180.f);  // (:107)
    static const int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};   // orient_kernel's table
    for (int i = 0; i < 16; ++i) {
        P->umax[i] = h->umax[i];
        if (h->umax[i] != kUmax[i]) return ORBX_ERR_BAD_ARGS;
    }
    while (P->fast_warps > 1 && orbx::fast_smem_bytes(*P) > 200 * 1024) --P->fast_warps;
    if (orbx::fast_smem_bytes(*P) > 200 * 1024 || orbx::octree_smem_bytes(*P) > 200 * 1024) return ORBX_ERR_BAD_GEOMETRY;
    return ORBX_OK;
}

void free_geometry(orbx_handle* h) {
    h->bow_slot.clear();
    for (size_t i = 0; i < h->graphs.size(); ++i)
        if (h->graphs[i].exec) cudaGraphExecDestroy(h->graphs[i].exec);
    h->graphs.clear();
    cudaFree(h->d_plan); cudaFree(h->d_taps); cudaFree(h->d_input); cudaFree(h->d_pyr); cudaFree(h->d_blur);
    cudaFree(h->d_cand); cudaFree(h->d_cand_sorted); cudaFree(h->d_kept); cudaFree(h->d_key_node);
    cudaFree(h->d_cell_rec); cudaFree(h->d_counters); cudaFree(h->d_angles); cudaFree(h->d_out_kp);
    cudaFree(h->d_out_desc);
    cudaFreeHost(h->h_counters); cudaFreeHost(h->h_out_kp); cudaFreeHost(h->h_out_desc); cudaFreeHost(h->h_pyr);
    cudaFreeHost(h->h_input);
    cudaFree(h->d_color); cudaFreeHost(h->h_color); h->d_color = h->h_color = 0; h->color_bytes = 0;
    cudaFree(h->d_un_xy); cudaFree(h->d_un_start); cudaFree(h->d_un_items); cudaFree(h->d_un_frames);
    cudaFreeHost(h->h_un_xy); cudaFreeHost(h->h_un_start); cudaFreeHost(h->h_un_items); cudaFreeHost(h->h_un_frames);
    cudaFree(h->d_sp); cudaFreeHost(h->h_sp); cudaFree(h->d_sp_out); cudaFreeHost(h->h_sp_out);
    cudaFree(h->d_bow); cudaFreeHost(h->h_bow); cudaFree(h->d_bow_val); cudaFreeHost(h->h_bow_val);
    h->d_bow = h->h_bow = 0; h->d_bow_val = h->h_bow_val = 0;
    h->d_sp = h->h_sp = 0; h->sp_bytes = 0; h->d_sp_out = h->h_sp_out = 0; h->sp_out_ints = 0;
    h->d_un_xy = h->h_un_xy = 0; h->d_un_start = h->d_un_items = h->d_un_frames = h->h_un_start = h->h_un_items = h->h_un_frames = 0;
    cudaFree(h->d_st_u); cudaFree(h->d_st_depth); cudaFree(h->d_st_sad); cudaFree(h->d_st_pairs); cudaFree(h->d_st_rows); cudaFree(h->d_st_bucket); h->d_st_rows = 0; h->d_st_bucket = 0;
    cudaFreeHost(h->h_st); cudaFreeHost(h->h_st_pairs);
    h->d_st_u = h->d_st_depth = h->h_st = 0; h->d_st_sad = h->d_st_pairs = h->h_st_pairs = 0;
    h->d_plan = 0; h->d_taps = 0; h->d_input = h->d_pyr = h->d_blur = 0;
    h->d_cand = h->d_cand_sorted = h->d_kept = 0; h->d_key_node = 0; h->d_cell_rec = 0; h->d_counters = 0;
    h->d_angles = 0; h->d_out_kp = 0; h->d_out_desc = 0;
    h->h_counters = 0; h->h_out_kp = 0; h->h_out_desc = 0; h->h_pyr = 0; h->h_input = 0;
    h->have_plan = false;
    h->pyramid_valid = false;
    h->last_n = 0;
}

int ensure_geometry(orbx_handle* h, int w, int hgt) {
    if (h->have_plan && h->plan.width == w && h->plan.height == hgt) return ORBX_OK;
    OrbxPlan P;
    std::vector<OrbxTap> taps;
    int rc = build_plan(h, w, hgt, &P, &taps);
    if (rc != ORBX_OK) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    free_geometry(h);
    h->plan = P;
    h->taps.swap(taps);
    const size_t B = (size_t)h->cfg.max_batch;
    h->in_pitch = round_up(w, 64) + 64;      // room to adopt the caller's own row stride (1-D H2D copies)
    CK(h, cudaMalloc(&h->d_plan, sizeof(OrbxPlan)));
    CK(h, cudaMalloc(&h->d_taps, sizeof(OrbxTap) * (h->taps.size() + 1)));
    CK(h, cudaMalloc(&h->d_input, B * hgt * h->in_pitch));
    CK(h, cudaMalloc(&h->d_pyr, B * P.slab_bytes));
    CK(h, cudaMalloc(&h->d_cand, B * P.cand_per_frame * 4));
    CK(h, cudaMalloc(&h->d_cand_sorted, B * P.cand_per_frame * 4));
    CK(h, cudaMalloc(&h->d_key_node, B * P.cand_per_frame * 2));
    CK(h, cudaMalloc(&h->d_cell_rec, B * P.cells_per_frame * sizeof(uint2)));
    CK(h, cudaMalloc(&h->d_counters, sizeof(int) * h->counters_count()));
    CK(h, cudaMalloc(&h->d_kept, B * P.kept_per_frame * 4));
    CK(h, cudaMalloc(&h->d_angles, B * P.kept_per_frame * 4));
    CK(h, cudaMalloc(&h->d_out_kp, B * P.kept_per_frame * sizeof(orbx_keypoint)));
    CK(h, cudaMalloc(&h->d_out_desc, B * P.kept_per_frame * 32));
    CK(h, cudaMallocHost(&h->h_counters, sizeof(int) * h->counters_count()));
    CK(h, cudaMallocHost(&h->h_out_kp, B * P.kept_per_frame * sizeof(orbx_keypoint)));
    CK(h, cudaMallocHost(&h->h_out_desc, B * P.kept_per_frame * 32));
    CK(h, cudaMallocHost(&h->h_input, B * hgt * h->in_pitch));
    if (h->cfg.download_pyramid) CK(h, cudaMallocHost(&h->h_pyr, B * P.slab_bytes));
    h->fast_maps.resize(orbx::fast_maps_bytes());
    h->desc_maps.resize(orbx::fast_maps_bytes());
    h->pyr_maps.resize(orbx::fast_maps_bytes());
    if (orbx::build_pyr_tile_maps(h->plan, h->d_pyr, h->cfg.max_batch, h->pyr_maps.data()) != 0 ||
        orbx::build_fast_maps(h->plan, h->d_pyr, h->cfg.max_batch, h->fast_maps.data()) != 0 ||
        orbx::build_describe_maps(h->plan, h->d_pyr, h->cfg.max_batch, h->desc_maps.data()) != 0) {
        h->last_error = "cuTensorMapEncodeTiled failed";
        return ORBX_ERR_CUDA;
    }
    CK(h, cudaMemcpyAsync(h->d_plan, &h->plan, sizeof(OrbxPlan), cudaMemcpyHostToDevice, h->stream));
    if (!h->taps.empty())
        CK(h, cudaMemcpyAsync(h->d_taps, h->taps.data(), sizeof(OrbxTap) * h->taps.size(), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemsetAsync(h->d_pyr, 0, B * P.slab_bytes, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->have_plan = true;
    return ORBX_OK;
}

void timing_collect(orbx_handle* h, int upto_pending) {
    // accumulate the oldest `upto_pending` recorded event sets (their work has to be complete)
    while (upto_pending-- > 0 && h->ev_pending > 0) {
        const int idx = (h->ev_head - h->ev_pending + kTimingRing * 2) % kTimingRing;
        cudaEventSynchronize(h->ev[idx][ST_COUNT]);
        for (int s = 0; s < ST_COUNT; ++s) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->ev[idx][s], h->ev[idx][s + 1]) == cudaSuccess) h->stage_ms[s] += ms;
        }
        --h->ev_pending;
    }
}

// Enqueues the whole path for frames [f0, f0 + n) on stream st.  All per-frame arrays are frame-major, so a chunk is
// addressed by offsetting the base pointers; only the TMA tile fetch needs the absolute frame index (frame0).
int enqueue_frames(orbx_handle* h, int f0, int n, const uint8_t* d_imgs, size_t pitch, size_t frame_stride, int chunk,
                   cudaStream_t st) {
    const OrbxPlan& P = h->plan;
    const int L = P.nlevels;
    cudaEvent_t* ev = 0;
    if (h->timing) {
        if (h->ev_pending == kTimingRing) timing_collect(h, 1);
        ev = h->ev[h->ev_head];
    }
    uint8_t* pyr = h->d_pyr + (size_t)f0 * P.slab_bytes;
    uint32_t* cand = h->d_cand + (size_t)f0 * P.cand_per_frame;
    uint32_t* cand_sorted = h->d_cand_sorted + (size_t)f0 * P.cand_per_frame;
    uint16_t* key_node = h->d_key_node + (size_t)f0 * P.cand_per_frame;
    uint2* cell_rec = h->d_cell_rec + (size_t)f0 * P.cells_per_frame;
    uint32_t* kept = h->d_kept + (size_t)f0 * P.kept_per_frame;
    float* angles = h->d_angles + (size_t)f0 * P.kept_per_frame;
    float* out_kp = h->d_out_kp + (size_t)f0 * P.kept_per_frame * 7;
    uint8_t* out_desc = h->d_out_desc + (size_t)f0 * P.kept_per_frame * 32;
    int* level_counts = h->d_level_counts() + f0 * L;
    int* sorted_counts = h->d_sorted_counts() + f0 * L;
    int* kept_counts = h->d_kept_counts() + f0 * L;
    int* status = h->d_status() + f0;
    int* retry_counts = h->d_retry_counts() + f0 * L;
    // Two schedules.  With per-stage timing on, every kernel runs alone on `st`, so its duration is its own.
    // Otherwise the dependency graph is exploited with a side stream:
    //   st : pyramid L0..1 | pyramid L2..  (small, latency-bound levels)  | FAST L2.. (strips)            | octree | describe
    //   aux:               | FAST L0..1 (needs only those levels)         | FAST of levels with big cells |
    int si = 0;
    for (int i = 1; i < 4; ++i) if (st == h->ks[i]) si = i;
    const bool overlap = ev == 0;
    const int ls = (overlap && L > 2) ? 2 : 0;                       // levels [0, ls) get their own FAST launch
    cudaStream_t ax = h->aux[si];
    int* wc = h->d_work_counter() + ORBX_MAXL * chunk;                // FAST work counters of this chunk, one per level
    if (ev) CK(h, cudaEventRecord(ev[ST_PYRAMID], st));
    for (int l = 0; l < L; ++l) {
        orbx::launch_pyr_level(h->d_plan, P, l, n, h->num_sms, d_imgs, pitch, frame_stride, pyr, h->d_taps, st, h->pyr_maps.data(), f0);
        if (ls && l == ls - 1) {
            CK(h, cudaEventRecord(h->ev_low[si], st));
            CK(h, cudaStreamWaitEvent(ax, h->ev_low[si], 0));
            for (int seg = 0; seg < 6; seg += 2)                         // levels 0-1: strips, then big cells and tall cells (if any)
                CK(h, orbx::launch_fast(h->d_plan, P, h->fast_maps.data(), h->d_taps, f0, n, seg, 1, h->num_sms, cand, cell_rec,
                                        level_counts, wc, status, retry_counts, ax));
            CK(h, cudaEventRecord(h->ev_fast_low[si], ax));
        }
    }
    if (ev) CK(h, cudaEventRecord(ev[ST_FAST], st));
    if (ls) {
        // the (small) levels with tall cells, or cells too large for fast_strips_kernel, run beside it on the side stream
        // (a single frame's critical path is one strip per CTA either way: 17 us each, measured)
        if (P.seg_count[3] + P.seg_count[5] > 0) {
            CK(h, cudaEventRecord(h->ev_low[si], st));
            CK(h, cudaStreamWaitEvent(ax, h->ev_low[si], 0));
            for (int seg = 5; seg >= 3; seg -= 2)
                CK(h, orbx::launch_fast(h->d_plan, P, h->fast_maps.data(), h->d_taps, f0, n, seg, 1, h->num_sms, cand, cell_rec,
                                        level_counts, wc, status, retry_counts, ax));
            CK(h, cudaEventRecord(h->ev_fast_low[si], ax));
        }
        CK(h, orbx::launch_fast(h->d_plan, P, h->fast_maps.data(), h->d_taps, f0, n, 1, 1, h->num_sms, cand, cell_rec, level_counts,
                                wc, status, retry_counts, st));
        CK(h, cudaStreamWaitEvent(st, h->ev_fast_low[si], 0));           // the quadtree needs every level's candidates
    } else {
        for (int seg = 0; seg < 6; seg += 2)                             // all strip levels, all big-cell levels, all tall-cell levels
            CK(h, orbx::launch_fast(h->d_plan, P, h->fast_maps.data(), h->d_taps, f0, n, seg, 2, h->num_sms, cand, cell_rec,
                                    level_counts, wc, status, retry_counts, st));
    }
    int fast_launches = 0;
    for (int seg = 0; seg < 6; seg += 2)
        fast_launches += ls ? (P.seg_count[seg] > 0) + (P.seg_count[seg + 1] > 0) : (P.seg_count[seg] + P.seg_count[seg + 1] > 0);
    if (ev) CK(h, cudaEventRecord(ev[ST_OCTREE], st));
    CK(h, orbx::launch_octree(h->d_plan, P, n, cand, cell_rec, cand_sorted, key_node, sorted_counts, kept, kept_counts,
                              status, st));
    if (ev) CK(h, cudaEventRecord(ev[ST_DESCRIBE], st));
    CK(h, orbx::launch_describe(h->d_plan, P, h->desc_maps.data(), f0, n, h->num_sms, kept, kept_counts, angles, out_kp, out_desc,
                                st));
    if (ev) {
        CK(h, cudaEventRecord(ev[ST_COUNT], st));
        h->ev_head = (h->ev_head + 1) % kTimingRing;
        ++h->ev_pending;
        h->stage_launches[ST_PYRAMID] += L;
        for (int s = ST_FAST; s < ST_COUNT; ++s) h->stage_launches[s] += s == ST_FAST ? fast_launches : 1;
    }
    h->launches += L + 2 + fast_launches;
    CK(h, cudaGetLastError());
    return ORBX_OK;
}

// The launch sequence of a sub-batch (nlevels + 4 kernels, two streams, three events) is identical from call to call
// for a given (first frame, frame count, input pointer, pitch, stream): the second time a sequence is seen it is
// stream-captured into a CUDA graph and from then on replayed with ONE launch.  For small images the path is bound by
// host launch overhead (640x480: 56 launches per 32-frame call), so this is where end-to-end throughput and single-frame
// latency come from.  Off while per-stage timing is on (events between kernels) and with ORBX_NO_GRAPHS=1.
int enqueue_frames_graphed(orbx_handle* h, int f0, int n, const uint8_t* d_imgs, size_t pitch, size_t frame_stride, int chunk,
                           cudaStream_t st) {
    if (!h->use_graphs || h->timing) return enqueue_frames(h, f0, n, d_imgs, pitch, frame_stride, chunk, st);
    int si = 0;
    for (int i = 1; i < 4; ++i) if (st == h->ks[i]) si = i;
    orbx_handle::GraphEntry* e = 0;
    for (size_t i = 0; i < h->graphs.size(); ++i) {
        orbx_handle::GraphEntry& g = h->graphs[i];
        if (g.f0 == f0 && g.n == n && g.chunk == chunk && g.si == si && g.pitch == pitch && g.frame_stride == frame_stride &&
            g.d_imgs == d_imgs) { e = &g; break; }
    }
    if (!e) {                                  // first sight: run eagerly (also performs the one-time kernel attribute set-up)
        if (h->graphs.size() < 64) {
            orbx_handle::GraphEntry g = {f0, n, chunk, si, pitch, frame_stride, d_imgs, 0, 0};
            h->graphs.push_back(g);
        }
        return enqueue_frames(h, f0, n, d_imgs, pitch, frame_stride, chunk, st);
    }
    if (!e->exec) {                            // second sight: capture
        const long long before = h->launches;
        cudaGraph_t graph = 0;
        CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_frames(h, f0, n, d_imgs, pitch, frame_stride, chunk, st);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != ORBX_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            h->use_graphs = false;             // capture is not possible here: stay on the eager path for good
            h->launches = before;
            return enqueue_frames(h, f0, n, d_imgs, pitch, frame_stride, chunk, st);
        }
        e->launches = (int)(h->launches - before);
        h->launches = before;
        const cudaError_t ie = cudaGraphInstantiate(&e->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            e->exec = 0;
            cudaGetLastError();
            h->use_graphs = false;
            return enqueue_frames(h, f0, n, d_imgs, pitch, frame_stride, chunk, st);
        }
    }
    CK(h, cudaGraphLaunch(e->exec, st));
    h->launches += e->launches;                // kernels the replay launches
    return ORBX_OK;
}

int enqueue_pipeline(orbx_handle* h, int n, const uint8_t* d_imgs, size_t pitch, size_t frame_stride) {
    CK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(int) * h->counters_count(), h->stream));
    // Two half-batches on the two kernel streams: the latency-bound tail of one half (octree, orientation,
    // descriptors) overlaps the issue-bound head (pyramid, FAST) of the other.  With per-stage timing on, one
    // stream runs everything so each kernel's duration is its own.
    const char* e_ch = getenv("ORBX_DEVICE_CHUNKS");          // tuning override
    int nchunks = (!h->timing && n >= 8) ? (e_ch ? atoi(e_ch) : h->cfg.device_chunks > 0 ? h->cfg.device_chunks : 2) : 1;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > kMaxChunks) nchunks = kMaxChunks;
    if (nchunks > n) nchunks = n;
    const int ns = nchunks < h->nks ? nchunks : h->nks;
    if (ns > 1) {
        CK(h, cudaEventRecord(h->ev_clear, h->stream));
        for (int i = 1; i < ns; ++i) CK(h, cudaStreamWaitEvent(h->ks[i], h->ev_clear, 0));
    }
    for (int k = 0; k < nchunks; ++k) {
        const int f0 = (int)((long long)n * k / nchunks), f1 = (int)((long long)n * (k + 1) / nchunks);
        int rc = enqueue_frames_graphed(h, f0, f1 - f0, d_imgs + (size_t)f0 * frame_stride, pitch, frame_stride, k, h->ks[k % ns]);
        if (rc != ORBX_OK) return rc;
    }
    for (int i = 1; i < ns; ++i) {      // join: everything recorded on the handle's stream after this call covers all sub-batches
        CK(h, cudaEventRecord(h->ev_join[i], h->ks[i]));
        CK(h, cudaStreamWaitEvent(h->stream, h->ev_join[i], 0));
    }
    h->last_n = n;
    ++h->gen;
    h->pyramid_valid = false;
    return ORBX_OK;
}

int finish_results(orbx_handle* h, int n, orbx_result* results);

int fetch(orbx_handle* h, int n, orbx_result* results) {
    const OrbxPlan& P = h->plan;
    cudaStream_t st = h->stream;
    const size_t kpf = (size_t)P.kept_per_frame;
    CK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * h->counters_count(), cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_out_kp, h->d_out_kp, n * kpf * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_out_desc, h->d_out_desc, n * kpf * 32, cudaMemcpyDeviceToHost, st));
    if (h->cfg.download_pyramid)
        CK(h, cudaMemcpyAsync(h->h_pyr, h->d_pyr, (size_t)n * P.slab_bytes, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    return finish_results(h, n, results);
}

int finish_results(orbx_handle* h, int n, orbx_result* results) {
    const OrbxPlan& P = h->plan;
    const size_t kpf = (size_t)P.kept_per_frame;
    h->pyramid_valid = h->cfg.download_pyramid != 0;
    int rc = ORBX_OK;
    for (int f = 0; f < n; ++f) {
        int total = 0;
        for (int l = 0; l < P.nlevels; ++l) total += h->h_kept_counts()[f * P.nlevels + l];
        const int dev = h->h_status()[f];
        int fs = ORBX_OK;
        if (dev & ORBX_DEV_CAND_OVERFLOW) fs = ORBX_ERR_CANDIDATE_OVERFLOW;
        else if (dev & ORBX_DEV_NODE_OVERFLOW) fs = ORBX_ERR_CUDA;
        if (fs != ORBX_OK) { rc = fs; total = 0; }
        if (results) {
            results[f].n = total;
            results[f].status = fs;
            results[f].kps = reinterpret_cast<const orbx_keypoint*>(h->h_out_kp) + f * kpf;
            results[f].desc = h->h_out_desc + f * kpf * 32;
        }
    }
    if (rc == ORBX_ERR_CUDA) h->last_error = "octree node capacity exceeded (internal invariant broken)";
    return rc;
}

}  // namespace

extern "C" {

const char* orbx_version(void) { return "orbx-b200 0.2 (sm_100a)"; }

int orbx_bind_thread_to_device(int device, int* node_out, int* ncpus_out) {
    if (node_out) *node_out = -1;
    if (ncpus_out) *ncpus_out = 0;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return ORBX_ERR_NO_DEVICE;
    }
    for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE* f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    if (node_out) *node_out = node;
    if (node < 0) return ORBX_OK;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return ORBX_OK;
    char list[4096] = {0};
    const bool got = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!got) return ORBX_OK;
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return ORBX_OK;
    int n = 0;
    for (char* p = list; *p;) {                       // "0-31,64-95"
        char* e;
        const long a = strtol(p, &e, 10);
        if (e == p) break;
        long b = a;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET((int)c, &allowed)) { CPU_SET((int)c, &want); ++n; }
        p = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    if (n > 0 && sched_setaffinity(0, sizeof want, &want) == 0 && ncpus_out) *ncpus_out = n;
    return ORBX_OK;
}

const char* orbx_strerror(int s) {
    switch (s) {
        case ORBX_OK: return "ok";
        case ORBX_ERR_BAD_ARGS: return "bad arguments";
        case ORBX_ERR_BAD_GEOMETRY: return "unsupported image geometry (level < 62 px, > 4096 px, or degenerate aspect ratio)";
        case ORBX_ERR_CUDA: return "CUDA error";
        case ORBX_ERR_CANDIDATE_OVERFLOW: return "FAST candidate buffer overflow (lower candidate_divisor)";
        case ORBX_ERR_NO_DEVICE: return "no usable CUDA device";
        case ORBX_ERR_OUT_OF_MEMORY: return "out of device or pinned memory";
        case ORBX_ERR_EMPTY_IMAGE: return "empty image: outputs untouched";
        default: return "unknown status";
    }
}

const char* orbx_last_cuda_error(orbx_handle* h) { return h ? h->last_error.c_str() : ""; }

int orbx_create(const orbx_config* cfg, orbx_handle** out) {
    if (!cfg || !out) return ORBX_ERR_BAD_ARGS;
    *out = 0;
    // scale factors above 2 would make a lane's 4 output pixels read past the 19-px border of the source row
    if (cfg->nlevels < 1 || cfg->nlevels > ORBX_MAXL || cfg->nfeatures < 1 || !(cfg->scale_factor > 1.0f) || cfg->scale_factor > 2.0f ||
        cfg->ini_th_fast < 1 || cfg->min_th_fast < 1 || cfg->ini_th_fast > 254 || cfg->min_th_fast > 254 ||
        cfg->max_batch < 1)
        return ORBX_ERR_BAD_ARGS;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return ORBX_ERR_NO_DEVICE;
    }
    orbx_handle* h = new orbx_handle();
    h->cfg = *cfg;
    h->launches = 0;
    h->have_plan = false;
    h->d_plan = 0; h->d_taps = 0; h->d_input = h->d_pyr = h->d_blur = 0;
    h->d_cand = h->d_cand_sorted = h->d_kept = 0; h->d_key_node = 0; h->d_cell_rec = 0; h->d_counters = 0;
    h->d_angles = 0; h->d_out_kp = 0; h->d_out_desc = 0;
    h->h_counters = 0; h->h_out_kp = 0; h->h_out_desc = 0; h->h_pyr = 0; h->h_input = 0;
    h->d_st_u = h->d_st_depth = h->h_st = 0; h->d_st_sad = h->d_st_pairs = h->h_st_pairs = 0; h->d_st_rows = 0; h->d_st_bucket = 0;
    h->last_n = 0; h->pyramid_valid = false;
    h->use_graphs = getenv("ORBX_NO_GRAPHS") == nullptr;
    h->d_un_xy = h->h_un_xy = 0; h->d_un_start = h->d_un_items = h->d_un_frames = h->h_un_start = h->h_un_items = h->h_un_frames = 0;
    h->d_color = h->h_color = 0; h->color_bytes = 0;
    h->d_sp = h->h_sp = 0; h->sp_bytes = 0; h->d_sp_out = h->h_sp_out = 0; h->sp_out_ints = 0;
    h->d_bow = h->h_bow = 0; h->d_bow_val = h->h_bow_val = 0;
    h->d_fr = h->h_fr = 0; h->fr_bytes = 0;
    h->timing = false; h->ev_created = false; h->ev_head = 0; h->ev_pending = 0;
    memset(h->stage_ms, 0, sizeof h->stage_ms);
    memset(h->stage_launches, 0, sizeof h->stage_launches);
    h->stream = 0;
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    h->h2d_stream = h->d2h_stream = 0;
    h->ks[0] = h->stream;
    for (int i = 1; i < 4; ++i) {
        h->ks[i] = 0;
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->ks[i], cudaStreamNonBlocking);
    }
    {
        const char* e_ns = getenv("ORBX_KERNEL_STREAMS");
        h->nks = e_ns ? atoi(e_ns) : 2;
        if (h->nks < 1) h->nks = 1;
        if (h->nks > 4) h->nks = 4;
    }
    for (int i = 0; i < 4; ++i) {
        h->aux[i] = 0;
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_low[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fast_low[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_clear, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_stereo, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking);
    for (int i = 0; i < kMaxChunks && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
    if (e != cudaSuccess) {
        delete h;
        return ORBX_ERR_CUDA;
    }
    build_ctor_tables(h);
    *out = h;
    return ORBX_OK;
}

int orbx_destroy(orbx_handle* h) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    for (int i = 1; i < 4; ++i) cudaStreamSynchronize(h->ks[i]);
    for (int i = 0; i < 4; ++i) cudaStreamSynchronize(h->aux[i]);
    cudaStreamSynchronize(h->h2d_stream);
    cudaStreamSynchronize(h->d2h_stream);
    free_geometry(h);
    cudaFree(h->d_fr); cudaFreeHost(h->h_fr);
    if (h->ev_created)
        for (int i = 0; i < kTimingRing; ++i)
            for (int s = 0; s <= ST_COUNT; ++s) cudaEventDestroy(h->ev[i][s]);
    for (int i = 0; i < kMaxChunks; ++i) { cudaEventDestroy(h->ev_h2d[i]); cudaEventDestroy(h->ev_done[i]); }
    cudaEventDestroy(h->ev_clear);
    cudaEventDestroy(h->ev_stereo);
    for (int i = 0; i < 4; ++i) {
        cudaEventDestroy(h->ev_low[i]);
        cudaEventDestroy(h->ev_fast_low[i]); cudaEventDestroy(h->ev_join[i]); cudaStreamDestroy(h->aux[i]);
    }
    for (int i = 1; i < 4; ++i) cudaStreamDestroy(h->ks[i]);
    cudaStreamDestroy(h->h2d_stream);
    cudaStreamDestroy(h->d2h_stream);
    cudaStreamDestroy(h->stream);
    delete h;
    return ORBX_OK;
}

int orbx_extract_device(orbx_handle* h, int n, const uint8_t* d_imgs, int width, int height, size_t pitch,
                        size_t frame_stride) {
    if (!h || !d_imgs || n < 1 || n > h->cfg.max_batch || pitch < (size_t)width) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    int rc = ensure_geometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    return enqueue_pipeline(h, n, d_imgs, pitch, frame_stride);
}

int orbx_fetch_results(orbx_handle* h, int n, orbx_result* results) {
    if (!h || !h->have_plan || n < 1 || n > h->last_n) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    return fetch(h, n, results);
}

int orbx_extract_batch(orbx_handle* h, int n, const uint8_t* const* imgs, int width, int height,
                       const size_t* strides, orbx_result* results) {
    if (!h || !imgs || !results || n < 1 || n > h->cfg.max_batch) return ORBX_ERR_BAD_ARGS;
    if (width == 0 || height == 0) return ORBX_ERR_EMPTY_IMAGE;       // (:1046-1047)
    if (width < 0 || height < 0) return ORBX_ERR_BAD_ARGS;
    for (int i = 0; i < n; ++i)
        if (!imgs[i] || (strides && strides[i] < (size_t)width)) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    int rc = ensure_geometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    const OrbxPlan& P = h->plan;
    // The staging buffer adopts the caller's row stride when all frames share one that fits: frames then go up as plain
    // 1-D copies (a strided 2-D cudaMemcpy of a 752x480 frame measured ~10x slower than the 1-D copy of the same bytes).
    size_t pitch = (size_t)round_up(width, 16);
    {
        const size_t s0 = strides ? strides[0] : (size_t)width;
        bool same = s0 <= (size_t)h->in_pitch;
        for (int i = 1; i < n && same; ++i) same = (strides ? strides[i] : (size_t)width) == s0;
        if (same) pitch = s0;
    }
    const size_t fbytes = (size_t)height * pitch;
    const size_t kpf = (size_t)P.kept_per_frame;
    // Sub-batches pipeline the three engines: H2D of chunk k+1, kernels of chunk k and D2H of chunk k-1 overlap.
    int nchunks = n >= 16 ? 4 : (n >= 4 ? 2 : 1);
    if (nchunks > kMaxChunks) nchunks = kMaxChunks;
    CK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(int) * h->counters_count(), h->stream));
    // a single sub-batch (latency mode) stays on one stream: no cross-stream events on the critical path
    const bool one = nchunks == 1;
    cudaStream_t h2d = one ? h->stream : h->h2d_stream, d2h = one ? h->stream : h->d2h_stream;
    const int ns = nchunks < h->nks ? nchunks : h->nks;
    if (!one) {
        CK(h, cudaEventRecord(h->ev_clear, h->stream));
        for (int i = 1; i < ns; ++i) CK(h, cudaStreamWaitEvent(h->ks[i], h->ev_clear, 0));
    }
    for (int k = 0; k < nchunks; ++k) {
        cudaStream_t cs = h->ks[k % ns];
        const int f0 = (int)((long long)n * k / nchunks), f1 = (int)((long long)n * (k + 1) / nchunks);
        for (int i = f0; i < f1;) {
            const size_t stride = strides ? strides[i] : (size_t)width;
            cudaPointerAttributes attr;
            bool pinned = cudaPointerGetAttributes(&attr, imgs[i]) == cudaSuccess && attr.type == cudaMemoryTypeHost;
            if (!pinned) cudaGetLastError();
            if (pinned && stride == pitch) {
                // same pitch on both sides: plain 1-D copies, and frames that are contiguous in the caller's
                // pinned buffer (a ring of camera frames) go as ONE copy
                int j = i + 1;
                while (j < f1 && imgs[j] == imgs[j - 1] + fbytes && (strides ? strides[j] : (size_t)width) == stride) ++j;
                // the ABI promises `width` valid bytes in a row, not `stride`: the last row of the run ends at its last pixel
                CK(h, cudaMemcpyAsync(h->d_input + i * fbytes, imgs[i], (size_t)(j - i) * fbytes - (pitch - (size_t)width),
                                      cudaMemcpyHostToDevice, h2d));
                i = j;
                continue;
            }
            const uint8_t* src = imgs[i];
            size_t spitch = stride;
            if (!pinned) {     // pageable caller memory: stage through the handle's pinned buffer
                uint8_t* stg = h->h_input + i * fbytes;
                for (int y = 0; y < height; ++y) memcpy(stg + (size_t)y * pitch, imgs[i] + (size_t)y * stride, (size_t)width);
                CK(h, cudaMemcpyAsync(h->d_input + i * fbytes, stg, fbytes, cudaMemcpyHostToDevice, h2d));
            } else {
                CK(h, cudaMemcpy2DAsync(h->d_input + i * fbytes, pitch, src, spitch, (size_t)width, (size_t)height,
                                        cudaMemcpyHostToDevice, h2d));
            }
            ++i;
        }
        if (!one) {
            CK(h, cudaEventRecord(h->ev_h2d[k], h2d));
            CK(h, cudaStreamWaitEvent(cs, h->ev_h2d[k], 0));
        }
        rc = enqueue_frames_graphed(h, f0, f1 - f0, h->d_input + f0 * fbytes, pitch, fbytes, k, cs);
        if (rc != ORBX_OK) return rc;
        if (!one) {
            CK(h, cudaEventRecord(h->ev_done[k], cs));
            CK(h, cudaStreamWaitEvent(d2h, h->ev_done[k], 0));
        }
        CK(h, cudaMemcpyAsync(h->h_out_kp + f0 * kpf * 7, h->d_out_kp + f0 * kpf * 7, (f1 - f0) * kpf * sizeof(orbx_keypoint),
                              cudaMemcpyDeviceToHost, d2h));
        CK(h, cudaMemcpyAsync(h->h_out_desc + f0 * kpf * 32, h->d_out_desc + f0 * kpf * 32, (f1 - f0) * kpf * 32,
                              cudaMemcpyDeviceToHost, d2h));
        if (h->cfg.download_pyramid)
            CK(h, cudaMemcpyAsync(h->h_pyr + (size_t)f0 * P.slab_bytes, h->d_pyr + (size_t)f0 * P.slab_bytes,
                                  (size_t)(f1 - f0) * P.slab_bytes, cudaMemcpyDeviceToHost, d2h));
    }
    CK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * h->counters_count(), cudaMemcpyDeviceToHost, d2h));
    CK(h, cudaStreamSynchronize(d2h));
    if (!one)
        for (int i = 0; i < ns; ++i) CK(h, cudaStreamSynchronize(h->ks[i]));
    h->last_n = n;
    ++h->gen;
    return finish_results(h, n, results);
}

int orbx_extract(orbx_handle* h, const uint8_t* img, int width, int height, size_t stride, orbx_result* result) {
    const uint8_t* imgs[1] = {img};
    const size_t strides[1] = {stride};
    return orbx_extract_batch(h, 1, imgs, width, height, strides, result);
}


namespace {
int stereo_enqueue(orbx_handle* L, orbx_handle* R, int npairs, const int* left_frames, const int* right_frames, float mbf,
                   float mb) {
    if (!L || !R || npairs < 1 || npairs > L->cfg.max_batch || !L->have_plan || !R->have_plan) return ORBX_ERR_BAD_ARGS;
    if (!(mb > 0.f) || !(mbf > 0.f)) return ORBX_ERR_BAD_ARGS;
    // row bands of at most 24 rows (2 * ceil(2 * scale) + 2) and at most 15 keypoints per thread of the filter kernel
    if (L->sf[L->cfg.nlevels - 1] > 5.4f || L->plan.kept_per_frame > 15 * 1024 || L->plan.kept_per_frame > 65535) return ORBX_ERR_BAD_GEOMETRY;
    const OrbxPlan& P = L->plan;
    // both eyes must have been extracted with the same constructor arguments and image size on the same device
    if (L->cfg.device != R->cfg.device || R->plan.width != P.width || R->plan.height != P.height ||
        R->plan.nlevels != P.nlevels || R->plan.kept_per_frame != P.kept_per_frame || R->plan.slab_bytes != P.slab_bytes ||
        R->cfg.scale_factor != L->cfg.scale_factor)
        return ORBX_ERR_BAD_ARGS;
    if (L->d_st_u && R->cfg.max_batch > L->st_right_cap) return ORBX_ERR_BAD_ARGS;   // bands buffer was sized for a smaller right handle
    for (int i = 0; i < npairs; ++i) {
        const int fl = left_frames ? left_frames[i] : i, fr = right_frames ? right_frames[i] : i;
        if (fl < 0 || fl >= L->last_n || fr < 0 || fr >= R->last_n) return ORBX_ERR_BAD_ARGS;
        for (int k = 0; k < i; ++k)            // results are stored per left frame and row tables per right frame: each
            if ((left_frames ? left_frames[k] : k) == fl || (right_frames ? right_frames[k] : k) == fr)   // may appear once per call
                return ORBX_ERR_BAD_ARGS;
    }
    CK(L, cudaSetDevice(L->cfg.device));
    const size_t B = (size_t)L->cfg.max_batch, kpf = (size_t)P.kept_per_frame;
    if (!L->d_st_u) {
        CK(L, cudaMalloc(&L->d_st_u, B * kpf * 4));
        CK(L, cudaMalloc(&L->d_st_depth, B * kpf * 4));
        CK(L, cudaMalloc(&L->d_st_sad, B * kpf * 4));
        CK(L, cudaMalloc(&L->d_st_pairs, B * 2 * sizeof(int)));
        CK(L, cudaMalloc(&L->d_st_rows, (size_t)R->cfg.max_batch * (P.height + 1) * sizeof(int)));
        CK(L, cudaMalloc(&L->d_st_bucket, (size_t)R->cfg.max_batch * orbx::stereo_bucket_entries(P) * sizeof(uint16_t)));
        L->st_right_cap = R->cfg.max_batch;
        L->st_pairs_n = 0;
        CK(L, cudaMallocHost(&L->h_st, 2 * B * kpf * 4));
        CK(L, cudaMallocHost(&L->h_st_pairs, B * 2 * sizeof(int)));
    }
    cudaStream_t st = L->stream;
    // the pair list rarely changes between calls (same batch layout every time): upload it only when it does
    bool same = L->st_pairs_n == npairs;
    for (int i = 0; i < npairs && same; ++i)
        same = L->h_st_pairs[2 * i] == (left_frames ? left_frames[i] : i) && L->h_st_pairs[2 * i + 1] == (right_frames ? right_frames[i] : i);
    if (!same) {
        CK(L, cudaStreamSynchronize(st));                                  // a previous upload from h_st_pairs is complete
        for (int i = 0; i < npairs; ++i) {
            L->h_st_pairs[2 * i] = left_frames ? left_frames[i] : i;
            L->h_st_pairs[2 * i + 1] = right_frames ? right_frames[i] : i;
        }
        L->st_pairs_n = npairs;
        CK(L, cudaMemcpyAsync(L->d_st_pairs, L->h_st_pairs, (size_t)npairs * 2 * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (R != L) {                                                          // the right eye's extraction may still be in flight
        CK(L, cudaEventRecord(L->ev_stereo, R->stream));
        CK(L, cudaStreamWaitEvent(st, L->ev_stereo, 0));
    }
    L->st_gen = L->gen;
    CK(L, orbx::launch_stereo(L->d_plan, P, L->num_sms, L->d_pyr, L->d_out_kp, L->d_out_desc, L->d_kept_counts(), R->d_pyr,
                              R->d_out_kp, R->d_out_desc, R->d_kept_counts(), L->d_st_pairs, npairs, mbf, mb, L->d_st_u,
                              L->d_st_depth, L->d_st_sad, L->d_st_rows, L->d_st_bucket, st));
    L->launches += 3;
    return ORBX_OK;
}
}  // namespace

int orbx_stereo_match_device(orbx_handle* left, orbx_handle* right, int npairs, const int* left_frames, const int* right_frames,
                             float mbf, float mb) {
    return stereo_enqueue(left, right, npairs, left_frames, right_frames, mbf, mb);
}

int orbx_stereo_fetch(orbx_handle* left, int npairs, const int* left_frames, orbx_stereo_result* results) {
    if (!left || !results || npairs < 1 || !left->d_st_u || npairs > left->cfg.max_batch) return ORBX_ERR_BAD_ARGS;
    orbx_handle* h = left;
    CK(h, cudaSetDevice(h->cfg.device));
    const OrbxPlan& P = h->plan;
    const size_t B = (size_t)h->cfg.max_batch, kpf = (size_t)P.kept_per_frame;
    const size_t n = (size_t)h->last_n;
    cudaStream_t st = h->stream;
    CK(h, cudaMemcpyAsync(h->h_st, h->d_st_u, n * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_st + B * kpf, h->d_st_depth, n * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * h->counters_count(), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int i = 0; i < npairs; ++i) {
        const int f = left_frames ? left_frames[i] : i;
        if (f < 0 || f >= h->last_n) return ORBX_ERR_BAD_ARGS;
        int total = 0;
        for (int l = 0; l < P.nlevels; ++l) total += h->h_kept_counts()[f * P.nlevels + l];
        results[i].n = total;
        results[i].u_right = h->h_st + (size_t)f * kpf;
        results[i].depth = h->h_st + B * kpf + (size_t)f * kpf;
    }
    return ORBX_OK;
}

int orbx_stereo_match(orbx_handle* left, orbx_handle* right, int npairs, const int* left_frames, const int* right_frames,
                      float mbf, float mb, orbx_stereo_result* results) {
    if (!results) return ORBX_ERR_BAD_ARGS;
    const int rc = stereo_enqueue(left, right, npairs, left_frames, right_frames, mbf, mb);
    if (rc != ORBX_OK) return rc;
    return orbx_stereo_fetch(left, npairs, left_frames, results);
}

// ---- cvtColor in front of the path: colour frames are converted to gray on the device into the
// handle's gray staging buffer, then the normal launch sequence runs on it.
namespace {
int color_channels(int format) { return (format == ORBX_BGR8 || format == ORBX_RGB8) ? 3 : (format == ORBX_BGRA8 || format == ORBX_RGBA8) ? 4 : 0; }
}

int orbx_extract_device_color(orbx_handle* h, int n, const uint8_t* d_imgs, int width, int height, size_t pitch, size_t frame_stride,
                              int format) {
    const int ch = h ? color_channels(format) : 0;
    if (!h || !d_imgs || ch == 0 || n < 1 || n > h->cfg.max_batch || pitch < (size_t)width * ch) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    int rc = ensure_geometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    const size_t gp = (size_t)round_up(width, 16);
    CK(h, orbx::launch_cvt_gray(d_imgs, pitch, frame_stride, width, height, n, format, h->d_input, gp, gp * height, h->stream));
    h->launches += 1;
    return enqueue_pipeline(h, n, h->d_input, gp, gp * height);
}

int orbx_extract_batch_color(orbx_handle* h, int n, const uint8_t* const* imgs, int width, int height, const size_t* strides,
                             int format, orbx_result* results) {
    const int ch = h ? color_channels(format) : 0;
    if (!h || !imgs || !results || ch == 0 || n < 1 || n > h->cfg.max_batch) return ORBX_ERR_BAD_ARGS;
    if (width == 0 || height == 0) return ORBX_ERR_EMPTY_IMAGE;
    if (width < 0 || height < 0) return ORBX_ERR_BAD_ARGS;
    const size_t rowb = (size_t)width * ch;
    for (int i = 0; i < n; ++i)
        if (!imgs[i] || (strides && strides[i] < rowb)) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    int rc = ensure_geometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    const size_t cp = (size_t)round_up((int)rowb, 16), cf = cp * height;
    if (h->color_bytes < cf * h->cfg.max_batch) {
        CK(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_color); cudaFreeHost(h->h_color);
        h->d_color = h->h_color = 0;
        h->color_bytes = cf * h->cfg.max_batch;
        CK(h, cudaMalloc(&h->d_color, h->color_bytes));
        CK(h, cudaMallocHost(&h->h_color, h->color_bytes));
    }
    cudaStream_t st = h->stream;
    for (int i = 0; i < n; ++i) {
        const size_t stride = strides ? strides[i] : rowb;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, imgs[i]) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();
        if (pinned && stride == cp) {
            CK(h, cudaMemcpyAsync(h->d_color + i * cf, imgs[i], cf - (cp - rowb), cudaMemcpyHostToDevice, st));   // last row: rowb bytes
        } else if (pinned) {
            CK(h, cudaMemcpy2DAsync(h->d_color + i * cf, cp, imgs[i], stride, rowb, (size_t)height, cudaMemcpyHostToDevice, st));
        } else {                                                           // pageable: through the pinned colour staging
            uint8_t* stg = h->h_color + i * cf;
            for (int y = 0; y < height; ++y) memcpy(stg + (size_t)y * cp, imgs[i] + (size_t)y * stride, rowb);
            CK(h, cudaMemcpyAsync(h->d_color + i * cf, stg, cf, cudaMemcpyHostToDevice, st));
        }
    }
    rc = orbx_extract_device_color(h, n, h->d_color, width, height, cp, cf, format);
    if (rc != ORBX_OK) return rc;
    return fetch(h, n, results);
}

// ---- Frame::UndistortKeyPoints + AssignFeaturesToGrid on the device-resident keypoints
namespace {
// cv::undistortPoints(K, D, R = I, P = K) of one point, the published 5-iteration algorithm in double (the kernel
// evaluates the same expression tree); used here for the four image corners of Frame::ComputeImageBounds (:436-464).
void undistort_point_host(const double* cam, double u, double v, float* ox, float* oy) {
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    const double k1 = cam[4], k2 = cam[5], p1 = cam[6], p2 = cam[7], k3 = cam[8];
    const double ifx = 1. / fx, ify = 1. / fy;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
        const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0. * r2 + 0. * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    *ox = (float)(fx * x + 0 * y + cx);
    *oy = (float)(0 * x + fy * y + cy);
}
}  // namespace

int orbx_undistort_grid(orbx_handle* h, int nframes, const int* frames, const float* K4, const float* dist, int ndist,
                        orbx_grid_result* results) {
    if (!h || !K4 || !dist || !results || nframes < 1 || nframes > h->cfg.max_batch || ndist < 4 || ndist > 5 || !h->have_plan)
        return ORBX_ERR_BAD_ARGS;
    if (!(K4[0] != 0.f) || !(K4[1] != 0.f)) return ORBX_ERR_BAD_ARGS;
    for (int i = 0; i < nframes; ++i) {
        const int f = frames ? frames[i] : i;
        if (f < 0 || f >= h->last_n) return ORBX_ERR_BAD_ARGS;
        for (int k = 0; k < i; ++k)
            if ((frames ? frames[k] : k) == f) return ORBX_ERR_BAD_ARGS;      // results are stored per frame
    }
    CK(h, cudaSetDevice(h->cfg.device));
    const OrbxPlan& P = h->plan;
    const size_t B = (size_t)h->cfg.max_batch, kpf = (size_t)P.kept_per_frame, NC = 64 * 48 + 1;
    if (!h->d_un_xy) {
        CK(h, cudaMalloc(&h->d_un_xy, B * kpf * 8));
        CK(h, cudaMalloc(&h->d_un_start, B * NC * 4));
        CK(h, cudaMalloc(&h->d_un_items, B * kpf * 4));
        CK(h, cudaMalloc(&h->d_un_frames, B * 4));
        CK(h, cudaMallocHost(&h->h_un_xy, B * kpf * 8));
        CK(h, cudaMallocHost(&h->h_un_start, B * NC * 4));
        CK(h, cudaMallocHost(&h->h_un_items, B * kpf * 4));
        CK(h, cudaMallocHost(&h->h_un_frames, B * 4));
    }
    // Frame::ComputeImageBounds (:436-464) and the grid scale of the Frame constructors (:155-156)
    double cam[9] = {K4[0], K4[1], K4[2], K4[3], dist[0], dist[1], dist[2], dist[3], ndist > 4 ? dist[4] : 0.f};
    const int distorted = dist[0] != 0.0f;                                   // mDistCoef.at<float>(0) != 0.0 (:406, :438)
    float bounds[4];
    if (distorted) {
        float cx[4], cy[4];
        const double W = (double)(float)P.width, H = (double)(float)P.height;
        undistort_point_host(cam, 0., 0., &cx[0], &cy[0]);
        undistort_point_host(cam, W, 0., &cx[1], &cy[1]);
        undistort_point_host(cam, 0., H, &cx[2], &cy[2]);
        undistort_point_host(cam, W, H, &cx[3], &cy[3]);
        bounds[0] = cx[0] < cx[2] ? cx[0] : cx[2];                             // mnMinX = min(corner 0, corner 2)
        bounds[1] = cx[1] > cx[3] ? cx[1] : cx[3];                             // mnMaxX
        bounds[2] = cy[0] < cy[1] ? cy[0] : cy[1];                             // mnMinY
        bounds[3] = cy[2] > cy[3] ? cy[2] : cy[3];                             // mnMaxY
    } else {
        bounds[0] = 0.f; bounds[1] = (float)P.width; bounds[2] = 0.f; bounds[3] = (float)P.height;
    }
    const float grid[4] = {bounds[0], bounds[2], 64.f / (bounds[1] - bounds[0]), 48.f / (bounds[3] - bounds[2])};
    for (int k = 0; k < 4; ++k) h->un_bounds[k] = bounds[k];
    cudaStream_t st = h->stream;
    CK(h, cudaStreamSynchronize(st));                                        // the pinned staging of a previous call is free
    for (int i = 0; i < nframes; ++i) h->h_un_frames[i] = frames ? frames[i] : i;
    CK(h, cudaMemcpyAsync(h->d_un_frames, h->h_un_frames, (size_t)nframes * 4, cudaMemcpyHostToDevice, st));
    h->un_gen = h->gen;
    CK(h, orbx::launch_undistort_grid(h->d_plan, P, h->d_out_kp, h->d_kept_counts(), h->d_un_frames, nframes, cam, distorted, grid,
                                      h->d_un_xy, h->d_un_start, h->d_un_items, st));
    h->launches += 1;
    const size_t n = (size_t)h->last_n;
    CK(h, cudaMemcpyAsync(h->h_un_xy, h->d_un_xy, n * kpf * 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_un_start, h->d_un_start, n * NC * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_un_items, h->d_un_items, n * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * h->counters_count(), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int i = 0; i < nframes; ++i) {
        const int f = h->h_un_frames[i];
        int total = 0;
        for (int l = 0; l < P.nlevels; ++l) total += h->h_kept_counts()[f * P.nlevels + l];
        results[i].n = total;
        results[i].xy_un = h->h_un_xy + (size_t)f * kpf * 2;
        results[i].cell_start = h->h_un_start + (size_t)f * NC;
        results[i].cell_items = h->h_un_items + (size_t)f * kpf;
        results[i].n_in_grid = results[i].cell_start[NC - 1];
        for (int k = 0; k < 4; ++k) results[i].bounds[k] = bounds[k];
    }
    return ORBX_OK;
}

// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono), src/ORBmatcher.cc:1328-1470, and
// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129: one staging / launch path.
namespace {
struct SpHostQuery {
    int cur_frame, n;
    const float* w3;           // n x 3: world position | (mTrackProjX, mTrackProjY, mTrackProjXR)
    const uint8_t* desc;       // n x 32
    const int32_t* obs;        // Observations(); < 0: no map point
    const uint8_t* flag;       // mvbOutlier (skip where set) | in-view flag (skip where clear); may be NULL
    const int32_t* oct;        // keypoint octave | mnTrackScaleLevel
    const float* ang;          // keypoint angle | mTrackViewCos
    const int32_t* cur_obs;    // local variant: what the frame's keypoints hold; NULL = nothing
    const float *Tc, *Tl;      // LastFrame variant: the two poses
};

// mode 0: LastFrame variant, 1: local-map variant, 2: KeyFrame variant (Relocalization, match_th = ORBdist)
int sp_enqueue(orbx_handle* h, int mode, int nq, const SpHostQuery* q, const float* K4, float mbf, float mb, float th, float nnratio,
               int mono, int check_orientation, int use_stereo, int match_th = 0) {
    const bool local = mode == 1, held = mode != 0;                      // held: cur_obs names what the frame's keypoints hold at entry
    if (!h || !q || nq < 1 || nq > h->cfg.max_batch || !h->have_plan) return ORBX_ERR_BAD_ARGS;
    if (!h->d_un_xy || h->un_gen != h->gen) return ORBX_ERR_BAD_ARGS;          // orbx_undistort_grid has to run first, for THIS extraction (mvKeysUn, mGrid)
    if (use_stereo && (!h->d_st_u || h->st_gen != h->gen)) return ORBX_ERR_BAD_ARGS;   // mvuRight comes from orbx_stereo_match on this handle, same extraction
    const OrbxPlan& P = h->plan;
    const size_t kpf = (size_t)P.kept_per_frame;
    if (kpf > 65535) return ORBX_ERR_BAD_ARGS;
    int cap = 1;
    for (int i = 0; i < nq; ++i) {
        if (q[i].cur_frame < 0 || q[i].cur_frame >= h->last_n || q[i].n < 0) return ORBX_ERR_BAD_ARGS;
        if (q[i].n && (!q[i].w3 || !q[i].desc || (!q[i].obs && mode != 2) || !q[i].oct || !q[i].ang)) return ORBX_ERR_BAD_ARGS;
        if (mode == 2 && q[i].n && (!q[i].flag || !q[i].Tc)) return ORBX_ERR_BAD_ARGS;
        if (q[i].n > cap) cap = q[i].n;
        // A level is validated only for the points the kernels will process: the reference leaves mnTrackScaleLevel
        // uninitialised until Frame::isInFrustum accepts a point (src/Frame.cc:321, src/MapPoint.cc:32-73), so a skipped
        // point may legitimately carry garbage there; it is staged as level 0 below.
        for (int k = 0; k < q[i].n; ++k) {
            const bool processed = mode == 1 ? !(q[i].flag && !q[i].flag[k])
                                 : mode == 2 ? (q[i].flag && q[i].flag[k]) : (!(q[i].flag && q[i].flag[k]) && q[i].obs[k] >= 0);
            if (processed && (q[i].oct[k] < 0 || q[i].oct[k] >= P.nlevels)) return ORBX_ERR_BAD_ARGS;
        }
    }
    cap = (cap + 3) & ~3;
    if ((size_t)(cap + P.kept_per_frame) * sizeof(int) > 200 * 1024) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    const size_t qbytes = (orbx::search_projection_query_bytes() + 15) & ~(size_t)15;
    int list_cap = 32;                                               // entries a point's candidate list may hold (test knob: small
    if (const char* e = getenv("ORBX_SP_LIST_CAP")) {                // values force the full-search path of the resolve kernel)
        const int v = atoi(e);
        if (v >= 1 && v <= 32) list_cap = v;
    }
    // staging layout: [queries][mp_desc nq*cap*32][w3 nq*cap*3 f32][obs nq*cap][octave nq*cap][angle nq*cap][cur_obs nq*kpf];
    // device only, behind it: [candidate counts nq*cap][candidate lists nq*cap*32 u32]
    const size_t o_desc = (size_t)nq * qbytes, o_world = o_desc + (size_t)nq * cap * 32, o_obs = o_world + (size_t)nq * cap * 12,
                 o_oct = o_obs + (size_t)nq * cap * 4, o_ang = o_oct + (size_t)nq * cap * 4, o_cur = o_ang + (size_t)nq * cap * 4,
                 staged = o_cur + (held ? (size_t)nq * kpf * 4 : 0), o_cnt = (staged + 15) & ~(size_t)15,
                 o_list = o_cnt + (size_t)nq * cap * 4, total = o_list + (size_t)nq * cap * 32 * 4;
    cudaStream_t st = h->stream;
    if (held)                                                        // F.N of the frames (cur_obs is sized by it), whatever was fetched so far
        CK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * h->counters_count(), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));                                // the staging of a previous call is free
    if (total > h->sp_bytes) {
        cudaFree(h->d_sp); cudaFreeHost(h->h_sp); h->d_sp = h->h_sp = 0; h->sp_bytes = 0;
        CK(h, cudaMalloc(&h->d_sp, total));
        CK(h, cudaMallocHost(&h->h_sp, total));
        h->sp_bytes = total;
    }
    const size_t out_ints = (size_t)h->cfg.max_batch * (kpf + 2);
    if (!h->d_sp_out) {
        CK(h, cudaMalloc(&h->d_sp_out, out_ints * 4));
        CK(h, cudaMallocHost(&h->h_sp_out, out_ints * 4));
        h->sp_out_ints = out_ints;
    }
    for (int i = 0; i < nq; ++i) {
        const SpHostQuery& Q = q[i];
        int fwd = 0, bwd = 0;
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
        if (mode == 2) {                                             // only the current pose is needed (:1476-1477)
            const float* Tc = Q.Tc;
            const float Rc[9] = {Tc[0], Tc[1], Tc[2], Tc[4], Tc[5], Tc[6], Tc[8], Tc[9], Tc[10]};
            memcpy(R, Rc, sizeof R);
            t[0] = Tc[3]; t[1] = Tc[7]; t[2] = Tc[11];
        } else if (!local) {
            // twc = -Rcw.t() * tcw (general gemm path: double products and sum), tlc = Rlw * twc + tlw (the 3x3 path: float
            // products and sums, "+ C" in double), bForward / bBackward (:1337-1349)
            const float* Tc = Q.Tc;
            const float* Tl = Q.Tl;
            float twc[3];
            for (int r = 0; r < 3; ++r) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += (double)Tc[4 * k + r] * (double)Tc[4 * k + 3];
                twc[r] = (float)(s * -1.0);
            }
            volatile float p0 = Tl[8] * twc[0], p1 = Tl[9] * twc[1], p2 = Tl[10] * twc[2];     // volatile: no contraction
            volatile float t0 = p0 + p1;
            t0 = t0 + p2;
            const float tlcz = (float)((double)t0 + (double)Tl[11]);
            fwd = tlcz > mb && !mono;
            bwd = -tlcz > mb && !mono;
            const float Rc[9] = {Tc[0], Tc[1], Tc[2], Tc[4], Tc[5], Tc[6], Tc[8], Tc[9], Tc[10]};
            memcpy(R, Rc, sizeof R);
            t[0] = Tc[3]; t[1] = Tc[7]; t[2] = Tc[11];
        }
        orbx::search_projection_fill_query(h->h_sp + (size_t)i * qbytes, R, t, Q.n, Q.cur_frame, fwd, bwd);
        const size_t n = (size_t)Q.n, b = (size_t)i * cap;
        if (n) {
            memcpy(h->h_sp + o_desc + b * 32, Q.desc, n * 32);
            memcpy(h->h_sp + o_world + b * 12, Q.w3, n * 12);
            int* obs = reinterpret_cast<int*>(h->h_sp + o_obs) + b;
            if (local)                                               // mbTrackInView && !isBad (:54-58)
                for (size_t k = 0; k < n; ++k) obs[k] = (Q.flag && !Q.flag[k]) ? -1 : (Q.obs[k] < 0 ? 0 : Q.obs[k]);
            else if (mode == 2)                                      // a good, not yet found point in range (:1491-1495, :1516-1518); every match claims (:1540)
                for (size_t k = 0; k < n; ++k) obs[k] = Q.flag[k] ? 1 : -1;
            else                                                     // pMP && !mvbOutlier (:1356-1360)
                for (size_t k = 0; k < n; ++k) obs[k] = (Q.flag && Q.flag[k]) ? -1 : Q.obs[k];
            int* oct = reinterpret_cast<int*>(h->h_sp + o_oct) + b;
            for (size_t k = 0; k < n; ++k) oct[k] = (obs[k] < 0 || Q.oct[k] < 0 || Q.oct[k] >= P.nlevels) ? 0 : Q.oct[k];
            memcpy(h->h_sp + o_ang + b * 4, Q.ang, n * 4);
        }
        if (held) {
            int* co = reinterpret_cast<int*>(h->h_sp + o_cur) + (size_t)i * kpf;
            if (Q.cur_obs) {
                int N = 0;                                           // CurrentFrame.N as of the last fetch of the counters
                for (int l = 0; l < P.nlevels; ++l) N += h->h_kept_counts()[Q.cur_frame * P.nlevels + l];
                memcpy(co, Q.cur_obs, (size_t)N * 4);
                for (size_t k = (size_t)N; k < kpf; ++k) co[k] = -1;
            } else {
                for (size_t k = 0; k < kpf; ++k) co[k] = -1;
            }
        }
    }
    const float K1[4] = {1.f, 1.f, 0.f, 0.f};
    CK(h, cudaMemcpyAsync(h->d_sp, h->h_sp, staged, cudaMemcpyHostToDevice, st));
    CK(h, orbx::launch_search_projection(h->d_plan, P, mode, nq, h->d_sp, K4 ? K4 : K1, h->un_bounds, mbf, th, nnratio, check_orientation,
                                         match_th, cap, list_cap, reinterpret_cast<const float*>(h->d_sp + o_world), h->d_sp + o_desc,
                                         reinterpret_cast<const int*>(h->d_sp + o_obs), reinterpret_cast<const int*>(h->d_sp + o_oct),
                                         reinterpret_cast<const float*>(h->d_sp + o_ang), h->d_out_kp, h->d_out_desc, h->d_kept_counts(),
                                         h->d_un_xy, h->d_un_start, h->d_un_items, use_stereo ? h->d_st_u : nullptr,
                                         held ? reinterpret_cast<const int*>(h->d_sp + o_cur) : nullptr,
                                         reinterpret_cast<uint32_t*>(h->d_sp + o_list), reinterpret_cast<int*>(h->d_sp + o_cnt), h->d_sp_out,
                                         h->d_sp_out + (size_t)h->cfg.max_batch * P.kept_per_frame, st));
    h->launches += 2;
    return ORBX_OK;
}

int sp_fetch(orbx_handle* h, int nq, const int* frames, orbx_projection_result* results) {
    if (!h || !results || nq < 1 || nq > h->cfg.max_batch || !h->d_sp_out) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    const OrbxPlan& P = h->plan;
    const size_t kpf = (size_t)P.kept_per_frame, B = (size_t)h->cfg.max_batch;
    cudaStream_t st = h->stream;
    CK(h, cudaMemcpyAsync(h->h_sp_out, h->d_sp_out, (size_t)nq * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_sp_out + B * kpf, h->d_sp_out + B * kpf, (size_t)nq * 2 * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * h->counters_count(), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int i = 0; i < nq; ++i) {
        const int f = frames[i];
        if (f < 0 || f >= h->last_n) return ORBX_ERR_BAD_ARGS;
        int total = 0;
        for (int l = 0; l < P.nlevels; ++l) total += h->h_kept_counts()[f * P.nlevels + l];
        results[i].n = total;
        results[i].nmatches = h->h_sp_out[B * kpf + 2 * i];
        results[i].rounds = h->h_sp_out[B * kpf + 2 * i + 1];
        results[i].match = h->h_sp_out + (size_t)i * kpf;
    }
    return ORBX_OK;
}

int sp_convert(int nq, const orbx_projection_query* q, std::vector<SpHostQuery>& out) {
    if (!q || nq < 1) return ORBX_ERR_BAD_ARGS;
    out.resize(nq);
    for (int i = 0; i < nq; ++i) {
        SpHostQuery& o = out[i];
        o.cur_frame = q[i].cur_frame; o.n = q[i].n_last; o.w3 = q[i].world_pos; o.desc = q[i].mp_desc; o.obs = q[i].mp_obs;
        o.flag = q[i].outlier; o.oct = q[i].octave; o.ang = q[i].angle; o.cur_obs = nullptr; o.Tc = q[i].Tcw_cur; o.Tl = q[i].Tcw_last;
    }
    return ORBX_OK;
}

int lp_convert(int nq, const orbx_local_points_query* q, std::vector<SpHostQuery>& out) {
    if (!q || nq < 1) return ORBX_ERR_BAD_ARGS;
    out.resize(nq);
    for (int i = 0; i < nq; ++i) {
        SpHostQuery& o = out[i];
        o.cur_frame = q[i].cur_frame; o.n = q[i].n_points; o.w3 = q[i].proj_xy_xr; o.desc = q[i].mp_desc; o.obs = q[i].mp_obs;
        o.flag = q[i].in_view; o.oct = q[i].scale_level; o.ang = q[i].view_cos; o.cur_obs = q[i].cur_obs; o.Tc = o.Tl = nullptr;
    }
    return ORBX_OK;
}
}  // namespace

int orbx_search_by_projection_device(orbx_handle* h, int nqueries, const orbx_projection_query* queries, const float* K4, float mbf,
                                     float mb, float th, int mono, int check_orientation, int use_stereo) {
    if (!K4) return ORBX_ERR_BAD_ARGS;
    std::vector<SpHostQuery> q;
    const int rc = sp_convert(nqueries, queries, q);
    if (rc != ORBX_OK) return rc;
    return sp_enqueue(h, 0, nqueries, q.data(), K4, mbf, mb, th, 0.f, mono, check_orientation, use_stereo);
}

int orbx_search_by_projection_fetch(orbx_handle* h, int nqueries, const orbx_projection_query* queries, orbx_projection_result* results) {
    if (!queries || nqueries < 1) return ORBX_ERR_BAD_ARGS;
    std::vector<int> frames(nqueries);
    for (int i = 0; i < nqueries; ++i) frames[i] = queries[i].cur_frame;
    return sp_fetch(h, nqueries, frames.data(), results);
}

int orbx_search_by_projection(orbx_handle* h, int nqueries, const orbx_projection_query* queries, const float* K4, float mbf, float mb,
                              float th, int mono, int check_orientation, int use_stereo, orbx_projection_result* results) {
    if (!results) return ORBX_ERR_BAD_ARGS;
    const int rc = orbx_search_by_projection_device(h, nqueries, queries, K4, mbf, mb, th, mono, check_orientation, use_stereo);
    if (rc != ORBX_OK) return rc;
    return orbx_search_by_projection_fetch(h, nqueries, queries, results);
}

int orbx_search_local_points_device(orbx_handle* h, int nqueries, const orbx_local_points_query* queries, float th, float nnratio,
                                    int use_stereo) {
    std::vector<SpHostQuery> q;
    const int rc = lp_convert(nqueries, queries, q);
    if (rc != ORBX_OK) return rc;
    return sp_enqueue(h, 1, nqueries, q.data(), nullptr, 0.f, 0.f, th, nnratio, 1, 0, use_stereo);
}

int orbx_search_local_points(orbx_handle* h, int nqueries, const orbx_local_points_query* queries, float th, float nnratio,
                             int use_stereo, orbx_projection_result* results) {
    if (!results) return ORBX_ERR_BAD_ARGS;
    const int rc = orbx_search_local_points_device(h, nqueries, queries, th, nnratio, use_stereo);
    if (rc != ORBX_OK) return rc;
    std::vector<int> frames(nqueries);
    for (int i = 0; i < nqueries; ++i) frames[i] = queries[i].cur_frame;
    return sp_fetch(h, nqueries, frames.data(), results);
}

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1472-1599 (Tracking::Relocalization)
int orbx_search_by_projection_kf_device(orbx_handle* h, int nqueries, const orbx_keyframe_projection_query* queries, const float* K4,
                                        float th, int orb_dist, int check_orientation) {
    if (!queries || nqueries < 1 || !K4 || orb_dist < 0 || orb_dist > 255) return ORBX_ERR_BAD_ARGS;     // bestDist starts at 256 (:1532)
    std::vector<SpHostQuery> q(nqueries);
    for (int i = 0; i < nqueries; ++i) {
        SpHostQuery& o = q[i];
        o.cur_frame = queries[i].cur_frame; o.n = queries[i].n_points; o.w3 = queries[i].world_pos; o.desc = queries[i].mp_desc;
        o.obs = nullptr; o.flag = queries[i].search; o.oct = queries[i].pred_level; o.ang = queries[i].kf_angle;
        o.cur_obs = queries[i].cur_held; o.Tc = queries[i].Tcw_cur; o.Tl = nullptr;
    }
    return sp_enqueue(h, 2, nqueries, q.data(), K4, 0.f, 0.f, th, 0.f, 1, check_orientation, 0, orb_dist);
}

int orbx_search_by_projection_kf(orbx_handle* h, int nqueries, const orbx_keyframe_projection_query* queries, const float* K4, float th,
                                 int orb_dist, int check_orientation, orbx_projection_result* results) {
    if (!results) return ORBX_ERR_BAD_ARGS;
    const int rc = orbx_search_by_projection_kf_device(h, nqueries, queries, K4, th, orb_dist, check_orientation);
    if (rc != ORBX_OK) return rc;
    std::vector<int> frames(nqueries);
    for (int i = 0; i < nqueries; ++i) frames[i] = queries[i].cur_frame;
    return sp_fetch(h, nqueries, frames.data(), results);
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), src/ORBmatcher.cc:405-520
// (Tracking::MonocularInitialization, src/Tracking.cc:600)
int orbx_search_for_initialization(orbx_handle* h, int nq, const orbx_initialization_query* q, float nnratio, int check_orientation,
                                   int window, orbx_initialization_result* results) {
    if (!h || !q || !results || nq < 1 || nq > h->cfg.max_batch || !h->have_plan || window < 0) return ORBX_ERR_BAD_ARGS;
    if (!h->d_un_xy || h->un_gen != h->gen) return ORBX_ERR_BAD_ARGS;          // orbx_undistort_grid has to run first, for THIS extraction (F2's mvKeysUn, mGrid)
    const OrbxPlan& P = h->plan;
    const size_t kpf = (size_t)P.kept_per_frame;
    if (kpf > 65535) return ORBX_ERR_BAD_ARGS;
    int cap = 1;
    for (int i = 0; i < nq; ++i) {
        if (q[i].cur_frame < 0 || q[i].cur_frame >= h->last_n || q[i].n1 < 0) return ORBX_ERR_BAD_ARGS;
        if (q[i].n1 && (!q[i].octave1 || !q[i].angle1 || !q[i].desc1 || !q[i].prev_matched)) return ORBX_ERR_BAD_ARGS;
        if (q[i].n1 > cap) cap = q[i].n1;
    }
    cap = (cap + 3) & ~3;
    if (kpf * 2 * sizeof(int) > 200 * 1024) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    int lc = 96;                                                     // entries of a keypoint's candidate list (test knob: small values
    if (const char* e = getenv("ORBX_INIT_LIST_CAP")) {              // force the window re-scan of the resolve kernel)
        const int v = atoi(e);
        if (v >= 1 && v <= 4096) lc = v;
    }
    // layout: [queries nq*8][desc1 nq*cap*32][prev nq*cap*8][octave nq*cap*4][angle nq*cap*4]            (staged, host -> device)
    //         [counts nq*cap*4][lists nq*cap*lc*4][bins nq*cap]                                          (device only)
    //         [matches12 nq*cap*4][prev out nq*cap*8][nmatches nq*4]                                     (device -> host)
    const size_t n = (size_t)nq * cap;
    const size_t o_desc = ((size_t)nq * 8 + 15) & ~(size_t)15, o_prev = o_desc + n * 32, o_oct = o_prev + n * 8, o_ang = o_oct + n * 4,
                 staged = o_ang + n * 4, o_cnt = (staged + 15) & ~(size_t)15, o_list = o_cnt + n * 4, o_bin = o_list + n * lc * 4,
                 o_m12 = (o_bin + n + 15) & ~(size_t)15, o_pout = o_m12 + n * 4, o_nm = o_pout + n * 8, total = o_nm + (size_t)nq * 4;
    cudaStream_t st = h->stream;
    CK(h, cudaStreamSynchronize(st));                                // the staging of a previous call is free
    if (total > h->sp_bytes) {
        cudaFree(h->d_sp); cudaFreeHost(h->h_sp); h->d_sp = h->h_sp = 0; h->sp_bytes = 0;
        CK(h, cudaMalloc(&h->d_sp, total));
        CK(h, cudaMallocHost(&h->h_sp, total));
        h->sp_bytes = total;
    }
    for (int i = 0; i < nq; ++i) {
        int* qi = reinterpret_cast<int*>(h->h_sp) + 2 * i;
        qi[0] = q[i].n1;
        qi[1] = q[i].cur_frame;
        const size_t m = (size_t)q[i].n1, b = (size_t)i * cap;
        if (!m) continue;
        memcpy(h->h_sp + o_desc + b * 32, q[i].desc1, m * 32);
        memcpy(h->h_sp + o_prev + b * 8, q[i].prev_matched, m * 8);
        memcpy(h->h_sp + o_oct + b * 4, q[i].octave1, m * 4);
        memcpy(h->h_sp + o_ang + b * 4, q[i].angle1, m * 4);
    }
    CK(h, cudaMemcpyAsync(h->d_sp, h->h_sp, staged, cudaMemcpyHostToDevice, st));
    CK(h, orbx::launch_search_init(h->d_plan, P, nq, h->d_sp, h->un_bounds, (float)window, nnratio, check_orientation, cap, lc,
                                   reinterpret_cast<const float*>(h->d_sp + o_prev), h->d_sp + o_desc,
                                   reinterpret_cast<const int*>(h->d_sp + o_oct), reinterpret_cast<const float*>(h->d_sp + o_ang),
                                   h->d_out_kp, h->d_out_desc, h->d_kept_counts(), h->d_un_xy, h->d_un_start, h->d_un_items,
                                   reinterpret_cast<uint32_t*>(h->d_sp + o_list), reinterpret_cast<int*>(h->d_sp + o_cnt), h->d_sp + o_bin,
                                   reinterpret_cast<int*>(h->d_sp + o_m12), reinterpret_cast<float*>(h->d_sp + o_pout),
                                   reinterpret_cast<int*>(h->d_sp + o_nm), st));
    h->launches += 2;
    CK(h, cudaMemcpyAsync(h->h_sp + o_m12, h->d_sp + o_m12, total - o_m12, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int i = 0; i < nq; ++i) {
        results[i].n1 = q[i].n1;
        results[i].nmatches = reinterpret_cast<const int*>(h->h_sp + o_nm)[i];
        results[i].matches12 = reinterpret_cast<const int32_t*>(h->h_sp + o_m12) + (size_t)i * cap;
        results[i].prev_matched = reinterpret_cast<const float*>(h->h_sp + o_pout) + (size_t)i * cap * 2;
    }
    return ORBX_OK;
}

// ---- Frame::ComputeBoW: the vocabulary lives in HBM, independent of any handle.
struct orbx_vocabulary {
    int device, n_nodes, L;
    int *d_child_start, *d_child_items, *d_word;
    uint8_t* d_desc;
    double* d_weight;
};

// Frame::isInFrustum (src/Frame.cc:269-325) + MapPoint::PredictScale (src/MapPoint.cc:402-417) over lists of map points.
namespace {
// MapPoint::PredictScale's level for a ratio, evaluated with THIS process's libm -- the one the reference would call
int frustum_level_host(float ratio, float lsf, int nlevels) {
    volatile float l = logf(ratio);                    // `log(ratio)` with a float argument under `using namespace std` (:410)
    volatile float qv = l / lsf;
    const float c = ceilf(qv);
    if (!(c == c) || c > 3.0e38f || c < -3.0e38f) return 0;      // NaN / inf -> int is undefined in the reference (x86: INT_MIN -> 0)
    return c < 0.f ? 0 : (c >= (float)nlevels ? nlevels - 1 : (int)c);
}
// T[k] = the smallest positive finite float whose level reaches k (bisection over the bit pattern; levels are monotone in the ratio)
void frustum_thresholds(float lsf, int nlevels, float* T) {
    T[0] = 0.f;
    for (int k = 1; k < nlevels; ++k) {
        uint32_t lo = 1u, hi = 0x7f7fffffu;            // level(lo) < k <= level(hi) is the invariant
        float f;
        memcpy(&f, &hi, 4);
        if (frustum_level_host(f, lsf, nlevels) < k) { const uint32_t inf = 0x7f800000u; memcpy(&T[k], &inf, 4); continue; }
        memcpy(&f, &lo, 4);
        if (frustum_level_host(f, lsf, nlevels) >= k) { memcpy(&T[k], &lo, 4); continue; }
        while (hi - lo > 1u) {
            const uint32_t mid = lo + (hi - lo) / 2u;
            memcpy(&f, &mid, 4);
            if (frustum_level_host(f, lsf, nlevels) >= k) hi = mid; else lo = mid;
        }
        memcpy(&T[k], &hi, 4);
    }
}
}  // namespace

int orbx_is_in_frustum(orbx_handle* h, int nq, const orbx_frustum_query* q, const float* K4, float mbf, const float* bounds,
                       float log_scale_factor, float viewing_cos_limit, orbx_frustum_result* results) {
    if (!h || !q || !K4 || !bounds || !results || nq < 1 || nq > 65535 || !(log_scale_factor > 0.f) || !(log_scale_factor < 3.0e38f))
        return ORBX_ERR_BAD_ARGS;
    size_t N = 0;
    int max_n = 0;
    for (int i = 0; i < nq; ++i) {
        if (q[i].n_points < 0) return ORBX_ERR_BAD_ARGS;
        if (q[i].n_points && (!q[i].world_pos || !q[i].normal || !q[i].min_dist || !q[i].max_dist)) return ORBX_ERR_BAD_ARGS;
        N += (size_t)q[i].n_points;
        if (q[i].n_points > max_n) max_n = q[i].n_points;
    }
    if (N > ((size_t)1 << 30)) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    const size_t qb = (orbx::frustum_query_bytes() + 15) & ~(size_t)15, Np = (N + 15) & ~(size_t)15;
    // [queries][world N x 3][normal N x 3][min N][max N][consider N] | [proj N x 3][level N][view cos N][in view N]
    const size_t o_world = (size_t)nq * qb, o_norm = o_world + Np * 12, o_min = o_norm + Np * 12, o_max = o_min + Np * 4,
                 o_cons = o_max + Np * 4, o_proj = o_cons + Np, o_lvl = o_proj + Np * 12, o_vc = o_lvl + Np * 4, o_in = o_vc + Np * 4,
                 total = o_in + Np;
    cudaStream_t st = h->stream;
    CK(h, cudaStreamSynchronize(st));                                // the staging of a previous call is free
    if (total > h->fr_bytes) {
        cudaFree(h->d_fr); cudaFreeHost(h->h_fr); h->d_fr = h->h_fr = 0; h->fr_bytes = 0;
        CK(h, cudaMalloc(&h->d_fr, total));
        CK(h, cudaMallocHost(&h->h_fr, total));
        h->fr_bytes = total;
    }
    size_t off = 0;
    for (int i = 0; i < nq; ++i) {
        const float* T = q[i].Tcw;
        const float R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]}, t[3] = {T[3], T[7], T[11]};
        float Ow[3];                                                 // mOw = -mRcw.t() * mtcw (:266): general gemm path, double products and sum
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)R[3 * k + r] * (double)t[k];
            Ow[r] = (float)(s * -1.0);
        }
        const size_t n = (size_t)q[i].n_points;
        orbx::frustum_fill_query(h->h_fr + (size_t)i * qb, R, t, Ow, (int)n, (int)off);
        if (n) {
            memcpy(h->h_fr + o_world + off * 12, q[i].world_pos, n * 12);
            memcpy(h->h_fr + o_norm + off * 12, q[i].normal, n * 12);
            memcpy(h->h_fr + o_min + off * 4, q[i].min_dist, n * 4);
            memcpy(h->h_fr + o_max + off * 4, q[i].max_dist, n * 4);
            if (q[i].consider) memcpy(h->h_fr + o_cons + off, q[i].consider, n);
            else memset(h->h_fr + o_cons + off, 1, n);
        }
        off += n;
    }
    float thr[ORBX_MAXL];
    frustum_thresholds(log_scale_factor, h->cfg.nlevels, thr);
    if (N) {
        CK(h, cudaMemcpyAsync(h->d_fr, h->h_fr, o_proj, cudaMemcpyHostToDevice, st));
        CK(h, orbx::launch_frustum(nq, max_n, h->d_fr, K4, bounds, mbf, viewing_cos_limit, h->cfg.nlevels, thr, h->d_fr + o_cons,
                                   reinterpret_cast<const float*>(h->d_fr + o_world), reinterpret_cast<const float*>(h->d_fr + o_norm),
                                   reinterpret_cast<const float*>(h->d_fr + o_min), reinterpret_cast<const float*>(h->d_fr + o_max),
                                   h->d_fr + o_in, reinterpret_cast<float*>(h->d_fr + o_proj), reinterpret_cast<int*>(h->d_fr + o_lvl),
                                   reinterpret_cast<float*>(h->d_fr + o_vc), st));
        ++h->launches;
        CK(h, cudaMemcpyAsync(h->h_fr + o_proj, h->d_fr + o_proj, total - o_proj, cudaMemcpyDeviceToHost, st));
        CK(h, cudaStreamSynchronize(st));
    }
    off = 0;
    for (int i = 0; i < nq; ++i) {
        const size_t n = (size_t)q[i].n_points;
        results[i].n = (int)n;
        results[i].in_view = h->h_fr + o_in + off;
        results[i].proj_xy_xr = reinterpret_cast<const float*>(h->h_fr + o_proj) + off * 3;
        results[i].scale_level = reinterpret_cast<const int32_t*>(h->h_fr + o_lvl) + off;
        results[i].view_cos = reinterpret_cast<const float*>(h->h_fr + o_vc) + off;
        int c = 0;
        for (size_t k = 0; k < n; ++k) c += results[i].in_view[k];
        results[i].n_in_view = c;
        off += n;
    }
    return ORBX_OK;
}

int orbx_vocabulary_create(int device, int n_nodes, int L, const int32_t* child_start, const int32_t* child_items,
                           const uint8_t* node_desc, const double* node_weight, const int32_t* node_word, orbx_vocabulary** out) {
    if (!out || !child_start || !child_items || !node_desc || !node_weight || !node_word || n_nodes < 2 || L < 1 || L > 32)
        return ORBX_ERR_BAD_ARGS;
    // a tree in DBoW2's sense: every node but the root is the child of exactly one node, the root has children
    if (child_start[0] != 0 || child_start[n_nodes] != n_nodes - 1 || child_start[1] == 0) return ORBX_ERR_BAD_ARGS;
    std::vector<unsigned char> seen((size_t)n_nodes, 0);
    for (int i = 0; i < n_nodes; ++i) {
        if (child_start[i + 1] < child_start[i]) return ORBX_ERR_BAD_ARGS;
        for (int c = child_start[i]; c < child_start[i + 1]; ++c) {
            const int id = child_items[c];
            if (id < 1 || id >= n_nodes || seen[id]) return ORBX_ERR_BAD_ARGS;
            seen[id] = 1;
        }
        if (child_start[i + 1] - child_start[i] > (1 << 20)) return ORBX_ERR_BAD_ARGS;
    }
    if (cudaSetDevice(device) != cudaSuccess) return ORBX_ERR_CUDA;
    orbx_vocabulary* v = new (std::nothrow) orbx_vocabulary();
    if (!v) return ORBX_ERR_BAD_ARGS;
    v->device = device; v->n_nodes = n_nodes; v->L = L;
    v->d_child_start = v->d_child_items = v->d_word = 0; v->d_desc = 0; v->d_weight = 0;
    const size_t n = (size_t)n_nodes;
    cudaError_t e = cudaMalloc(&v->d_child_start, (n + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_child_items, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_word, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_desc, n * 32);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_weight, n * 8);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_start, child_start, (n + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_items, child_items, (n - 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_word, node_word, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, node_desc, n * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_weight, node_weight, n * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { orbx_vocabulary_destroy(v); return ORBX_ERR_CUDA; }
    *out = v;
    return ORBX_OK;
}

int orbx_vocabulary_destroy(orbx_vocabulary* v) {
    if (!v) return ORBX_OK;
    cudaSetDevice(v->device);
    cudaFree(v->d_child_start); cudaFree(v->d_child_items); cudaFree(v->d_word); cudaFree(v->d_desc); cudaFree(v->d_weight);
    delete v;
    return ORBX_OK;
}

static int bow_enqueue(orbx_handle* h, const orbx_vocabulary* voc, int nframes, const int* frames, int levelsup) {
    if (!h || !voc || nframes < 1 || nframes > h->cfg.max_batch || !h->have_plan || voc->device != h->cfg.device || levelsup < 0)
        return ORBX_ERR_BAD_ARGS;
    for (int i = 0; i < nframes; ++i) {
        const int f = frames ? frames[i] : i;
        if (f < 0 || f >= h->last_n) return ORBX_ERR_BAD_ARGS;
    }
    CK(h, cudaSetDevice(h->cfg.device));
    const OrbxPlan& P = h->plan;
    const size_t B = (size_t)h->cfg.max_batch, kpf = (size_t)P.kept_per_frame;
    if (kpf > 8192) return ORBX_ERR_BAD_ARGS;
    const size_t words = 5 * B * kpf + 3 * B;
    if (!h->d_bow) {
        CK(h, cudaMalloc(&h->d_bow, words * 4));
        CK(h, cudaMallocHost(&h->h_bow, words * 4));
        CK(h, cudaMalloc(&h->d_bow_val, B * kpf * 8));
        CK(h, cudaMallocHost(&h->h_bow_val, B * kpf * 8));
    }
    cudaStream_t st = h->stream;
    CK(h, cudaStreamSynchronize(st));
    unsigned* hf = h->h_bow + 5 * B * kpf + 2 * B;
    for (int i = 0; i < nframes; ++i) hf[i] = (unsigned)(frames ? frames[i] : i);
    h->bow_slot.assign(B, -1);
    h->bow_gen = h->gen;
    for (int i = 0; i < nframes; ++i) h->bow_slot[hf[i]] = i;
    unsigned* df = h->d_bow + 5 * B * kpf + 2 * B;
    CK(h, cudaMemcpyAsync(df, hf, (size_t)nframes * 4, cudaMemcpyHostToDevice, st));
    CK(h, orbx::launch_compute_bow(h->d_plan, P, voc->d_child_start, voc->d_child_items, voc->d_desc, voc->d_weight, voc->d_word,
                                   voc->n_nodes, voc->L, reinterpret_cast<const int*>(df), nframes, levelsup, h->d_out_desc,
                                   h->d_kept_counts(), reinterpret_cast<int*>(h->d_bow), reinterpret_cast<int*>(h->d_bow + B * kpf),
                                   h->d_bow + 2 * B * kpf, h->d_bow_val, h->d_bow + 3 * B * kpf, h->d_bow + 4 * B * kpf,
                                   reinterpret_cast<int*>(h->d_bow + 5 * B * kpf), st));
    h->launches += 2;
    return ORBX_OK;
}

int orbx_compute_bow_device(orbx_handle* h, const orbx_vocabulary* voc, int nframes, const int* frames, int levelsup) {
    return bow_enqueue(h, voc, nframes, frames, levelsup);
}

int orbx_compute_bow(orbx_handle* h, const orbx_vocabulary* voc, int nframes, const int* frames, int levelsup, orbx_bow_result* results) {
    if (!results) return ORBX_ERR_BAD_ARGS;
    const int rc = bow_enqueue(h, voc, nframes, frames, levelsup);
    if (rc != ORBX_OK) return rc;
    const OrbxPlan& P = h->plan;
    const size_t B = (size_t)h->cfg.max_batch, kpf = (size_t)P.kept_per_frame, n = (size_t)nframes;
    cudaStream_t st = h->stream;
    CK(h, cudaMemcpyAsync(h->h_bow + 2 * B * kpf, h->d_bow + 2 * B * kpf, n * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_bow + 3 * B * kpf, h->d_bow + 3 * B * kpf, n * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_bow + 4 * B * kpf, h->d_bow + 4 * B * kpf, n * kpf * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_bow + 5 * B * kpf, h->d_bow + 5 * B * kpf, n * 2 * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(h->h_bow_val, h->d_bow_val, n * kpf * 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int i = 0; i < nframes; ++i) {
        const int* cnt = reinterpret_cast<const int*>(h->h_bow + 5 * B * kpf) + 2 * i;
        results[i].n_words = cnt[0];
        results[i].word_ids = h->h_bow + 2 * B * kpf + (size_t)i * kpf;
        results[i].word_values = h->h_bow_val + (size_t)i * kpf;
        results[i].n_features = cnt[1];
        results[i].fv_nodes = h->h_bow + 3 * B * kpf + (size_t)i * kpf;
        results[i].fv_features = h->h_bow + 4 * B * kpf + (size_t)i * kpf;
    }
    return ORBX_OK;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), src/ORBmatcher.cc:159-288
int orbx_search_by_bow_device(orbx_handle* h, int nq, const orbx_bow_match_query* q, float nnratio, int check_orientation) {
    if (!h || !q || nq < 1 || nq > h->cfg.max_batch || !h->have_plan || !h->d_bow) return ORBX_ERR_BAD_ARGS;
    const OrbxPlan& P = h->plan;
    const size_t kpf = (size_t)P.kept_per_frame, B = (size_t)h->cfg.max_batch;
    if (kpf > 65535) return ORBX_ERR_BAD_ARGS;
    int cap = 4;
    for (int i = 0; i < nq; ++i) {
        const orbx_bow_match_query& Q = q[i];
        if (h->bow_gen != h->gen || Q.cur_frame < 0 || Q.cur_frame >= h->last_n || (size_t)Q.cur_frame >= h->bow_slot.size() ||
            h->bow_slot[Q.cur_frame] < 0)
            return ORBX_ERR_BAD_ARGS;                               // orbx_compute_bow has to run for the frame first (F.mFeatVec)
        if (Q.n_kf < 0 || Q.n_kf_fv < 0 || Q.n_kf_fv > Q.n_kf) return ORBX_ERR_BAD_ARGS;
        if (Q.n_kf && (!Q.kf_desc || !Q.kf_valid || !Q.kf_angle)) return ORBX_ERR_BAD_ARGS;
        if (Q.n_kf_fv && (!Q.kf_fv_nodes || !Q.kf_fv_features)) return ORBX_ERR_BAD_ARGS;
        for (int k = 0; k < Q.n_kf_fv; ++k) {
            if (Q.kf_fv_features[k] >= (unsigned)Q.n_kf) return ORBX_ERR_BAD_ARGS;
            if (k && Q.kf_fv_nodes[k] < Q.kf_fv_nodes[k - 1]) return ORBX_ERR_BAD_ARGS;      // std::map order
        }
        if (Q.n_kf > cap) cap = Q.n_kf;
    }
    cap = (cap + 3) & ~3;
    CK(h, cudaSetDevice(h->cfg.device));
    const size_t qbytes = (orbx::search_bow_query_bytes() + 15) & ~(size_t)15;
    // staging: [queries][kf_desc nq*cap*32][kf_angle nq*cap f32][fv nodes nq*cap][fv features nq*cap][kf_valid nq*cap u8]
    const size_t o_desc = (size_t)nq * qbytes, o_ang = o_desc + (size_t)nq * cap * 32, o_fn = o_ang + (size_t)nq * cap * 4,
                 o_ff = o_fn + (size_t)nq * cap * 4, o_val = o_ff + (size_t)nq * cap * 4, total = o_val + (size_t)nq * cap;
    cudaStream_t st = h->stream;
    CK(h, cudaStreamSynchronize(st));
    if (total > h->sp_bytes) {
        cudaFree(h->d_sp); cudaFreeHost(h->h_sp); h->d_sp = h->h_sp = 0; h->sp_bytes = 0;
        CK(h, cudaMalloc(&h->d_sp, total));
        CK(h, cudaMallocHost(&h->h_sp, total));
        h->sp_bytes = total;
    }
    if (!h->d_sp_out) {
        const size_t out_ints = B * (kpf + 2);
        CK(h, cudaMalloc(&h->d_sp_out, out_ints * 4));
        CK(h, cudaMallocHost(&h->h_sp_out, out_ints * 4));
        h->sp_out_ints = out_ints;
    }
    for (int i = 0; i < nq; ++i) {
        const orbx_bow_match_query& Q = q[i];
        orbx::search_bow_fill_query(h->h_sp + (size_t)i * qbytes, Q.cur_frame, h->bow_slot[Q.cur_frame], Q.n_kf, Q.n_kf_fv);
        const size_t b = (size_t)i * cap;
        if (Q.n_kf) {
            memcpy(h->h_sp + o_desc + b * 32, Q.kf_desc, (size_t)Q.n_kf * 32);
            memcpy(h->h_sp + o_ang + b * 4, Q.kf_angle, (size_t)Q.n_kf * 4);
            memcpy(h->h_sp + o_val + b, Q.kf_valid, (size_t)Q.n_kf);
        }
        if (Q.n_kf_fv) {
            memcpy(h->h_sp + o_fn + b * 4, Q.kf_fv_nodes, (size_t)Q.n_kf_fv * 4);
            memcpy(h->h_sp + o_ff + b * 4, Q.kf_fv_features, (size_t)Q.n_kf_fv * 4);
        }
    }
    CK(h, cudaMemcpyAsync(h->d_sp, h->h_sp, total, cudaMemcpyHostToDevice, st));
    CK(h, orbx::launch_search_bow(h->d_plan, P, nq, h->d_sp, cap, nnratio, check_orientation, h->d_sp + o_desc, h->d_sp + o_val,
                                  reinterpret_cast<const float*>(h->d_sp + o_ang), reinterpret_cast<const unsigned*>(h->d_sp + o_fn),
                                  reinterpret_cast<const unsigned*>(h->d_sp + o_ff), h->d_out_kp, h->d_out_desc, h->d_kept_counts(),
                                  h->d_bow + 3 * B * kpf, h->d_bow + 4 * B * kpf, reinterpret_cast<const int*>(h->d_bow + 5 * B * kpf),
                                  h->d_sp_out, h->d_sp_out + B * kpf, st));
    h->launches += 1;
    return ORBX_OK;
}

int orbx_search_by_bow(orbx_handle* h, int nqueries, const orbx_bow_match_query* queries, float nnratio, int check_orientation,
                       orbx_projection_result* results) {
    if (!results) return ORBX_ERR_BAD_ARGS;
    const int rc = orbx_search_by_bow_device(h, nqueries, queries, nnratio, check_orientation);
    if (rc != ORBX_OK) return rc;
    std::vector<int> frames(nqueries);
    for (int i = 0; i < nqueries; ++i) frames[i] = queries[i].cur_frame;
    return sp_fetch(h, nqueries, frames.data(), results);
}

int orbx_fast_stats(orbx_handle* h, int frame, int* candidates, int* retries) {
    if (!h || !h->have_plan || frame < 0 || frame >= h->last_n) return ORBX_ERR_BAD_ARGS;
    const int L = h->plan.nlevels;
    for (int l = 0; l < L; ++l) {
        if (candidates) candidates[l] = h->h_level_counts()[frame * L + l];
        if (retries) retries[l] = h->h_retry_counts()[frame * L + l];
    }
    return ORBX_OK;
}

int orbx_alloc_host(size_t bytes, void** out) {
    if (!out) return ORBX_ERR_BAD_ARGS;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); *out = 0; return ORBX_ERR_OUT_OF_MEMORY; }
    return ORBX_OK;
}

int orbx_free_host(void* p) {
    if (cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return ORBX_ERR_CUDA; }
    return ORBX_OK;
}

int orbx_pyramid_level(orbx_handle* h, int frame, int level, const uint8_t** image, int* width, int* height, size_t* step) {
    if (!h || !h->have_plan || !h->pyramid_valid || frame < 0 || frame >= h->last_n || level < 0 || level >= h->plan.nlevels)
        return ORBX_ERR_BAD_ARGS;
    const OrbxLevel& L = h->plan.lv[level];
    if (image) *image = h->h_pyr + (size_t)frame * h->plan.slab_bytes + L.plane_off + (size_t)ORBX_EDGE * L.pitch + ORBX_XO;
    if (width) *width = L.w;
    if (height) *height = L.h;
    if (step) *step = (size_t)L.pitch;
    return ORBX_OK;
}

int orbx_scale_tables(orbx_handle* h, const float** scale, const float** inv_scale, const float** sigma2, const float** inv_sigma2) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    if (scale) *scale = h->sf;
    if (inv_scale) *inv_scale = h->inv_sf;
    if (sigma2) *sigma2 = h->sigma2;
    if (inv_sigma2) *inv_sigma2 = h->inv_sigma2;
    return ORBX_OK;
}

int orbx_get_levels(orbx_handle* h) { return h ? h->cfg.nlevels : ORBX_ERR_BAD_ARGS; }
float orbx_get_scale_factor(orbx_handle* h) { return h ? (float)(double)h->cfg.scale_factor : 0.f; }

int orbx_level_quotas(orbx_handle* h, int* quotas, int* umax16) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    if (quotas) for (int i = 0; i < h->cfg.nlevels; ++i) quotas[i] = h->quotas[i];
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = h->umax[i];
    return ORBX_OK;
}

int orbx_level_sizes(orbx_handle* h, int width, int height, int* widths, int* heights) {
    if (!h || !widths || !heights) return ORBX_ERR_BAD_ARGS;
    for (int l = 0; l < h->cfg.nlevels; ++l) level_size(h, width, height, l, &widths[l], &heights[l]);
    return ORBX_OK;
}

int orbx_stage_dump(orbx_handle* h, int frame, int level, int stage, void* out, size_t cap, size_t* bytes) {
    if (!h || !h->have_plan || frame < 0 || frame >= h->last_n || level < 0 || level >= h->plan.nlevels || !bytes)
        return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    CK(h, cudaStreamSynchronize(h->stream));
    const OrbxPlan& P = h->plan;
    const OrbxLevel& L = P.lv[level];
    const int nl = P.nlevels;
    int counts[2];
    CK(h, cudaMemcpy(&counts[0], h->d_sorted_counts() + frame * nl + level, sizeof(int), cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(&counts[1], h->d_kept_counts() + frame * nl + level, sizeof(int), cudaMemcpyDeviceToHost));
    switch (stage) {
        case ORBX_STAGE_PYRAMID: {
            const size_t wp = (size_t)L.w + 2 * ORBX_EDGE;
            *bytes = wp * L.rows;
            if (out && cap >= *bytes)
                CK(h, cudaMemcpy2D(out, wp, h->d_pyr + (size_t)frame * P.slab_bytes + L.plane_off + (ORBX_XO - ORBX_EDGE), L.pitch, wp, L.rows, cudaMemcpyDeviceToHost));
            return ORBX_OK;
        }
        case ORBX_STAGE_BLURRED: {
            *bytes = (size_t)L.w * L.h;
            if (out && cap >= *bytes) {
                // Diagnostics only: the product path blurs just the keypoint patches (describe_kernel); the whole-level
                // blur of the reference (:1085-1086) is produced here on demand with the same arithmetic.
                if (!h->d_blur) {
                    CK(h, cudaMalloc(&h->d_blur, P.slab_bytes));
                    CK(h, cudaMemsetAsync(h->d_blur, 0, P.slab_bytes, h->stream));
                }
                orbx::launch_blur(h->d_plan, P, 1, h->num_sms, h->d_pyr + (size_t)frame * P.slab_bytes, h->d_blur, h->stream);
                CK(h, cudaStreamSynchronize(h->stream));
                CK(h, cudaMemcpy2D(out, L.w, h->d_blur + L.plane_off + (size_t)ORBX_EDGE * L.pitch + ORBX_XO, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
            }
            return ORBX_OK;
        }
        case ORBX_STAGE_CANDIDATES:
        case ORBX_STAGE_KEPT: {
            const bool isc = stage == ORBX_STAGE_CANDIDATES;
            const int n = isc ? counts[0] : counts[1];
            *bytes = (size_t)n * 12;
            if (out && cap >= *bytes && n > 0) {
                std::vector<uint32_t> tmp((size_t)n);
                const uint32_t* src = isc ? h->d_cand_sorted + (size_t)frame * P.cand_per_frame + L.cand_off
                                          : h->d_kept + (size_t)frame * P.kept_per_frame + L.kept_off;
                CK(h, cudaMemcpy(tmp.data(), src, (size_t)n * 4, cudaMemcpyDeviceToHost));
                int32_t* o = (int32_t*)out;
                for (int i = 0; i < n; ++i) { o[3 * i] = ORBX_PX(tmp[i]); o[3 * i + 1] = ORBX_PY(tmp[i]); o[3 * i + 2] = ORBX_PR(tmp[i]); }
            }
            return ORBX_OK;
        }
        case ORBX_STAGE_ANGLES: {
            const int n = counts[1];
            *bytes = (size_t)n * 4;
            if (out && cap >= *bytes && n > 0)
                CK(h, cudaMemcpy(out, h->d_angles + (size_t)frame * P.kept_per_frame + L.kept_off, (size_t)n * 4, cudaMemcpyDeviceToHost));
            return ORBX_OK;
        }
        default: return ORBX_ERR_BAD_ARGS;
    }
}

void* orbx_stream(orbx_handle* h) { return h ? (void*)h->stream : 0; }

int orbx_synchronize(orbx_handle* h) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    CK(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

int orbx_stage_timing_enable(orbx_handle* h, int enable) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    CK(h, cudaSetDevice(h->cfg.device));
    if (enable && !h->ev_created) {
        for (int i = 0; i < kTimingRing; ++i)
            for (int s = 0; s <= ST_COUNT; ++s) CK(h, cudaEventCreate(&h->ev[i][s]));
        h->ev_created = true;
    }
    if (h->ev_pending) timing_collect(h, h->ev_pending);
    h->timing = enable != 0;
    memset(h->stage_ms, 0, sizeof h->stage_ms);
    memset(h->stage_launches, 0, sizeof h->stage_launches);
    return ORBX_OK;
}

int orbx_stage_timing_read(orbx_handle* h, int cap, const char** names, float* ms, int* launches) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    if (h->ev_pending) timing_collect(h, h->ev_pending);
    for (int s = 0; s < ST_COUNT && s < cap; ++s) {
        if (names) names[s] = kStageNames[s];
        if (ms) ms[s] = (float)h->stage_ms[s];
        if (launches) launches[s] = h->stage_launches[s];
    }
    return ST_COUNT;
}

long long orbx_launch_count(orbx_handle* h) { return h ? h->launches : 0; }

// B_alg = sum_l [A_max(l-1,0) + P_l] + 3 * sum_l A_l + nfeatures * 1321   
long long orbx_algorithmic_bytes(orbx_handle* h, int width, int height) {
    if (!h) return ORBX_ERR_BAD_ARGS;
    long long total = 0, prevA = 0;
    for (int l = 0; l < h->cfg.nlevels; ++l) {
        int lw, lh;
        level_size(h, width, height, l, &lw, &lh);
        const long long A = (long long)lw * lh, Pl = (long long)(lw + 38) * (lh + 38);
        total += (l == 0 ? A : prevA) + Pl + 3 * A;
        prevA = A;
    }
    return total + (long long)h->cfg.nfeatures * (749 + 512 + 32 + 28);
}

}  // extern "C"
