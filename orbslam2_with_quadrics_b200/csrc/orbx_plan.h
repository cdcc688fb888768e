// Device-visible plan of one extractor geometry: per-level sizes, HBM layout, cell grid,
// quotas.  Built on the host (orbx_api.cu) from the five ORBextractor constructor
// arguments and the image size, following src/ORBextractor.cc:410-470 (tables),
// :1107-1115 (level sizes) and :771-787 (cell grid) of the reference.
#pragma once
#include <stdint.h>

#define ORBX_XO 32          // plane column of level pixel x = 0 (19 used by the border; 32 keeps rows 32-byte aligned)
#define ORBX_EDGE 19        // EDGE_THRESHOLD, src/ORBextractor.cc:74
#define ORBX_BOX 16         // minBorderX/Y = EDGE_THRESHOLD - 3, :773-774
#define ORBX_MAXL 16
#define ORBX_MAX_ROOTS 256

// Candidate / keypoint record: x | y << 12 | response << 24 (x, y < 4096; response = FAST score <= 254)
#define ORBX_PACK(x, y, r) ((uint32_t)(x) | ((uint32_t)(y) << 12) | ((uint32_t)(r) << 24))
#define ORBX_PX(p) ((int)((p) & 0xfffu))
#define ORBX_PY(p) ((int)(((p) >> 12) & 0xfffu))
#define ORBX_PR(p) ((int)((p) >> 24))

struct OrbxLevel {
    int w, h;                 // level size
    int pitch;                // plane row pitch in bytes (multiple of 64)
    int rows;                 // h + 38
    long long plane_off;      // byte offset of the padded plane inside one frame's pyramid slab
    int maxBX, maxBY;         // w - 16, h - 16 (box [16, maxB) is where FAST runs)
    int nColsV, nRowsV;       // cells that survive the skip rules (:794, :803)
    int wCell, hCell;
    int cell_base;            // first cell of this level inside a frame (cell_rec index)
    int strips_x;             // FAST work items: strips of strip_nc cells per cell row
    int strip_nc;             // cells per strip of this level (a strip's scoring pixels fit 32 tile words)
    int strip_ok;             // 1: fast_strips_kernel's dense strip path applies (cells <= 32 wide, <= 40 tall), 0: cell by cell
    int strip_tall;           // 1: cells of 33 .. 40 rows (fast_strips_kernel<BW, 5>, own table segment)
    int wcell_recip;          // 65536 / wCell + 1: (n * wcell_recip) >> 16 == n / wCell for n < 256
    int cand_off, cand_cap;   // this level's region in a frame's candidate buffers (entries)
    int quota;                // mnFeaturesPerLevel[l]
    int nIni;                 // root nodes of DistributeOctTree (:543)
    float hX;                 // (:545)
    int kept_off, kept_cap;   // this level's slots in a frame's kept-keypoint arrays
    float scale;              // mvScaleFactor[l]
    float kp_size;            // (float)(int)(31 * scale) (:837)
    int xtab_off, ytab_off;   // resize tap tables of this level (entries)
    int blur_tile_base;       // first blur tile of this level inside a frame
    int blur_tiles_x;
    int resize_wide;          // 1: the 4 pixels of a lane span more than two source words (large scale factors)
    int col8_off, ngroups8;   // pyr_resize8_kernel: this level's OrbxCol8 records (offset in OrbxTap units), groups of 8 plane columns
    int yrow_off;             // per PLANE row row-tap entries (border reflection already applied), OrbxTap units
    int resize8_ok;           // 0: a half-group's source window exceeds 8 bytes (scale factor > 2): 4-pixel kernel instead
    int resize_tile_ok;       // 1: a CTA's source footprint fits pyr_resize8_tile_kernel's TMA tile (scale factors up to ~1.2)
};

struct OrbxPlan {
    int nlevels;
    int width, height;
    int cells_per_frame;
    int cand_per_frame;       // entries
    int kept_per_frame;       // slots (= capacity of a frame's output arrays)
    int blur_tiles_per_frame;
    int node_cap;             // octree node capacity (max over levels)
    int max_cell_w, max_cell_h;   // largest FAST window (incl. the 6-px overlap)
    int fast_bw, fast_bh;         // TMA box of a fast_strips_kernel tile (bw: 96 / 128 / 160, bh = ORBX_FS_BH)
    int cells_bw, cells_bh;       // TMA box of a fast_cells_kernel tile (levels with cells > 32 x 32, ORBX_FAST_LEGACY=1)
    int fast_nc, fast_nb, fast_warps;   // max cells per tile, tile buffers per warp (legacy kernel: 1 or 2), warps per CTA
    int fast_legacy;              // 1: fast_cells_kernel (ORBX_FAST_LEGACY=1, A/B measurements)
    int strips_per_frame;
    int strip_tab_off;            // strip table (level | cell row << 4 | first cell << 16 per strip of a frame) inside the tap tables, OrbxTap units
    int strip_rec_off;            // OrbxStripRec of every strip of a frame (same order as the strip table), OrbxTap units, 16-byte aligned
    int seg_first[6], seg_count[6];   // table segments: strips of levels 0-1, strips of levels 2+, big-cell levels 0-1, big-cell levels 2+, tall-cell strips of levels 0-1, of levels 2+
    int ini_th, min_th;
    long long slab_bytes;     // one frame's pyramid slab
    float atan_p1, atan_p3, atan_p5, atan_p7;   // cv::fastAtan2 coefficients (float products)
    float factor_pi;          // (float)(CV_PI/180.f) (:107)
    int umax[16];             // (:454-469)
    OrbxLevel lv[ORBX_MAXL];
};

// Everything fast_strips_kernel needs to know about one strip of a frame (frame-independent, built with the plan).
// "tile byte" = byte offset inside the strip's TMA tile row; the strip's scoring pixels (window x in [3, sw - 3)) are
// tile byte columns [lo, hi).
struct OrbxStripRec {
    uint32_t tile_xy;         // TMA start: plane column (multiple of 16) | plane row << 16
    uint32_t shape;           // level | cells << 4 | scoring rows (0: window smaller than 7 x 7) << 8 | delta0 << 16 | lo / 4 << 24
    uint32_t cols;            // lo | hi << 16
    uint32_t cellw;           // wCell | wcell_recip << 16
    uint32_t cell0;           // cell_rec index of the strip's first cell inside a frame
    uint32_t origin;          // level x | y << 16 of the first scoring pixel (:822-823 offsets included)
    uint32_t frame;           // filled in by the kernel
    uint32_t window;          // strip window width | height << 16
};

// Column plan of one lane of pyr_resize8_kernel: 8 adjacent plane columns (from plane column 8 + 8 * group), as two
// halves of 4.  Each half reads 3 aligned source words starting `base` bytes after source pixel x = 0; shifting by `sh`
// bits leaves the 8 source bytes from the half's smallest tap offset on in two registers (U, V); a pixel pair's two
// source byte pairs are then ONE PRMT (sel) and each pixel's row sum ONE IDP.2A with its packed weights (coef).
struct OrbxCol8 {
    int base[2];              // byte offsets (multiples of 4) of the halves' first source word
    uint32_t sh;              // bit shift of half 0 | half 1 << 8 (0, 8, 16 or 24)
    uint32_t pad;
    uint32_t sel[4];          // PRMT selectors of pixel pairs (0,1) (2,3) | (4,5) (6,7)
    uint32_t coef[8];         // c0 | c1 << 16 (11-bit fixed point weights, App. A-1)
};

// One bilinear tap pair of cv::resize INTER_LINEAR 8U
struct OrbxTap {
    int ofs;                  // first source index; second is min(ofs + 1, ssize - 1)
    short c0, c1;             // 11-bit fixed-point weights
};
