/*
 * orbx.h -- C ABI of the B200-native ORB feature front end.
 *
 * This is the drop-in boundary for ONE path of yxqc/ORBSLAM2_with_quadrics:
 * ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111,
 * src/ORBextractor.cc:410-470 and :1043-1132).  Everything behind these entry
 * points runs as hand-written sm_100a CUDA kernels; there is no CPU fallback.
 * The OpenCV-typed C++ class that ORB-SLAM2 links against
 * (orbslam2_with_quadrics_b200/cpp/ORBextractor.{h,cc}) and the Python mirror
 * (orbslam2_with_quadrics_b200/extractor.py) are thin adapters over this header.
 * INTEGRATION.md shows the reference-side change (CMakeLists.txt:57,76-82).
 *
 * Conventions: plain C, POD only, no exceptions cross the boundary.  Every
 * function returns 0 (ORBX_OK) or a negative orbx_status.  A handle owns one
 * CUDA stream plus its device and pinned-host buffers and, like the reference
 * object (src/Frame.cc:78-81), is NOT re-entrant: use one handle per host
 * thread; distinct handles run concurrently (stereo = two handles, two streams).
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ORBX_API __attribute__((visibility("default")))
#else
#define ORBX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE_THRESHOLD 19 /* src/ORBextractor.cc:74 */

typedef enum orbx_status {
    ORBX_OK = 0,
    ORBX_ERR_BAD_ARGS = -1,           /* null pointer, n > max_batch, non-positive sizes ...            */
    ORBX_ERR_BAD_GEOMETRY = -2,       /* a level narrower than 62 px, aspect ratio with 0 root nodes,   */
                                      /* or an image larger than 4096 px     */
    ORBX_ERR_CUDA = -3,               /* a CUDA runtime call failed; see orbx_last_cuda_error()         */
    ORBX_ERR_CANDIDATE_OVERFLOW = -4, /* more FAST corners than the candidate buffer holds              */
    ORBX_ERR_NO_DEVICE = -5,          /* no CUDA device / device ordinal out of range                   */
    ORBX_ERR_OUT_OF_MEMORY = -6,
    ORBX_ERR_EMPTY_IMAGE = -7         /* orbx_extract*: w*h == 0; outputs untouched (:1046-1047)        */
} orbx_status;

/* The five constructor arguments of ORBextractor (src/ORBextractor.cc:410-414) + placement. */
typedef struct orbx_config {
    int nfeatures;
    float scale_factor;   /* 1 < scale_factor <= 2 */
    int nlevels;
    int ini_th_fast;
    int min_th_fast;
    int device;           /* CUDA device ordinal                                                        */
    int max_batch;        /* frames processed per launch sequence (>= 1)                                */
    int download_pyramid; /* 1: copy the padded pyramid to pinned host memory on every extract, as the  */
                          /*    public mvImagePyramid contract requires (read by src/Frame.cc:563-580); */
                          /* 0: leave it in HBM (monocular / RGB-D callers never read it)               */
    int candidate_divisor;/* per-level candidate capacity = level_pixels / divisor + 1024 (0 -> 8)      */
    int device_chunks;    /* orbx_extract_device: sub-batches run on concurrent kernel streams (0 -> 2; 1..4).  */
                          /* 1 when several handles are driven back to back (double buffering): their launch  */
                          /* sequences already overlap, and whole batches make the larger launches            */
    int reserved[6];
} orbx_config;

/* Same field order and size (28 bytes) as cv::KeyPoint. */
typedef struct orbx_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

/* Result of one frame; pointers are library-owned pinned host memory, valid until the
 * next extract call on the same handle.  desc is n x 32 bytes, row i <-> kps[i]
 * (read by consumers as int32[8], src/ORBmatcher.cc:1649-1650). */
typedef struct orbx_result {
    int n;
    int status; /* per-frame orbx_status */
    const orbx_keypoint* kps;
    const uint8_t* desc;
} orbx_result;

typedef struct orbx_handle orbx_handle;

/* ORBextractor::ORBextractor (src/ORBextractor.cc:410-470). */
ORBX_API int orbx_create(const orbx_config* cfg, orbx_handle** out);
/* ~ORBextractor (include/ORBextractor.h:54). */
ORBX_API int orbx_destroy(orbx_handle* h);

/* ORBextractor::operator() (src/ORBextractor.cc:1043-1105) on one host image (CV_8UC1, any
 * stride).  Blocks until the outputs are in host memory. */
ORBX_API int orbx_extract(orbx_handle* h, const uint8_t* img, int width, int height, size_t stride, orbx_result* result);

/* The same for n <= max_batch frames of identical size in one launch sequence.  One (width, height) per call by design: the
 * plan (level sizes, cell grid, tap tables, tensor maps, HBM layout) is built per geometry and every kernel of the sequence
 * covers all frames of the batch with it, so a per-frame w[] / h[] (SURVEY.md 8b's sketch) would mean one launch sequence per
 * distinct size anyway.  Streams of different sizes use one handle per geometry -- in ORB-SLAM2 a camera is one
 * ORBextractor, so that is the reference's own arrangement (src/Tracking.cc:121-127); a handle that sees a new size rebuilds
 * its plan (supported, costs a reallocation). */
ORBX_API int orbx_extract_batch(orbx_handle* h, int n, const uint8_t* const* imgs, int width, int height,
                       const size_t* strides, orbx_result* results);

/* Device-resident variant for roofline timing: n images already in HBM at
 * d_imgs + i*frame_stride, row pitch `pitch`.  Only enqueues work on the handle's stream. */
ORBX_API int orbx_extract_device(orbx_handle* h, int n, const uint8_t* d_imgs, int width, int height, size_t pitch,
                        size_t frame_stride);
/* Copies the results of the last orbx_extract_device to host and waits for them. */
ORBX_API int orbx_fetch_results(orbx_handle* h, int n, orbx_result* results);

/* ---- Frame::ComputeStereoMatches (reference src/Frame.cc:466-640) on the device-resident results of a left and a
 * right extraction: the immediate consumer of the path in the stereo configurations.  Running it here means the
 * pyramids never have to leave HBM (download_pyramid = 0).  Pair i matches frame left_frames[i] of the last extract on
 * `left` against frame right_frames[i] of the last extract on `right` (NULL index arrays mean i); `left` and `right`
 * may be the same handle (a batch that holds both eyes).  Both handles must share constructor arguments, image size
 * and device.  mbf = baseline * fx, mb = baseline (Frame::mbf, Frame::mb, src/Frame.cc:97-98).  Thresholds are the
 * reference's ORBmatcher::TH_HIGH / TH_LOW (src/ORBmatcher.cc:37-38).
 * Result: mvuRight / mvDepth, one float per LEFT keypoint in keypoint order, -1 where there is no match; pinned host
 * memory owned by `left`, valid until its next stereo call. */
typedef struct orbx_stereo_result {
    int n;                 /* left keypoints */
    const float* u_right;  /* mvuRight */
    const float* depth;    /* mvDepth  */
} orbx_stereo_result;
ORBX_API int orbx_stereo_match(orbx_handle* left, orbx_handle* right, int npairs, const int* left_frames,
                               const int* right_frames, float mbf, float mb, orbx_stereo_result* results);
/* The same split in two for device-side timing: enqueue only / copy back and wait. */
ORBX_API int orbx_stereo_match_device(orbx_handle* left, orbx_handle* right, int npairs, const int* left_frames,
                                      const int* right_frames, float mbf, float mb);
ORBX_API int orbx_stereo_fetch(orbx_handle* left, int npairs, const int* left_frames, orbx_stereo_result* results);

/* ---- cvtColor in front of the path (reference src/Tracking.cc:172-197, :212-225, :242-255: every GrabImage* converts
 * RGB/BGR(A) to gray on the CPU before the Frame is built).  These variants take the colour frame and convert it on the
 * device with OpenCV 4.x's 8U arithmetic, gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15, bit-identical to cv::cvtColor.
 * strides / pitch are in BYTES of the colour rows. */
typedef enum orbx_pixel_format {
    ORBX_GRAY8 = 0,
    ORBX_BGR8 = 1,  /* CV_BGR2GRAY  */
    ORBX_RGB8 = 2,  /* CV_RGB2GRAY  */
    ORBX_BGRA8 = 3, /* CV_BGRA2GRAY */
    ORBX_RGBA8 = 4  /* CV_RGBA2GRAY */
} orbx_pixel_format;
ORBX_API int orbx_extract_batch_color(orbx_handle* h, int n, const uint8_t* const* imgs, int width, int height,
                                      const size_t* strides, int format, orbx_result* results);
ORBX_API int orbx_extract_device_color(orbx_handle* h, int n, const uint8_t* d_imgs, int width, int height, size_t pitch,
                                       size_t frame_stride, int format);

/* ---- Frame::UndistortKeyPoints + Frame::AssignFeaturesToGrid (reference src/Frame.cc:404-434, :230-245, PosInGrid
 * :382-392, ComputeImageBounds :436-464) on the keypoints the last extract left in HBM: the step after the path in
 * every Frame constructor (:88-92, :141-145, :198-202).  K4 = {fx, fy, cx, cy} (Frame::mK), dist = mDistCoef
 * (k1, k2, p1, p2[, k3]; ndist = 4 or 5).  cv::undistortPoints is restated as the published 5-iteration algorithm in
 * double, bit-identical to OpenCV 4.13; k1 == 0 copies the points, as the reference does.
 * Result per frame: mvKeysUn[i].pt as n x 2 floats (all other KeyPoint fields equal mvKeys[i]), mGrid as CSR
 * (cell gx * 48 + gy = mGrid[gx][gy], 64 x 48 cells, keypoint indices in push_back order) and mnMinX, mnMaxX, mnMinY,
 * mnMaxY.  Pinned host memory owned by the handle, valid until its next orbx_undistort_grid call. */
typedef struct orbx_grid_result {
    int n;                     /* keypoints of the frame */
    int n_in_grid;             /* entries of cell_items (keypoints that fall outside the grid are dropped, :387-388) */
    const float* xy_un;        /* n x 2 */
    const int32_t* cell_start; /* 64 * 48 + 1 */
    const int32_t* cell_items;
    float bounds[4];           /* mnMinX, mnMaxX, mnMinY, mnMaxY */
} orbx_grid_result;
ORBX_API int orbx_undistort_grid(orbx_handle* h, int nframes, const int* frames, const float* K4, const float* dist, int ndist,
                                 orbx_grid_result* results);

/* ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (reference
 * src/ORBmatcher.cc:1328-1470, with Frame::GetFeaturesInArea src/Frame.cc:327-380, DescriptorDistance
 * src/ORBmatcher.cc:1647-1663 and ComputeThreeMaxima :1601-1642): the descriptor consumer of every tracked frame
 * (Tracking::TrackWithMotionModel, src/Tracking.cc:885, :891).  The CurrentFrame side is what this handle holds in HBM:
 * frame `cur_frame` of the last extract, its mvKeysUn / mGrid from orbx_undistort_grid (which must have run for that
 * frame; its K4 must be the K4 passed here) and, with use_stereo = 1, its mvuRight from orbx_stereo_match on this
 * handle (use_stereo = 0: every mvuRight is -1, the monocular Frame).  The LastFrame side comes from the caller's map.
 * As at both call sites of the reference, CurrentFrame.mvpMapPoints starts all NULL.
 * Result: match[i2] = index i of the LastFrame keypoint whose map point CurrentFrame.mvpMapPoints[i2] holds at
 * return, -1 for NULL; nmatches = the function's return value (it counts overwritten matches, as the reference does).
 * Pinned host memory owned by the handle, valid until its next search call. */
typedef struct orbx_projection_query {
    int cur_frame;            /* frame of this handle's last extract */
    int n_last;               /* LastFrame.N */
    const float* world_pos;   /* n_last x 3: pMP->GetWorldPos() (ignored where there is no map point) */
    const uint8_t* mp_desc;   /* n_last x 32: pMP->GetDescriptor() */
    const int32_t* mp_obs;    /* n_last: pMP->Observations(); < 0 where LastFrame.mvpMapPoints[i] is NULL */
    const uint8_t* outlier;   /* n_last: LastFrame.mvbOutlier[i]; NULL = none */
    const int32_t* octave;    /* n_last: LastFrame.mvKeys[i].octave */
    const float* angle;       /* n_last: LastFrame.mvKeysUn[i].angle */
    float Tcw_cur[16];        /* CurrentFrame.mTcw, row-major 4 x 4 */
    float Tcw_last[16];       /* LastFrame.mTcw */
} orbx_projection_query;
typedef struct orbx_projection_result {
    int n;                    /* CurrentFrame.N */
    int nmatches;
    int rounds;               /* fixed-point rounds the device needed (diagnostic) */
    const int32_t* match;     /* n */
} orbx_projection_result;
ORBX_API int orbx_search_by_projection(orbx_handle* h, int nqueries, const orbx_projection_query* queries, const float* K4, float mbf,
                                       float mb, float th, int mono, int check_orientation, int use_stereo,
                                       orbx_projection_result* results);
/* The same split in two for device-side timing: stage + enqueue only / copy back and wait. */
ORBX_API int orbx_search_by_projection_device(orbx_handle* h, int nqueries, const orbx_projection_query* queries, const float* K4,
                                              float mbf, float mb, float th, int mono, int check_orientation, int use_stereo);
ORBX_API int orbx_search_by_projection_fetch(orbx_handle* h, int nqueries, const orbx_projection_query* queries,
                                             orbx_projection_result* results);

/* ---- ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (reference src/ORBmatcher.cc:45-129
 * with RadiusByViewingCos :131-137): the matcher of Tracking::SearchLocalPoints (src/Tracking.cc:1184-1194, ORBmatcher(0.8),
 * th = 1 / 3 for RGB-D / 5 after a relocalisation), every tracked frame.  The map points carry the tracking fields
 * Frame::isInFrustum (src/Frame.cc:255-325) left on them; the frame side is the handle's device-resident state as for
 * orbx_search_by_projection.  Unlike there, the frame already holds map points: cur_obs names them.
 * Result (orbx_projection_result): match[i2] = index into vpMapPoints of the point F.mvpMapPoints[i2] was SET to by
 * this call, -1 = left as it was; nmatches = the function's return value. */
typedef struct orbx_local_points_query {
    int cur_frame;             /* frame of this handle's last extract */
    int n_points;              /* vpMapPoints.size() */
    const uint8_t* in_view;    /* n: pMP->mbTrackInView && !pMP->isBad(); NULL = all */
    const float* proj_xy_xr;   /* n x 3: mTrackProjX, mTrackProjY, mTrackProjXR */
    const int32_t* scale_level;/* n: mnTrackScaleLevel */
    const float* view_cos;     /* n: mTrackViewCos */
    const uint8_t* mp_desc;    /* n x 32: pMP->GetDescriptor() */
    const int32_t* mp_obs;     /* n: pMP->Observations() */
    const int32_t* cur_obs;    /* F.N: < 0 where F.mvpMapPoints[i2] is NULL, else its Observations(); NULL = all NULL */
} orbx_local_points_query;
ORBX_API int orbx_search_local_points(orbx_handle* h, int nqueries, const orbx_local_points_query* queries, float th, float nnratio,
                                      int use_stereo, orbx_projection_result* results);
ORBX_API int orbx_search_local_points_device(orbx_handle* h, int nqueries, const orbx_local_points_query* queries, float th,
                                             float nnratio, int use_stereo);

/* ---- Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) (reference src/Frame.cc:269-325) with MapPoint::PredictScale(dist,
 * Frame*) and GetMin/MaxDistanceInvariance (src/MapPoint.cc:402-417, :373-383): the loop of Tracking::SearchLocalPoints
 * (src/Tracking.cc:1165-1178, viewingCosLimit 0.5) that leaves the tracking fields on every local map point, i.e. the producer of
 * orbx_local_points_query's in_view / proj_xy_xr / scale_level / view_cos -- the result arrays can be handed to
 * orbx_search_local_points as they are.  A query is one Frame pose and its list of map points; nothing of the handle's extraction
 * state is read (only its device, stream and nlevels), so the call may precede the frame's extraction.
 *   consider[i] = 0: the point is not handed to isInFrustum (already seen in this frame, src/Tracking.cc:1169, or isBad(), :1171);
 *                    its in_view is 0.
 *   bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY} (orbx_grid_result.bounds), log_scale_factor = Frame::mfLogScaleFactor.
 * PredictScale goes through the host libm's logf; the library evaluates it with the calling process's own logf (per level, the
 * smallest ratio that reaches it) and the device only compares, so the levels are the reference's bit for bit.
 * Result arrays: pinned host memory owned by the handle, valid until its next orbx_is_in_frustum call; fields of points that are
 * not in view are 0. */
typedef struct orbx_frustum_query {
    int n_points;              /* mvpLocalMapPoints.size() */
    const uint8_t* consider;   /* n: see above; NULL = all */
    const float* world_pos;    /* n x 3: pMP->GetWorldPos() */
    const float* normal;       /* n x 3: pMP->GetNormal() */
    const float* min_dist;     /* n: mfMinDistance (GetMinDistanceInvariance()'s 0.8f is applied by the library) */
    const float* max_dist;     /* n: mfMaxDistance (likewise 1.2f; PredictScale reads the raw value) */
    float Tcw[16];             /* Frame::mTcw, row-major 4 x 4 */
} orbx_frustum_query;
typedef struct orbx_frustum_result {
    int n;                     /* = n_points */
    int n_in_view;             /* nToMatch of Tracking::SearchLocalPoints */
    const uint8_t* in_view;    /* n: mbTrackInView */
    const float* proj_xy_xr;   /* n x 3: mTrackProjX, mTrackProjY, mTrackProjXR */
    const int32_t* scale_level;/* n: mnTrackScaleLevel */
    const float* view_cos;     /* n: mTrackViewCos */
} orbx_frustum_result;
ORBX_API int orbx_is_in_frustum(orbx_handle* h, int nqueries, const orbx_frustum_query* queries, const float* K4, float mbf,
                                const float* bounds, float log_scale_factor, float viewing_cos_limit, orbx_frustum_result* results);

/* ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
 * (reference src/ORBmatcher.cc:1472-1599): the projection matcher of Tracking::Relocalization (src/Tracking.cc:1452 with th 10 /
 * ORBdist 100, :1466 with th 3 / ORBdist 64, ORBmatcher(0.9, true)).  The frame side is the handle's device-resident state
 * as for orbx_search_by_projection (no mvuRight test in this overload).  The KeyFrame side comes from the caller's map; the two
 * scalar steps that go through the host's libm and the map's mutexes stay with the caller, who stages their result per point:
 *   search[i]     = pMP && !pMP->isBad() && !sAlreadyFound.count(pMP) && dist3D within [GetMinDistanceInvariance(),
 *                   GetMaxDistanceInvariance()] (:1491-1495, :1510-1518; dist3D = cv::norm(x3Dw - Ow), Ow = -Rcw.t()*tcw)
 *   pred_level[i] = pMP->PredictScale(dist3D, &CurrentFrame) (:1520; read only where search[i] is set)
 * Projection, image-bounds test, window search over levels pred_level-1 .. +1, "any holder blocks" (:1540-1541), the ORBdist
 * test, rotation histogram and culling run on the device.
 * Result (orbx_projection_result): match[i2] = index i of the KeyFrame point CurrentFrame.mvpMapPoints[i2] was SET to by this
 * call and kept after the orientation check, -1 = left as it was; nmatches = the function's return value. */
typedef struct orbx_keyframe_projection_query {
    int cur_frame;             /* frame of this handle's last extract */
    int n_points;              /* pKF->GetMapPointMatches().size() */
    const uint8_t* search;     /* n: see above */
    const float* world_pos;    /* n x 3: pMP->GetWorldPos() (ignored where search[i] is 0) */
    const int32_t* pred_level; /* n */
    const uint8_t* mp_desc;    /* n x 32: pMP->GetDescriptor() */
    const float* kf_angle;     /* n: pKF->mvKeysUn[i].angle */
    const int32_t* cur_held;   /* F.N: > 0 where CurrentFrame.mvpMapPoints[i2] is not NULL at entry; NULL = all NULL */
    float Tcw_cur[16];         /* CurrentFrame.mTcw, row-major 4 x 4 */
} orbx_keyframe_projection_query;
ORBX_API int orbx_search_by_projection_kf(orbx_handle* h, int nqueries, const orbx_keyframe_projection_query* queries, const float* K4,
                                          float th, int orb_dist, int check_orientation, orbx_projection_result* results);
ORBX_API int orbx_search_by_projection_kf_device(orbx_handle* h, int nqueries, const orbx_keyframe_projection_query* queries,
                                                 const float* K4, float th, int orb_dist, int check_orientation);

/* ---- ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12,
 * int windowSize) (reference src/ORBmatcher.cc:405-520): the matcher of Tracking::MonocularInitialization (src/Tracking.cc:597-600,
 * ORBmatcher(0.9, true), windowSize 100).  F2 (the current frame) is the handle's device-resident state: frame `cur_frame` of
 * the last extract with its mvKeysUn / mGrid from orbx_undistort_grid.  F1 (mInitialFrame, which the caller keeps) comes as
 * arrays over its undistorted keypoints.  The sequential rules of the reference loop -- a later F1 keypoint takes an F2
 * keypoint from an earlier one when its distance is strictly smaller, and only candidates that survive that filter enter the
 * ratio test -- are replayed in order on the device.
 * Result: matches12[i1] = vnMatches12[i1] (index in F2 or -1), nmatches = the return value, prev_matched = vbPrevMatched as
 * updated at :515-517.  Pinned host memory owned by the handle, valid until its next search call. */
typedef struct orbx_initialization_query {
    int cur_frame;             /* F2: frame of this handle's last extract */
    int n1;                    /* F1.mvKeysUn.size() */
    const int32_t* octave1;    /* n1: F1.mvKeysUn[i1].octave */
    const float* angle1;       /* n1: F1.mvKeysUn[i1].angle */
    const uint8_t* desc1;      /* n1 x 32: F1.mDescriptors */
    const float* prev_matched; /* n1 x 2: vbPrevMatched at entry */
} orbx_initialization_query;
typedef struct orbx_initialization_result {
    int n1;
    int nmatches;
    const int32_t* matches12;  /* n1 */
    const float* prev_matched; /* n1 x 2 */
} orbx_initialization_result;
ORBX_API int orbx_search_for_initialization(orbx_handle* h, int nqueries, const orbx_initialization_query* queries, float nnratio,
                                            int check_orientation, int window, orbx_initialization_result* results);

/* ---- Frame::ComputeBoW (reference src/Frame.cc:395-402): DBoW2's TemplatedVocabulary::transform(features, BowVector&,
 * FeatureVector&, levelsup = 4) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194, :1217-1259) with the ORB
 * vocabulary's settings -- TF-IDF weighting, L1 scoring -- on the descriptors the last extract left in HBM.
 * The vocabulary is uploaded once: node i (0 = root) has the children child_items[child_start[i] .. child_start[i+1]) in
 * DBoW2's order (m_nodes[i].children), a 32-byte descriptor, a weight and, for leaves, a word id (ignored for inner
 * nodes).  child_items lists every node but the root exactly once.  L = the vocabulary's depth (m_L).
 * Result per frame: the BowVector as (word id, value) in std::map order, values L1-normalised doubles bit-identical to
 * the reference's; the FeatureVector as (node id, feature index) pairs in std::map order with each node's features in
 * push_back order.  Stopped words (weight 0) appear in neither (:1157).  Pinned host memory owned by the handle, valid
 * until its next orbx_compute_bow call. */
typedef struct orbx_vocabulary orbx_vocabulary;
ORBX_API int orbx_vocabulary_create(int device, int n_nodes, int L, const int32_t* child_start, const int32_t* child_items,
                                    const uint8_t* node_desc, const double* node_weight, const int32_t* node_word,
                                    orbx_vocabulary** out);
ORBX_API int orbx_vocabulary_destroy(orbx_vocabulary* voc);
typedef struct orbx_bow_result {
    int n_words;                 /* mBowVec.size() */
    const uint32_t* word_ids;
    const double* word_values;
    int n_features;              /* features listed in mFeatVec */
    const uint32_t* fv_nodes;
    const uint32_t* fv_features;
} orbx_bow_result;
ORBX_API int orbx_compute_bow(orbx_handle* h, const orbx_vocabulary* voc, int nframes, const int* frames, int levelsup,
                              orbx_bow_result* results);
ORBX_API int orbx_compute_bow_device(orbx_handle* h, const orbx_vocabulary* voc, int nframes, const int* frames, int levelsup);

/* ---- ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (reference
 * src/ORBmatcher.cc:159-288; Tracking::TrackReferenceKeyFrame src/Tracking.cc:771-776 with ORBmatcher(0.7, true),
 * Tracking::Relocalization :1376 with 0.75): matching inside common vocabulary nodes.  The Frame side is frame
 * `cur_frame` of this handle's last extract together with its FeatureVector from the last orbx_compute_bow (which must
 * have covered that frame); the KeyFrame side comes from the caller.
 * Result (orbx_projection_result): match[iF] = index of the KeyFrame feature whose map point vpMapPointMatches[iF]
 * holds, -1 = NULL; nmatches = the function's return value. */
typedef struct orbx_bow_match_query {
    int cur_frame;
    int n_kf;                      /* pKF->N */
    const uint8_t* kf_desc;        /* n_kf x 32: pKF->mDescriptors */
    const uint8_t* kf_valid;       /* n_kf: 0 = no map point, 1 = a good one, 2 = pMP->isBad() */
    const float* kf_angle;         /* n_kf: pKF->mvKeysUn[i].angle */
    int n_kf_fv;                   /* entries of pKF->mFeatVec */
    const uint32_t* kf_fv_nodes;   /* (node id, feature index) pairs in std::map / push_back order */
    const uint32_t* kf_fv_features;
} orbx_bow_match_query;
ORBX_API int orbx_search_by_bow(orbx_handle* h, int nqueries, const orbx_bow_match_query* queries, float nnratio,
                                int check_orientation, orbx_projection_result* results);
ORBX_API int orbx_search_by_bow_device(orbx_handle* h, int nqueries, const orbx_bow_match_query* queries, float nnratio,
                                       int check_orientation);

/* Pinned host buffers callers may fill with frames so that H2D copies are asynchronous DMA. */
ORBX_API int orbx_alloc_host(size_t bytes, void** out);
ORBX_API int orbx_free_host(void* p);

/* mvImagePyramid (include/ORBextractor.h:85): level `level` of frame `frame` of the last extract
 * as a host-readable plane.  *image points at pixel (0,0) of the w x h level; the 19-pixel
 * BORDER_REFLECT_101 frame around it is valid (src/ORBextractor.cc:1113-1128).  Needs
 * download_pyramid = 1. */
ORBX_API int orbx_pyramid_level(orbx_handle* h, int frame, int level, const uint8_t** image, int* width, int* height,
                       size_t* step);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:66-83): nlevels floats each, owned by the handle. */
ORBX_API int orbx_scale_tables(orbx_handle* h, const float** scale, const float** inv_scale, const float** sigma2,
                      const float** inv_sigma2);
ORBX_API int orbx_get_levels(orbx_handle* h);                 /* GetLevels (:63)        */
ORBX_API float orbx_get_scale_factor(orbx_handle* h);         /* GetScaleFactor (:66)   */
/* mnFeaturesPerLevel (src/ORBextractor.cc:435-446) and umax (:454-469). */
ORBX_API int orbx_level_quotas(orbx_handle* h, int* quotas /* nlevels */, int* umax16 /* 16 */);
/* Level sizes for an input of width x height (src/ORBextractor.cc:1111-1112). */
ORBX_API int orbx_level_sizes(orbx_handle* h, int width, int height, int* widths, int* heights);

/* Intermediate results of the last extract, for stage-level parity tests. */
typedef enum orbx_stage {
    ORBX_STAGE_PYRAMID = 0,    /* padded plane, (h+38) x (w+38) bytes, step w+38                        */
    ORBX_STAGE_CANDIDATES = 1, /* int32 (x, y, response) triples in the reference's push order, box     */
                               /* coordinates (src/ORBextractor.cc:818-826)                             */
    ORBX_STAGE_KEPT = 2,       /* int32 (x, y, response) triples after DistributeOctTree, level coords  */
    ORBX_STAGE_ANGLES = 3,     /* float32 per kept keypoint (IC_Angle, :77-104)                          */
    ORBX_STAGE_BLURRED = 4     /* blurred level, h x w bytes, step w (:1085-1086)                        */
} orbx_stage;
/* Writes up to cap bytes to out; *bytes receives the full size of the stage output. */
ORBX_API int orbx_stage_dump(orbx_handle* h, int frame, int level, int stage, void* out, size_t cap, size_t* bytes);

/* The handle's cudaStream_t (for event timing from the caller's side) and a blocking wait on it. */
ORBX_API void* orbx_stream(orbx_handle* h);
ORBX_API int orbx_synchronize(orbx_handle* h);

/* Per-stage device time of the extract calls since the last reset (CUDA events on the handle's
 * stream; enabling it adds event records between kernels).  names/ms arrays hold up to cap entries;
 * returns the number of stages. */
ORBX_API int orbx_stage_timing_enable(orbx_handle* h, int enable);
ORBX_API int orbx_stage_timing_read(orbx_handle* h, int cap, const char** names, float* ms, int* launches);

/* Workload statistics of frame `frame` of the last extract whose results were fetched: FAST corners handed to the quadtree
 * per level (vToDistributeKeys.size(), src/ORBextractor.cc:818-826) and cells re-run at minThFAST per level (:812).  Either
 * output may be NULL; nlevels ints each. */
ORBX_API int orbx_fast_stats(orbx_handle* h, int frame, int* candidates, int* retries);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
ORBX_API long long orbx_launch_count(orbx_handle* h);

/* Algorithmic bytes per image of the whole path. */
ORBX_API long long orbx_algorithmic_bytes(orbx_handle* h, int width, int height);

ORBX_API const char* orbx_strerror(int status);
ORBX_API const char* orbx_last_cuda_error(orbx_handle* h);
ORBX_API const char* orbx_version(void);

/* ---- Host placement for multi-GPU boxes (one camera stream per GPU, fed from host memory).  Binds the
 * CALLING thread to the CPUs of the NUMA node the device hangs off (sysfs: /sys/bus/pci/devices/<bus id>/numa_node and
 * /sys/devices/system/node/node<k>/cpulist), intersected with the CPUs the process may use; threads and pinned
 * allocations (orbx_create, orbx_alloc_host) made afterwards from this thread inherit the placement (first touch).  Call
 * it once per feeding thread before orbx_create.  *node = the device's NUMA node (-1: unknown / single node),
 * *ncpus = CPUs the thread is now bound to (0: affinity left unchanged, e.g. the container exposes no CPU of that node).
 * The reference has no counterpart: its extractor threads run wherever the scheduler puts them (src/Frame.cc:78-81). */
ORBX_API int orbx_bind_thread_to_device(int device, int* node, int* ncpus);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
