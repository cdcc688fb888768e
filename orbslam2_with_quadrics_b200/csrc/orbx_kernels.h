// Launch wrappers of the stage kernels (orbx_kernels.cu), used by the C-ABI layer (orbx_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "orbx_plan.h"

#ifndef ORBX_FAST_WARPS
#define ORBX_FAST_WARPS 8
#endif
// fast_strips_kernel shape (compile-time; -D overrides are for A/B builds)
#ifndef ORBX_FS_WARPS
#define ORBX_FS_WARPS 2          // warps per strip (= per CTA): 1, 2 or 4.  Measured at 64 x 1080p (ms): 1 -> 1.118, 2 -> 0.891, 4 -> 0.993
#endif
#ifndef ORBX_FS_NBUF
#define ORBX_FS_NBUF 1           // tile buffers per CTA.  2 hides the TMA latency but costs resident CTAs: 0.991 vs 0.891 ms
#endif
#define ORBX_FS_BH 38            // tile rows of a strip: cells of <= 32 scoring rows + the 6-px frame
#define ORBX_FS_BH_TALL 46       // ... of 33 .. 40 scoring rows (fast_strips_kernel<BW, 5>)
#ifndef ORBX_FS_QCAP
#define ORBX_FS_QCAP 1536        // survivor queue entries of a strip (>= 1280: the per-cell path needs 32 x 40).  1536 lets 14 CTAs share an SM (0.790 vs 0.806 ms with 2048 / 12 CTAs); 1024 sends cluttered strips to the per-cell path (1.00 ms)
#endif
#define ORBX_OT_THREADS 1024
#define ORBX_OT_KEYCAP 8192     // candidates of one level kept in shared memory by the octree kernel (6 bytes each)

// per-frame device status bits
#define ORBX_DEV_CAND_OVERFLOW 1
#define ORBX_DEV_NODE_OVERFLOW 2

namespace orbx {

// tile_maps: build_pyr_tile_maps' descriptors (launches with full-length strips read their source through TMA tiles), or NULL
void launch_pyr_level(const OrbxPlan* d_plan, const OrbxPlan& hp, int l, int nframes, int num_sms, const uint8_t* imgs,
                      size_t img_pitch, size_t img_frame_stride, uint8_t* pyr, const OrbxTap* taps, cudaStream_t st,
                      const void* tile_maps = nullptr, int frame0 = 0);
int build_pyr_tile_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, void* out_maps);
// cvtColor(RGB/BGR(A) -> GRAY), src/Tracking.cc:172-255; format = orbx_pixel_format (1..4)
cudaError_t launch_cvt_gray(const uint8_t* src, size_t src_pitch, size_t src_frame_stride, int w, int h, int nframes, int format,
                            uint8_t* dst, size_t dst_pitch, size_t dst_frame_stride, cudaStream_t st);
size_t fast_smem_bytes(const OrbxPlan& hp);
int build_fast_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, void* out_maps);
size_t fast_maps_bytes();
cudaError_t launch_fast(const OrbxPlan* d_plan, const OrbxPlan& hp, const void* maps, const OrbxTap* taps, int frame0, int nframes,
                        int seg, int nseg, int num_sms, uint32_t* cand, uint2* cell_rec, int* level_counts, int* work_counters,
                        int* status, int* retry_counts, cudaStream_t st);
size_t octree_smem_bytes(const OrbxPlan& hp);
cudaError_t launch_octree(const OrbxPlan* d_plan, const OrbxPlan& hp, int nframes, const uint32_t* cand,
                          const uint2* cell_rec, uint32_t* cand_sorted, uint16_t* key_node, int* sorted_counts,
                          uint32_t* kept, int* kept_counts, int* status, cudaStream_t st);
void launch_blur(const OrbxPlan* d_plan, const OrbxPlan& hp, int nframes, int num_sms, const uint8_t* pyr,
                 uint8_t* blur, cudaStream_t st);
int build_describe_maps(const OrbxPlan& hp, uint8_t* d_pyr, int max_frames, void* out_maps);
cudaError_t launch_describe(const OrbxPlan* d_plan, const OrbxPlan& hp, const void* maps, int frame0, int nframes, int num_sms,
                            const uint32_t* kept, const int* kept_counts, float* angles, float* out_kp, uint8_t* out_desc,
                            cudaStream_t st);

// Frame::ComputeStereoMatches (src/Frame.cc:466-640) on the device-resident outputs of two extractions
cudaError_t launch_stereo(const OrbxPlan* d_plan, const OrbxPlan& hp, int num_sms, const uint8_t* pyrL, const float* kpL,
                          const uint8_t* descL, const int* countsL, const uint8_t* pyrR, const float* kpR, const uint8_t* descR,
                          const int* countsR, const int* d_pairs, int npairs, float mbf, float mb, float* u_right, float* depth,
                          int* sad, int* row_start, uint16_t* bucket, cudaStream_t st);
size_t frustum_query_bytes();
void frustum_fill_query(void* dst, const float* Rcw, const float* tcw, const float* Ow, int n, int off);
cudaError_t launch_frustum(int nq, int max_n, const void* d_queries, const float* K4, const float* bounds, float mbf, float cos_limit,
                           int nlevels, const float* thresholds, const uint8_t* consider, const float* world, const float* normal,
                           const float* min_dist, const float* max_dist, uint8_t* in_view, float* proj, int* level, float* view_cos,
                           cudaStream_t st);
size_t search_projection_query_bytes();
void search_projection_fill_query(void* dst, const float* Rcw, const float* tcw, int n_last, int frame, int fwd, int bwd);
cudaError_t launch_search_projection(const OrbxPlan* d_plan, const OrbxPlan& hp, int mode, int nq, const void* d_queries, const float* K4,
                                     const float* bounds, float mbf, float th, float nnratio, int check_ori, int match_th, int cap, int list_cap,
                                     const float* world, const uint8_t* mp_desc, const int* mp_obs, const int* last_octave,
                                     const float* last_angle, const float* kp, const uint8_t* desc, const int* kept_counts,
                                     const float* xy_un, const int* cell_start, const int* cell_items, const float* u_right,
                                     const int* cur_obs, uint32_t* cand_list, int* cand_count, int* match_out, int* stats_out,
                                     cudaStream_t st);
cudaError_t launch_compute_bow(const OrbxPlan* d_plan, const OrbxPlan& hp, const int* voc_child_start, const int* voc_child_items,
                               const uint8_t* voc_desc, const double* voc_weight, const int* voc_word, int n_nodes, int L,
                               const int* d_frames, int nframes, int levelsup, const uint8_t* desc, const int* kept_counts, int* leaf,
                               int* nid, unsigned* word_ids, double* word_values, unsigned* fv_nodes, unsigned* fv_features,
                               int* counts_out, cudaStream_t st);
cudaError_t launch_search_init(const OrbxPlan* d_plan, const OrbxPlan& hp, int nq, const void* d_queries, const float* bounds, float window,
                               float nnratio, int check_ori, int cap, int list_cap, const float* prev, const uint8_t* desc1,
                               const int* octave1, const float* angle1, const float* kp, const uint8_t* desc, const int* kept_counts,
                               const float* xy_un, const int* cell_start, const int* cell_items, uint32_t* cand_list, int* cand_count,
                               uint8_t* bin_scratch, int* match_out, float* prev_out, int* stats_out, cudaStream_t st);
size_t search_bow_query_bytes();
void search_bow_fill_query(void* dst, int frame, int slot, int n_kf, int n_kf_fv);
cudaError_t launch_search_bow(const OrbxPlan* d_plan, const OrbxPlan& hp, int nq, const void* d_queries, int cap, float nnratio,
                              int check_ori, const uint8_t* kf_desc, const uint8_t* kf_valid, const float* kf_angle,
                              const unsigned* kf_fv_nodes, const unsigned* kf_fv_features, const float* kp, const uint8_t* desc,
                              const int* kept_counts, const unsigned* f_fv_nodes, const unsigned* f_fv_features, const int* bow_counts,
                              int* match_out, int* stats_out, cudaStream_t st);
size_t stereo_bucket_entries(const OrbxPlan& hp);     // uint16 entries of one right frame's row table

// Frame::UndistortKeyPoints + AssignFeaturesToGrid (src/Frame.cc:404-434, :230-245) on the device-resident keypoints
cudaError_t launch_undistort_grid(const OrbxPlan* d_plan, const OrbxPlan& hp, const float* kp, const int* kept_counts,
                                  const int* d_frames, int nframes, const double* cam, int distorted, const float* grid, float* xy_un,
                                  int* cell_start, int* cell_items, cudaStream_t st);

}  // namespace orbx
