// Internal launch interface between the C-ABI orchestration (api.cu) and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "common.cuh"

namespace f3dgs {

extern std::atomic<unsigned long long> g_launches;  // kernels launched by this library (api.cu)

struct ViewParams {
    int P, D, M, C;
    int W, H;
    uint32_t grid_x, grid_y;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    float scale_modifier;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
};

// ---- preprocess.cu
void launch_preprocess_fwd(const ViewParams& vp, const float* means3D, const float* scales,
                           const float* rotations, const float* opacities, const float* shs,
                           const float* cov3D_precomp, const float* colors_precomp, bool prefiltered,
                           int* radii, SplatRec* rec, float* cov3D, uint8_t* clamped,
                           uint32_t* tiles_touched, cudaStream_t s);

void launch_preprocess_bwd(const ViewParams& vp, const float* means3D, const int* radii, const float* shs,
                           const uint8_t* clamped, const float* scales, const float* rotations,
                           const float* cov3D, const float* dL_dmean2D, const float* dL_dconic,
                           float* dL_dmean3D, const float* dL_dcolor, float* dL_dcov3D, float* dL_dsh,
                           float* dL_dscale, float* dL_drot, const float* dL_dz, cudaStream_t s,
                           bool accumulate = false, float* grad_accum = nullptr, float* denom = nullptr);

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                         cudaStream_t s);

// ---- binning.cu
void launch_duplicate_keys(int P, const SplatRec* rec, const uint32_t* offsets, const int* radii,
                           uint32_t grid_x, uint32_t grid_y, uint64_t* keys, uint32_t* values,
                           cudaStream_t s);
void launch_tile_ranges(int R, const uint64_t* sorted_keys, uint2* ranges, cudaStream_t s);

// ---- composite_fwd.cu
// returns cudaSuccess or the launch error
cudaError_t launch_composite_fwd(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                 const SplatRec* rec, const float* features, const float* bg,
                                 float* final_T, uint32_t* n_contrib, float* out_color,
                                 float* out_feature, float* out_depth, int* work_counter, cudaStream_t s);

// composite_fwd_tc.cu: the same contract with the feature contraction on the tensor cores (tcgen05, 3xTF32);
// needs C % 4 == 0 and a 16-byte aligned feature matrix
cudaError_t launch_composite_fwd_tc(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                    const SplatRec* rec, const float* features, const float* bg,
                                    float* final_T, uint32_t* n_contrib, float* out_color,
                                    float* out_feature, float* out_depth, int* work_counter, cudaStream_t s);

// geometric-gradient kernel of the two-kernel backward (composite_bwd.cu): alpha-only, two CTAs per SM; with list pointers it
// also emits the per-(tile, block) blend weights for launch_feature_bwd
cudaError_t launch_composite_bwd_geom_slim(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                           const SplatRec* rec, const float* bg, const float* final_T,
                                           const uint32_t* n_contrib, const float* dL_dpix, const float* dL_ddepth,
                                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                                           float* dL_dz, int* work_counter, cudaStream_t s, float* list_w = nullptr,
                                           uint2* list_meta = nullptr, uint32_t* list_cnt = nullptr);

// ---- feature_bwd.cu: feature gradient from the instance lists (second kernel of the two-kernel backward)
cudaError_t launch_feature_bwd(const ViewParams& vp, const uint2* ranges, const float* list_w, const uint2* list_meta,
                               const uint32_t* list_cnt, const float* dL_dfeat_pix, float* dL_dfeature,
                               int* work_counter, cudaStream_t s, bool use_tc = false);

// ---- composite_bwd.cu
cudaError_t launch_composite_bwd(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                 const SplatRec* rec, const float* bg, const float* final_T,
                                 const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dfeat_pix,
                                 const float* dL_ddepth, float* dL_dmean2D, float* dL_dconic,
                                 float* dL_dopacity, float* dL_dcolor, float* dL_dfeature, float* dL_dz,
                                 int* work_counter, cudaStream_t s);

// ---- feature_head.cu
cudaError_t launch_feature_resize_fwd(int C, int H, int W, int Hg, int Wg, const float* fm, const float* gt,
                                      float grad_scale, float* out, float* loss_sum, cudaStream_t s);
cudaError_t launch_feature_resize_bwd(int C, int H, int W, int Hg, int Wg, const float* dout, float* dfm, cudaStream_t s);

}  // namespace f3dgs
