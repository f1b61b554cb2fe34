// gn.cuh -- launcher of the stage-5 kernel (see gn.cu)
#pragma once
#include "svo_internal.cuh"

namespace svo {

struct GnScratch {            // per-slot scratch, [max_seqs][kp_cap] unless noted
    float4* prev;             // previous frame: left x, left y, right x, response of the left keypoint
    float4* obs;              // current frame: left x, left y, right x, right y
    double* lmk;              // [..][3] triangulated landmarks
    uint8_t* mask;            // survivors
    double* residual;         // == Pools::residual
    int32_t* cur_idx;         // external path: tracked_pairs[i].second
    uint32_t* occ;            // [max_seqs][occ_words] occupancy grid of the aux NMS
    int occ_words;
    double* ini_pose;         // [max_seqs][6] custom initial pose (getChangeInPose)
};

cudaError_t configure_gn(const Params& p);
// getProjectedCoords (common.cpp:415-470) on device arrays
cudaError_t launch_project(const int4* d_matches, const float2* d_left, const float2* d_right, const int32_t* d_order, int n,
                           const Camera& cam, const double* d_pose, float4* d_out, cudaStream_t st);
cudaError_t launch_gn(const Pools& P, const GnScratch& G, int n, int external, int use_custom_pose,
                      int assume_fast_kps, cudaStream_t st, int* launches, StepTimer* tm = nullptr);
}  // namespace svo
