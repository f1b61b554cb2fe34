// ransac.cuh -- launcher of the F-matrix RANSAC kernel (see ransac.cu)
#pragma once
#include <vector>
#include "svo_internal.cuh"

namespace svo {

struct RansacScratch {
    float2* p1;        // [n_slots*2][kp_cap] points of the previous frame (left-left, right-right problems)
    float2* p2;        // [n_slots*2][kp_cap] points of the current frame
    int32_t* n_pts;    // [n_slots*2]
    int32_t* n_iter;   // [n_slots*2] RANSAC / LMedS iterations OpenCV's loop executes for this problem (statistics)
    int kp_cap;
    int32_t* slots;    // == Pools::slots (batch item -> slot)
    uint8_t* inl;      // == Pools::inl    [n_slots*2][kp_cap] output masks
    int32_t* n_inl;    // == Pools::n_inl  [n_slots*2] inlier counts
};

cudaError_t alloc_ransac(RansacScratch& R, int n_slots, int kp_cap, std::vector<void*>& allocs);
cudaError_t configure_ransac(int kp_cap);
// masks for every slot of the running batch: P.inl / P.n_inl (stage4_match_consecutive.cpp:171-241)
cudaError_t launch_ransac(const Pools& P, const RansacScratch& R, int n, cudaStream_t st, int* launches,
                          StepTimer* tm = nullptr);
// stand-alone: host points -> host mask (svo_fundmat_ransac)
cudaError_t ransac_host_points(const RansacScratch& R, int scratch_slot, const float* p1, const float* p2, int n,
                               uint8_t* mask, int* n_inliers, cudaStream_t st);
// stand-alone run7Point on n_sets samples of 7 correspondences (svo_fundmat_7point)
cudaError_t fm7_host_points(const float* p1, const float* p2, int n_sets, double* F, int32_t* n_models, cudaStream_t st);
}  // namespace svo
