// gn.cuh -- launcher of the stage-5 kernel (see gn.cu)
#pragma once
#include "svo_internal.cuh"

namespace svo {

struct GnScratch {            // per-slot scratch, [max_seqs][kp_cap] unless noted
    float4* prev;             // previous frame: left x, left y, right x, response of the left keypoint
    float4* obs;              // current frame: left x, left y, right x, right y
    double* lmk;              // [..][3] triangulated landmarks
    uint8_t* mask;            // survivors
    uint8_t* mask0;           // survivors of the aux NMS before the first loop removes outliers (optimization.txt, S5:477-492)
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
// `external` of launch_gn: where the four point lists of stage5_optimization.cpp:419-461 come from
enum { GN_FROM_TRACKED = 0,     // gathered by the kernel from the slot's tracked pairs (nOctaves == 1)
       GN_EXTERNAL = 1,         // written to GnScratch by the caller: getChangeInPose, no processNewImagePair gating
       GN_FLATTENED = 2 };      // written to GnScratch by launch_flatten_octaves (nOctaves > 1), gating as GN_FROM_TRACKED
// `assume_fast_kps` of launch_gn: 1 = the points are stage 2's keypoints (distinct, never 8-adjacent: with 1-px cells the
// aux NMS of S5:465-474 keeps them all), 0 = run the aux NMS; GN_MASK_READY is set internally once aux_nms_kernel ran
enum { GN_MASK_READY = 2 };
// nOctaves > 1: the tracked pairs of all octaves into G.prev / G.obs / G.cur_idx at the scale of octave 0
// (x 2^octave, stage5_optimization.cpp:422-456); PG = the Pools of the flattened problem (kp_cap = sum over the octaves)
cudaError_t launch_flatten_octaves(const Pools& PG, const GnScratch& G, const OctaveList& O, int n, cudaStream_t st, int* launches);
cudaError_t launch_gn(const Pools& P, const GnScratch& G, int n, int external, int use_custom_pose,
                      int assume_fast_kps, cudaStream_t st, int* launches, StepTimer* tm = nullptr);
}  // namespace svo
