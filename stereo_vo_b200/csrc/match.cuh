// match.cuh -- launchers of the stage-3 / stage-4 kernels (see match.cu)
#pragma once
#include "svo_internal.cuh"

namespace svo {
cudaError_t configure_match(const Params& p);
cudaError_t launch_match_lr(const Pools& P, int n, cudaStream_t st, int* launches, StepTimer* tm = nullptr);
cudaError_t launch_track_match(const Pools& P, int n, cudaStream_t st, int* launches, StepTimer* tm = nullptr);
cudaError_t launch_consistency(const Pools& P, int n, cudaStream_t st, int* launches, StepTimer* tm = nullptr);
// nOctaves > 1: the resetIds() renumbering across all octaves (process_new_image_pair.cpp:254-267), before stage 3
cudaError_t launch_reset_ids_octaves(const Pools& P, const OctaveList& O, int n, cudaStream_t st, int* launches);
cudaError_t launch_hamming_raw(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int reps, uint32_t* d_out,
                               cudaStream_t st);
}  // namespace svo
