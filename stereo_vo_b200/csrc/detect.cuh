// detect.cuh -- launchers of the stage-2 kernels (see detect.cu)
#pragma once
#include "svo_internal.cuh"

namespace svo {
int detect_smem_bytes(const Params& p);
cudaError_t configure_detect(const Params& p);
cudaError_t launch_begin_step(const Pools& P, int n, const uint8_t* d_repeat, cudaStream_t st);
cudaError_t launch_detect(const Pools& P, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                          size_t image_pitch, int pitch_out, uint32_t* occ_scratch, int occ_words_per_img,
                          uint32_t* sort_scratch, uint2* list_scratch, uint32_t* key_scratch, cudaStream_t st, int* launches,
                          StepTimer* tm = nullptr);
}  // namespace svo
