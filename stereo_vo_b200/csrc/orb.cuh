// orb.cuh -- device side of the dmORB detector (cv::ORB::detectAndCompute, stage2_detect.cpp:458-496): declarations.
#pragma once
#include <vector>
#include "svo_internal.cuh"
#include "orb_core.cuh"

namespace svo {

struct OrbLevelDev {
    int32_t w, h, pitch, nfeat;       // level size, row pitch of its planes (multiple of 16), nfeaturesPerLevel
    float scale;                      // layerScale
    uint8_t* imgL;                    // [max_seqs][pitch*h]   level image per batch item (levels >= 1; level 0 = the caller's image)
    uint8_t* imgR;
    uint8_t* blur;                    // [max_seqs*2][pitch*h] GaussianBlur(7x7, 2) of the level, per (slot, side)
    uint16_t* tile_seg;               // FAST corners of the level as fast_blur_kernel's tile lists (svo_internal.cuh)
    uint16_t* tile_list;
    const uint16_t *xofs, *xc1, *yofs, *yc1;   // INTER_LINEAR_EXACT tables from level - 1 (left tap, weight of the right tap / 256)
};

struct OrbDev {
    int32_t n_levels;
    int32_t n_extract;                // nfeatures handed to cv::ORB::create (S2:460-464)
    int32_t cap;                      // candidates per level and image (= max_corners)
    int32_t raw_cap;                  // keypoints per image that detectAndCompute may return
    int32_t lvl_cap;                  // ... per level
    OrbLevelDev L[kOrbMaxLevels];
    // per (slot, side, level) scratch of one step
    uint32_t* idx;                    // [..][cap]     the keypoint vector of the level being filtered (indices)
    float* resp;                      // [..][cap]     response per candidate: FAST score, then Harris
    uint32_t* pos;                    // [..][cap]     y<<16 | x in level coordinates
    uint32_t* tmp;                    // [..][cap]     partition scratch
    float4* lraw;                     // [..][lvl_cap] the level's keypoints {pt.x, pt.y, response, angle}
    int32_t* n_lraw;                  // [..]
    // per (slot, side)
    float4* raw;                      // [max_seqs*2][raw_cap] detectAndCompute's keypoints {pt.x, pt.y, response, angle}
    int32_t* raw_oct;                 // [max_seqs*2][raw_cap] octave
    int32_t* n_raw;                   // [max_seqs*2]
};

cudaError_t alloc_orb(OrbDev& D, const Params& p, int n_slots, int n_levels, int n_extract, std::vector<void*>& allocs);
// detectAndCompute + m_non_max_sup + m_update_indexes for the batch items [0, n) of P.slots: fills P.kpf / P.kpo / P.kp / P.desc /
// P.n_kp of the current frame of every slot.  d_left / d_right: the caller's images (by batch item).
cudaError_t launch_orb_detect(const Pools& P, const OrbDev& D, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                              size_t image_pitch, int item0, uint32_t* occ_scratch, int occ_words_per_img, uint32_t* sort_scratch,
                              uint2* list_scratch, cudaStream_t st, int* launches);
// detect.cu: fast_blur_kernel alone on the images of one level (blur plane + corner lists of PL)
cudaError_t launch_fast_blur_only(const Pools& PL, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                                  size_t image_pitch, int pitch_out, cudaStream_t st);

}  // namespace svo
