// prepare.cuh -- stage 1 (grayscale conversion + stereo rectification), see prepare.cu
#pragma once
#include "svo_internal.cuh"

namespace svo {

struct PrepareArgs {
    const uint8_t* srcL; const uint8_t* srcR;   // source images of the batch, item i at + i * src_image_pitch
    size_t src_stride, src_image_pitch;         // bytes per source row / per source image
    uint8_t* dstL; uint8_t* dstR;               // gray (rectified) planes, item i at + i * dst_image_pitch
    size_t dst_image_pitch;
    int W, H, pitch_out;                        // pitch_out: row pitch of the output planes (multiple of 16)
    const short2* map_xy[2];                    // CV_16SC2 map of cv::initUndistortRectifyMap, per side (rectify only)
    const uint16_t* map_a[2];                   // CV_16UC1 interpolation-table indices, per side
    uint8_t* tmpL; uint8_t* tmpR;               // optional scratch planes (layout of dst) for colour + rectification
};

cudaError_t launch_prepare(const PrepareArgs& A, int n, int channels, bool rectify, cudaStream_t st, int* launches);

// One level of CImagePyramid::buildPyramidFast (stage1_rectify.cpp:82-83): (sw x sh) gray planes -> (sw>>1 x sh>>1).
struct PyramidArgs {
    const uint8_t* srcL; const uint8_t* srcR;   // item i at + i * src_image_pitch
    size_t src_stride, src_image_pitch;
    uint8_t* dstL; uint8_t* dstR;               // planes with a row pitch that is a multiple of 16
    size_t dst_stride, dst_image_pitch;
    int sw, sh;                                 // source size; destination = (sw >> 1) x (sh >> 1)
};
enum { PYR_MRPT_AUTO = 0, PYR_AVG2X2 = 1, PYR_PYRDOWN = 2 };    // svo_config.pyr_mode
// the method MRPT 1.x picks for a source of this width: its SSE2 2x2 average needs width % 16 == 0, else cvPyrDown
inline int pyramid_method(int pyr_mode, int src_width) {
    if (pyr_mode == PYR_MRPT_AUTO) return (src_width % 16 == 0) ? PYR_AVG2X2 : PYR_PYRDOWN;
    return pyr_mode;
}
cudaError_t launch_pyramid_level(const PyramidArgs& A, int n, int method, cudaStream_t st, int* launches);

}  // namespace svo
