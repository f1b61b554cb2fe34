// microbench.cuh -- integer-pipe issue-rate probes (see microbench.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace svo {
cudaError_t measure_pipe_peaks(double* popc_per_s, double* lop3_per_s, cudaStream_t st);
}
