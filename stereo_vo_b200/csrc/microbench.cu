// microbench.cu -- register-only issue-rate probes for the integer pipes that bound the Hamming kernels
// ("measure the POPC/LOP3 peak with a register-only microbench on the same box").
// Not on the product path; used by bench.py / tools to state the match stage's binding roof from measurement.
#include "microbench.cuh"

namespace svo {

constexpr int MB_THREADS = 256;
constexpr int MB_ITERS = 4096;

__global__ void __launch_bounds__(MB_THREADS) popc_probe_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
#pragma unroll 1
    for (int i = 0; i < MB_ITERS; ++i) {
        a0 = __popc(a0) + seed; a1 = __popc(a1) + seed; a2 = __popc(a2) + seed; a3 = __popc(a3) + seed;
        a4 = __popc(a4) + seed; a5 = __popc(a5) + seed; a6 = __popc(a6) + seed; a7 = __popc(a7) + seed;
    }
    out[blockIdx.x * MB_THREADS + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

__global__ void __launch_bounds__(MB_THREADS) lop3_probe_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    const uint32_t b = seed * 0x9E3779B9u, c = ~seed;
#pragma unroll 1
    for (int i = 0; i < MB_ITERS; ++i) {
        a0 = (a0 & a1) ^ b; a1 = (a1 & a2) ^ c; a2 = (a2 & a3) ^ b; a3 = (a3 & a4) ^ c;
        a4 = (a4 & a5) ^ b; a5 = (a5 & a6) ^ c; a6 = (a6 & a7) ^ b; a7 = (a7 & a0) ^ c;
    }
    out[blockIdx.x * MB_THREADS + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

// ops per second (lane operations) of POPC and of a dependent-free LOP3 stream
cudaError_t measure_pipe_peaks(double* popc_per_s, double* lop3_per_s, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8;
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, (size_t)blocks * MB_THREADS * 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms_p = 0, ms_l = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        popc_probe_kernel<<<blocks, MB_THREADS, 0, st>>>(d, 12345u + rep);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_p, e0, e1);
        cudaEventRecord(e0, st);
        lop3_probe_kernel<<<blocks, MB_THREADS, 0, st>>>(d, 12345u + rep);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_l, e0, e1);
    }
    const double ops = (double)blocks * MB_THREADS * MB_ITERS * 8.0;
    *popc_per_s = ops / (ms_p * 1e-3);
    *lop3_per_s = ops / (ms_l * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return cudaGetLastError();
}

}  // namespace svo
