// Register-resident FP64 peak micro-benchmarks: the roofline denominators of the step kernel
// ("FP64 DMMA peak to be measured by a register-resident DMMA.8x8x4 micro-benchmark").
#include "common.cuh"

namespace aceqd {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 8 independent accumulator chains per warp; 512 flop per DMMA.
__global__ void __launch_bounds__(1024) k_dmma_peak(double* sink, int iters) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

// 16 independent DFMA chains per thread; 2 flop per FMA per lane.
__global__ void __launch_bounds__(256) k_dfma_peak(double* sink, int iters) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * blockIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

int launch_fp64_peak(int kind, int iters, double* sink_dev, int* blocks, int* threads,
                     cudaStream_t s, LaunchLog* log) {
    int dev = 0, sms = 0;
    ACEQD_CUDA(cudaGetDevice(&dev));
    ACEQD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    *blocks = sms * 4;
    *threads = 256;
    if (kind >= 10) {   // occupancy probes: ONE block per SM with (kind - 10) warps per SM sub-partition
        *blocks = sms;
        *threads = 128 * (kind - 10);
        if (*threads < 128 || *threads > 1024) {
            set_error("aceqd_fp64_peak: kind %d out of range", kind);
            return ACEQD_ERR_ARG;
        }
        k_dmma_peak<<<*blocks, *threads, 0, s>>>(sink_dev, iters);
    } else if (kind == 0)
        k_dmma_peak<<<*blocks, *threads, 0, s>>>(sink_dev, iters);
    else
        k_dfma_peak<<<*blocks, *threads, 0, s>>>(sink_dev, iters);
    ++log->count;
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
