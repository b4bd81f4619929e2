// Micro-benchmark: how fast can ONE CTA per SM stream an L2-resident buffer into shared memory with cp.async.bulk
// (the step kernels' PT chunk ring)?  Sweeps bytes per copy, ring depth and the number of issuing lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_ingest.bin l2_ingest.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// warp 0 = producer (lanes 0..split-1 issue one piece each), warps 1..8 = consumers that touch a few words of
// every chunk and release it (consumer cost ~0: this measures the ingest path alone)
__global__ void __launch_bounds__(288, 1) k_ingest(const double* buf, size_t buf_bytes, int chunk_bytes, int stages, int split,
                                                   int n_chunks, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* ring = smem + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full = smem_u32(bars), empty = smem_u32(bars + 16);
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t n_in_buf = buf_bytes / chunk_bytes;
    double acc = 0.0;
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int piece = chunk_bytes / split;
        for (int i = 0; i < n_chunks; ++i) {
            const size_t src = ((size_t)blockIdx.x * 7919 + i) % n_in_buf * chunk_bytes;
            if (lane == 0) {
                mbar_wait(empty + 8 * stage, phase ^ 1u);
                mbar_expect_tx(full + 8 * stage, chunk_bytes);
            }
            __syncwarp();
            if (lane < split)
                bulk_g2s(smem_u32(ring + (size_t)stage * chunk_bytes + (size_t)lane * piece),
                         reinterpret_cast<const unsigned char*>(buf) + src + (size_t)lane * piece, piece, full + 8 * stage);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_chunks; ++i) {
            mbar_wait(full + 8 * stage, phase);
            acc += reinterpret_cast<const double*>(ring + (size_t)stage * chunk_bytes)[lane + 32 * (warp - 1)];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + 8 * stage);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}

int main() {
    const size_t buf_bytes = 4u << 20;   // 4 MB: one chi=128 PT slice (9 blocks) is 2.4 MB -> L2 resident
    double *buf, *sink;
    cudaMalloc(&buf, buf_bytes);
    cudaMalloc(&sink, 64);
    cudaMemset(buf, 0, buf_bytes);
    cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    printf("chunk_bytes stages split grid  GB/s_per_SM  B/clk/SM(1.965GHz)  TB/s_total\n");
    const int chunks[] = {2304, 4608, 8704, 16896, 33792};
    for (int grid : {n_sm, 74, 8})
        for (int cb : chunks)
            for (int stages : {2, 3, 4, 6})
                for (int split : {1, 4}) {
                    if ((size_t)cb * stages + 256 > 200 * 1024) continue;
                    if ((cb / split) % 16) continue;
                    const int n_chunks = (int)((64u << 20) / cb);   // 64 MB per CTA
                    const size_t smem = 256 + (size_t)cb * stages;
                    k_ingest<<<grid, 288, smem>>>(buf, buf_bytes, cb, stages, split, n_chunks / 8, sink);
                    cudaEventRecord(e0);
                    k_ingest<<<grid, 288, smem>>>(buf, buf_bytes, cb, stages, split, n_chunks, sink);
                    cudaEventRecord(e1);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("error: %s\n", cudaGetErrorString(e));
                        return 1;
                    }
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    const double per_sm = (double)cb * n_chunks / (ms * 1e-3) / 1e9;
                    printf("%6d %2d %d %3d  %8.1f  %6.1f  %6.2f\n", cb, stages, split, grid, per_sm, per_sm / 1.965,
                           per_sm * grid / 1e3);
                }
    return 0;
}
