// Micro-benchmark 2: do cp.async.bulk copies issued back to back overlap?  One CTA; `lanes` lanes of warp 0 (mode 0) or
// lane 0 of `lanes` different warps (mode 1) each issue `per_lane` copies of `bytes` into distinct shared-memory
// destinations, all completing on ONE mbarrier; thread 0 measures clock64 from before the first issue until the
// barrier completes.  Prints cycles per configuration (median of 20 repetitions).
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(1024, 1) k_lat(const unsigned char* buf, int bytes, int lanes, int per_lane, int mode,
                                                 long long* out, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    unsigned char* dst = smem + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b = smem_u32(bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        long long t0 = clock64();
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * lanes * per_lane) : "memory");
        __syncthreads();
        const int me = mode == 0 ? (warp == 0 && lane < lanes ? lane : -1) : (lane == 0 && warp < lanes ? warp : -1);
        if (me >= 0) {
            for (int i = 0; i < per_lane; ++i) {
                const int k = me * per_lane + i;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(dst + (size_t)k * bytes)),
                             "l"(buf + ((size_t)(r * 64 + k) * bytes) % (4u << 20)), "r"(bytes), "r"(b)
                             : "memory");
            }
        }
        long long t1 = clock64();
        if (tid == 0) {
            asm volatile(
                "{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(b),
                "r"(phase)
                : "memory");
            long long t2 = clock64();
            out[2 * r] = t1 - t0;
            out[2 * r + 1] = t2 - t0;
        }
        phase ^= 1u;
    }
}

int main() {
    unsigned char* buf;
    long long* out;
    cudaMalloc(&buf, 8u << 20);
    cudaMemset(buf, 1, 8u << 20);
    cudaMalloc(&out, 1024);
    cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int reps = 21;
    printf("mode bytes lanes per_lane total_KB  issue_cyc  done_cyc  B/clk\n");
    for (int mode : {0, 1})
        for (int bytes : {1024, 4096, 16384})
            for (int lanes : {1, 2, 4, 8, 16})
                for (int per_lane : {1, 2, 4, 8}) {
                    const size_t total = (size_t)bytes * lanes * per_lane;
                    if (total > 192 * 1024) continue;
                    k_lat<<<1, 1024, 128 + total>>>(buf, bytes, lanes, per_lane, mode, out, reps);
                    if (cudaDeviceSynchronize() != cudaSuccess) {
                        printf("error %s\n", cudaGetErrorString(cudaGetLastError()));
                        return 1;
                    }
                    long long h[2 * reps];
                    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                    std::vector<long long> a, d;
                    for (int r = 1; r < reps; ++r) {
                        a.push_back(h[2 * r]);
                        d.push_back(h[2 * r + 1]);
                    }
                    std::sort(a.begin(), a.end());
                    std::sort(d.begin(), d.end());
                    printf("%d %6d %2d %2d %6.1f  %6lld  %6lld  %6.1f\n", mode, bytes, lanes, per_lane, total / 1024.0,
                           a[a.size() / 2], d[d.size() / 2], (double)total / d[d.size() / 2]);
                }
    return 0;
}
