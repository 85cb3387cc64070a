// Micro-benchmark: cycles per tcgen05.mma (M=128 or 64, K=16, fp16) as a function of N, operand layout
// (no-swizzle / 128B swizzle descriptors), A source (smem / TMEM) and accumulator reuse.
// Development aid: informs the tile shapes of csrc/blobnet_tc.cuh.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
struct Args { int n, m, layout, a_tmem, n_acc, iters, a_stride_rows; long long *out; };

__global__ void __launch_bounds__(128, 1) k(Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    const uint32_t barp = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barp));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tslot;
    if (threadIdx.x < 32 && elect_one()) {
        const uint32_t idesc = make_idesc(a.m, a.n);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 64 * 1024;
        // no-swizzle: LBO = rows*16 (next 8-channel block), SBO = 128.  128B swizzle: LBO unused(1), SBO = 1024
        const uint64_t adesc0 = a.layout == 0 ? make_desc(sA, 4096 * 4, 128, 0) : make_desc(sA, 16, 1024, 2);
        const uint64_t bdesc = a.layout == 0 ? make_desc(sB, (uint32_t)a.n * 16, 128, 0) : make_desc(sB, 16, 1024, 2);
        long long t0 = clock64();
        // 8 descriptors precomputed (shifted starts like the taps of a conv); the loop body is 8 back-to-back MMAs
        uint64_t ad[8];
        uint32_t dd[8];
        for (int j = 0; j < 8; j++) { ad[j] = adesc0 + (uint64_t)(j * a.a_stride_rows); dd[j] = tm + (uint32_t)((j % a.n_acc) * a.n); }
        for (int it = 0; it < a.iters; it += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (a.a_tmem) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dd[j]), "r"(tm + 448u), "l"(bdesc), "r"(idesc), "r"(1u));
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd[j]), "l"(ad[j]), "l"(bdesc), "r"(idesc), "r"(1u));
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp));
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(barp), "r"(0u));
        long long t1 = clock64();
        if (blockIdx.x == 0) a.out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
    long long *d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("%-6s %-4s %-8s %-7s %-6s %-8s %10s\n", "M", "N", "layout", "A", "n_acc", "a_shift", "cyc/mma");
    for (int grid : {1, 148})
        for (int m : {128, 64})
            for (int layout : {0, 2})
                for (int a_tmem : {0, 1})
                    for (int n_acc : {1, 2})
                        for (int shift : {0, 1})
                            for (int n : {16, 32, 64, 128, 256}) {
                                if (n * n_acc > 448) continue;
                                if (grid == 148 && (n_acc != 1 || shift != 1)) continue;
                                if (a_tmem && shift) continue;
                                Args a{n, m, layout, a_tmem, n_acc, iters, shift ? 4 : 0, d};
                                k<<<grid, 128, 200 * 1024>>>(a);
                                cudaError_t e = cudaDeviceSynchronize();
                                if (e != cudaSuccess) { printf("error %s (m=%d n=%d layout=%d a_tmem=%d)\n", cudaGetErrorString(e), m, n, layout, a_tmem); return 1; }
                                long long c;
                                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                                printf("%-6d %-4d %-8s %-7s %-6d %-8d %10.1f  grid=%d\n", m, n, layout ? "sw128" : "none", a_tmem ? "tmem" : "smem", n_acc, shift, (double)c / iters, grid);
                            }
    return 0;
}
