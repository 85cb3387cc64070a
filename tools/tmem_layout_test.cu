// Development aid: empirical register <-> (TMEM lane, column) mapping of tcgen05.ld.16x256b, checked against the
// layout csrc/blobnet_enc.cuh assumes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout_test tmem_layout_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(128, 1) k(uint32_t *out) {
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tslot;
    const uint32_t row = warp * 32 + lane;
    for (int c0 = 0; c0 < 16; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; j++) v[j] = row * 1000 + c0 + j;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(tm + ((uint32_t)(warp * 32) << 16) + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int half = 0; half < 2; half++) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tm + ((uint32_t)(warp * 32 + half * 16) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) out[((warp * 2 + half) * 32 + lane) * 8 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32));
}
int main() {
    uint32_t *d, h[4 * 2 * 32 * 8];
    cudaMalloc(&d, sizeof(h));
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int w = 0; w < 4; w++) for (int half = 0; half < 2; half++) for (int i = 0; i < 32; i++) for (int j = 0; j < 8; j++) {
        // assumed: reg 4k + {0,1} = row i/4, columns 8k + 2*(i%4) + {0,1}; reg 4k + {2,3} = row i/4 + 8
        const int kblk = j >> 2, rr = (j >> 1) & 1, cc = j & 1;
        const uint32_t want = (uint32_t)(w * 32 + half * 16 + i / 4 + 8 * rr) * 1000 + 8 * kblk + 2 * (i % 4) + cc;
        const uint32_t got = h[((w * 2 + half) * 32 + i) * 8 + j];
        if (got != want && bad++ < 20) printf("w%d half%d lane%d reg%d: got %u want %u\n", w, half, i, j, got, want);
    }
    printf("tcgen05.ld.16x256b layout: %s (%d mismatches)\n", bad ? "DIFFERENT" : "as assumed", bad);
    return 0;
}
