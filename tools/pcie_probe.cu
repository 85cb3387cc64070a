// Development aid: which way of bringing 123 MB of host frames to HBM is fastest on this box?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pcie_probe tools/pcie_probe.cu
//   copy engine (cudaMemcpyAsync) from default / write-combined pinned memory, split over 1..4 streams,
//   and a zero-copy kernel (SMs read mapped pinned memory directly).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void zero_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

int main() {
    const size_t n = 123494400;
    void *h_def, *h_wc, *d, *d2, *h_out;
    CK(cudaHostAlloc(&h_def, n, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_wc, n, cudaHostAllocMapped | cudaHostAllocWriteCombined));
    CK(cudaHostAlloc(&h_out, n, cudaHostAllocDefault));
    CK(cudaMalloc(&d, n));
    CK(cudaMalloc(&d2, n));
    memset(h_def, 1, n);
    memset(h_wc, 2, n);
    cudaStream_t st[4], so;
    for (auto &s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 10;
    for (int wc = 0; wc < 2; wc++) {
        const char *src = (const char *)(wc ? h_wc : h_def);
        for (int with_d2h = 0; with_d2h < 2; with_d2h++)
            for (int ns = 1; ns <= 4; ns *= 2) {
                float best = 1e9f;
                for (int r = 0; r < reps; r++) {
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0, st[0]));
                    for (int k = 1; k < ns; k++) CK(cudaStreamWaitEvent(st[k], e0, 0));
                    size_t part = (n / ns + 15) & ~size_t(15);
                    for (int k = 0; k < ns; k++) {
                        size_t off = k * part, len = off + part > n ? n - off : part;
                        CK(cudaMemcpyAsync((char *)d + off, src + off, len, cudaMemcpyHostToDevice, st[k]));
                    }
                    if (with_d2h) CK(cudaMemcpyAsync(h_out, d2, n / 3, cudaMemcpyDeviceToHost, so));
                    for (int k = 1; k < ns; k++) {
                        cudaEvent_t ek; CK(cudaEventCreateWithFlags(&ek, cudaEventDisableTiming));
                        CK(cudaEventRecord(ek, st[k])); CK(cudaStreamWaitEvent(st[0], ek, 0)); CK(cudaEventDestroy(ek));
                    }
                    CK(cudaEventRecord(e1, st[0]));
                    CK(cudaDeviceSynchronize());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                printf("memcpy  %-7s streams=%d d2h=%d : %.3f ms  %.1f GB/s\n", wc ? "WC" : "default", ns, with_d2h, best, n / best / 1e6);
            }
        for (int grid = 148; grid <= 148 * 16; grid *= 2) {
            const void *dp;
            CK(cudaHostGetDevicePointer((void **)&dp, (void *)src, 0));
            float best = 1e9f;
            for (int r = 0; r < reps; r++) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, st[0]));
                zero_copy_kernel<<<grid, 256, 0, st[0]>>>((const uint4 *)dp, (uint4 *)d, n / 16);
                CK(cudaEventRecord(e1, st[0]));
                CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            printf("kernel  %-7s grid=%-5d       : %.3f ms  %.1f GB/s\n", wc ? "WC" : "default", grid, best, n / best / 1e6);
        }
    }
    return 0;
}
