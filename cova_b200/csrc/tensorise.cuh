// Tensorisation kernels: decoder metadata quads -> (a) the stacked RGBA window image the reference's
// metapreprocess element emits, (b) the BlobNet input in the phase-plane fp16 layout of common.cuh.
//
// Reference: cova-rs/gst-plugins/src/metapreprocess/imp.rs:307-320 (out[0:S] = current frame,
// out[kS:(k+1)S] = frame t-k) and the nvinfer pre-process + Reshape that follow it
// (config/blobnet/*.txt:7,9 - RGBA -> planar RGB, scale 1, byte 3 dropped;
//  utils/train-blobnet.py:113-116 - (3, T*H, W) -> (3, T, H, W)); utils/model/preprocessing.py:5-8
// (clip(x, 0, 6); the 1/6 is folded into the first convolution's weights).
#pragma once
#include "common.cuh"

namespace cova {

// (a) stacked RGBA: pure gather/copy, HBM bound.  One element = one 32-bit macroblock quad, or one
// 128-bit vector of four quads when the frame size allows it.
template <typename V>
__global__ void __launch_bounds__(256) stack_rgba_kernel(const V *__restrict__ frames, const int *__restrict__ newest,
                                                         V *__restrict__ out, int vec_per_frame, int T,
                                                         long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int e = (int)(i % vec_per_frame);
        long long r = i / vec_per_frame;
        int k = (int)(r % T);
        int n = (int)(r / T);
        long long src = (long long)(newest[n] - k) * vec_per_frame + e;
        out[i] = __ldg(frames + src);
    }
}

// (b) BlobNet input layout.  Thread <-> (window n, row y, column pair x2, time t); the four t of one
// (y, x2) are adjacent lanes so that the 16-byte rows they write are contiguous (64 B), and a warp
// reads 8 consecutive 8-byte pixel pairs from each of the 4 frames.
__global__ void __launch_bounds__(256) tensorise_x0_kernel(const uint32_t *__restrict__ frames,
                                                           const int *__restrict__ newest, uint4 *__restrict__ x0,
                                                           Geom g, int n_windows) {
    const int Wh = g.Wh;
    const long long total = (long long)n_windows * g.H * Wh * kT;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int px_per_frame = g.H * g.W;
    for (; i < total; i += stride) {
        int t = (int)(i & 3);
        long long r = i >> 2;
        int x2 = (int)(r % Wh);
        r /= Wh;
        int y = (int)(r % g.H);
        int n = (int)(r / g.H);
        const uint32_t *f = frames + (long long)(newest[n] - t) * px_per_frame + (long long)y * g.W;
        int x = 2 * x2;
        uint32_t q0 = __ldg(f + x);
        uint32_t q1 = (x + 1 < g.W) ? __ldg(f + x + 1) : 0u;
        long long pos = geom_pos(g, n, y >> 1, x2, t);
        int ph = (y & 1) << 1;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            if (b == 1 && x + 1 >= g.W) break;
            uint32_t q = b ? q1 : q0;
            // bytes: 0 mb_weight, 1 |mv_x|, 2 |mv_y|, 3 stale (dropped).  clip(., 0, 6) of a u8 = min(., 6)
            __half2 c01 = __floats2half2_rn((float)min(q & 0xffu, 6u), (float)min((q >> 8) & 0xffu, 6u));
            __half2 c2z = __floats2half2_rn((float)min((q >> 16) & 0xffu, 6u), 0.f);
            uint4 row;
            row.x = *reinterpret_cast<uint32_t *>(&c01);
            row.y = *reinterpret_cast<uint32_t *>(&c2z);
            row.z = 0u;
            row.w = 0u;
            x0[geom_row(g, 0, ph | b, pos)] = row;
        }
    }
}

}  // namespace cova
