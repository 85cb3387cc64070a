// Shared definitions of the cova_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cova_b200.h"

namespace cova {

// ------------------------------------------------------------------------------------------------
// error plumbing: nothing throws or aborts across the C ABI
// ------------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
inline int set_err(int code, const char *fmt, const char *a = "", const char *b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
#define COVA_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return cova::set_err(COVA_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Activation layout in HBM ("phase planes").
//
// A logical activation [n][t][c][y][x] (n windows, t in [0,Tn), c channels, H x W) is stored as fp16
// rows of 8 channels (16 bytes), one row per position, in 4 phase planes (a = y&1, b = x&1) per
// channel block:
//     row(cb, ph, pos)        = (cb*4 + ph) * Lp + pos                       [units of 16 bytes]
//     pos(n, y2, x2, t)       = guard + ((n*S + y2*P + x2) * Tn + t)         y2 = y>>1, x2 = x>>1
//     P = ceil(W/2) + 1, R = ceil(H/2) + 1, S = R*P
// Every phase plane carries one shared zero column (x2 = Wh) and one shared zero row (y2 = Hh), so
// that a spatial shift of the operand is a constant offset of `pos`: (dy2*P + dx2)*Tn.  A 3x3 "same"
// convolution followed by 2x2 max-pooling, and a stride-2 transposed convolution, both become sums of
// GEMMs over *contiguous* 128-position slices of these planes - no im2col.  T is interleaved
// innermost so that the four frames of a window sit in four adjacent TMEM lanes (PointWiseTN).
// Rows outside the logical extent are zero and are never written.
// ------------------------------------------------------------------------------------------------
struct Geom {
    int H, W;        // logical extent
    int Hh, Wh;      // phase-plane extent
    int P, R, S;     // pitch, rows, positions per (window, phase)
    int Tn;          // interleaved time planes: 4 (encoder side) or 1 (decoder side)
    int CB;          // channel blocks (8 channels each)
    int halo;        // (P+1)*Tn : largest |shift| any operand uses
    long long guard; // zero rows in front of position 0 (>= halo)
    long long Lp;    // rows per (cb, phase) plane
    int N;           // window capacity
};

constexpr int kTileM = 128;
constexpr int kGroupAlign = 8 * kTileM;  // plane length slack so that any tiles-per-stage <= 8 may over-read

inline Geom make_geom(int H, int W, int C, int Tn, int N) {
    Geom g;
    g.H = H; g.W = W;
    g.Hh = (H + 1) / 2; g.Wh = (W + 1) / 2;
    g.P = g.Wh + 1; g.R = g.Hh + 1; g.S = g.R * g.P;
    g.Tn = Tn; g.CB = (C + 7) / 8;
    g.halo = (g.P + 1) * Tn;
    g.guard = ((g.halo + 7) / 8) * 8;
    long long m = (long long)N * g.S * Tn;
    m = ((m + kGroupAlign - 1) / kGroupAlign) * kGroupAlign + kGroupAlign;   // a whole tile group may over-read
    g.Lp = g.guard + m + g.guard;
    g.N = N;
    return g;
}
__host__ __device__ inline long long geom_rows(const Geom &g) { return (long long)g.CB * 4 * g.Lp; }
__host__ __device__ inline long long geom_pos(const Geom &g, int n, int y2, int x2, int t) {
    return g.guard + ((long long)n * g.S + (long long)y2 * g.P + x2) * g.Tn + t;
}
__host__ __device__ inline long long geom_row(const Geom &g, int cb, int ph, long long pos) {
    return ((long long)cb * 4 + ph) * g.Lp + pos;
}
// row index (16-byte units) of logical element (n, t, cb, y, x)
__host__ __device__ inline long long geom_row_of(const Geom &g, int n, int t, int cb, int y, int x) {
    return geom_row(g, cb, ((y & 1) << 1) | (x & 1), geom_pos(g, n, y >> 1, x >> 1, t));
}

// BlobNet architecture constants (reference utils/train-blobnet.py:57-69)
constexpr int kT = 4;
constexpr int kEncCin[4] = {3, 16, 32, 64};
constexpr int kEncCout[4] = {16, 32, 64, 128};
constexpr int kDecCin[4] = {128, 128, 64, 32};
constexpr int kDecCout[4] = {64, 32, 16, 16};
constexpr float kBnEps = 1e-3f;  // Keras BatchNormalization default (reference encoder.py:45-48)

// packed fp32 pairs (SASS FFMA2 / FADD2 / FMUL2: two fp32 operations per issue slot; ptxas folds {x, x} pairs
// into broadcast operands)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the path is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, fires `launch_dependents` on entry and executes
// `pdl_wait()` (griddepcontrol.wait: the preceding kernel of the stream has completed and flushed) after its own
// prologue (barrier init, TMEM allocation, weights -> shared memory / TMEM) and before it touches any activation
// buffer.  A persistent CTA of layer k+1 therefore starts on an SM the moment layer k's CTA leaves it, and the
// launch latency and prologue of k+1 hide behind the tail of k.  EVERY thread of EVERY kernel in the chain waits:
// completion of k then implies completion of k-1, k-2, ... (buffers are re-used from step to step).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// `pdl` is a per-handle property (cova_pipeline_set_debug bit 5 clears it for that handle: plain stream-ordered launches)
constexpr int kDbgNoPdl = 32;

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// exact unsigned division by an invariant d >= 2 (Granlund-Montgomery, round-up variant)
struct FastDiv { uint32_t m, s; };
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) l++;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s = l - 1;
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t x, FastDiv f) {
    const uint32_t t = __umulhi(f.m, x);
    return (t + ((x - t) >> 1)) >> f.s;
}

struct alignas(16) Row8 {  // one 16-byte row: 8 fp16 channels
    __half v[8];
};

}  // namespace cova
