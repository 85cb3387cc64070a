// C ABI of libcova_b200.so (see include/cova_b200.h).  One translation unit: kernels + host plumbing.
// There is NO CPU fallback anywhere in this file: without a CUDA device every constructor fails with
// COVA_E_NODEVICE.
#include <ctype.h>
#include <math.h>
#include <sched.h>

#include <algorithm>
#include <mutex>

// The fp32 CUDA-core validation kernels (COVA_IMPL_SIMT: layer-by-layer comparison partner of the tcgen05 kernels in the
// tests) are only compiled into libcova_b200_val.so (-DCOVA_VALIDATION); the shipped library holds the hot path alone.
#ifdef COVA_VALIDATION
#include "blobnet_simt.cuh"
#endif
#include "blobnet_tc.cuh"
#include "blobnet_enc.cuh"
#include "blobnet_enc1.cuh"
#include "ccl.cuh"
#include "common.cuh"
#include "tensorise.cuh"
#include "weights_pack.cuh"

namespace cova {
thread_local char g_err[512] = "";

static int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return set_err(COVA_E_NODEVICE, "no CUDA device (%s); cova_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_err(COVA_E_INVAL, "device index out of range");
    COVA_CUDA(cudaSetDevice(device));
    return COVA_OK;
}

// newest-frame ring used by the metapreprocess element: slot of frame t-k = (head - k) mod T
__global__ void __launch_bounds__(256) stack_ring_kernel(const uint32_t *__restrict__ ring, uint32_t *__restrict__ out,
                                                         int words_per_frame, int T, int head) {
    const int total = words_per_frame * T;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int k = i / words_per_frame, e = i - k * words_per_frame;
        int slot = (head - k + T) % T;
        out[i] = ring[(size_t)slot * words_per_frame + e];
    }
}
}  // namespace cova

using namespace cova;

// =================================================================================================
// misc
// =================================================================================================
extern "C" const char *cova_version(void) { return "cova_b200 0.1 (sm_100a)"; }
extern "C" const char *cova_last_error(void) { return g_err; }
extern "C" const char *cova_strerror(int code) {
    switch (code) {
        case COVA_OK: return "ok";
        case COVA_DROPPED: return "dropped (no output for this input)";
        case COVA_E_INVAL: return "invalid argument";
        case COVA_E_CUDA: return "CUDA error";
        case COVA_E_NOMEM: return "out of memory";
        case COVA_E_TOOSMALL: return "output buffer too small";
        case COVA_E_WEIGHTS: return "bad weight container";
        case COVA_E_UNSUPPORTED: return "unsupported configuration";
        case COVA_E_NODEVICE: return "no CUDA device (no CPU fallback)";
        case COVA_E_NUMERIC: return "numerical failure in the host tracker";
        case COVA_E_STATE: return "inconsistent frame-selection state";
        default: return "unknown error";
    }
}
extern "C" int cova_host_alloc(void **out, size_t bytes) {
    if (!out || !bytes) return set_err(COVA_E_INVAL, "null argument");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(e == cudaErrorMemoryAllocation ? COVA_E_NOMEM : COVA_E_NODEVICE, "cudaMallocHost: %s", cudaGetErrorString(e)); }
    return COVA_OK;
}
extern "C" void cova_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}
// One process per GPU (DESIGN.md "Multi-GPU"): the frames of a rank travel host -> device at PCIe rate, so its pinned
// buffers must live on the NUMA node the GPU hangs off - with 8 ranks on a two-socket box the inter-socket link is
// otherwise shared by up to 4 x 55 GB/s of remote reads.  Restricting the calling thread to the GPU's local CPUs
// (sysfs local_cpulist) before it allocates makes first-touch placement do that without libnuma.
extern "C" int cova_bind_host_to_device(int device, int *numa_node, int *n_cpus) {
    if (numa_node) *numa_node = -1;
    if (n_cpus) *n_cpus = 0;
    char bus[32] = "";
    cudaError_t e = cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(COVA_E_NODEVICE, "cudaDeviceGetPCIBusId: %s", cudaGetErrorString(e)); }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    const std::string base = std::string("/sys/bus/pci/devices/") + bus;
    if (FILE *f = fopen((base + "/numa_node").c_str(), "r")) {
        int node = -1;
        if (fscanf(f, "%d", &node) == 1 && numa_node) *numa_node = node;
        fclose(f);
    }
    FILE *f = fopen((base + "/local_cpulist").c_str(), "r");
    if (!f) return COVA_OK;                                   // no topology information: leave the affinity alone
    char list[4096] = "";
    const bool got = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!got) return COVA_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    int count = 0;
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); count++; }
    }
    if (!count) return COVA_OK;
    cpu_set_t cur;
    if (sched_getaffinity(0, sizeof(cur), &cur) == 0) {        // stay inside what the launcher (cgroup, taskset) allows
        cpu_set_t both;
        CPU_AND(&both, &set, &cur);
        if (CPU_COUNT(&both) == 0) return COVA_OK;
        set = both;
        count = CPU_COUNT(&both);
    }
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return COVA_OK;
    if (n_cpus) *n_cpus = count;
    return COVA_OK;
}
extern "C" int cova_device_count(int *n) {
    if (!n) return set_err(COVA_E_INVAL, "null argument");
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
    *n = c;
    return COVA_OK;
}

// =================================================================================================
// metapreprocess element
// =================================================================================================
struct cova_metapreprocess {
    int device;
    uint32_t w_mb, h_mb, timestep, gamma, gamma_idx, n_prev;
    int head;             // ring slot of the newest stored frame
    size_t S;             // size_per_buf (imp.rs:233)
    uint8_t *d_ring = nullptr, *d_out = nullptr;
    cudaStream_t stream = nullptr;
};

extern "C" int cova_metapreprocess_new(cova_metapreprocess **out, int device, uint32_t width_px, uint32_t height_px,
                                       uint32_t timestep, uint32_t gamma) {
    if (!out) return set_err(COVA_E_INVAL, "null out");
    *out = nullptr;
    if (timestep < 1 || gamma < 1) return set_err(COVA_E_INVAL, "timestep and gamma are u32 >= 1");
    if (width_px / 16 == 0 || height_px / 16 == 0) return set_err(COVA_E_INVAL, "frame smaller than one macroblock");
    int rc = check_device(device);
    if (rc) return rc;
    auto *mp = new (std::nothrow) cova_metapreprocess();
    if (!mp) return set_err(COVA_E_NOMEM, "host allocation failed");
    mp->device = device;
    mp->w_mb = width_px / 16;   // imp.rs:262-268: integer division
    mp->h_mb = height_px / 16;
    mp->timestep = timestep; mp->gamma = gamma; mp->gamma_idx = 0; mp->n_prev = 0; mp->head = -1;
    mp->S = (size_t)mp->w_mb * mp->h_mb * 4;
    cudaError_t e = cudaMalloc(&mp->d_ring, mp->S * timestep);
    if (e == cudaSuccess) e = cudaMalloc(&mp->d_out, mp->S * timestep);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&mp->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        cova_metapreprocess_free(mp);
        return set_err(COVA_E_CUDA, "metapreprocess allocation: %s", cudaGetErrorString(e));
    }
    *out = mp;
    return COVA_OK;
}
extern "C" void cova_metapreprocess_free(cova_metapreprocess *mp) {
    if (!mp) return;
    cudaSetDevice(mp->device);
    if (mp->d_ring) cudaFree(mp->d_ring);
    if (mp->d_out) cudaFree(mp->d_out);
    if (mp->stream) cudaStreamDestroy(mp->stream);
    delete mp;
}
extern "C" int cova_metapreprocess_set_gamma(cova_metapreprocess *mp, uint32_t gamma) {
    if (!mp || gamma < 1) return set_err(COVA_E_INVAL, "gamma is a u32 >= 1");
    mp->gamma = gamma;
    return COVA_OK;
}
extern "C" int cova_metapreprocess_out_caps(const cova_metapreprocess *mp, uint32_t *width, uint32_t *height, size_t *size) {
    if (!mp) return set_err(COVA_E_INVAL, "null handle");
    if (width) *width = mp->w_mb;
    if (height) *height = mp->h_mb * mp->timestep;
    if (size) *size = mp->S * mp->timestep;
    return COVA_OK;
}
extern "C" int cova_metapreprocess_transform(cova_metapreprocess *mp, const uint8_t *inbuf, size_t in_len, uint8_t *outbuf,
                                             size_t out_cap) {
    if (!mp || !inbuf) return set_err(COVA_E_INVAL, "null argument");
    if (in_len < mp->S) return set_err(COVA_E_INVAL, "input buffer shorter than size_per_buf");
    COVA_CUDA(cudaSetDevice(mp->device));
    const int T = (int)mp->timestep;
    // every branch of imp.rs:302-330 stores the incoming buffer as the newest one
    const int slot = (mp->head + 1) % T;
    COVA_CUDA(cudaMemcpyAsync(mp->d_ring + (size_t)slot * mp->S, inbuf, mp->S, cudaMemcpyHostToDevice, mp->stream));
    mp->head = slot;
    if (mp->n_prev < mp->timestep - 1) {          // imp.rs:302-305
        mp->n_prev++;
        COVA_CUDA(cudaStreamSynchronize(mp->stream));
        return COVA_DROPPED;
    }
    if (mp->gamma_idx != 0) {                      // imp.rs:325-330
        mp->gamma_idx--;
        COVA_CUDA(cudaStreamSynchronize(mp->stream));
        return COVA_DROPPED;
    }
    if (!outbuf || out_cap < mp->S * T) {
        // keep element state consistent with "this buffer was consumed", but tell the caller
        mp->gamma_idx = mp->gamma - 1;
        cudaStreamSynchronize(mp->stream);
        return set_err(COVA_E_TOOSMALL, "output buffer smaller than the RGBA stack");
    }
    const int words = (int)(mp->S / 4);
    int blocks = std::min(1024, (words * T + 255) / 256);
    stack_ring_kernel<<<blocks, 256, 0, mp->stream>>>(reinterpret_cast<const uint32_t *>(mp->d_ring),
                                                      reinterpret_cast<uint32_t *>(mp->d_out), words, T, mp->head);
    COVA_CUDA(cudaGetLastError());
    COVA_CUDA(cudaMemcpyAsync(outbuf, mp->d_out, mp->S * T, cudaMemcpyDeviceToHost, mp->stream));
    COVA_CUDA(cudaStreamSynchronize(mp->stream));
    mp->gamma_idx = mp->gamma - 1;                 // imp.rs:323
    return COVA_OK;
}

// =================================================================================================
// CCL launch shared by the bboxcc element and the pipeline
// =================================================================================================
struct CclBuffers {
    int H = 0, W = 0, nbx = 0, nby = 0, nb = 0, max_masks = 0;
    uint8_t *d_masks = nullptr, *d_blob = nullptr;
    size_t blob_cap = 0;
    unsigned long long *d_cursor = nullptr, *d_offsets = nullptr, *d_lens = nullptr;
    int32_t *d_labels = nullptr, *d_stats = nullptr, *d_nlabels = nullptr;
    size_t smem = 0;
    int threads = 0;
    bool compact = false;     // 13-byte-per-block shared-memory layout (ccl.cuh): two CTAs per SM at 4K
};

constexpr int kCclSmemLimit = tc::kSmemLimit - 1024;   // dynamic part: the kernel also has a few bytes of static shared memory

static int ccl_alloc(CclBuffers &b, int H, int W, int max_masks, bool own_masks) {
    b.H = H; b.W = W; b.nbx = (W + 1) / 2; b.nby = (H + 1) / 2; b.nb = b.nbx * b.nby; b.max_masks = max_masks;
    b.threads = ccl_threads_for(b.nb);
    // The compact layout would give 4K masks two CTAs per SM (204 KB -> 106 KB), but measured it does not pay: its
    // compare-and-swap statistics cost more than the second CTA brings (4K network masks 0.40 -> 0.44 ms per 1024, dense
    // 0.83 -> 0.86; only empty / all-ones masks gain).  It is used where the wide layout does not fit at all.
    b.compact = false;
    static const char *force = getenv("COVA_CCL_COMPACT");         // development knob: "0" / "1"
    if (force && (force[0] == '0' || force[0] == '1')) b.compact = force[0] == '1';
    b.smem = ccl_smem_bytes(b.nb, b.threads, b.compact);
    if (b.smem > (size_t)kCclSmemLimit && !b.compact) { b.compact = true; b.smem = ccl_smem_bytes(b.nb, b.threads, true); }
    if (b.smem > (size_t)kCclSmemLimit || b.nb >= 0x7FFF)       // block indices travel in 15 bits of the foreground lists
        return set_err(COVA_E_UNSUPPORTED, "mask grid too large for the shared-memory CCL kernel");
    b.blob_cap = (size_t)max_masks * (8 + 24 * (size_t)b.nb);
    if (own_masks) COVA_CUDA(cudaMalloc(&b.d_masks, (size_t)max_masks * H * W));
    COVA_CUDA(cudaMalloc(&b.d_blob, b.blob_cap));
    COVA_CUDA(cudaMalloc(&b.d_cursor, 2 * sizeof(unsigned long long)));
    COVA_CUDA(cudaMalloc(&b.d_offsets, (size_t)max_masks * sizeof(unsigned long long)));
    COVA_CUDA(cudaMalloc(&b.d_lens, (size_t)max_masks * sizeof(unsigned long long)));
    return COVA_OK;
}
static void ccl_free(CclBuffers &b, bool own_masks) {
    if (own_masks && b.d_masks) cudaFree(b.d_masks);
    if (b.d_blob) cudaFree(b.d_blob);
    if (b.d_cursor) cudaFree(b.d_cursor);
    if (b.d_offsets) cudaFree(b.d_offsets);
    if (b.d_lens) cudaFree(b.d_lens);
    if (b.d_labels) cudaFree(b.d_labels);
    if (b.d_stats) cudaFree(b.d_stats);
    if (b.d_nlabels) cudaFree(b.d_nlabels);
}
static int ccl_launch(CclBuffers &b, const uint8_t *d_masks, int n, uint32_t cc_threshold, bool want_labels, cudaStream_t st,
                      size_t first = 0, bool reset_cursor = true, bool pdl = true) {
    if (n <= 0) return COVA_OK;
    CclArgs a;
    a.masks = d_masks + first * (size_t)b.H * b.W; a.H = b.H; a.W = b.W; a.nbx = b.nbx; a.nby = b.nby;
    a.area_thresh = (int)cc_threshold;   // `settings.cc_threshold as i32` (imp.rs:248)
    a.blob = b.d_blob; a.blob_cap = b.blob_cap; a.cursor = b.d_cursor; a.offsets = b.d_offsets + first; a.lens = b.d_lens + first;
    a.labels = want_labels ? b.d_labels : nullptr;
    a.stats = want_labels ? b.d_stats : nullptr;
    a.n_labels = want_labels ? b.d_nlabels : nullptr;
    a.div_nbx = make_fastdiv((uint32_t)std::max(2, b.nbx));
    a.step_by = b.threads / b.nbx; a.step_bx = b.threads % b.nbx;
    // merge phase: tile scan (ccl.cuh, 2a/2b; about one tile of 32 columns per warp) for large grids, the concurrent
    // union-find over all foreground blocks for small ones.  Measured per 1024 4K masks (tools/ccl_timing.py): network masks
    // 0.35 -> 0.31 ms, dense network masks 0.86 -> 0.66, dense noise 0.88 -> 0.74 (all-ones 0.35 -> 0.55, empty 0.09 -> 0.11);
    // at 720p / 1080p the row-serial scan loses to the union-find (0.160 -> 0.189 / 0.156 -> 0.191 ms), so it is off there.
    // COVA_CCL_SCAN=0/1 forces either (development knob; both pass the golden and stress suites).
    static const char *scan_env = getenv("COVA_CCL_SCAN");
    a.scan = b.nb >= 4096;
    if (scan_env && (scan_env[0] == '0' || scan_env[0] == '1')) a.scan = scan_env[0] == '1';
    a.tiles_x = (b.nbx + 31) / 32;
    a.tiles_y = std::max(1, std::min(b.nby, (b.threads / 32) / a.tiles_x));
    a.tile_rows = (b.nby + a.tiles_y - 1) / a.tiles_y;
    a.tiles_y = (b.nby + a.tile_rows - 1) / a.tile_rows;
    a.div_tile_rows = make_fastdiv((uint32_t)std::max(2, a.tile_rows));
    if (reset_cursor) COVA_CUDA(cudaMemsetAsync(b.d_cursor, 0, 2 * sizeof(unsigned long long), st));
    // the attribute is per function AND per device, last write wins: handles of different grids (or on different devices,
    // or on different threads) share ccl_bbox_kernel, so it is set for every launch and always to the same value, the
    // architectural maximum (see tc::try_launch)
    auto kernel = b.compact ? ccl_bbox_kernel<true> : ccl_bbox_kernel<false>;
    COVA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCclSmemLimit));
    COVA_CUDA(launch_pdl(kernel, dim3((unsigned)n), dim3((unsigned)b.threads), b.smem, st, pdl, a));
    return COVA_OK;
}

// =================================================================================================
// bboxcc element
// =================================================================================================
struct cova_bboxcc {
    int device;
    uint32_t width, height, cc_threshold;
    CclBuffers b;
    cudaStream_t stream = nullptr;
};

extern "C" int cova_bboxcc_new(cova_bboxcc **out, int device, uint32_t width, uint32_t height, uint32_t cc_threshold) {
    if (!out) return set_err(COVA_E_INVAL, "null out");
    *out = nullptr;
    if (!width || !height) return set_err(COVA_E_INVAL, "empty mask");
    int rc = check_device(device);
    if (rc) return rc;
    auto *cc = new (std::nothrow) cova_bboxcc();
    if (!cc) return set_err(COVA_E_NOMEM, "host allocation failed");
    cc->device = device; cc->width = width; cc->height = height; cc->cc_threshold = cc_threshold;
    rc = ccl_alloc(cc->b, (int)height, (int)width, 1, true);
    if (rc == COVA_OK && cudaStreamCreateWithFlags(&cc->stream, cudaStreamNonBlocking) != cudaSuccess)
        rc = set_err(COVA_E_CUDA, "stream creation failed");
    if (rc) { cova_bboxcc_free(cc); return rc; }
    *out = cc;
    return COVA_OK;
}
extern "C" void cova_bboxcc_free(cova_bboxcc *cc) {
    if (!cc) return;
    cudaSetDevice(cc->device);
    ccl_free(cc->b, true);
    if (cc->stream) cudaStreamDestroy(cc->stream);
    delete cc;
}
extern "C" int cova_bboxcc_set_cc_threshold(cova_bboxcc *cc, uint32_t t) {
    if (!cc) return set_err(COVA_E_INVAL, "null handle");
    cc->cc_threshold = t;
    return COVA_OK;
}
extern "C" int cova_bboxcc_get_cc_threshold(const cova_bboxcc *cc, uint32_t *t) {
    if (!cc || !t) return set_err(COVA_E_INVAL, "null argument");
    *t = cc->cc_threshold;
    return COVA_OK;
}
extern "C" size_t cova_bboxcc_max_out_size(const cova_bboxcc *cc) { return cc ? 8 + 24 * (size_t)cc->b.nb : 0; }

extern "C" int cova_bboxcc_transform_ip(cova_bboxcc *cc, const uint8_t *mask, size_t mask_len, uint8_t *out, size_t out_cap,
                                        size_t *out_len) {
    if (!cc || !mask || !out_len) return set_err(COVA_E_INVAL, "null argument");
    // process.rs:14-15: reshape(1, height) -> rows of len/height bytes; the element is configured per caps
    if (mask_len != (size_t)cc->width * cc->height) return set_err(COVA_E_INVAL, "mask length != width*height of the caps");
    COVA_CUDA(cudaSetDevice(cc->device));
    COVA_CUDA(cudaMemcpyAsync(cc->b.d_masks, mask, mask_len, cudaMemcpyHostToDevice, cc->stream));
    int rc = ccl_launch(cc->b, cc->b.d_masks, 1, cc->cc_threshold, false, cc->stream);
    if (rc) return rc;
    // One synchronisation per buffer: the length and as much of the blob as the caller's buffer (or the worst case of this
    // grid) can hold travel back together; the single mask's blob starts at offset 0 of the arena.
    unsigned long long len = 0;
    const size_t spec = out ? std::min(out_cap, cc->b.blob_cap) : 0;
    COVA_CUDA(cudaMemcpyAsync(&len, cc->b.d_lens, sizeof(len), cudaMemcpyDeviceToHost, cc->stream));
    if (spec) COVA_CUDA(cudaMemcpyAsync(out, cc->b.d_blob, spec, cudaMemcpyDeviceToHost, cc->stream));
    COVA_CUDA(cudaStreamSynchronize(cc->stream));
    *out_len = (size_t)len;
    if (!out || out_cap < len) return set_err(COVA_E_TOOSMALL, "serialized boxes need a larger buffer");
    return COVA_OK;
}

extern "C" int cova_bboxcc_labels(cova_bboxcc *cc, const uint8_t *mask, size_t mask_len, int32_t *labels, int32_t *stats,
                                  int32_t *n_labels) {
    if (!cc || !mask || !labels || !stats || !n_labels) return set_err(COVA_E_INVAL, "null argument");
    if (mask_len != (size_t)cc->width * cc->height) return set_err(COVA_E_INVAL, "mask length != width*height");
    COVA_CUDA(cudaSetDevice(cc->device));
    CclBuffers &b = cc->b;
    if (!b.d_labels) {
        COVA_CUDA(cudaMalloc(&b.d_labels, mask_len * sizeof(int32_t)));
        COVA_CUDA(cudaMalloc(&b.d_stats, (size_t)(b.nb + 1) * 5 * sizeof(int32_t)));
        COVA_CUDA(cudaMalloc(&b.d_nlabels, sizeof(int32_t)));
    }
    COVA_CUDA(cudaMemcpyAsync(b.d_masks, mask, mask_len, cudaMemcpyHostToDevice, cc->stream));
    int rc = ccl_launch(b, b.d_masks, 1, cc->cc_threshold, true, cc->stream);
    if (rc) return rc;
    COVA_CUDA(cudaMemcpyAsync(n_labels, b.d_nlabels, sizeof(int32_t), cudaMemcpyDeviceToHost, cc->stream));
    COVA_CUDA(cudaMemcpyAsync(labels, b.d_labels, mask_len * sizeof(int32_t), cudaMemcpyDeviceToHost, cc->stream));
    COVA_CUDA(cudaStreamSynchronize(cc->stream));
    COVA_CUDA(cudaMemcpyAsync(stats, b.d_stats, (size_t)(*n_labels) * 5 * sizeof(int32_t), cudaMemcpyDeviceToHost, cc->stream));
    COVA_CUDA(cudaStreamSynchronize(cc->stream));
    return COVA_OK;
}

// =================================================================================================
// fused batch pipeline
// =================================================================================================
struct DevLayer {
    uint4 *wpack = nullptr;
    float *epi = nullptr;
    int n_cols = 0;
};

constexpr int kSlots = 4;

struct cova_pipeline {
    int device = 0, n_sms = 0;
    uint32_t W = 0, H = 0, T = 0, gamma = 1, max_streams = 0, max_fps = 0, max_windows = 0, flags = 0, impl = 0;
    uint32_t cc_threshold = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    size_t frame_bytes = 0;              // bytes of one frame in the pool: 4 per macroblock, or 2 with COVA_FLAG_INPUT_PACKED16
    bool packed = false;
    uint8_t *d_frames = nullptr;
    int *d_newest = nullptr;
    uint32_t cur_streams = 0, cur_fps = 0, cur_windows = 0, table_streams = 0, table_fps = 0, table_first = 0;
    // index (inside a device chain) of the newest frame of the chain's first window, and windows per chain: T - 1 and
    // (fps - T) / gamma + 1 for a chain that starts with an empty window; a CONTINUED chain (submit_host2) starts with the
    // carried T - 1 frames of its stream and its first window sits where the stream's gamma phase puts it
    uint32_t cur_first = 0, cur_wps = 0;
    // per-stream state for CONTINUED batches (metapreprocess/imp.rs:38-42: prev_buffers and gamma_idx live as long as the
    // stream): the last T - 1 frames of every stream id on the device, frames seen so far on the host
    uint8_t *d_carry = nullptr;
    std::vector<uint64_t> n_seen, id_mark;
    uint64_t id_epoch = 0;
    // A batch is processed in chunks of whole chains: the activation buffers are sized for ONE chunk, and
    // process_host() overlaps the H2D copy of chunk c+1, the kernels of chunk c and the D2H of chunk c-1.
    uint32_t chunk_streams = 0;          // chains per chunk (capacity)
    uint32_t ck_stream0 = 0, ck_n_streams = 0, ck_window0 = 0, ck_windows = 0;   // chunk being processed
    cudaStream_t s_in = nullptr, s_out = nullptr;
    // Four batch slots (frame pool + box arena + events): submit_host(k+3) can be queued before collect_host(k).  A
    // batch's host-visible latency is H2D + kernels + D2H of the boxes (whose size the host must first learn): measured
    // 2.25 + 2.05 + 0.65 ms for 8192 windows of 720p.  With k batches ahead the loop sustains one batch per
    // max(H2D, kernels, latency / k): two ahead = 2.47 ms (measured 2.46), three ahead = the PCIe time, 2.25 ms.
    // The synchronous entry points use slot 0.
    struct Slot {
        uint8_t *d_frames = nullptr;
        CclBuffers ccl;
        std::vector<cudaEvent_t> ev_in, ev_done;
        unsigned long long *h_cursor = nullptr;      // pinned: per-chunk cursor snapshots (2 words each)
        uint32_t n_streams = 0, fps = 0, n_windows = 0, wps = 0;
        bool busy = false, failed = false, ready = false;
        uint32_t *h_ids = nullptr, *d_ids = nullptr;   // stream ids of a submit_host2 batch (pinned staging + device copy)
        std::vector<uint64_t> win_pts;                 // per window: PTS of its newest frame / stream id (submit_host2)
        std::vector<uint32_t> win_ids;
    } slot[kSlots];
    int cur_slot = 0, next_submit = 0, next_collect = 0;
    int sizes_h[5], sizes_w[5];          // extents: [0] input, [1..4] encoder outputs
    Geom gx[4];                          // X0..X3 (Tn = 4): inputs of enc1..enc4
    Geom gd[4];                          // D0in..D3in (Tn = 1): inputs of dec0..dec3
    Geom gx0f, gp1;                      // per-FRAME first-conv input (x-pair packed) and pooled output (Tn = 1, N = frames)
    uint4 *x0f = nullptr, *p1 = nullptr;
    std::vector<int> h_newest;
    uint4 *x[4] = {nullptr, nullptr, nullptr, nullptr};
    uint4 *d[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *d_mask = nullptr, *d_stacked = nullptr;
    float *d_logits = nullptr;
    HostWeights hw;
    float *d_wraw = nullptr;
    DevLayer enc[4], dec[4];
    DevLayer encs[4];                    // blocks 2..4, weights-stationary kernel (blobnet_enc.cuh)
    CclBuffers ccl;
    unsigned int *d_watchdog = nullptr;
    unsigned long long *h_pinned = nullptr;   // cursor + overflow, then offsets/lens staging
    bool profiling = false;
    std::vector<cudaEvent_t> events;
    std::vector<std::string> names, last_names;
    std::vector<float> last_ms;
    size_t ev_used = 0;
    uint64_t launches = 0;
    int dbg = 0;
};

static uint32_t windows_per_stream(uint32_t fps, uint32_t T, uint32_t gamma) {
    if (fps < T) return 0;
    return (fps - T) / gamma + 1;
}
// windows of a device chain of `fps` frames whose first window has its newest frame at index `first`
static uint32_t windows_from(uint32_t fps, uint32_t first, uint32_t gamma) { return fps > first ? (fps - 1 - first) / gamma + 1 : 0; }

static void prof_mark(cova_pipeline *p, const char *name) {
    if (!p->profiling) return;
    if (p->ev_used >= p->events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        p->events.push_back(e);
    }
    cudaEventRecord(p->events[p->ev_used++], p->stream);
    p->names.push_back(name);
}

template <typename T>
static int dev_upload(T **dst, const void *src, size_t bytes) {
    COVA_CUDA(cudaMalloc(dst, bytes));
    COVA_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return COVA_OK;
}

static int upload_layer(DevLayer &dl, const PackedLayer *halves, int n_halves) {
    size_t hb = halves[0].b.size() * sizeof(__half);
    COVA_CUDA(cudaMalloc(&dl.wpack, hb * n_halves));
    for (int h = 0; h < n_halves; h++)
        COVA_CUDA(cudaMemcpy(reinterpret_cast<char *>(dl.wpack) + hb * h, halves[h].b.data(), hb, cudaMemcpyHostToDevice));
    int rc = dev_upload(&dl.epi, halves[0].epi.data(), halves[0].epi.size() * sizeof(float));
    dl.n_cols = halves[0].n_cols;
    return rc;
}

extern "C" int cova_pipeline_new(cova_pipeline **out, int device, uint32_t w_mb, uint32_t h_mb, uint32_t timestep,
                                 uint32_t gamma, uint32_t max_streams, uint32_t max_fps, const void *weights,
                                 size_t weights_len, uint32_t cc_threshold, uint32_t flags) {
    if (!out) return set_err(COVA_E_INVAL, "null out");
    *out = nullptr;
    if (timestep != (uint32_t)kT) return set_err(COVA_E_UNSUPPORTED, "BlobNet is built for timestep = 4 (utils/train-blobnet.py:58)");
    if (gamma < 1 || !max_streams || max_fps < timestep) return set_err(COVA_E_INVAL, "need gamma >= 1, max_streams >= 1, max_frames_per_stream >= timestep");
    if (w_mb < 16 || h_mb < 16) return set_err(COVA_E_UNSUPPORTED, "macroblock grid must be at least 16x16 for four 2x poolings");
    if ((flags & 0xffu) > COVA_IMPL_SIMT) return set_err(COVA_E_INVAL, "unknown implementation selector");
    if (flags & COVA_FLAG_INPUT_PACKED16) {
        if (flags & COVA_FLAG_KEEP_STACKED) return set_err(COVA_E_INVAL, "the stacked RGBA windows cannot be rebuilt from packed input (byte 3 and values above 6 are gone)");
        if ((flags & 0xffu) == COVA_IMPL_SIMT) return set_err(COVA_E_UNSUPPORTED, "the validation kernels read the 4-byte input format");
        if (w_mb & 1) return set_err(COVA_E_UNSUPPORTED, "packed input needs an even macroblock-grid width");
    }
#ifndef COVA_VALIDATION
    if ((flags & 0xffu) == COVA_IMPL_SIMT)
        return set_err(COVA_E_UNSUPPORTED, "the validation kernels (COVA_IMPL_SIMT) are not part of this build; use libcova_b200_val.so");
#endif
    int rc = check_device(device);
    if (rc) return rc;
    auto *p = new (std::nothrow) cova_pipeline();
    if (!p) return set_err(COVA_E_NOMEM, "host allocation failed");
    p->device = device; p->W = w_mb; p->H = h_mb; p->T = timestep; p->gamma = gamma;
    p->max_streams = max_streams; p->max_fps = max_fps; p->flags = flags; p->impl = flags & 0xffu;
    p->cc_threshold = cc_threshold;
    p->max_windows = max_streams * windows_per_stream(max_fps, timestep, gamma);
    auto fail = [&](int code) { cova_pipeline_free(p); return code; };
    if ((rc = parse_weights(weights, weights_len, p->hw))) return fail(rc);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(set_err(COVA_E_CUDA, "cudaGetDeviceProperties failed"));
    p->n_sms = prop.multiProcessorCount;
    if (p->impl == COVA_IMPL_TCGEN05 && prop.major != 10)
        return fail(set_err(COVA_E_UNSUPPORTED, "the tcgen05 path needs an sm_100 device"));
    if (cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_err(COVA_E_CUDA, "stream creation failed"));
    p->stream = p->own_stream;

    {   // chunking: ~1024 windows per chunk, at most 16 chunks; tiny pipelines stay single-chunk
        const uint32_t wps = windows_per_stream(max_fps, timestep, gamma);
        uint32_t chunks = std::min<uint32_t>(16, std::max<uint32_t>(1, p->max_windows / 1024));
        const uint32_t hint = (flags >> 16) & 0xffu;
        if (hint) chunks = hint;
        chunks = std::min(chunks, max_streams);
        p->chunk_streams = (max_streams + chunks - 1) / chunks;
        chunks = (max_streams + p->chunk_streams - 1) / p->chunk_streams;
        if (cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) != cudaSuccess)
            return fail(set_err(COVA_E_CUDA, "stream creation failed"));
        for (auto &sl : p->slot) {
            sl.ev_in.resize(chunks); sl.ev_done.resize(chunks);
            for (uint32_t c = 0; c < chunks; c++)
                if (cudaEventCreateWithFlags(&sl.ev_in[c], cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&sl.ev_done[c], cudaEventDisableTiming) != cudaSuccess)
                    return fail(set_err(COVA_E_CUDA, "event creation failed"));
        }
        (void)wps;
    }
    const int N = (int)(p->chunk_streams * windows_per_stream(max_fps, timestep, gamma));   // windows per chunk
    const int NB = (int)p->max_windows;                                                       // windows per batch
    p->sizes_h[0] = (int)h_mb; p->sizes_w[0] = (int)w_mb;
    for (int i = 1; i <= 4; i++) { p->sizes_h[i] = (p->sizes_h[i - 1] + 1) / 2; p->sizes_w[i] = (p->sizes_w[i - 1] + 1) / 2; }
    // encoder inputs (Tn = 4): X0 (8 ch incl. padding), X1 (16), X2 (32), X3 (64)
    const int xc[4] = {8, 16, 32, 64};
    for (int i = 0; i < 4; i++) p->gx[i] = make_geom(p->sizes_h[i], p->sizes_w[i], xc[i], kT, N);
    // decoder inputs (Tn = 1): D0in = enc4 t0 (128 @ s4), D1in (128 @ s3), D2in (64 @ s2), D3in (32 @ s1)
    for (int i = 0; i < 4; i++) p->gd[i] = make_geom(p->sizes_h[4 - i], p->sizes_w[4 - i], kDecCin[i], 1, N);
    auto alloc_zero = [&](uint4 **ptr, const Geom &g) -> int {
        size_t bytes = (size_t)geom_rows(g) * 16;
        COVA_CUDA(cudaMalloc(ptr, bytes));
        COVA_CUDA(cudaMemset(*ptr, 0, bytes));
        return COVA_OK;
    };
    for (int i = 0; i < 4; i++) {
        // X0 in window layout is only consumed by the validation kernels; the tcgen05 path works per frame
        if (i > 0 || p->impl == COVA_IMPL_SIMT)
            if ((rc = alloc_zero(&p->x[i], p->gx[i]))) return fail(rc);
        if ((rc = alloc_zero(&p->d[i], p->gd[i]))) return fail(rc);
    }
    const int F = (int)(p->chunk_streams * max_fps);
    p->gx0f = make_geom(p->sizes_h[0], p->sizes_w[0], 8, 1, F);
    p->gp1 = make_geom(p->sizes_h[1], p->sizes_w[1], kEncCout[0], 1, F);
    if ((rc = alloc_zero(&p->x0f, p->gx0f))) return fail(rc);
    if ((rc = alloc_zero(&p->p1, p->gp1))) return fail(rc);
    p->packed = (flags & COVA_FLAG_INPUT_PACKED16) != 0;
    p->frame_bytes = (size_t)w_mb * h_mb * (p->packed ? 2 : 4);
    cudaError_t e = cudaMalloc(&p->slot[0].d_frames, p->frame_bytes * max_streams * max_fps);
    p->d_frames = p->slot[0].d_frames;
    if (e == cudaSuccess) e = cudaMalloc(&p->d_newest, sizeof(int) * std::max(1, N));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_mask, (size_t)std::max(1, NB) * w_mb * h_mb);
    if (e == cudaSuccess) e = cudaMemset(p->d_mask, 0, (size_t)std::max(1, NB) * w_mb * h_mb);
    if (e == cudaSuccess && (flags & COVA_FLAG_KEEP_LOGITS)) e = cudaMalloc(&p->d_logits, sizeof(float) * (size_t)NB * w_mb * h_mb);
    if (e == cudaSuccess && (flags & COVA_FLAG_KEEP_STACKED)) e = cudaMalloc(&p->d_stacked, p->frame_bytes * timestep * (size_t)NB);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_watchdog, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p->d_watchdog, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMallocHost(&p->h_pinned, sizeof(unsigned long long) * (2 + 2 * (size_t)std::max(1, NB) + 2 * 64));
    if (e != cudaSuccess) return fail(set_err(COVA_E_CUDA, "pipeline allocation: %s", cudaGetErrorString(e)));
    p->slot[0].ccl.d_masks = p->d_mask;
    if ((rc = ccl_alloc(p->slot[0].ccl, (int)h_mb, (int)w_mb, std::max(1, NB), false))) return fail(rc);
    p->slot[0].ready = true;
    p->ccl = p->slot[0].ccl;
    for (auto &sl : p->slot)
        if (cudaMallocHost(&sl.h_cursor, sizeof(unsigned long long) * 2 * 256) != cudaSuccess)
            return fail(set_err(COVA_E_CUDA, "pinned allocation failed"));

    // weights: raw fp32 for the validation kernels, packed fp16 operand blocks for the tcgen05 path
    if ((rc = dev_upload(&p->d_wraw, p->hw.storage.data(), p->hw.storage.size() * sizeof(float)))) return fail(rc);
    for (int i = 0; i < 4; i++) {
        PackedLayer pl;
        pack_encoder(p->hw, i, pl);
        if ((rc = upload_layer(p->enc[i], &pl, 1))) return fail(rc);
        if (i >= 1) {
            PackedLayer ws;
            pack_encoder_ws(p->hw, i, i == 1 ? 2 : 1, i <= 2 ? 2 : 1, ws);
            if ((rc = upload_layer(p->encs[i], &ws, 1))) return fail(rc);
        }
    }
    for (int i = 0; i < 4; i++) {
        const int nsplit = i == 0 ? 2 : 1;
        PackedLayer pl[2];
        for (int h = 0; h < nsplit; h++) pack_decoder(p->hw, i, nsplit, h, pl[h]);
        if ((rc = upload_layer(p->dec[i], pl, nsplit))) return fail(rc);
    }
    *out = p;
    return COVA_OK;
}

extern "C" void cova_pipeline_free(cova_pipeline *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 4; i++) {
        if (p->x[i]) cudaFree(p->x[i]);
        if (i == 0 && p->x0f) cudaFree(p->x0f);
        if (i == 0 && p->p1) cudaFree(p->p1);
        if (p->d[i]) cudaFree(p->d[i]);
        if (p->enc[i].wpack) cudaFree(p->enc[i].wpack);
        if (p->enc[i].epi) cudaFree(p->enc[i].epi);
        if (p->encs[i].wpack) cudaFree(p->encs[i].wpack);
        if (p->encs[i].epi) cudaFree(p->encs[i].epi);
        if (p->dec[i].wpack) cudaFree(p->dec[i].wpack);
        if (p->dec[i].epi) cudaFree(p->dec[i].epi);
    }
    for (auto &sl : p->slot) {
        if (sl.d_frames) cudaFree(sl.d_frames);
        ccl_free(sl.ccl, false);
        if (sl.h_cursor) cudaFreeHost(sl.h_cursor);
        if (sl.h_ids) cudaFreeHost(sl.h_ids);
        if (sl.d_ids) cudaFree(sl.d_ids);
        for (auto e : sl.ev_in) cudaEventDestroy(e);
        for (auto e : sl.ev_done) cudaEventDestroy(e);
    }
    if (p->d_newest) cudaFree(p->d_newest);
    if (p->d_carry) cudaFree(p->d_carry);
    if (p->d_mask) cudaFree(p->d_mask);
    if (p->d_logits) cudaFree(p->d_logits);
    if (p->d_stacked) cudaFree(p->d_stacked);
    if (p->d_wraw) cudaFree(p->d_wraw);
    if (p->d_watchdog) cudaFree(p->d_watchdog);
    if (p->h_pinned) cudaFreeHost(p->h_pinned);
    for (auto e : p->events) cudaEventDestroy(e);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    cudaGetLastError();
    delete p;
}

extern "C" int cova_pipeline_set_cc_threshold(cova_pipeline *p, uint32_t t) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    p->cc_threshold = t;
    return COVA_OK;
}
extern "C" int cova_pipeline_set_stream(cova_pipeline *p, void *s) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    p->stream = s ? reinterpret_cast<cudaStream_t>(s) : p->own_stream;
    return COVA_OK;
}
extern "C" int cova_pipeline_n_windows(const cova_pipeline *p, uint32_t n_streams, uint32_t fps, uint32_t *n) {
    if (!p || !n) return set_err(COVA_E_INVAL, "null argument");
    *n = n_streams * windows_per_stream(fps, p->T, p->gamma);
    return COVA_OK;
}

static int set_batch_shape(cova_pipeline *p, uint32_t n_streams, uint32_t fps, uint32_t first = 0xffffffffu) {
    if (!n_streams || n_streams > p->max_streams || fps > p->max_fps)
        return set_err(COVA_E_INVAL, "batch exceeds max_streams / max_frames_per_stream of the pipeline");
    if (first == 0xffffffffu) first = p->T - 1;                   // chains that start with an empty window
    const uint32_t wps = windows_from(fps, first, p->gamma);
    if ((uint64_t)n_streams * wps > p->max_windows) return set_err(COVA_E_INVAL, "batch produces more windows than the pipeline was sized for");
    p->cur_streams = n_streams; p->cur_fps = fps; p->cur_windows = n_streams * wps;
    p->cur_first = first; p->cur_wps = wps;
    const uint32_t tstreams = std::min(n_streams, p->chunk_streams);
    if (p->table_streams != tstreams || p->table_fps != fps || p->table_first != first) {
        // chunk-relative table: window k of a chunk -> index (inside the chunk's frames) of its newest frame
        std::vector<int> &newest = p->h_newest;
        newest.assign(std::max<size_t>(1, (size_t)tstreams * wps), 0);
        size_t k = 0;
        for (uint32_t s = 0; s < tstreams; s++)
            for (uint32_t w = 0; w < wps; w++) newest[k++] = (int)(s * fps + first + w * p->gamma);
        if (k) COVA_CUDA(cudaMemcpyAsync(p->d_newest, newest.data(), sizeof(int) * k, cudaMemcpyHostToDevice, p->stream));
        COVA_CUDA(cudaStreamSynchronize(p->stream));
        p->table_streams = tstreams; p->table_fps = fps; p->table_first = first;
    }
    return COVA_OK;
}

static void use_slot(cova_pipeline *p, int k);
static uint32_t n_chunks_of(const cova_pipeline *p) { return (p->cur_streams + p->chunk_streams - 1) / p->chunk_streams; }
static void select_chunk(cova_pipeline *p, uint32_t c) {
    const uint32_t wps = p->cur_wps;
    p->ck_stream0 = c * p->chunk_streams;
    p->ck_n_streams = std::min(p->chunk_streams, p->cur_streams - p->ck_stream0);
    p->ck_window0 = p->ck_stream0 * wps;
    p->ck_windows = p->ck_n_streams * wps;
}
static int require_single_chunk(cova_pipeline *p) {
    if (p->cur_streams > p->chunk_streams)
        return set_err(COVA_E_UNSUPPORTED, "stage-wise calls need a batch that fits one chunk; use cova_pipeline_run / process_host");
    select_chunk(p, 0);
    return COVA_OK;
}

extern "C" int cova_pipeline_load_frames(cova_pipeline *p, const uint8_t *frames, uint32_t n_streams, uint32_t fps, int is_device) {
    if (!p || !frames) return set_err(COVA_E_INVAL, "null argument");
    COVA_CUDA(cudaSetDevice(p->device));
    int rc = set_batch_shape(p, n_streams, fps);
    if (rc) return rc;
    use_slot(p, 0);
    COVA_CUDA(cudaMemcpyAsync(p->d_frames, frames, p->frame_bytes * n_streams * fps,
                              is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    return COVA_OK;
}

extern "C" int cova_pipeline_load_masks(cova_pipeline *p, const uint8_t *masks, uint32_t n, int is_device) {
    if (!p || !masks) return set_err(COVA_E_INVAL, "null argument");
    if (!n || n > p->max_windows) return set_err(COVA_E_INVAL, "mask batch exceeds the pipeline's window capacity");
    COVA_CUDA(cudaSetDevice(p->device));
    p->cur_windows = n; p->cur_streams = 0;
    use_slot(p, 0);
    COVA_CUDA(cudaMemcpyAsync(p->d_mask, masks, (size_t)n * p->W * p->H, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    return COVA_OK;
}

static int tensorise_chunk(cova_pipeline *p) {
    const int N = (int)p->ck_windows;
    if (!N) return COVA_OK;
    const uint8_t *frames = p->d_frames + (size_t)p->ck_stream0 * p->cur_fps * p->frame_bytes;
    if (p->d_stacked) {
        const size_t S = p->frame_bytes;
        uint8_t *dst = p->d_stacked + (size_t)p->ck_window0 * p->T * S;
        if (S % 16 == 0) {
            long long total = (long long)N * p->T * (S / 16);
            int blocks = (int)std::min<long long>((total + 255) / 256, (long long)p->n_sms * 16);
            stack_rgba_kernel<uint4><<<blocks, 256, 0, p->stream>>>(reinterpret_cast<const uint4 *>(frames), p->d_newest,
                                                                   reinterpret_cast<uint4 *>(dst), (int)(S / 16), (int)p->T, total);
        } else {
            long long total = (long long)N * p->T * (S / 4);
            int blocks = (int)std::min<long long>((total + 255) / 256, (long long)p->n_sms * 16);
            stack_rgba_kernel<uint32_t><<<blocks, 256, 0, p->stream>>>(reinterpret_cast<const uint32_t *>(frames), p->d_newest,
                                                                      reinterpret_cast<uint32_t *>(dst), (int)(S / 4), (int)p->T, total);
        }
        COVA_CUDA(cudaGetLastError());
        p->launches++;
        prof_mark(p, "stack_rgba");
    }
    {   // per-frame first-conv input (tcgen05 path)
        const int F = (int)(p->ck_n_streams * p->cur_fps);
        long long total = (long long)F * p->H * p->gx0f.Wh;
        int blocks = (int)std::min<long long>((total + 255) / 256, (long long)p->n_sms * 32);
        if (total >= (1ll << 32)) return set_err(COVA_E_UNSUPPORTED, "chunk too large for the frame tensorisation kernel (32-bit index)");
        COVA_CUDA(launch_pdl(p->packed ? tensorise_frames_kernel<true> : tensorise_frames_kernel<false>, dim3((unsigned)blocks), dim3(256), 0,
                             p->stream, !(p->dbg & kDbgNoPdl), reinterpret_cast<const uint32_t *>(frames), p->x0f, p->gx0f, F,
                             make_fastdiv((uint32_t)std::max(2, p->gx0f.Wh)), make_fastdiv((uint32_t)std::max<uint32_t>(2u, p->H))));
        p->launches++;
        prof_mark(p, "tensorise_frames");
    }
#ifdef COVA_VALIDATION
    if (p->x[0]) {   // window-layout input of the validation kernels
        long long total = (long long)N * p->H * p->gx[0].Wh * kT;
        int blocks = (int)std::min<long long>((total + 255) / 256, (long long)p->n_sms * 32);
        tensorise_x0_kernel<<<blocks, 256, 0, p->stream>>>(reinterpret_cast<const uint32_t *>(frames), p->d_newest, p->x[0], p->gx[0], N);
        COVA_CUDA(cudaGetLastError());
        p->launches++;
        prof_mark(p, "tensorise_x0");
    }
#endif
    return COVA_OK;
}

extern "C" int cova_pipeline_tensorise(cova_pipeline *p) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    if (!p->cur_windows) return COVA_OK;
    COVA_CUDA(cudaSetDevice(p->device));
    int rc = require_single_chunk(p);
    return rc ? rc : tensorise_chunk(p);
}

// ---- layer launchers -------------------------------------------------------------------------------
static void crop_for(int in_extent, int target, int &crop_lo) {
    int pad = (2 * in_extent + 2) - target;     // decoder.py:42-59: (pad//2 + pad%2, pad//2)
    crop_lo = pad / 2 + pad % 2;
}

#ifndef COVA_VALIDATION
static int simt_layer(cova_pipeline *, int) {
    return set_err(COVA_E_UNSUPPORTED, "the validation kernels (COVA_IMPL_SIMT) are not part of this build; use libcova_b200_val.so");
}
#else
static int simt_layer(cova_pipeline *p, int layer) {
    const int N = (int)p->ck_windows;
    const float *w = p->d_wraw;
    if (layer == 0 && !p->x[0]) return set_err(COVA_E_UNSUPPORTED, "validation kernel of layer 0 needs a pipeline created with COVA_IMPL_SIMT");
    if (layer < 4) {
        const int i = layer;
        SimtEncArgs a;
        a.in = reinterpret_cast<const Row8 *>(p->x[i]); a.gin = p->gx[i];
        a.out = i < 3 ? reinterpret_cast<Row8 *>(p->x[i + 1]) : nullptr;
        a.gout = i < 3 ? p->gx[i + 1] : p->gx[3];
        // skip of enc(i+1) feeds dec(3-i): enc1 -> D3in, enc2 -> D2in, enc3 -> D1in, enc4 -> D0in (whole input)
        a.out2 = reinterpret_cast<Row8 *>(p->d[3 - i]); a.gout2 = p->gd[3 - i];
        a.out2_cb = i < 3 ? kDecCout[2 - i] / 8 : 0;
        a.w = w + p->hw.off_enc[i][0]; a.b = w + p->hw.off_enc[i][1]; a.gamma = w + p->hw.off_enc[i][2];
        a.beta = w + p->hw.off_enc[i][3]; a.mean = w + p->hw.off_enc[i][4]; a.var = w + p->hw.off_enc[i][5];
        a.w1 = w + p->hw.off_enc[i][6]; a.w2 = w + p->hw.off_enc[i][7];
        a.Cin = kEncCin[i]; a.Cout = kEncCout[i]; a.N = N;
        a.in_scale = i == 0 ? 1.0f / 6.0f : 1.0f;
        long long total = (long long)N * (a.gin.H / 2) * (a.gin.W / 2) * (a.Cout / 8);
        simt_encoder_kernel<<<(unsigned)((total + 127) / 128), 128, 0, p->stream>>>(a);
        COVA_CUDA(cudaGetLastError());
        p->launches++;
        prof_mark(p, i == 0 ? "simt_enc1" : i == 1 ? "simt_enc2" : i == 2 ? "simt_enc3" : "simt_enc4");
        return COVA_OK;
    }
    const int i = layer - 4;
    SimtDecArgs a;
    a.in = reinterpret_cast<const Row8 *>(p->d[i]); a.gin = p->gd[i];
    a.out = i < 3 ? reinterpret_cast<Row8 *>(p->d[i + 1]) : nullptr;
    a.gout = i < 3 ? p->gd[i + 1] : p->gd[3];
    a.w = w + p->hw.off_dec[i][0]; a.b = w + p->hw.off_dec[i][1];
    a.gamma = w + p->hw.off_dec[i][2]; a.beta = w + p->hw.off_dec[i][3]; a.mean = w + p->hw.off_dec[i][4]; a.var = w + p->hw.off_dec[i][5];
    a.head_w = w + p->hw.off_head[0]; a.head_b = w + p->hw.off_head[1];
    a.mask = p->d_mask + (size_t)p->ck_window0 * p->W * p->H;
    a.logits = p->d_logits ? p->d_logits + (size_t)p->ck_window0 * p->W * p->H : nullptr;
    a.Cin = kDecCin[i]; a.Cout = kDecCout[i]; a.N = N;
    a.Ht = p->sizes_h[3 - i]; a.Wt = p->sizes_w[3 - i];
    crop_for(a.gin.H, a.Ht, a.crop_t);
    crop_for(a.gin.W, a.Wt, a.crop_l);
    if (i < 3) {
        long long total = (long long)N * a.Ht * a.Wt * (a.Cout / 8);
        simt_decoder_kernel<<<(unsigned)((total + 127) / 128), 128, 0, p->stream>>>(a);
    } else {
        long long total = (long long)N * a.Ht * a.Wt;
        simt_head_kernel<<<(unsigned)((total + 127) / 128), 128, 0, p->stream>>>(a);
    }
    COVA_CUDA(cudaGetLastError());
    p->launches++;
    prof_mark(p, i == 0 ? "simt_dec0" : i == 1 ? "simt_dec1" : i == 2 ? "simt_dec2" : "simt_dec3_head");
    return COVA_OK;
}
#endif

// Candidate tile configurations in order of preference; a configuration that can double-buffer its strips
// (ring depth >= 2) beats an earlier one that cannot.
template <class C0, class... Cs>
static int launch_fit(const tc::LayerParams &lp, int n_sms, cudaStream_t st, int min_stage) {
    cudaError_t err = cudaSuccess;
    if (tc::try_launch<C0>(lp, n_sms, st, err, min_stage)) {
        if (err != cudaSuccess) return set_err(COVA_E_CUDA, "tcgen05 layer launch: %s", cudaGetErrorString(err));
        return COVA_OK;
    }
    if constexpr (sizeof...(Cs) > 0) return launch_fit<Cs...>(lp, n_sms, st, min_stage);
    else return COVA_E_UNSUPPORTED;
}
template <class... Cs>
static int launch_first_fit(const tc::LayerParams &lp, int n_sms, cudaStream_t st) {
    int rc = launch_fit<Cs...>(lp, n_sms, st, 2);
    if (rc == COVA_E_UNSUPPORTED) rc = launch_fit<Cs...>(lp, n_sms, st, 1);
    if (rc == COVA_E_UNSUPPORTED) return set_err(COVA_E_UNSUPPORTED, "no tcgen05 tile configuration fits shared memory for this macroblock grid");
    return rc;
}

template <class C0, class... Cs>
static int launch_fit_e(const tc::LayerParams &lp, int n_sms, cudaStream_t st, int min_stage) {
    cudaError_t err = cudaSuccess;
    if (tcs::try_launch_e<C0>(lp, n_sms, st, err, min_stage)) {
        if (err != cudaSuccess) return set_err(COVA_E_CUDA, "tcgen05 encoder launch: %s", cudaGetErrorString(err));
        return COVA_OK;
    }
    if constexpr (sizeof...(Cs) > 0) return launch_fit_e<Cs...>(lp, n_sms, st, min_stage);
    else return COVA_E_UNSUPPORTED;
}
template <class... Cs>
static int launch_first_fit_e(const tc::LayerParams &lp, int n_sms, cudaStream_t st) {
    int rc = launch_fit_e<Cs...>(lp, n_sms, st, 2);
    if (rc == COVA_E_UNSUPPORTED) rc = launch_fit_e<Cs...>(lp, n_sms, st, 1);
    return rc;
}

static int tc_layer(cova_pipeline *p, int layer) {
    using namespace tc;
    const int N = (int)p->ck_windows;
    LayerParams lp;
    memset(&lp, 0, sizeof(lp));
    lp.N = N; lp.watchdog = p->d_watchdog; lp.dbg = p->dbg;
    int rc;
    if (layer < 4) {
        const int i = layer;
        lp.in = p->x[i]; lp.gin = p->gx[i];
        lp.out = i < 3 ? p->x[i + 1] : nullptr;
        lp.gout = i < 3 ? p->gx[i + 1] : p->gx[3];
        lp.out2 = p->d[3 - i]; lp.gout2 = p->gd[3 - i];
        lp.out2_cb = i < 3 ? kDecCout[2 - i] / 8 : 0;
        lp.wpack = p->enc[i].wpack; lp.epi = p->enc[i].epi;
        memcpy(lp.tn_w1, p->hw.enc[i].tn_w1, 64);
        memcpy(lp.tn_w2, p->hw.enc[i].tn_w2, 64);
        lp.nsplit = 1;
        lp.bn_nonneg = 1;
        for (int c = 0; c < kEncCout[i]; c++)
            if (!(p->hw.enc[i].gamma[c] >= 0.f)) lp.bn_nonneg = 0;
        //                                  MODE   CIN_CB NCOLS TPS KCH COUT
        if (i == 0) {
            // first block: conv+ReLU+BN+pool once per FRAME, then PointWiseTN gathers the 4 frames of every window
            const int F = (int)(p->ck_n_streams * p->cur_fps);
            if (!(p->dbg & 16)) {
                // fused: frame-level conv + PointWiseTN over a register ring of 4 frames, one kernel (blobnet_enc1.cuh);
                // dbg bit 4 forces the two-kernel path below (also the fallback for grids the fused kernel cannot take)
                LayerParams fp = lp;
                fp.in = p->x0f; fp.gin = p->gx0f; fp.out = p->x[1]; fp.gout = p->gx[1]; fp.N = F;
                tc1::Enc1Extra ex;
                memset(&ex, 0, sizeof(ex));
                ex.n_chains = (int)p->ck_n_streams; ex.fps = (int)p->cur_fps; ex.wps = (int)p->cur_wps;
                ex.gamma = (int)p->gamma; ex.first = (int)p->cur_first;
                cudaError_t err = cudaSuccess;
                if (tc1::try_launch_enc1(fp, ex, p->n_sms, p->stream, err)) {
                    if (err != cudaSuccess) return set_err(COVA_E_CUDA, "fused enc1 launch: %s", cudaGetErrorString(err));
                    p->launches++;
                    prof_mark(p, "tc_enc1_fused");
                    return COVA_OK;
                }
            }
            lp.in = p->x0f; lp.gin = p->gx0f; lp.out = p->p1; lp.gout = p->gp1; lp.out2 = nullptr; lp.N = F;
            rc = launch_first_fit<Cfg<MODE_ENCF, 1, 16, 8, 1, 16>, Cfg<MODE_ENCF, 1, 16, 4, 1, 16>, Cfg<MODE_ENCF, 1, 16, 2, 1, 16>,
                                  Cfg<MODE_ENCF, 1, 16, 1, 1, 16>>(lp, p->n_sms, p->stream);
            if (rc) return rc;
            p->launches++;
            prof_mark(p, "tc_enc1_conv");
            TnArgs ta;
            ta.p1 = p->p1; ta.gp1 = p->gp1; ta.x1 = p->x[1]; ta.gx1 = p->gx[1]; ta.skip = p->d[3]; ta.gskip = p->gd[3];
            ta.skip_cb = kDecCout[2] / 8; ta.newest = p->d_newest; ta.n_windows = N;
            ta.wps = p->cur_wps; ta.fps = p->cur_fps; ta.gamma = p->gamma; ta.first = p->cur_first; ta.CB = kEncCout[0] / 8;
            memcpy(ta.w1, p->hw.enc[0].tn_w1, 64);
            memcpy(ta.w2, p->hw.enc[0].tn_w2, 64);
            long long total = (long long)ta.CB * 4 * N * p->gx[1].S;
            if (total >= (1ll << 32)) return set_err(COVA_E_UNSUPPORTED, "chunk too large for the PointWiseTN gather kernel (32-bit index)");
            // 64 CTAs per SM queued (4-5 resident) - measured sweep per 8192 windows at 720p: resident-only grid 0.49 ms,
            // x16 0.43, x64 0.39, one iteration per thread 0.45
            static const int tn_factor = getenv("COVA_TN_GRID") ? atoi(getenv("COVA_TN_GRID")) : 64;   // development knob
            int blocks = (int)std::min<long long>((total + 255) / 256, (long long)p->n_sms * tn_factor);
            pointwise_tn_kernel<<<blocks, 256, 0, p->stream>>>(ta);
            COVA_CUDA(cudaGetLastError());
            p->launches++;
            prof_mark(p, "enc1_pointwise_tn");
            return COVA_OK;
        }
        // block 3 (Cout = 64) stays on the positions-as-M kernel: measured 0.335 ms vs 0.376 ms per 8192 windows at 720p
        // (profiles/r1c_layer_timing.txt); dbg bit 3 forces the weights-stationary variant for experiments
        if (!(p->dbg & 4) && (i != 2 || (p->dbg & 8))) {
            // weights-stationary kernel first; the positions-as-M kernel below remains the fallback for grids
            // whose strips do not fit shared memory
            LayerParams ws = lp;
            ws.wpack = p->encs[i].wpack; ws.epi = p->encs[i].epi;
            using tcs::ECfg;
            if (i == 1) rc = launch_first_fit_e<ECfg<2, 32, 2, 2, 192, 3>, ECfg<2, 32, 2, 2, 192, 2>, ECfg<2, 32, 2, 2, 192, 1>>(ws, p->n_sms, p->stream);
            else if (i == 2) rc = launch_first_fit_e<ECfg<4, 64, 1, 2, 128, 2>, ECfg<4, 64, 1, 2, 128, 1>>(ws, p->n_sms, p->stream);
            else rc = launch_first_fit_e<ECfg<8, 128, 1, 1, 64, 2>, ECfg<8, 128, 1, 1, 64, 1>>(ws, p->n_sms, p->stream);
            if (rc != COVA_E_UNSUPPORTED) {
                if (rc) return rc;
                p->launches++;
                prof_mark(p, i == 1 ? "tc_enc2" : i == 2 ? "tc_enc3" : "tc_enc4");
                return COVA_OK;
            }
        }
        if (i == 1) rc = launch_first_fit<Cfg<MODE_ENC, 2, 32, 4, 2, 32>, Cfg<MODE_ENC, 2, 32, 2, 2, 32>, Cfg<MODE_ENC, 2, 32, 1, 2, 32>>(lp, p->n_sms, p->stream);
        else if (i == 2) rc = launch_first_fit<Cfg<MODE_ENC, 4, 64, 2, 4, 64>, Cfg<MODE_ENC, 4, 64, 1, 4, 64>, Cfg<MODE_ENC, 4, 64, 1, 2, 64>>(lp, p->n_sms, p->stream);
        else rc = launch_first_fit<Cfg<MODE_ENC, 8, 128, 1, 2, 128>>(lp, p->n_sms, p->stream);
        if (rc) return rc;
        p->launches++;
        prof_mark(p, i == 1 ? "tc_enc2" : i == 2 ? "tc_enc3" : "tc_enc4");
        return COVA_OK;
    }
    const int i = layer - 4;
    lp.in = p->d[i]; lp.gin = p->gd[i];
    lp.out = i < 3 ? p->d[i + 1] : nullptr;
    lp.gout = i < 3 ? p->gd[i + 1] : p->gd[3];
    lp.wpack = p->dec[i].wpack; lp.epi = p->dec[i].epi;
    lp.nsplit = i == 0 ? 2 : 1;
    lp.psplit = (p->dbg & 64) ? 0 : 1;     // dbg bit 6: whole-tile accumulators for dec0 / dec1 (experiments)
    lp.Ht = p->sizes_h[3 - i]; lp.Wt = p->sizes_w[3 - i];
    crop_for(lp.gin.H, lp.Ht, lp.crop_t);
    crop_for(lp.gin.W, lp.Wt, lp.crop_l);
    lp.mask = p->d_mask + (size_t)p->ck_window0 * p->W * p->H;
    lp.logits = p->d_logits ? p->d_logits + (size_t)p->ck_window0 * p->W * p->H : nullptr;
    if (i == 0) rc = launch_first_fit<Cfg<MODE_DEC, 16, 128, 1, 2, 64>>(lp, p->n_sms, p->stream);
    else if (i == 1) rc = launch_first_fit<Cfg<MODE_DEC, 16, 128, 1, 2, 32>>(lp, p->n_sms, p->stream);
    else if (i == 2) rc = launch_first_fit<Cfg<MODE_DEC, 8, 64, 1, 4, 16>, Cfg<MODE_DEC, 8, 64, 1, 2, 16>>(lp, p->n_sms, p->stream);
    else rc = launch_first_fit<Cfg<MODE_HEAD, 4, 16, 2, 4, 16>, Cfg<MODE_HEAD, 4, 16, 1, 4, 16>, Cfg<MODE_HEAD, 4, 16, 1, 2, 16>>(lp, p->n_sms, p->stream);
    if (rc) return rc;
    p->launches++;
    prof_mark(p, i == 0 ? "tc_dec0" : i == 1 ? "tc_dec1" : i == 2 ? "tc_dec2" : "tc_dec3_head");
    return COVA_OK;
}

extern "C" int cova_pipeline_run_layer(cova_pipeline *p, int layer, uint32_t impl) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    if (layer < 0 || layer > 7 || impl > COVA_IMPL_SIMT) return set_err(COVA_E_INVAL, "layer must be 0..7, impl 0 or 1");
    if (!p->cur_windows) return COVA_OK;
    COVA_CUDA(cudaSetDevice(p->device));
    int rc = require_single_chunk(p);
    if (rc) return rc;
    return impl == COVA_IMPL_SIMT ? simt_layer(p, layer) : tc_layer(p, layer);
}

static int blobnet_chunk(cova_pipeline *p) {
    if (!p->ck_windows) return COVA_OK;
    for (int layer = 0; layer < 8; layer++) {
        int rc = p->impl == COVA_IMPL_SIMT ? simt_layer(p, layer) : tc_layer(p, layer);
        if (rc) return rc;
    }
    return COVA_OK;
}

extern "C" int cova_pipeline_blobnet(cova_pipeline *p) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    if (!p->cur_windows) return COVA_OK;
    COVA_CUDA(cudaSetDevice(p->device));
    int rc = require_single_chunk(p);
    return rc ? rc : blobnet_chunk(p);
}

static int ccl_range(cova_pipeline *p, size_t first, int n, bool reset_cursor) {
    int rc = ccl_launch(p->ccl, p->d_mask, n, p->cc_threshold, false, p->stream, first, reset_cursor, !(p->dbg & kDbgNoPdl));
    if (rc) return rc;
    if (n > 0) {
        p->launches++;
        prof_mark(p, "ccl_bbox");
    }
    return COVA_OK;
}

extern "C" int cova_pipeline_ccl(cova_pipeline *p) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    COVA_CUDA(cudaSetDevice(p->device));
    if (!p->cur_windows) {
        COVA_CUDA(cudaMemsetAsync(p->ccl.d_cursor, 0, 2 * sizeof(unsigned long long), p->stream));
        return COVA_OK;
    }
    return ccl_range(p, 0, (int)p->cur_windows, true);
}

static void prof_begin(cova_pipeline *p) {
    if (!p->profiling) return;
    p->ev_used = 0;
    p->names.clear();
    prof_mark(p, "start");
}

// all kernels of one chunk, on p->stream
static int run_chunk(cova_pipeline *p, uint32_t c) {
    select_chunk(p, c);
    // the box arena's cursor is reset ahead of the batch's first kernel, not between the last layer and the CCL
    // kernel: a memset node there would break the chain of programmatic dependent launches (common.cuh)
    if (c == 0) COVA_CUDA(cudaMemsetAsync(p->ccl.d_cursor, 0, 2 * sizeof(unsigned long long), p->stream));
    int rc = tensorise_chunk(p);
    if (!rc) rc = blobnet_chunk(p);
    if (!rc) rc = ccl_range(p, p->ck_window0, (int)p->ck_windows, false);
    return rc;
}

extern "C" int cova_pipeline_run(cova_pipeline *p) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    COVA_CUDA(cudaSetDevice(p->device));
    prof_begin(p);
    if (!p->cur_streams) return cova_pipeline_ccl(p);      // masks were loaded directly
    if (!p->cur_windows) {
        COVA_CUDA(cudaMemsetAsync(p->ccl.d_cursor, 0, 2 * sizeof(unsigned long long), p->stream));
        return COVA_OK;
    }
    const uint32_t nc = n_chunks_of(p);
    for (uint32_t c = 0; c < nc; c++) {
        int rc = run_chunk(p, c);
        if (rc) return rc;
    }
    return COVA_OK;
}

static int sync_stream(cova_pipeline *p, cudaStream_t st) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        unsigned int code = 0;
        cudaMemcpy(&code, p->d_watchdog, sizeof(code), cudaMemcpyDeviceToHost);
        char extra[64];
        snprintf(extra, sizeof(extra), " (barrier watchdog code %u)", code);
        return set_err(COVA_E_CUDA, "stream synchronize: %s%s", cudaGetErrorString(e), extra);
    }
    return COVA_OK;
}

extern "C" int cova_pipeline_sync(cova_pipeline *p) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    COVA_CUDA(cudaSetDevice(p->device));
    int rc = sync_stream(p, p->stream);
    if (rc) return rc;
    if (p->profiling && p->ev_used > 1) {
        // per-kernel device time, summed over the chunks of the batch
        p->last_ms.clear();
        p->last_names.clear();
        for (size_t i = 1; i < p->ev_used; i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p->events[i - 1], p->events[i]);
            size_t k = 0;
            for (; k < p->last_names.size(); k++)
                if (p->last_names[k] == p->names[i]) break;
            if (k == p->last_names.size()) { p->last_names.push_back(p->names[i]); p->last_ms.push_back(0.f); }
            p->last_ms[k] += ms;
        }
        p->ev_used = 0;
        p->names.clear();
    }
    return COVA_OK;
}

extern "C" int cova_pipeline_fetch_boxes(cova_pipeline *p, uint8_t *blob, size_t blob_cap, size_t *blob_len, uint64_t *offsets, uint64_t *lens) {
    if (!p || !blob_len) return set_err(COVA_E_INVAL, "null argument");
    COVA_CUDA(cudaSetDevice(p->device));
    const size_t n = p->cur_windows;
    unsigned long long *hp = p->h_pinned;
    COVA_CUDA(cudaMemcpyAsync(hp, p->ccl.d_cursor, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    if (n) {
        COVA_CUDA(cudaMemcpyAsync(hp + 2, p->ccl.d_offsets, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
        COVA_CUDA(cudaMemcpyAsync(hp + 2 + n, p->ccl.d_lens, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    }
    int rc = cova_pipeline_sync(p);
    if (rc) return rc;
    if (hp[1]) return set_err(COVA_E_CUDA, "device box arena overflow (internal sizing error)");
    *blob_len = (size_t)hp[0];
    if (offsets) memcpy(offsets, hp + 2, n * sizeof(uint64_t));
    if (lens) memcpy(lens, hp + 2 + n, n * sizeof(uint64_t));
    if (!blob || blob_cap < hp[0]) return set_err(COVA_E_TOOSMALL, "box blob needs a larger buffer");
    if (hp[0]) {
        COVA_CUDA(cudaMemcpyAsync(blob, p->ccl.d_blob, (size_t)hp[0], cudaMemcpyDeviceToHost, p->stream));
        COVA_CUDA(cudaStreamSynchronize(p->stream));
    }
    return COVA_OK;
}

static void use_slot(cova_pipeline *p, int k) {
    p->cur_slot = k;
    p->d_frames = p->slot[k].d_frames;
    p->ccl = p->slot[k].ccl;
}
static int ensure_slot(cova_pipeline *p, int k) {
    auto &sl = p->slot[k];
    if (sl.ready) return COVA_OK;
    if (!sl.d_frames) COVA_CUDA(cudaMalloc(&sl.d_frames, p->frame_bytes * p->max_streams * p->max_fps));
    if (!sl.ccl.d_blob) {
        sl.ccl.d_masks = p->d_mask;
        int rc = ccl_alloc(sl.ccl, (int)p->H, (int)p->W, std::max<int>(1, (int)p->max_windows), false);
        if (rc) { ccl_free(sl.ccl, false); sl.ccl = CclBuffers(); return rc; }    // a later submit retries from scratch
    }
    sl.ready = true;
    return COVA_OK;
}

// Per-stream carry-over for CONTINUED batches.  carry[id] holds the last T - 1 frames of stream `id`, newest last.
//   mode 0 (restore): device chain s gets the newest h carried frames of stream ids[s] in front of its new frames;
//   mode 1 (save):    the last min(fps_dev, T - 1) frames of device chain s become the carry of stream ids[s].
__global__ void __launch_bounds__(256) carry_kernel(uint32_t *__restrict__ pool, uint32_t *__restrict__ carry, const uint32_t *__restrict__ ids,
                                                    int n_streams, int words_per_frame, int fps_dev, int h, int t1, int mode) {
    const int k = mode ? min(fps_dev, t1) : h;                     // frames moved per stream
    const long long per_stream = (long long)k * words_per_frame, total = per_stream * n_streams;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(i / per_stream);
        const long long e = i - (long long)s * per_stream;
        uint32_t *c = carry + ((long long)ids[s] * t1 + (t1 - k)) * words_per_frame + e;
        uint32_t *f = pool + ((long long)s * fps_dev + (mode ? fps_dev - k : 0)) * words_per_frame + e;
        if (mode) *c = *f; else *f = *c;
    }
}

// Host frames in, boxes out, asynchronously.  Three streams: the H2D copy of chunk c+1, the kernels of chunk c and
// the D2H copy of earlier boxes overlap; with up to four batches in flight (submit k+3 before collect k) the copies
// of one batch hide behind the kernels of another and the H2D engine is never idle.  A chunk's boxes occupy one contiguous range of the slot's device
// arena (the cursor is only reset at the start of a batch), so each chunk needs exactly one blob copy of exactly
// the bytes it produced.
//
// submit_host2 adds the stream identity the reference's elements have implicitly (one element instance per stream,
// State.prev_buffers / gamma_idx alive for the whole stream, metapreprocess/imp.rs:38-42,302-330): with
// COVA_SUBMIT_CONTINUE the chains of this batch continue the streams `stream_ids` from where their previous batch left
// them - the carried T - 1 frames are put in front of the new ones on the device and the gamma phase continues, so a
// stream cut into batches yields exactly the windows of the uncut stream.
static int submit_impl(cova_pipeline *p, const uint8_t *frames, uint32_t n_streams, uint32_t fps, const uint32_t *stream_ids,
                       const uint64_t *pts, uint32_t flags, bool stateful) {
    if (!p || !frames) return set_err(COVA_E_INVAL, "null argument");
    if (flags & ~COVA_SUBMIT_CONTINUE) return set_err(COVA_E_INVAL, "unknown submit flag");
    COVA_CUDA(cudaSetDevice(p->device));
    const int k = p->next_submit;
    if (p->slot[k].busy) return set_err(COVA_E_INVAL, "four batches are already in flight: collect one first");
    int rc = ensure_slot(p, k);
    if (rc) return rc;
    if (!n_streams || n_streams > p->max_streams) return set_err(COVA_E_INVAL, "batch exceeds max_streams of the pipeline");
    auto &sl = p->slot[k];
    const uint32_t t1 = p->T - 1;
    uint32_t h = 0, first = t1;
    if (stateful) {
        if (!fps) return set_err(COVA_E_INVAL, "a stream batch needs at least one frame per stream");
        if (p->n_seen.empty()) { p->n_seen.assign(p->max_streams, 0); p->id_mark.assign(p->max_streams, 0); }
        p->id_epoch++;                                              // "seen in this batch" marks without clearing the array
        for (uint32_t s = 0; s < n_streams; s++) {
            const uint32_t id = stream_ids ? stream_ids[s] : s;
            if (id >= p->max_streams) return set_err(COVA_E_INVAL, "stream id out of range (ids are 0 .. max_streams-1)");
            if (p->id_mark[id] == p->id_epoch) return set_err(COVA_E_INVAL, "a stream id appears twice in one batch");
            p->id_mark[id] = p->id_epoch;
        }
        if (flags & COVA_SUBMIT_CONTINUE) {
            // the kernels take one window shape per batch: every chain must carry the same number of frames and sit at the
            // same gamma phase, i.e. the streams of a batch advance in lock-step (start new streams in a batch of their own)
            for (uint32_t s = 0; s < n_streams; s++) {
                const uint64_t seen = p->n_seen[stream_ids ? stream_ids[s] : s];
                const uint32_t hs = (uint32_t)std::min<uint64_t>(seen, t1);
                const uint64_t base = seen - hs;                    // frames of the stream in front of the device chain
                // smallest index i' >= T-1 of the device chain with (base + i' - (T-1)) % gamma == 0 (imp.rs:321-330)
                const uint32_t fs = t1 + (uint32_t)((p->gamma - base % p->gamma) % p->gamma);
                if (s == 0) { h = hs; first = fs; }
                else if (hs != h || fs != first)
                    return set_err(COVA_E_INVAL, "streams of one CONTINUED batch must have the same history length and gamma phase");
            }
        }
        if (!p->d_carry && t1) {
            COVA_CUDA(cudaMalloc(&p->d_carry, (size_t)p->max_streams * t1 * p->frame_bytes));
            COVA_CUDA(cudaMemsetAsync(p->d_carry, 0, (size_t)p->max_streams * t1 * p->frame_bytes, p->stream));
        }
        if (!sl.h_ids) {
            COVA_CUDA(cudaMallocHost(&sl.h_ids, sizeof(uint32_t) * p->max_streams));
            COVA_CUDA(cudaMalloc(&sl.d_ids, sizeof(uint32_t) * p->max_streams));
        }
    }
    const uint32_t fps_dev = fps + h;                               // frames per chain on the device
    if (fps_dev > p->max_fps)
        return set_err(COVA_E_INVAL, "carried frames + new frames exceed max_frames_per_stream of the pipeline");
    if ((rc = set_batch_shape(p, n_streams, fps_dev, first))) return rc;
    use_slot(p, k);
    sl.n_streams = n_streams; sl.fps = fps_dev; sl.n_windows = p->cur_windows; sl.wps = p->cur_wps;
    sl.win_pts.clear(); sl.win_ids.clear();
    if (stateful) {
        // window w of chain s has its newest frame at device index first + w*gamma = new-frame index first + w*gamma - h
        sl.win_ids.resize(sl.n_windows);
        if (pts) sl.win_pts.resize(sl.n_windows);
        for (uint32_t s = 0; s < n_streams; s++) {
            const uint32_t id = stream_ids ? stream_ids[s] : s;
            sl.h_ids[s] = id;
            for (uint32_t w = 0; w < sl.wps; w++) {
                sl.win_ids[(size_t)s * sl.wps + w] = id;
                if (pts) sl.win_pts[(size_t)s * sl.wps + w] = pts[(size_t)s * fps + (first + w * p->gamma - h)];
            }
            p->n_seen[id] = ((flags & COVA_SUBMIT_CONTINUE) ? p->n_seen[id] : 0) + fps;
        }
    }
    sl.busy = true; sl.failed = false;
    p->next_submit = (k + 1) % kSlots;
    if (!stateful && !p->cur_windows) return COVA_OK;              // chains shorter than a window: nothing to compute
    // from here on a failure leaves the slot marked failed: the matching collect reports it instead of reading events that
    // were never recorded
    auto fail = [&](int code) { sl.failed = true; return code; };
#define COVA_CUDA_SLOT(expr)                                                                                        \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess) return fail(cova::set_err(COVA_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)));    \
    } while (0)
    const uint32_t nc = n_chunks_of(p);
    const size_t src_chain = p->frame_bytes * fps, dst_chain = p->frame_bytes * fps_dev, lead = p->frame_bytes * h;
    for (uint32_t c = 0; c < nc; c++) {
        const size_t s0 = (size_t)c * p->chunk_streams, ns = std::min<size_t>(p->chunk_streams, n_streams - s0);
        if (!h) COVA_CUDA_SLOT(cudaMemcpyAsync(sl.d_frames + s0 * dst_chain, frames + s0 * src_chain, ns * src_chain, cudaMemcpyHostToDevice, p->s_in));
        else COVA_CUDA_SLOT(cudaMemcpy2DAsync(sl.d_frames + s0 * dst_chain + lead, dst_chain, frames + s0 * src_chain, src_chain, src_chain, ns,
                                              cudaMemcpyHostToDevice, p->s_in));
        COVA_CUDA_SLOT(cudaEventRecord(sl.ev_in[c], p->s_in));
    }
    if (stateful) COVA_CUDA_SLOT(cudaMemcpyAsync(sl.d_ids, sl.h_ids, sizeof(uint32_t) * n_streams, cudaMemcpyHostToDevice, p->stream));
    const int wpf = (int)(p->frame_bytes / 4);
    for (uint32_t c = 0; c < nc; c++) {
        COVA_CUDA_SLOT(cudaStreamWaitEvent(p->stream, sl.ev_in[c], 0));
        if (stateful && t1) {
            const size_t s0 = (size_t)c * p->chunk_streams;
            const int ns = (int)std::min<size_t>(p->chunk_streams, n_streams - s0);
            uint32_t *pool = reinterpret_cast<uint32_t *>(sl.d_frames + s0 * dst_chain);
            const int blocks = std::min(p->n_sms * 8, std::max(1, (int)(((long long)ns * t1 * wpf + 255) / 256)));
            if (h) {
                carry_kernel<<<blocks, 256, 0, p->stream>>>(pool, reinterpret_cast<uint32_t *>(p->d_carry), sl.d_ids + s0, ns, wpf, (int)fps_dev, (int)h, (int)t1, 0);
                p->launches++;
            }
            carry_kernel<<<blocks, 256, 0, p->stream>>>(pool, reinterpret_cast<uint32_t *>(p->d_carry), sl.d_ids + s0, ns, wpf, (int)fps_dev, (int)h, (int)t1, 1);
            p->launches++;
            COVA_CUDA_SLOT(cudaGetLastError());
        }
        if (p->cur_windows) {
            if ((rc = run_chunk(p, c))) return fail(rc);
            COVA_CUDA_SLOT(cudaMemcpyAsync(sl.h_cursor + 2 * c, sl.ccl.d_cursor, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
        }
        COVA_CUDA_SLOT(cudaEventRecord(sl.ev_done[c], p->stream));
    }
#undef COVA_CUDA_SLOT
    return COVA_OK;
}

extern "C" int cova_pipeline_submit_host(cova_pipeline *p, const uint8_t *frames, uint32_t n_streams, uint32_t fps) {
    return submit_impl(p, frames, n_streams, fps, nullptr, nullptr, 0, false);
}
extern "C" int cova_pipeline_submit_host2(cova_pipeline *p, const uint8_t *frames, uint32_t n_streams, uint32_t fps,
                                          const uint32_t *stream_ids, const uint64_t *pts, uint32_t flags) {
    return submit_impl(p, frames, n_streams, fps, stream_ids, pts, flags, true);
}
extern "C" int cova_pipeline_reset_streams(cova_pipeline *p, const uint32_t *stream_ids, uint32_t n) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    if (p->n_seen.empty()) return COVA_OK;
    if (!stream_ids) { std::fill(p->n_seen.begin(), p->n_seen.end(), 0); return COVA_OK; }
    for (uint32_t i = 0; i < n; i++) {
        if (stream_ids[i] >= p->max_streams) return set_err(COVA_E_INVAL, "stream id out of range");
        p->n_seen[stream_ids[i]] = 0;
    }
    return COVA_OK;
}

static int collect_impl(cova_pipeline *p, uint8_t *blob, size_t blob_cap, size_t *blob_len, uint64_t *offsets, uint64_t *lens,
                        uint32_t *n_windows, uint32_t *win_stream_ids, uint64_t *win_pts) {
    if (!p || !blob_len) return set_err(COVA_E_INVAL, "null argument");
    COVA_CUDA(cudaSetDevice(p->device));
    const int k = p->next_collect;
    auto &sl = p->slot[k];
    if (!sl.busy) return set_err(COVA_E_INVAL, "no batch in flight");
    sl.busy = false;
    p->next_collect = (k + 1) % kSlots;
    if (n_windows) *n_windows = sl.n_windows;
    *blob_len = 0;
    if (sl.failed) {
        sl.failed = false;
        cudaStreamSynchronize(p->stream);
        cudaGetLastError();
        return set_err(COVA_E_CUDA, "the submit of this batch failed part-way; its results are void");
    }
    if (win_stream_ids && !sl.win_ids.empty()) memcpy(win_stream_ids, sl.win_ids.data(), sl.win_ids.size() * sizeof(uint32_t));
    if (win_pts && !sl.win_pts.empty()) memcpy(win_pts, sl.win_pts.data(), sl.win_pts.size() * sizeof(uint64_t));
    const uint32_t nc = (sl.n_streams + p->chunk_streams - 1) / p->chunk_streams;
    if (!sl.n_windows) {
        // nothing to copy, but the batch's device work (carry-over of a short CONTINUED batch) must have completed before
        // the caller may reuse its frame buffer
        if (nc && cudaEventSynchronize(sl.ev_done[nc - 1]) != cudaSuccess) return sync_stream(p, p->stream);
        return COVA_OK;
    }
    const uint32_t wps = sl.wps;
    unsigned long long prev = 0;
    bool too_small = false;
    for (uint32_t c = 0; c < nc; c++) {
        if (cudaEventSynchronize(sl.ev_done[c]) != cudaSuccess) return sync_stream(p, p->stream);
        if (sl.h_cursor[2 * c + 1]) return set_err(COVA_E_CUDA, "device box arena overflow (internal sizing error)");
        const unsigned long long end = sl.h_cursor[2 * c];
        const size_t w0 = (size_t)c * p->chunk_streams * wps;
        const size_t nw = (size_t)std::min<uint32_t>(p->chunk_streams, sl.n_streams - c * p->chunk_streams) * wps;
        if (offsets) COVA_CUDA(cudaMemcpyAsync(offsets + w0, sl.ccl.d_offsets + w0, nw * sizeof(uint64_t), cudaMemcpyDeviceToHost, p->s_out));
        if (lens) COVA_CUDA(cudaMemcpyAsync(lens + w0, sl.ccl.d_lens + w0, nw * sizeof(uint64_t), cudaMemcpyDeviceToHost, p->s_out));
        if (blob && end <= blob_cap) {
            if (end > prev) COVA_CUDA(cudaMemcpyAsync(blob + prev, sl.ccl.d_blob + prev, (size_t)(end - prev), cudaMemcpyDeviceToHost, p->s_out));
        } else {
            too_small = true;
        }
        prev = end;
    }
    COVA_CUDA(cudaStreamSynchronize(p->s_out));
    *blob_len = (size_t)prev;
    if (too_small) return set_err(COVA_E_TOOSMALL, "box blob needs a larger buffer");
    return COVA_OK;
}

extern "C" int cova_pipeline_collect_host(cova_pipeline *p, uint8_t *blob, size_t blob_cap, size_t *blob_len, uint64_t *offsets,
                                          uint64_t *lens, uint32_t *n_windows) {
    return collect_impl(p, blob, blob_cap, blob_len, offsets, lens, n_windows, nullptr, nullptr);
}
extern "C" int cova_pipeline_collect_host2(cova_pipeline *p, uint8_t *blob, size_t blob_cap, size_t *blob_len, uint64_t *offsets,
                                           uint64_t *lens, uint32_t *n_windows, uint32_t *win_stream_ids, uint64_t *win_pts) {
    return collect_impl(p, blob, blob_cap, blob_len, offsets, lens, n_windows, win_stream_ids, win_pts);
}

extern "C" int cova_pipeline_process_host(cova_pipeline *p, const uint8_t *frames, uint32_t n_streams, uint32_t fps, uint8_t *blob,
                                          size_t blob_cap, size_t *blob_len, uint64_t *offsets, uint64_t *lens, uint32_t *n_windows) {
    if (!p || !frames || !blob_len) return set_err(COVA_E_INVAL, "null argument");
    for (auto &sl : p->slot)
        if (sl.busy) return set_err(COVA_E_INVAL, "batches submitted asynchronously are still in flight");
    p->next_submit = p->next_collect = 0;
    prof_begin(p);
    int rc = cova_pipeline_submit_host(p, frames, n_streams, fps);
    if (!rc) rc = cova_pipeline_collect_host(p, blob, blob_cap, blob_len, offsets, lens, n_windows);
    else if (p->slot[0].busy) { p->slot[0].busy = false; p->slot[0].failed = false; p->next_submit = p->next_collect = 0; }
    if (!rc) rc = cova_pipeline_sync(p);
    return rc;
}

// ---- inspection ------------------------------------------------------------------------------------
extern "C" int cova_pipeline_read_stacked(cova_pipeline *p, uint8_t *out, size_t cap) {
    if (!p || !out) return set_err(COVA_E_INVAL, "null argument");
    if (!p->d_stacked) return set_err(COVA_E_INVAL, "pipeline was created without COVA_FLAG_KEEP_STACKED");
    size_t bytes = p->frame_bytes * p->T * p->cur_windows;
    if (cap < bytes) return set_err(COVA_E_TOOSMALL, "buffer too small");
    int rc = cova_pipeline_sync(p);
    if (rc) return rc;
    COVA_CUDA(cudaMemcpy(out, p->d_stacked, bytes, cudaMemcpyDeviceToHost));
    return COVA_OK;
}
extern "C" int cova_pipeline_read_mask(cova_pipeline *p, uint8_t *out, size_t cap) {
    if (!p || !out) return set_err(COVA_E_INVAL, "null argument");
    size_t bytes = (size_t)p->cur_windows * p->W * p->H;
    if (cap < bytes) return set_err(COVA_E_TOOSMALL, "buffer too small");
    int rc = cova_pipeline_sync(p);
    if (rc) return rc;
    COVA_CUDA(cudaMemcpy(out, p->d_mask, bytes, cudaMemcpyDeviceToHost));
    return COVA_OK;
}
extern "C" int cova_pipeline_read_logits(cova_pipeline *p, float *out, size_t cap_floats) {
    if (!p || !out) return set_err(COVA_E_INVAL, "null argument");
    if (!p->d_logits) return set_err(COVA_E_INVAL, "pipeline was created without COVA_FLAG_KEEP_LOGITS");
    size_t n = (size_t)p->cur_windows * p->W * p->H;
    if (cap_floats < n) return set_err(COVA_E_TOOSMALL, "buffer too small");
    int rc = cova_pipeline_sync(p);
    if (rc) return rc;
    COVA_CUDA(cudaMemcpy(out, p->d_logits, n * sizeof(float), cudaMemcpyDeviceToHost));
    return COVA_OK;
}

extern "C" int cova_pipeline_read_activation(cova_pipeline *p, int layer, float *out, size_t cap_floats, uint32_t shape[5]) {
    if (!p || !out || !shape) return set_err(COVA_E_INVAL, "null argument");
    if (layer < 0 || layer > 7) return set_err(COVA_E_INVAL, "layer must be 0..7");
    const bool enc_side = layer <= 3;
    const Geom &g = enc_side ? p->gx[layer] : p->gd[layer - 4];
    const uint4 *src = enc_side ? p->x[layer] : p->d[layer - 4];
    const int C = layer == 0 ? 3 : (enc_side ? kEncCout[layer - 1] : kDecCin[layer - 4]);
    const int N = (int)p->ck_windows, Tn = g.Tn;   // activations of the chunk processed last
    shape[0] = N; shape[1] = C; shape[2] = Tn; shape[3] = g.H; shape[4] = g.W;
    size_t n = (size_t)N * C * Tn * g.H * g.W;
    if (cap_floats < n) return set_err(COVA_E_TOOSMALL, "buffer too small");
    int rc = cova_pipeline_sync(p);
    if (rc) return rc;
    if (layer == 0 && !p->x[0]) {
        // BlobNet input as the tcgen05 path sees it: per-frame x-pair-packed rows + the window table
        const Geom &gf = p->gx0f;
        std::vector<Row8> host((size_t)geom_rows(gf));
        COVA_CUDA(cudaMemcpy(host.data(), p->x0f, host.size() * sizeof(Row8), cudaMemcpyDeviceToHost));
        for (int nn = 0; nn < N; nn++)
            for (int c = 0; c < 3; c++)
                for (int t = 0; t < kT; t++)
                    for (int y = 0; y < gf.H; y++)
                        for (int x = 0; x < gf.W; x++) {
                            const int f = p->h_newest[nn] - t;
                            const Row8 &r = host[(size_t)geom_row(gf, 0, (y & 1) << 1, geom_pos(gf, f, y >> 1, x >> 1, 0))];
                            out[((((size_t)nn * 3 + c) * kT + t) * gf.H + y) * gf.W + x] = __half2float(r.v[(x & 1) * 4 + c]);
                        }
        shape[2] = kT;
        return COVA_OK;
    }
    std::vector<Row8> host((size_t)geom_rows(g));
    COVA_CUDA(cudaMemcpy(host.data(), src, host.size() * sizeof(Row8), cudaMemcpyDeviceToHost));
    for (int nn = 0; nn < N; nn++)
        for (int c = 0; c < C; c++)
            for (int t = 0; t < Tn; t++)
                for (int y = 0; y < g.H; y++)
                    for (int x = 0; x < g.W; x++) {
                        const Row8 &r = host[(size_t)geom_row_of(g, nn, t, c >> 3, y, x)];
                        out[((((size_t)nn * C + c) * Tn + t) * g.H + y) * g.W + x] = __half2float(r.v[c & 7]);
                    }
    return COVA_OK;
}

extern "C" int cova_pipeline_set_debug(cova_pipeline *p, int flags) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    p->dbg = flags;
    // bit 5 (kDbgNoPdl): plain stream-ordered launches instead of programmatic dependent launch, for THIS handle only
    return COVA_OK;
}
extern "C" int cova_pipeline_launch_count(const cova_pipeline *p, uint64_t *count) {
    if (!p || !count) return set_err(COVA_E_INVAL, "null argument");
    *count = p->launches;
    return COVA_OK;
}
extern "C" int cova_pipeline_set_profiling(cova_pipeline *p, int enable) {
    if (!p) return set_err(COVA_E_INVAL, "null handle");
    p->profiling = enable != 0;
    p->ev_used = 0;
    p->names.clear();
    return COVA_OK;
}
extern "C" int cova_pipeline_last_timings(cova_pipeline *p, char *names, size_t names_cap, float *ms, uint32_t *n) {
    if (!p || !n) return set_err(COVA_E_INVAL, "null argument");
    uint32_t cap = *n;
    *n = (uint32_t)p->last_ms.size();
    std::string joined;
    for (size_t i = 0; i < p->last_names.size(); i++) { if (i) joined += ";"; joined += p->last_names[i]; }
    if (names && names_cap) { strncpy(names, joined.c_str(), names_cap - 1); names[names_cap - 1] = 0; }
    if (ms) for (uint32_t i = 0; i < std::min<uint32_t>(cap, *n); i++) ms[i] = p->last_ms[i];
    return COVA_OK;
}
