// Host-side packer for the 2-byte input format (include/cova_b200.h, COVA_FLAG_INPUT_PACKED16).  Plain C++17, no CUDA.
//
// The decoder writes 4 bytes per macroblock - mb_weight, |mv_x|, |mv_y| and one stale byte
// (third_parties/FFmpeg/libavcodec/h264_mb.c:822-855) - and BlobNet's first operation is clip(x, 0, 6)
// (utils/model/preprocessing.py:5-8), so 9 bits per macroblock carry everything the network can see.  Packing on the
// host halves the bytes that cross PCIe, which is what bounds the end-to-end rate of the path ("Measurement").
//     u16 = min(b0, 6) | min(b1, 6) << 3 | min(b2, 6) << 6
// A packer owns its worker threads for its lifetime (a thread start per batch would cost as much as the packing).
#pragma once
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace cova {
namespace host {

inline void pack_scalar(const uint8_t *q, uint16_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        const uint32_t a = q[4 * i], b = q[4 * i + 1], c = q[4 * i + 2];
        out[i] = (uint16_t)((a < 6 ? a : 6) | ((b < 6 ? b : 6) << 3) | ((c < 6 ? c : 6) << 6));
    }
}

#if defined(__x86_64__)
// 16 macroblocks per iteration: byte-wise min with 6, then (b0 + 8*b1) and (64*b2 + 0*b3) by one multiply-add of adjacent
// bytes, their sum by one multiply-add of adjacent words, and a saturating pack 32 -> 16 bits.
__attribute__((target("avx2"))) inline void pack_avx2(const uint8_t *q, uint16_t *out, size_t n) {
    const __m256i six = _mm256_set1_epi8(6);
    const __m256i coef = _mm256_set1_epi32(0x00400801);       // bytes 1, 8, 64, 0
    const __m256i ones = _mm256_set1_epi16(1);
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(q + 4 * i));
        __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(q + 4 * i + 32));
        a = _mm256_madd_epi16(_mm256_maddubs_epi16(_mm256_min_epu8(a, six), coef), ones);
        b = _mm256_madd_epi16(_mm256_maddubs_epi16(_mm256_min_epu8(b, six), coef), ones);
        __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);   // packus interleaves the 128-bit lanes
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(out + i), p);
    }
    pack_scalar(q + 4 * i, out + i, n - i);
}
inline bool have_avx2() { return __builtin_cpu_supports("avx2"); }
#else
inline bool have_avx2() { return false; }
#endif

inline void pack_range(const uint8_t *q, uint16_t *out, size_t n) {
#if defined(__x86_64__)
    if (have_avx2()) { pack_avx2(q, out, n); return; }
#endif
    pack_scalar(q, out, n);
}

class Packer {
   public:
    explicit Packer(unsigned n_threads) {
        if (!n_threads) n_threads = std::thread::hardware_concurrency();
        if (!n_threads) n_threads = 1;
        if (n_threads > 64) n_threads = 64;
        n_ = n_threads;
        for (unsigned t = 1; t < n_; t++) workers_.emplace_back([this, t] { loop(t); });   // the caller is worker 0
    }
    ~Packer() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    unsigned threads() const { return n_; }
    void pack(const uint8_t *q, uint16_t *out, size_t n) {
        if (n_ == 1 || n < 65536) { pack_range(q, out, n); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            q_ = q; out_ = out; total_ = n; pending_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        share(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

   private:
    void share(unsigned t) {
        // contiguous shares in multiples of 64 macroblocks (whole cache lines of output)
        const size_t per = ((total_ + n_ - 1) / n_ + 63) / 64 * 64;
        const size_t lo = std::min(total_, per * t), hi = std::min(total_, lo + per);
        if (hi > lo) pack_range(q_ + 4 * lo, out_ + lo, hi - lo);
    }
    void loop(unsigned t) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            share(t);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    unsigned n_ = 1;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const uint8_t *q_ = nullptr;
    uint16_t *out_ = nullptr;
    size_t total_ = 0;
    unsigned pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

}  // namespace host
}  // namespace cova
