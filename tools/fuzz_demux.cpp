// Sanitizer harness for the host-side MP4 / Annex-B parsers (csrc/gop_demux.hpp): mutates the moov fixture and checks
// under AddressSanitizer + UBSan that no mutation reads out of bounds.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -o /tmp/fuzz_demux tools/fuzz_demux.cpp && /tmp/fuzz_demux tests/golden/demo_1m_moov.bin
#include <stdio.h>
#include <stdlib.h>

#include <random>

#include "../cova_b200/csrc/gop_demux.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> moov;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) moov.insert(moov.end(), buf, buf + n);
    fclose(f);
    std::mt19937 rng(1);
    long ok = 0, bad = 0;
    const int iters = argc > 2 ? atoi(argv[2]) : 20000;
    for (int it = 0; it < iters; it++) {
        // exact-size heap copy: ASan sees any read past `len`
        size_t len = (it % 3 == 0) ? 16 + rng() % (moov.size() - 15) : moov.size();
        uint8_t *m = (uint8_t *)malloc(len);
        memcpy(m, moov.data(), len);
        const int flips = 1 + rng() % 6;
        for (int k = 0; k < flips; k++) {
            size_t at = rng() % len;
            if (it % 2) {   // aim at size / count fields: 4-byte big-endian values after a box tag
                static const char *tags[] = {"stsd", "stsz", "stco", "stsc", "stts", "ctts", "stss", "avc1", "avcC", "trak", "mdia"};
                const char *t = tags[rng() % 11];
                for (size_t i = 4; i + 24 < len; i++)
                    if (!memcmp(m + i, t, 4)) { at = i - 4 + 4 * (rng() % 6); break; }
            }
            const uint32_t v = (rng() % 4 == 0) ? 0xFFFFFFF0u + rng() % 16 : rng() % 64;
            for (int b = 0; b < 4 && at + b < len; b++) m[at + b] = (uint8_t)(v >> (24 - 8 * b));
        }
        std::vector<cova::host::Sample> out;
        cova::host::Mp4Info info;
        int rc = -3;
        try { rc = cova::host::mp4_video_samples(m, len, out, info); } catch (const std::bad_alloc &) { rc = -4; }
        (rc == 0 ? ok : bad)++;
        std::vector<cova::host::Sample> au;
        cova::host::annexb_frames(m, len, au);
        free(m);
    }
    printf("fuzz_demux: %ld parsed, %ld rejected, no sanitizer report\n", ok, bad);
    return 0;
}
