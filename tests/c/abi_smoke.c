/* Plain-C consumer of include/cova_b200.h: what a Rust / cgo / JNI host links against.  No Python, no CUDA headers.
 *   gcc -std=c11 -Wall -Wextra -Werror -Iinclude tests/c/abi_smoke.c -Lcova_b200 -lcova_b200 -Wl,-rpath,$PWD/cova_b200 -o /tmp/abi_smoke
 * Without a GPU it checks that every constructor fails with COVA_E_NODEVICE (no CPU fallback) and exercises the host-only
 * entry points (packer, gopsplit ranges, sorttracker).  With a GPU it runs the element shims and one small batch through
 * the fused path with all-zero weights refused / a synthetic container accepted, and prints "abi_smoke ok". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cova_b200.h"

#define CHECK(cond)                                                                     \
    do {                                                                                \
        if (!(cond)) { fprintf(stderr, "abi_smoke: %s failed at line %d: %s\n", #cond, __LINE__, cova_last_error()); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    printf("%s\n", cova_version());
    int n_dev = -1;
    CHECK(cova_device_count(&n_dev) == COVA_OK && n_dev >= 0);

    /* host-only pieces work with or without a device */
    cova_packer *pk = NULL;
    CHECK(cova_packer_new(&pk, 2) == COVA_OK && pk);
    uint8_t quads[4 * 40];
    uint16_t packed[40];
    for (int i = 0; i < 40; i++) { quads[4 * i] = (uint8_t)i; quads[4 * i + 1] = (uint8_t)(3 * i); quads[4 * i + 2] = 1; quads[4 * i + 3] = 0xAB; }
    CHECK(cova_packer_pack(pk, quads, packed, 40) == COVA_OK);
    for (int i = 0; i < 40; i++) {
        unsigned a = quads[4 * i] < 6 ? quads[4 * i] : 6, b = quads[4 * i + 1] < 6 ? quads[4 * i + 1] : 6;
        CHECK(packed[i] == (a | (b << 3) | (1u << 6)));
    }
    cova_packer_free(pk);
    uint32_t flags[10] = {0, 1, 1, 0, 1, 1, 1, 0, 1, 1};      /* key frames at 0, 3, 7 */
    uint64_t first[2], end[2];
    CHECK(cova_gopsplit_ranges(flags, 10, 2, first, end) == COVA_OK);
    CHECK(first[0] == 0 && end[0] == 3 && first[1] == 3 && end[1] == 10);   /* floor(3/2) GoPs, remainder to the last pad */
    cova_sorttracker *st = NULL;
    CHECK(cova_sorttracker_new(&st) == COVA_OK);
    CHECK(cova_sorttracker_set_property(st, "maxage", 5) == COVA_OK && cova_sorttracker_set_caps(st, 80, 45) == COVA_OK);
    uint8_t empty_boxes[8] = {0}, out[64];
    size_t out_len = 0;
    CHECK(cova_sorttracker_transform(st, empty_boxes, 8, 0, out, sizeof(out), &out_len) == COVA_OK && out_len == 8);
    cova_sorttracker_free(st);

    cova_metapreprocess *mp = NULL;
    cova_bboxcc *cc = NULL;
    if (n_dev == 0) {
        CHECK(cova_metapreprocess_new(&mp, 0, 1280, 720, 4, 1) == COVA_E_NODEVICE && !mp);
        CHECK(cova_bboxcc_new(&cc, 0, 80, 45, 30) == COVA_E_NODEVICE && !cc);
        CHECK(strstr(cova_last_error(), "no CPU fallback") != NULL);
        printf("abi_smoke ok (no device: constructors refuse, host-side entry points work)\n");
        return 0;
    }
    (void)argc; (void)argv;
    /* element shims on the device */
    CHECK(cova_metapreprocess_new(&mp, 0, 1280, 720, 4, 1) == COVA_OK);
    uint32_t w = 0, h = 0; size_t size = 0;
    CHECK(cova_metapreprocess_out_caps(mp, &w, &h, &size) == COVA_OK && w == 80 && h == 180 && size == 57600);
    uint8_t *frame = malloc(14400), *stack = malloc(57600);
    int n_ok = 0;
    for (int f = 0; f < 6; f++) {
        memset(frame, f + 1, 14400);
        int rc = cova_metapreprocess_transform(mp, frame, 14400, stack, 57600);
        CHECK(rc == COVA_OK || rc == COVA_DROPPED);
        if (rc == COVA_OK) { n_ok++; CHECK(stack[0] == f + 1 && stack[14400] == f && stack[3 * 14400] == f - 2); }   /* newest first */
    }
    CHECK(n_ok == 3);                                             /* the first timestep-1 buffers are dropped */
    cova_metapreprocess_free(mp);
    CHECK(cova_bboxcc_new(&cc, 0, 80, 45, 1) == COVA_OK);
    uint8_t *mask = calloc(3600, 1), blob[64];
    mask[80 * 10 + 20] = mask[80 * 10 + 21] = mask[80 * 11 + 21] = 1;    /* one 3-pixel component: box (20,10) 2x2 */
    size_t len = 0;
    CHECK(cova_bboxcc_transform_ip(cc, mask, 3600, blob, sizeof(blob), &len) == COVA_OK && len == 32);
    float v[5];
    memcpy(v, blob + 8, sizeof(v));
    CHECK(blob[0] == 1 && v[0] == 20.f && v[1] == 10.f && v[2] == 2.f && v[3] == 2.f && v[4] == 4.f);
    cova_bboxcc_free(cc);
    free(frame); free(stack); free(mask);
    printf("abi_smoke ok\n");
    return 0;
}
