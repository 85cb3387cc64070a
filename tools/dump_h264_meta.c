/* Fixture generator: decodes an H.264 file with the reference's PATCHED avdec_h264
 * (third_parties/FFmpeg, h264_mb.c:822-855 writes 4 bytes per macroblock at data[0]; h264_slice.c:2646-2662
 * skips pixel reconstruction for CABAC streams) and dumps, per decoded frame in output order, the first
 * (W/16)*(H/16)*4 bytes of plane 0 -- exactly the prefix metapreprocess reads (imp.rs:219-234,307-320) --
 * followed by the list of key-frame flags and PTS values.
 *
 * Runs only in the build container (the patched FFmpeg lives under /root/reference); the GPU box and the test
 * suite only ever see the committed fixture.  Built and driven by tools/make_demo_fixture.py.
 *
 * usage: dump_h264_meta in.mp4 out.bin [max_frames]
 * out.bin: int32 {n_frames, w_mb, h_mb, reserved} | n_frames * (w_mb*h_mb*4) u8 | n_frames * u8 key | n_frames * i64 pts
 */
#include <libavcodec/avcodec.h>
#include <libavformat/avformat.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void die(const char *what) {
    fprintf(stderr, "dump_h264_meta: %s\n", what);
    exit(1);
}

int main(int argc, char **argv) {
    if (argc < 3) die("usage: dump_h264_meta in.mp4 out.bin [max_frames]");
    long max_frames = argc > 3 ? atol(argv[3]) : (1L << 30);
    AVFormatContext *fmt = NULL;
    if (avformat_open_input(&fmt, argv[1], NULL, NULL) < 0) die("cannot open input");
    if (avformat_find_stream_info(fmt, NULL) < 0) die("no stream info");
    int vs = av_find_best_stream(fmt, AVMEDIA_TYPE_VIDEO, -1, -1, NULL, 0);
    if (vs < 0) die("no video stream");
    const AVCodec *dec = avcodec_find_decoder(fmt->streams[vs]->codecpar->codec_id);
    if (!dec) die("no decoder");
    AVCodecContext *ctx = avcodec_alloc_context3(dec);
    avcodec_parameters_to_context(ctx, fmt->streams[vs]->codecpar);
    ctx->thread_count = 1; /* pipeline/cova/pipeline.py:91-92: avdec_h264 max-threads=1 */
    if (avcodec_open2(ctx, dec, NULL) < 0) die("cannot open decoder");

    FILE *out = fopen(argv[2], "wb");
    if (!out) die("cannot open output");
    int32_t hdr[4] = {0, 0, 0, 0};
    fwrite(hdr, sizeof hdr, 1, out);

    size_t cap = 4096, n = 0;
    uint8_t *keys = malloc(cap);
    int64_t *pts = malloc(cap * sizeof(int64_t));
    AVPacket *pkt = av_packet_alloc();
    AVFrame *frm = av_frame_alloc();
    int w_mb = 0, h_mb = 0, draining = 0;
    while ((long)n < max_frames) {
        if (!draining) {
            int r = av_read_frame(fmt, pkt);
            if (r < 0) {
                avcodec_send_packet(ctx, NULL);
                draining = 1;
            } else {
                if (pkt->stream_index == vs && avcodec_send_packet(ctx, pkt) < 0) die("send_packet failed");
                av_packet_unref(pkt);
            }
        }
        for (;;) {
            int r = avcodec_receive_frame(ctx, frm);
            if (r == AVERROR(EAGAIN)) break;
            if (r == AVERROR_EOF) goto done;
            if (r < 0) die("receive_frame failed");
            if (!w_mb) {
                w_mb = frm->width / 16; /* metapreprocess/imp.rs:262-268: integer division */
                h_mb = frm->height / 16;
            }
            fwrite(frm->data[0], 1, (size_t)w_mb * h_mb * 4, out);
            if (n == cap) {
                cap *= 2;
                keys = realloc(keys, cap);
                pts = realloc(pts, cap * sizeof(int64_t));
            }
            keys[n] = (uint8_t)frm->key_frame;
            pts[n] = frm->pts;
            n++;
            av_frame_unref(frm);
            if ((long)n >= max_frames) goto done;
        }
    }
done:
    fwrite(keys, 1, n, out);
    fwrite(pts, sizeof(int64_t), n, out);
    hdr[0] = (int32_t)n;
    hdr[1] = w_mb;
    hdr[2] = h_mb;
    fseek(out, 0, SEEK_SET);
    fwrite(hdr, sizeof hdr, 1, out);
    fclose(out);
    fprintf(stderr, "dump_h264_meta: %zu frames, grid %dx%d\n", n, w_mb, h_mb);
    return 0;
}
