/* tools/swdec_cli.c — command-line decode harness over the H264SwDec* API
 * (Decoder/inc/H264SwDecApi.h:142-173 in the reference).
 *
 * The same source links against EITHER implementation of that API:
 *   oracle/_ref/refdec        -> oracle/_ref/libh264ref.so (unmodified reference)
 *   broadway_b200/bin/b200dec -> broadway_b200/libh264b200.so (this repo, CUDA)
 * which is the drop-in property in executable form.  Behaviour mirrors the
 * reference's DecTestBench whole-stream mode (DecTestBench.c:213-400): feed the
 * remaining buffer, advance by pStrmCurrPos, drain H264SwDecNextPicture after
 * every PIC_RDY, flush at end of stream.
 *
 *   swdec_cli [-m] [-o out.yuv] [-r reps] [-n maxpics] [-R] file.264
 *     -m  print "frame <i> <md5>" for every output picture (Y|U|V, MB aligned)
 *     -o  append raw I420 frames to a file
 *     -r  decode the stream <reps> times (timing; output only on first pass)
 *     -R  noOutputReordering=1
 *   last line: JSON {"frames":..,"seconds":..,"fps":..,"err_mbs":..,"width":..,"height":..}
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#ifdef USE_B200
#include "h264b200_swdec.h"   /* this repo's declaration of the same API */
#else
#include "H264SwDecApi.h"     /* the reference's own header */
#endif

/* ---- MD5 (RFC 1321), written for this harness ---- */
typedef struct { uint32_t h[4]; uint64_t len; unsigned char buf[64]; unsigned fill; } md5_t;
static const uint32_t MD5K[64] = {
 0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,
 0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
 0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
 0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
 0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,
 0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
 0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,
 0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
static const unsigned char MD5S[64] = {
 7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22, 5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
 4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23, 6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21};
static void md5_block(md5_t *m, const unsigned char *p)
{
    uint32_t w[16], a = m->h[0], b = m->h[1], c = m->h[2], d = m->h[3];
    int i;
    for (i = 0; i < 16; i++)
        w[i] = (uint32_t)p[4*i] | ((uint32_t)p[4*i+1] << 8) | ((uint32_t)p[4*i+2] << 16) | ((uint32_t)p[4*i+3] << 24);
    for (i = 0; i < 64; i++) {
        uint32_t f; int g;
        if (i < 16)      { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5*i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d;          g = (3*i + 5) & 15; }
        else             { f = c ^ (b | ~d);       g = (7*i) & 15; }
        f += a + MD5K[i] + w[g];
        a = d; d = c; c = b;
        b += (f << MD5S[i]) | (f >> (32 - MD5S[i]));
    }
    m->h[0] += a; m->h[1] += b; m->h[2] += c; m->h[3] += d;
}
static void md5_init(md5_t *m)
{ m->h[0]=0x67452301; m->h[1]=0xefcdab89; m->h[2]=0x98badcfe; m->h[3]=0x10325476; m->len=0; m->fill=0; }
static void md5_update(md5_t *m, const unsigned char *p, size_t n)
{
    m->len += n;
    if (m->fill) {
        while (n && m->fill < 64) { m->buf[m->fill++] = *p++; n--; }
        if (m->fill == 64) { md5_block(m, m->buf); m->fill = 0; }
    }
    while (n >= 64) { md5_block(m, p); p += 64; n -= 64; }
    while (n) { m->buf[m->fill++] = *p++; n--; }
}
static void md5_final(md5_t *m, char hex[33])
{
    uint64_t bits = m->len * 8; unsigned char pad = 0x80, z = 0, lenb[8]; int i;
    md5_update(m, &pad, 1);
    while (m->fill != 56) md5_update(m, &z, 1);
    for (i = 0; i < 8; i++) lenb[i] = (unsigned char)(bits >> (8*i));
    md5_update(m, lenb, 8);
    for (i = 0; i < 16; i++) sprintf(hex + 2*i, "%02x", (m->h[i >> 2] >> (8 * (i & 3))) & 0xff);
}

static double now_s(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

typedef struct { int md5; FILE *out; unsigned frames; unsigned err_mbs; size_t pic_size; } sink_t;

static void drain(H264SwDecInst inst, sink_t *s, int eos, int emit)
{
    H264SwDecPicture pic;
    while (H264SwDecNextPicture(inst, &pic, eos) == H264SWDEC_PIC_RDY) {
        if (emit) {
            if (s->md5) {
                md5_t m; char hex[33];
                md5_init(&m); md5_update(&m, (unsigned char *)pic.pOutputPicture, s->pic_size); md5_final(&m, hex);
                printf("frame %u %s id=%u idr=%u err=%u\n", s->frames, hex, pic.picId, pic.isIdrPicture, pic.nbrOfErrMBs);
            }
            if (s->out) fwrite(pic.pOutputPicture, 1, s->pic_size, s->out);
        }
        s->frames++;
        s->err_mbs += pic.nbrOfErrMBs;
    }
}

int main(int argc, char **argv)
{
    const char *in_name = NULL, *out_name = NULL;
    int md5 = 0, reps = 1, no_reorder = 0, i; unsigned maxpics = 0;
    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-m")) md5 = 1;
        else if (!strcmp(argv[i], "-R")) no_reorder = 1;
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out_name = argv[++i];
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) reps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-n") && i + 1 < argc) maxpics = (unsigned)atoi(argv[++i]);
        else in_name = argv[i];
    }
    if (!in_name) { fprintf(stderr, "usage: %s [-m] [-o out.yuv] [-r reps] [-n maxpics] [-R] file.264\n", argv[0]); return 2; }
    FILE *f = fopen(in_name, "rb");
    if (!f) { perror(in_name); return 2; }
    fseek(f, 0, SEEK_END); long size = ftell(f); fseek(f, 0, SEEK_SET);
    unsigned char *orig = malloc(size + 64), *work = malloc(size + 64);
    if (fread(orig, 1, size, f) != (size_t)size) { perror("read"); return 2; }
    fclose(f);
    memset(orig + size, 0, 64);

    unsigned total_frames = 0, total_err = 0, width = 0, height = 0;
    double t_total = 0;
    int rep, rc = 0;
    for (rep = 0; rep < reps; rep++) {
        sink_t s; memset(&s, 0, sizeof s);
        int emit = rep == 0;
        s.md5 = md5;
        if (emit && out_name) s.out = fopen(out_name, "wb");
        /* the decoder edits the buffer in place (emulation prevention removal) */
        memcpy(work, orig, size + 64);
        double t0 = now_s();
        H264SwDecInst inst;
        if (H264SwDecInit(&inst, no_reorder) != H264SWDEC_OK) { fprintf(stderr, "init failed\n"); return 1; }
        H264SwDecInput in; H264SwDecOutput out; H264SwDecInfo info;
        memset(&in, 0, sizeof in);
        in.pStream = work; in.dataLen = (u32)size; in.picId = 0;
        unsigned decoded = 0;
        while (in.dataLen > 0) {
            H264SwDecRet ret = H264SwDecDecode(inst, &in, &out);
            if (ret == H264SWDEC_HDRS_RDY_BUFF_NOT_EMPTY) {
                if (H264SwDecGetInfo(inst, &info) != H264SWDEC_OK) { rc = 1; break; }
                width = info.picWidth; height = info.picHeight;
                s.pic_size = (size_t)width * height * 3 / 2;
                in.dataLen -= (u32)(out.pStrmCurrPos - in.pStream); in.pStream = out.pStrmCurrPos;
            } else if (ret == H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY || ret == H264SWDEC_PIC_RDY) {
                if (ret == H264SWDEC_PIC_RDY) in.dataLen = 0;
                else { in.dataLen -= (u32)(out.pStrmCurrPos - in.pStream); in.pStream = out.pStrmCurrPos; }
                decoded++; in.picId = decoded;
                if (maxpics && decoded == maxpics) in.dataLen = 0;
                drain(inst, &s, 0, emit);
            } else if (ret == H264SWDEC_STRM_PROCESSED || ret == H264SWDEC_STRM_ERR) {
                in.dataLen = 0;
            } else { fprintf(stderr, "fatal decode error %d\n", (int)ret); rc = 1; break; }
        }
        drain(inst, &s, 1, emit);
        H264SwDecRelease(inst);
        t_total += now_s() - t0;
        if (s.out) fclose(s.out);
        total_frames += s.frames; total_err += s.err_mbs;
    }
    printf("{\"frames\": %u, \"seconds\": %.6f, \"fps\": %.3f, \"err_mbs\": %u, \"width\": %u, \"height\": %u}\n",
           total_frames, t_total, t_total > 0 ? total_frames / t_total : 0.0, total_err, width, height);
    free(orig); free(work);
    return rc ? rc : (total_err ? 1 : 0);
}
