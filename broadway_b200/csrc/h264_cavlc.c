/* h264_cavlc.c — CAVLC residual block decoding (ITU-T H.264 9.2).
 * Same job as the reference's h264bsdDecodeResidualBlockCavlc and its
 * DecodeCoeffToken / DecodeLevelPrefix / DecodeTotalZeros / DecodeRunBefore
 * helpers (h264bsd_cavlc.c:395-915), redesigned for speed: the code tables of
 * cavlc_tables.h are compiled at start-up into leading-zero indexed look-up
 * tables, levels are decoded from a 64-bit bit cache, and coefficients are
 * written straight into the 16 x int16 raster-order slot the GPU transform
 * kernel consumes (the reference writes i32 at zig-zag positions and
 * un-zig-zags later, h264bsd_cavlc.c:897-905, h264bsd_transform.c:118-153).
 */
#include <stdlib.h>
#include <stdio.h>
#include "h264_internal.h"
#include "cavlc_tables.h"

typedef struct { uint8_t len, tc, t1, pad; } ct_entry_t;

/* coeff_token, nC classes 0..2: index = leading_zeros*8 + (3 bits after the first 1) */
static ct_entry_t g_ct[3][16 * 8];
static ct_entry_t g_ct_cdc[256];          /* chroma DC: 8-bit direct */
static uint8_t g_tz[15][512][2];          /* total_zeros: [tc-1][9 bits] -> {len, value} */
static uint8_t g_tz_cdc[3][8][2];
static uint8_t g_rb[6][8][2];             /* run_before, zerosLeft 1..6: [zl-1][3 bits] -> {len, run} */
static int g_init;

static int bitlen(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

void h264_cavlc_init(void)
{
    int t, tc, t1, i;
    if (g_init) return;
    for (t = 0; t < 3; t++) for (tc = 0; tc <= 16; tc++) for (t1 = 0; t1 < 4; t1++) {
        vlc_code_t c = H264_COEFF_TOKEN[t][tc][t1];
        int lz, rem, s;
        if (!c.len) continue;
        lz = c.len - bitlen(c.code);
        rem = bitlen(c.code) - 1;
        if (rem > 3 || lz > 15) { fprintf(stderr, "h264b200: coeff_token table shape unexpected\n"); abort(); }
        for (s = 0; s < 8; s++) if ((s >> (3 - rem)) == (int)(c.code & ((1u << rem) - 1))) {
            ct_entry_t *e = &g_ct[t][lz * 8 + s];
            if (e->len) { fprintf(stderr, "h264b200: coeff_token LUT conflict\n"); abort(); }
            e->len = c.len; e->tc = (uint8_t)tc; e->t1 = (uint8_t)t1;
        }
    }
    for (tc = 0; tc <= 4; tc++) for (t1 = 0; t1 < 4; t1++) {
        vlc_code_t c = H264_COEFF_TOKEN_CHROMA_DC[tc][t1];
        if (!c.len) continue;
        for (i = 0; i < (1 << (8 - c.len)); i++) {
            ct_entry_t *e = &g_ct_cdc[(c.code << (8 - c.len)) | i];
            e->len = c.len; e->tc = (uint8_t)tc; e->t1 = (uint8_t)t1;
        }
    }
    for (tc = 1; tc <= 15; tc++) for (t = 0; t < 16; t++) {
        vlc_code_t c = H264_TOTAL_ZEROS[tc - 1][t];
        if (!c.len) continue;
        for (i = 0; i < (1 << (9 - c.len)); i++) {
            g_tz[tc - 1][(c.code << (9 - c.len)) | i][0] = c.len;
            g_tz[tc - 1][(c.code << (9 - c.len)) | i][1] = (uint8_t)t;
        }
    }
    for (tc = 1; tc <= 3; tc++) for (t = 0; t < 4; t++) {
        vlc_code_t c = H264_TOTAL_ZEROS_CHROMA_DC[tc - 1][t];
        if (!c.len) continue;
        for (i = 0; i < (1 << (3 - c.len)); i++) {
            g_tz_cdc[tc - 1][(c.code << (3 - c.len)) | i][0] = c.len;
            g_tz_cdc[tc - 1][(c.code << (3 - c.len)) | i][1] = (uint8_t)t;
        }
    }
    for (t = 1; t <= 6; t++) for (i = 0; i < 15; i++) {
        vlc_code_t c = H264_RUN_BEFORE[t - 1][i];
        int k;
        if (!c.len) continue;
        for (k = 0; k < (1 << (3 - c.len)); k++) {
            g_rb[t - 1][(c.code << (3 - c.len)) | k][0] = c.len;
            g_rb[t - 1][(c.code << (3 - c.len)) | k][1] = (uint8_t)i;
        }
    }
    g_init = 1;
}

int h264_cavlc_block(br_t *b, int nc, int max_coeff, int16_t *out, const uint8_t *scan)
{
    int tc, t1, i, sl, zeros_left, pos;
    int level[16];

    /* ---- coeff_token ---- */
    if (nc < 0) {
        ct_entry_t e = g_ct_cdc[br_peek(b, 8)];
        if (!e.len) return -1;
        br_skip(b, e.len); tc = e.tc; t1 = e.t1;
    } else if (nc < 8) {
        uint32_t v;
        int lz;
        ct_entry_t e;
        if (b->bits < 32) br_refill(b);
        v = (uint32_t)(b->cache >> 32);
        if (v < 0x10000u) return -1;              /* more than 15 leading zeros: no such code */
        lz = __builtin_clz(v);
        e = g_ct[nc < 2 ? 0 : nc < 4 ? 1 : 2][lz * 8 + ((v >> (28 - lz)) & 7)];
        if (!e.len) return -1;
        br_skip(b, e.len); tc = e.tc; t1 = e.t1;
    } else {
        uint32_t v = br_get(b, 6);
        if (v == 3) { tc = 0; t1 = 0; }
        else { tc = (int)(v >> 2) + 1; t1 = (int)(v & 3); if (t1 > tc) return -1; }
    }
    if (tc == 0) return 0;
    if (tc > max_coeff) return -1;

    /* ---- levels ---- */
    sl = (tc > 10 && t1 < 3) ? 1 : 0;
    if (t1) {
        uint32_t s = br_get(b, t1);
        for (i = 0; i < t1; i++) level[i] = ((s >> (t1 - 1 - i)) & 1) ? -1 : 1;
    }
    for (i = t1; i < tc; i++) {
        uint32_t v;
        int prefix, code, lv;
        if (b->bits < 32) br_refill(b);
        v = (uint32_t)(b->cache >> 32);
        if (v < 0x10000u) return -1;              /* level_prefix > 15: not Baseline (h264bsd_cavlc.c:513-514) */
        prefix = __builtin_clz(v);
        br_skip(b, prefix + 1);
        code = (prefix < 15 ? prefix : 15) << sl;
        if (sl > 0 || prefix >= 14) {
            int size = (prefix == 14 && sl == 0) ? 4 : prefix >= 15 ? 12 : sl;
            code += (int)br_get(b, size);
        }
        if (prefix >= 15 && sl == 0) code += 15;
        if (i == t1 && t1 < 3) code += 2;
        lv = (code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1;
        level[i] = lv;
        if (sl == 0) sl = 1;
        if ((lv < 0 ? -lv : lv) > (3 << (sl - 1)) && sl < 6) sl++;
    }

    /* ---- total_zeros ---- */
    if (tc < max_coeff) {
        if (nc < 0) {
            const uint8_t *e = g_tz_cdc[tc - 1][br_peek(b, 3)];
            if (!e[0]) return -1;
            br_skip(b, e[0]); zeros_left = e[1];
        } else {
            const uint8_t *e = g_tz[tc - 1][br_peek(b, 9)];
            if (!e[0]) return -1;
            br_skip(b, e[0]); zeros_left = e[1];
        }
        if (zeros_left + tc > max_coeff) return -1;
    } else zeros_left = 0;

    /* ---- run_before + placement (highest frequency first) ---- */
    pos = zeros_left + tc - 1;
    for (i = 0; i < tc; i++) {
        int run = 0;
        out[scan[pos]] = (int16_t)level[i];
        if (i == tc - 1) break;
        if (zeros_left > 0) {
            if (zeros_left <= 6) {
                const uint8_t *e = g_rb[zeros_left - 1][br_peek(b, 3)];
                br_skip(b, e[0]); run = e[1];
            } else {
                uint32_t v = br_peek(b, 11);
                if (v >> 8) { run = 7 - (int)(v >> 8); br_skip(b, 3); }
                else {
                    int lz;
                    if (!v) return -1;
                    lz = __builtin_clz(v) - 21;      /* leading zeros within the 11 bits */
                    run = lz + 4; br_skip(b, lz + 1);
                }
            }
            if (run > zeros_left) return -1;
            zeros_left -= run;
        }
        pos -= run + 1;
    }
    return tc;
}
