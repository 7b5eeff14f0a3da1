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
#include <stddef.h>
#include <string.h>
#include "h264_internal.h"
#include "cavlc_tables.h"

#include "h264_cavlc_inl.h"
#include "h264_consts.h"
#include "kp_types.h"

/* coeff_token, nC classes 0..2: index = leading_zeros*8 + (3 bits after the first 1) */
ct_entry_t g_ct[3][16 * 8];
ct_entry_t g_ct_cdc[256];          /* chroma DC: 8-bit direct */
uint8_t g_tz[15][512][2];          /* total_zeros: [tc-1][9 bits] -> {len, value} */
uint8_t g_tz_cdc[3][8][2];
uint8_t g_rb[7][8][2];             /* run_before, zerosLeft 1..6 and >6: [min(zl,7)-1][3 bits] -> {len, run} (len 0: code longer than 3 bits) */
int8_t  g_lvl[7][256][4];          /* level: [suffixLength][8 bits] -> {level, bits, next suffixLength} (bits 0: longer than 8) */
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
    for (t = 0; t < 7; t++) for (i = 1; i < 256; i++) {      /* level_prefix + level_suffix codes that fit in 8 bits */
        int prefix = 7 - (bitlen((unsigned)i) - 1), size = t, code, len;
        if (t == 0 && prefix >= 14) continue;
        len = prefix + 1 + size;
        if (len > 8) continue;
        code = (prefix << t) + ((i >> (8 - len)) & ((1 << size) - 1));
        g_lvl[t][i][0] = (int8_t)((code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1);
        g_lvl[t][i][1] = (int8_t)len;
        {   /* suffixLength after this level (9.2.2.1), valid when no levelCode adjustment applies */
            int a = g_lvl[t][i][0] < 0 ? -g_lvl[t][i][0] : g_lvl[t][i][0], ns = t ? t : 1;
            if (a > (3 << (ns - 1)) && ns < 6) ns++;
            g_lvl[t][i][2] = (int8_t)ns;
        }
    }
    for (i = 1; i < 8; i++) { g_rb[6][i][0] = 3; g_rb[6][i][1] = (uint8_t)(7 - i); }   /* zerosLeft > 6: 111 -> 0 ... 001 -> 6 */
    g_init = 1;
}


/* The same look-up tables, as one block for the device-side parser (kernel Kp, kp_core.h): the CUDA engine copies it to
 * HBM once; the CPU test build of kp_core.h reads it in place. */
void h264_kp_fill_tables(KpTables *t)
{
    h264_cavlc_init();
    memset(t, 0, sizeof *t);
    memcpy(t->ct, g_ct, sizeof t->ct);
    memcpy(t->ct_cdc, g_ct_cdc, sizeof t->ct_cdc);
    memcpy(t->tz, g_tz, sizeof t->tz);
    memcpy(t->tz_cdc, g_tz_cdc, sizeof t->tz_cdc);
    memcpy(t->rb, g_rb, sizeof t->rb);
    memcpy(t->lvl, g_lvl, sizeof t->lvl);
    memcpy(t->cbp_map, H264_CBP_MAP, sizeof t->cbp_map);
    memcpy(t->zigzag, H264_ZIGZAG4x4, sizeof t->zigzag);
    memcpy(t->raster_to_blk, H264_RASTER_TO_BLK, sizeof t->raster_to_blk);
    memcpy(t->qpc, H264_QPC, sizeof t->qpc);
    {
        int blk;
        for (blk = 0; blk < 16; blk++) t->lc_idx[blk] = (uint8_t)(9 + ((blk & 1) | ((blk >> 1) & 2)) + 8 * (((blk >> 1) & 1) | ((blk >> 2) & 2)));
        for (blk = 0; blk < 4; blk++) t->ident4[blk] = (uint8_t)blk;
        for (blk = 0; blk < 16; blk++) t->cbp_luma[blk] = (uint16_t)(((blk & 1) ? 0x000f : 0) | ((blk & 2) ? 0x00f0 : 0) | ((blk & 4) ? 0x0f00 : 0) | ((blk & 8) ? 0xf000 : 0));
        /* residual steps: 0 Intra16x16 DC (nC of block 0), 1..16 luma, 17/18 chroma DC, 19..26 chroma AC (grid cells 40.. = KpStage.cc) */
        {
            const uint32_t zz = (uint32_t)offsetof(KpTables, zigzag), id = (uint32_t)offsetof(KpTables, ident4);
            t->step_desc[0][0] = (uint32_t)t->lc_idx[0] | (8u << 8) | (0u << 12);
            t->step_desc[0][1] = zz | (16u << 16);
            for (blk = 0; blk < 16; blk++) {
                t->step_desc[1 + blk][0] = (uint32_t)t->lc_idx[blk] | (8u << 8) | (1u << 12) | ((uint32_t)blk << 16);
                t->step_desc[1 + blk][1] = zz | (16u << 16);                       /* Intra16x16: one less, from scan position 1 (kp_parse_residual) */
            }
            t->step_desc[17][0] = t->step_desc[18][0] = 8u | (4u << 8) | (2u << 12);
            t->step_desc[17][1] = id | (4u << 16);
            t->step_desc[18][1] = id | (4u << 16) | (4u << 24);                    /* Cr DC behind Cb DC in the same slot */
            for (blk = 0; blk < 8; blk++) {
                t->step_desc[19 + blk][0] = (uint32_t)(40 + 12 * (blk >> 2) + 5 + (blk & 1) + 4 * ((blk >> 1) & 1)) | (4u << 8) | (3u << 12) | ((uint32_t)(16 + blk) << 16);
                t->step_desc[19 + blk][1] = (zz + 1u) | (15u << 16);
            }
        }
    }
}
